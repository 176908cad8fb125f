// tests/helpers/adapter_gpu.cpp - TEST INFRASTRUCTURE.  The reference's own program flow (src/main.cpp:13-46: scene -> accelerator
// -> render_frame -> image) with ONE line changed - `using A = b200_accel<float>` instead of kd_tree_simd_accel - compiled
// against the UNMODIFIED reference headers where they lie (/root/reference/include) and linked with librt_b200.so.  Built in the
// build container by __graft_entry__.build() (tests/helpers/build_adapter_gpu.py) into tests/helpers/_bin/, which travels to the
// GPU box like the other built files; tests/test_gpu_parity.py::test_cpp_adapter_* run it there.
//
//   adapter_gpu <scene.rtsc> <out.bin> [rays.bin]
//
// out.bin: u32 height, u32 width, then height*width*3 floats of the image<float> render_frame returned (get_pixel order), then the
//          same through b200_frame_sequence (two frames), then height*width*3 bytes of b200_render_frame_rgb8, then - with
//          rays.bin (n x 6 floats) - for cull = 0, 1: n x { u32 hit, float distance, u, v, position[3], hit_normal[3], u32 mesh }
// stdout:  one JSON line with the wall-clock times (render_frame as src/main.cpp:16-20 times it, incl. the image<F> conversion).
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <memory>
#include <vector>

#include <raytracer/config.hpp>
#include <raytracer/scene/scene.hpp>
#include <raytracer/render/render.hpp>
#include <b200_accel.hpp>

#include "../../oracle/rtsc_ref_scene.hpp"

using A = b200_accel<float>;                                   // src/main.cpp:37 - the one edited line
static_assert(accelerator<A, float>, "render/accel/accel.hpp:8-12");

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void put_image(std::ofstream& out, const image<float>& img) {
    std::vector<float> row(img.get_width() * 3);
    for (std::size_t y = 0; y < img.get_height(); ++y) {
        for (std::size_t x = 0; x < img.get_width(); ++x) {
            const auto& c = img.get_pixel(y, x);
            row[3 * x] = c.red; row[3 * x + 1] = c.green; row[3 * x + 2] = c.blue;
        }
        out.write(reinterpret_cast<const char*>(row.data()), std::streamsize(row.size() * 4));
    }
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: adapter_gpu scene.rtsc out.bin [rays.bin]\n"); return 2; }
    try {
        static int tag;
        const auto sc = std::make_shared<const scene<float>>(scene_from_rtsc(argv[1], &tag));
        const double t0 = now();
        auto accelerator_ = A(sc);                                                             // src/main.cpp:41
        const double t_ctor = now() - t0;
        double t_first = now();
        auto img = render_frame<A, float>(accelerator_, scheduling_type::BUCKET_TILES);        // src/main.cpp:17
        t_first = now() - t_first;
        double t_frame = 1e30;
        for (int i = 0; i < 5; ++i) {
            const double t = now();
            img = render_frame<A, float>(accelerator_, scheduling_type::BUCKET_TILES);
            t_frame = std::min(t_frame, now() - t);
        }
        std::ofstream out(argv[2], std::ios::binary);
        const std::uint32_t hw[2] = {std::uint32_t(img.get_height()), std::uint32_t(img.get_width())};
        out.write(reinterpret_cast<const char*>(hw), 8);
        put_image(out, img);
        double t_seq = 0;
        {
            b200_frame_sequence<float> seq(accelerator_);
            seq.submit(); seq.submit();
            auto a = seq.next();
            auto b = seq.next();
            put_image(out, a);
            put_image(out, b);
            const int n = 8;
            const double t = now();
            seq.submit();
            for (int i = 1; i < n; ++i) { seq.submit(); (void)seq.next(); }
            (void)seq.next();
            t_seq = (now() - t) / n;
        }
        const auto rgb8 = b200_render_frame_rgb8(accelerator_);
        out.write(reinterpret_cast<const char*>(rgb8.data()), std::streamsize(rgb8.size()));
        std::size_t n_rays = 0;
        if (argc > 3) {
            std::ifstream in(argv[3], std::ios::binary);
            std::vector<char> raw((std::istreambuf_iterator<char>(in)), {});
            n_rays = raw.size() / 24;
            const float* r = reinterpret_cast<const float*>(raw.data());
            for (int cull = 0; cull < 2; ++cull)
                for (std::size_t i = 0; i < n_rays; ++i) {
                    const ray3<float> ray(vec3<float>{r[6 * i], r[6 * i + 1], r[6 * i + 2]}, vec3<float>{r[6 * i + 3], r[6 * i + 4], r[6 * i + 5]});
                    const auto h = cull ? accelerator_.intersect<true>(ray) : accelerator_.intersect<false>(ray);      // accel.hpp:8-12
                    std::uint32_t hit = h ? 1u : 0u, mesh = h ? std::uint32_t(h->mesh_idx) : 0u;
                    float f[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                    if (h) { f[0] = h->distance; f[1] = h->u; f[2] = h->v; f[3] = h->position.x; f[4] = h->position.y; f[5] = h->position.z;
                             f[6] = h->hit_normal.x; f[7] = h->hit_normal.y; f[8] = h->hit_normal.z; }
                    out.write(reinterpret_cast<const char*>(&hit), 4);
                    out.write(reinterpret_cast<const char*>(f), 36);
                    out.write(reinterpret_cast<const char*>(&mesh), 4);
                }
        }
        std::printf("{\"height\": %u, \"width\": %u, \"ctor_s\": %.6f, \"first_render_frame_s\": %.6f, \"render_frame_s\": %.6f, "
                    "\"sequence_frame_s\": %.6f, \"rays\": %zu}\n", hw[0], hw[1], t_ctor, t_first, t_frame, t_seq, n_rays);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "adapter_gpu: %s\n", e.what());
        return 1;
    }
}
