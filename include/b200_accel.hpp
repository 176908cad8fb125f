// b200_accel.hpp - header-only C++23 adapter that plugs the B200 backend (include/rt_b200.h, librt_b200.so) into
// simd-raytracer where kd_tree_simd_accel sits today.
//
// Include it AFTER the reference's own headers (it uses scene<F>, ray3<F>, hit<F>, image<F>, the material / texture
// variants and the config.hpp constants; it includes none of them itself):
//
//     #include <raytracer/render/render.hpp>
//     #include <b200_accel.hpp>
//     using A = b200_accel<float>;                       // was: kd_tree_simd_accel<F, static_cast<F>(epsilon)>   (src/main.cpp:37)
//     auto accelerator = A(std::make_shared<const scene<float>>(scene));                                      // (src/main.cpp:41)
//     auto image = render_frame<A, float>(accelerator, scheduling_type::BUCKET_TILES);                        // (src/main.cpp:17)
//
// What it provides, matched to the reference interface:
//   * the accelerator<A,F> concept (render/accel/accel.hpp:8-12): intersect<true|false>(ray) -> std::optional<hit<F>>,
//     noexcept and re-entrant like kd_tree_simd_accel::intersect (kd_tree_simd.hpp:187-264); one synchronous ray per
//     call goes through rt_trace_closest as a batch of one - a conformance path, not a fast one;
//   * the implicit `scene_ptr` member the render loops dereference (render/render.hpp:21,113,136);
//   * construction from std::shared_ptr<const scene<F>> like kd_tree_simd_accel's ctor (kd_tree_simd.hpp:100);
//   * an overload of render_frame for this accelerator type that runs the whole frame on the GPU (ray generation,
//     traversal, shading, shadow / reflection / refraction / GI loops) through rt_render_frame and returns image<F>,
//     so src/main.cpp's flow is unchanged.  The config.hpp constants are read here and passed as run-time parameters.
//
// F = float only: the device path is FP32 like the reference's shipped configuration (src/main.cpp:36).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <initializer_list>
#include <memory>
#include <optional>
#include <ostream>
#include <stdexcept>
#include <string>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <variant>
#include <vector>

#include "rt_b200.h"

template <typename F>
struct b200_accel {
    static_assert(std::is_same_v<F, float>, "the B200 backend computes in FP32");

    std::shared_ptr<const scene<F>> scene_ptr;          // render/render.hpp:21,113,136 read this member

    struct handle_deleter { void operator()(rt_scene* s) const noexcept { rt_scene_destroy(s); } };
    std::shared_ptr<rt_scene> handle;                   // shared: the reference copies accelerators by value in places
    std::vector<std::size_t> mesh_first_triangle;       // global triangle id -> (mesh, local triangle)
    // page-locked staging frame of the render_frame overload below (pageable memory halves the PCIe rate); shared like `handle`
    struct pinned_deleter { void operator()(float* p) const noexcept { rt_free_pinned(p); } };
    mutable std::shared_ptr<float> staging;

    explicit b200_accel(std::shared_ptr<const scene<F>> sp, std::uint32_t kd_max_depth = 8, std::uint32_t kd_max_leaf_size = 64,
                        int device = 0, std::uint32_t accel_build = RT_ACCEL_BUILD_HOST)
        : scene_ptr(std::move(sp)) {
        const scene<F>& sc = *scene_ptr;
        rt_scene_desc d{};
        d.background[0] = sc.config.background_color.red; d.background[1] = sc.config.background_color.green;
        d.background[2] = sc.config.background_color.blue;
        d.width = static_cast<std::uint32_t>(sc.config.image_width);
        d.height = static_cast<std::uint32_t>(sc.config.image_height);
        d.bucket_size = static_cast<std::uint32_t>(sc.config.bucket_size);
        d.camera_position[0] = sc.viewpoint.position.x; d.camera_position[1] = sc.viewpoint.position.y;
        d.camera_position[2] = sc.viewpoint.position.z;
        for (int i = 0; i < 9; ++i) d.camera_matrix[i] = sc.viewpoint.matrix.m[i];

        std::vector<rt_light_desc> lights;
        for (const auto& l : sc.lights) lights.push_back({{l.position.x, l.position.y, l.position.z}, l.intensity});

        // textures: the reference keys them by name (scene/scene.hpp:19); the ABI indexes them
        std::vector<rt_texture_desc> textures;
        std::vector<std::uint8_t> texels;
        std::unordered_map<std::string, std::int32_t> texture_index;
        for (const auto& [name, tv] : sc.textures) {
            rt_texture_desc t{};
            std::visit([&](const auto& tex) {
                using T = std::decay_t<decltype(tex)>;
                auto put = [](float* dst, const color<F>& c) { dst[0] = c.red; dst[1] = c.green; dst[2] = c.blue; };
                if constexpr (std::is_same_v<T, albedo_texture<F>>) { t.kind = RT_TEX_ALBEDO; put(t.c0, tex.albedo); }
                else if constexpr (std::is_same_v<T, edge_texture<F>>) {
                    t.kind = RT_TEX_EDGES; put(t.c0, tex.edge_color); put(t.c1, tex.inner_color); t.scalar = tex.edge_width;
                } else if constexpr (std::is_same_v<T, checker_texture<F>>) {
                    t.kind = RT_TEX_CHECKER; put(t.c0, tex.color_a); put(t.c1, tex.color_b); t.scalar = tex.square_size;
                } else {
                    // bitmap texels were stored as float(byte) * float(1/255) (scene/texture/bitmap.hpp:19-30): recover the bytes
                    t.kind = RT_TEX_BITMAP;
                    t.bmp_w = static_cast<std::uint32_t>(tex.texture.get_width());
                    t.bmp_h = static_cast<std::uint32_t>(tex.texture.get_height());
                    t.bmp_off = static_cast<std::uint32_t>(texels.size());
                    for (std::size_t r = 0; r < tex.texture.get_height(); ++r)
                        for (std::size_t c = 0; c < tex.texture.get_width(); ++c) {
                            const color<F>& p = tex.texture.get_pixel(r, c);
                            for (F ch : {p.red, p.green, p.blue}) texels.push_back(static_cast<std::uint8_t>(std::lround(ch * F(255))));
                        }
                }
            }, tv);
            texture_index.emplace(name, static_cast<std::int32_t>(textures.size()));
            textures.push_back(t);
        }

        std::vector<rt_material_desc> materials;
        for (const auto& mv : sc.materials) {
            rt_material_desc m{};
            m.texture = -1; m.ior = 1.0f;
            std::visit([&](const auto& mat) {
                using M = std::decay_t<decltype(mat)>;
                m.smooth_shading = mat.smooth_shading ? 1u : 0u;
                if constexpr (std::is_same_v<M, diffuse_material<F>>) m.kind = RT_MAT_DIFFUSE;
                else if constexpr (std::is_same_v<M, reflective_material<F>>) m.kind = RT_MAT_REFLECTIVE;
                else if constexpr (std::is_same_v<M, refractive_material<F>>) m.kind = RT_MAT_REFRACTIVE;
                else if constexpr (std::is_same_v<M, constant_material<F>>) m.kind = RT_MAT_CONSTANT;
                else m.kind = RT_MAT_TEXTURE;
                if constexpr (requires { mat.albedo; }) { m.albedo[0] = mat.albedo.red; m.albedo[1] = mat.albedo.green; m.albedo[2] = mat.albedo.blue; }
                if constexpr (requires { mat.ior; }) m.ior = mat.ior;
                if constexpr (requires { mat.texture; }) m.texture = texture_index.at(mat.texture);
            }, mv);
            materials.push_back(m);
        }

        std::vector<rt_mesh_desc> meshes;
        std::vector<std::vector<float>> vbuf(sc.meshes.size()), uvbuf(sc.meshes.size());
        std::vector<std::vector<std::uint32_t>> tbuf(sc.meshes.size());
        std::size_t first = 0;
        for (std::size_t i = 0; i < sc.meshes.size(); ++i) {
            const auto& mo = sc.meshes[i];
            for (const auto& v : mo.vertices) { vbuf[i].push_back(v.x); vbuf[i].push_back(v.y); vbuf[i].push_back(v.z); }
            for (const auto& uv : mo.uvs) { uvbuf[i].push_back(uv.x); uvbuf[i].push_back(uv.y); }
            for (const auto& t : mo.triangles) for (std::size_t k : t.vertex_indices) tbuf[i].push_back(static_cast<std::uint32_t>(k));
            rt_mesh_desc md{};
            md.material = static_cast<std::uint32_t>(mo.material_idx);
            md.n_vertices = static_cast<std::uint32_t>(mo.vertices.size());
            md.n_uvs = static_cast<std::uint32_t>(mo.uvs.size());
            md.n_triangles = static_cast<std::uint32_t>(mo.triangles.size());
            md.vertices = vbuf[i].data(); md.uvs = uvbuf[i].empty() ? nullptr : uvbuf[i].data(); md.triangles = tbuf[i].data();
            meshes.push_back(md);
            mesh_first_triangle.push_back(first);       // global id = position in the mesh-order concatenation (kd_tree_simd.hpp:103-111)
            first += mo.triangles.size();
        }
        mesh_first_triangle.push_back(first);

        d.n_lights = static_cast<std::uint32_t>(lights.size()); d.lights = lights.data();
        d.n_textures = static_cast<std::uint32_t>(textures.size()); d.textures = textures.data();
        d.n_materials = static_cast<std::uint32_t>(materials.size()); d.materials = materials.data();
        d.n_meshes = static_cast<std::uint32_t>(meshes.size()); d.meshes = meshes.data();
        d.n_texel_bytes = texels.size(); d.texels = texels.data();

        rt_build_opts o;
        rt_default_build_opts(&o);
        o.kd_max_depth = kd_max_depth; o.kd_max_leaf_size = kd_max_leaf_size; o.device = device; o.accel_build = accel_build;
        rt_scene* raw = nullptr;
        const int st = rt_scene_create(&d, &o, &raw);
        if (st != RT_OK) throw std::runtime_error(std::string("b200_accel: ") + rt_status_string(st) + ": " + rt_last_error());
        handle = std::shared_ptr<rt_scene>(raw, handle_deleter{});
    }

    // accel.intersect<bf>(ray) - kd_tree_simd.hpp:187-264.  A failed call reports a miss: the reference's query is noexcept.
    template <bool backface_culling>
    std::optional<hit<F>> intersect(const ray3<F>& ray) const noexcept {
        const float r[6] = {ray.origin.x, ray.origin.y, ray.origin.z, ray.direction.x, ray.direction.y, ray.direction.z};
        rt_hit h{};
        if (rt_trace_closest(handle.get(), r, 1, backface_culling ? 1 : 0, static_cast<float>(epsilon), 0u, &h) != RT_OK || h.tri < 0)
            return std::nullopt;
        // hit assembly as kd_tree_simd.hpp:234-263, from the host scene
        std::size_t mesh = 0;
        while (mesh_first_triangle[mesh + 1] <= static_cast<std::size_t>(h.tri)) ++mesh;
        const auto& mo = scene_ptr->meshes[mesh];
        const auto& tri = mo.triangles[static_cast<std::size_t>(h.tri) - mesh_first_triangle[mesh]];
        const F u = h.u, v = h.v, w = F(1.) - u - v;
        const auto [i0, i1, i2] = tri.vertex_indices;
        const vec3<F> n = normalized(u * mo.vertex_normals[i1] + v * mo.vertex_normals[i2] + w * mo.vertex_normals[i0]);
        return hit<F>{ray, ray.origin + h.t * ray.direction, n, tri.normal, tri.uvs, h.t, u, v, w, tri.mesh_idx};
    }
};

// config.hpp:6-17 as the run-time parameters of one frame
inline rt_params b200_config_params() {
    rt_params p;
    rt_default_params(&p);
    p.fov_degrees = fov_degrees;                                               // config.hpp:6
    p.epsilon = static_cast<float>(epsilon);                                   // config.hpp:8, narrowed as src/main.cpp:37
    p.shadow_bias = static_cast<float>(shadow_bias);                           // config.hpp:9
    p.reflection_bias = static_cast<float>(reflection_bias);                   // config.hpp:10
    p.refraction_bias = static_cast<float>(refraction_bias);                   // config.hpp:11
    p.samples_per_pixel = static_cast<std::uint32_t>(samples_per_pixel);       // config.hpp:13
    p.max_ray_depth = static_cast<std::uint32_t>(max_ray_depth);               // config.hpp:14
    p.diffuse_reflection_ray_count = static_cast<std::uint32_t>(diffuse_reflection_ray_count);   // config.hpp:15
    if (fixed_rng_seed) p.seed = static_cast<std::uint32_t>(*fixed_rng_seed);  // config.hpp:17
    return p;
}

// image<F> (scene/image.hpp:7-31: one std::vector<color<F>> per row) from a packed float frame.  color<F> is three F in a row
// (scene/color.hpp:3-7), so a row is one contiguous copy; the rows are filled by a few threads because at 1080p the cost is the
// page faults of 25 MB of fresh vectors, not the copy.
template <typename F>
image<F> b200_image_from_rgb(const float* rgb, std::size_t h, std::size_t w) {
    static_assert(std::is_same_v<F, float> && sizeof(color<F>) == 3 * sizeof(F) && std::is_trivially_copyable_v<color<F>>);
    std::vector<std::vector<color<F>>> pixels(h);
    auto fill = [&](std::size_t y0, std::size_t y1) {
        for (std::size_t y = y0; y < y1; ++y) {
            const color<F>* row = reinterpret_cast<const color<F>*>(rgb + y * w * 3);
            pixels[y].assign(row, row + w);
        }
    };
    const std::size_t n_threads = std::min<std::size_t>({8, std::max<std::size_t>(1, std::thread::hardware_concurrency()), std::max<std::size_t>(1, h * w / 200000)});
    if (n_threads <= 1) fill(0, h);
    else {
        std::vector<std::thread> pool;
        const std::size_t step = (h + n_threads - 1) / n_threads;
        for (std::size_t y0 = step; y0 < h; y0 += step) pool.emplace_back(fill, y0, std::min(h, y0 + step));
        fill(0, std::min(h, step));
        for (auto& t : pool) t.join();
    }
    return image<F>(h, w, std::move(pixels));
}

// render_frame for the B200 accelerator: same signature and result type as render/render.hpp:18-19, whole frame on device.
// (A more specialised overload than the generic template, so `render_frame<b200_accel<F>, F>(accel, schedule)` picks it.)
template <typename A, typename F>
requires std::is_same_v<A, b200_accel<F>>
image<F> render_frame(const b200_accel<F>& accel, const scheduling_type /* tiles are scheduled by the device */) {
    const rt_params p = b200_config_params();
    const std::size_t h = accel.scene_ptr->config.image_height, w = accel.scene_ptr->config.image_width;
    if (!accel.staging) accel.staging = std::shared_ptr<float>(static_cast<float*>(rt_alloc_pinned(h * w * 3 * sizeof(float))),
                                                              typename b200_accel<F>::pinned_deleter{});
    if (accel.staging) {
        // the queued path (one frame in flight): its kernels chain without per-launch timing events, the download is one DMA
        std::uint64_t ticket = 0;
        int st = rt_render_frame_begin(accel.handle.get(), &p, accel.staging.get(), &ticket);
        if (st == RT_OK) st = rt_frame_wait(accel.handle.get(), ticket);
        if (st != RT_OK) throw std::runtime_error(std::string("b200 render_frame: ") + rt_status_string(st) + ": " + rt_last_error());
        return b200_image_from_rgb<F>(accel.staging.get(), h, w);
    }
    std::vector<float> rgb(h * w * 3);                  // no page-locked memory to be had: the synchronous call
    const int st = rt_render_frame(accel.handle.get(), &p, rgb.data());
    if (st != RT_OK) throw std::runtime_error(std::string("b200 render_frame: ") + rt_status_string(st) + ": " + rt_last_error());
    return b200_image_from_rgb<F>(rgb.data(), h, w);
}

// ---- image output --------------------------------------------------------------------------------
// The reference quantises while it writes: write_ppm (io/image/ppm.hpp:7-25) turns every float channel into
// uint8(255.999 * clamp(c, 0, 1)) and prints ASCII P3 - at 4K that is 100 MB of text through operator<<.  Here the frame is
// quantised on the device by the same expression (rt_render_frame_rgb8: a quarter of the bytes over PCIe) and the file is
// written from the bytes: b200_write_ppm produces the reference's file byte for byte, b200_write_ppm_binary the P6 form.
template <typename F>
std::vector<std::uint8_t> b200_render_frame_rgb8(const b200_accel<F>& accel, const rt_params& p = b200_config_params()) {
    const std::size_t h = accel.scene_ptr->config.image_height, w = accel.scene_ptr->config.image_width;
    std::vector<std::uint8_t> rgb8(h * w * 3);
    const int st = rt_render_frame_rgb8(accel.handle.get(), &p, rgb8.data());
    if (st != RT_OK) throw std::runtime_error(std::string("b200 render_frame_rgb8: ") + rt_status_string(st) + ": " + rt_last_error());
    return rgb8;
}
// the bytes write_ppm would print for this image (host-side restatement of ppm.hpp:17-19, for callers that hold an image<F>)
template <typename F>
std::vector<std::uint8_t> b200_quantise(const image<F>& img) {
    std::vector<std::uint8_t> rgb8(img.get_height() * img.get_width() * 3);
    std::size_t k = 0;
    for (std::size_t y = 0; y < img.get_height(); ++y)
        for (std::size_t x = 0; x < img.get_width(); ++x) {
            const auto c = img.get_pixel(y, x);
            for (const F ch : {c.red, c.green, c.blue})
                rgb8[k++] = static_cast<std::uint8_t>(255.999 * std::clamp(ch, static_cast<F>(0.), static_cast<F>(1.)));
        }
    return rgb8;
}
// ASCII P3, identical to write_ppm's output for the same bytes: "r g b\t" per pixel, one line per row (ppm.hpp:8-24)
inline void b200_write_ppm(const std::uint8_t* rgb8, std::size_t width, std::size_t height, std::ostream& out) {
    char dec[256][4];
    std::uint8_t len[256];
    for (int v = 0; v < 256; ++v) len[v] = static_cast<std::uint8_t>(std::snprintf(dec[v], sizeof dec[v], "%d", v));
    out << "P3\n" << width << " " << height << "\n255\n";
    std::string line;
    line.reserve(width * 12 + 1);
    for (std::size_t y = 0; y < height; ++y) {
        line.clear();
        const std::uint8_t* px = rgb8 + y * width * 3;
        for (std::size_t x = 0; x < width; ++x, px += 3) {
            line.append(dec[px[0]], len[px[0]]); line.push_back(' ');
            line.append(dec[px[1]], len[px[1]]); line.push_back(' ');
            line.append(dec[px[2]], len[px[2]]); line.push_back('\t');
        }
        line.push_back('\n');
        out.write(line.data(), static_cast<std::streamsize>(line.size()));
    }
}
// binary P6: the same pixels in 3 bytes each
inline void b200_write_ppm_binary(const std::uint8_t* rgb8, std::size_t width, std::size_t height, std::ostream& out) {
    out << "P6\n" << width << " " << height << "\n255\n";
    out.write(reinterpret_cast<const char*>(rgb8), static_cast<std::streamsize>(width * height * 3));
}

// A caller that renders frame after frame (src/main.cpp:13-25 in a loop - the reference's animation outputs) keeps two
// frames in flight: submit() queues a frame without waiting for the device, next() returns the oldest one.  Frame i's
// PCIe download overlaps frame i+1's kernels (rt_render_frame_begin / rt_frame_wait), which halves the time per frame
// of the one-call render_frame above on the reference's scenes.
template <typename F>
class b200_frame_sequence {
    const b200_accel<F>& accel_;
    std::size_t h_, w_;
    float* buf_[2] = {nullptr, nullptr};
    std::uint64_t ticket_[2] = {0, 0};
    std::uint64_t submitted_ = 0, returned_ = 0;

public:
    explicit b200_frame_sequence(const b200_accel<F>& accel)
        : accel_(accel), h_(accel.scene_ptr->config.image_height), w_(accel.scene_ptr->config.image_width) {
        for (float*& b : buf_) {
            b = static_cast<float*>(rt_alloc_pinned(h_ * w_ * 3 * sizeof(float)));
            if (!b) throw std::runtime_error(std::string("b200_frame_sequence: ") + rt_last_error());
        }
    }
    b200_frame_sequence(const b200_frame_sequence&) = delete;
    b200_frame_sequence& operator=(const b200_frame_sequence&) = delete;
    ~b200_frame_sequence() {
        while (returned_ < submitted_) { rt_frame_wait(accel_.handle.get(), ticket_[returned_ & 1]); ++returned_; }
        for (float* b : buf_) rt_free_pinned(b);
    }
    std::size_t in_flight() const { return submitted_ - returned_; }
    // queue one frame (at most two in flight: call next() first when in_flight() == 2)
    void submit(const rt_params& p = b200_config_params()) {
        if (in_flight() == 2) throw std::logic_error("b200_frame_sequence: two frames are in flight, take one with next()");
        const int st = rt_render_frame_begin(accel_.handle.get(), &p, buf_[submitted_ & 1], &ticket_[submitted_ & 1]);
        if (st != RT_OK) throw std::runtime_error(std::string("b200 submit: ") + rt_status_string(st) + ": " + rt_last_error());
        ++submitted_;
    }
    // the oldest frame in flight, as render_frame would have returned it
    image<F> next() {
        if (!in_flight()) throw std::logic_error("b200_frame_sequence: no frame in flight");
        const int st = rt_frame_wait(accel_.handle.get(), ticket_[returned_ & 1]);
        if (st != RT_OK) throw std::runtime_error(std::string("b200 next: ") + rt_status_string(st) + ": " + rt_last_error());
        const float* rgb = buf_[returned_ & 1];
        ++returned_;
        return b200_image_from_rgb<F>(rgb, h_, w_);
    }
};
