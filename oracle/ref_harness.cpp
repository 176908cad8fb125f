// oracle/ref_harness.cpp - TEST INFRASTRUCTURE.  Compiles the UNMODIFIED reference headers where they lie under
// /root/reference/include into oracle/_ref/libref_<variant>.so (recipe: oracle/Makefile).  Nothing here is product
// code and nothing in the product may link or load it; it is the strongest checker we have for oracle/rt_oracle.c
// and, through bench.py --impl reference / cpu_baseline, the CPU arm that is timed beside the GPU.
//
// What is the reference and what is harness:
//   reference (untouched):  kd_tree_simd_accel (render/accel/kd_tree_simd.hpp:63-303), render_frame / is_occluded /
//                           color_hit (render/render.hpp:18-308), all scene / math types.
//   harness (this file):    RTSC reader -> reference scene<float> through the reference's own constructors
//                           (mirrors io/json/loader.hpp:149-265, which needs simdjson - absent here);
//                           stbi_load stub; a recording / counting accelerator wrapper (satisfies the
//                           `accelerator` concept, render/accel/accel.hpp:8-12); the 30-line traversal loop of
//                           kd_tree_simd.hpp:191-228 re-run over the accel's PUBLIC members only to recover the
//                           winning triangle index, which hit<F> does not carry (render/hit.hpp:9-21).
//
// Mandatory flag: -fstack-reuse=none (core/math/aabb3.hpp:79 binds std::minmax's reference pair to temporaries;
// g++ 13 miscompiles it at -O2+).

#include <algorithm>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <string>
#include <vector>

#include <raytracer/config.hpp>
#include <raytracer/scene/scene.hpp>
#include <raytracer/render/render.hpp>
#include <raytracer/render/accel/kd_tree_simd.hpp>

#ifndef RT_KD_MAX_DEPTH
#define RT_KD_MAX_DEPTH 8
#endif
#ifndef RT_KD_MAX_LEAF
#define RT_KD_MAX_LEAF 64
#endif

#include "rtsc_ref_scene.hpp"          // F, scene_from_rtsc, the stbi_load stub

using A = kd_tree_simd_accel<F, static_cast<F>(epsilon), RT_KD_MAX_DEPTH, RT_KD_MAX_LEAF>;

namespace {

// closest-hit query re-run over PUBLIC members (kd_tree_simd.hpp:191-228) so the triangle index can be reported
template <bool bf>
bool trace_with_index(const A& a, const ray3<F>& ray, F& t, F& u, F& v, int64_t& tri) {
    std::optional<A::hit_candidate> closest;
    std::vector<std::size_t> stack;
    stack.push_back(0);
    while (!stack.empty()) {
        const auto idx = stack.back();
        stack.pop_back();
        const auto& node = a.tree[idx];
        const F best_t = closest ? closest->t : A::MAX_F;
        auto bh = node.box.intersect(ray);
        if (!bh || best_t < bh->t_min) continue;
        if (node.start_idx == A::EMPTY) {
            if (node.child0 != A::EMPTY) stack.push_back(node.child0);
            if (node.child1 != A::EMPTY) stack.push_back(node.child1);
        } else {
            const auto c = a.template intersect_leaf<bf>(ray, node);
            if (!c) continue;
            const F best_t2 = closest ? closest->t : A::MAX_F;
            if (c->t < best_t2) closest = c;
        }
    }
    if (!closest) { tri = -1; t = u = v = 0; return false; }
    t = closest->t; u = closest->u; v = closest->v;
    tri = int64_t(a.triangle_packs[closest->pack_idx].triangle_indices[closest->lane]);
    return true;
}

struct query_record { float o[3], d[3]; float t, u, v; int32_t tri; uint32_t cull; };

// accelerator wrapper (concept: render/accel/accel.hpp:8-12 + implicit scene_ptr member, render.hpp:21,113,136)
struct tap_accel {
    std::shared_ptr<const scene<F>> scene_ptr;
    const A* inner = nullptr;
    mutable std::atomic<uint64_t> n_cull{0}, n_nocull{0}, n_cull_hit{0}, n_nocull_hit{0};
    mutable std::vector<query_record>* log = nullptr;
    mutable std::mutex log_mutex;
    std::size_t log_cap = 0;

    template <bool bf>
    std::optional<hit<F>> intersect(const ray3<F>& ray) const {
        auto h = inner->template intersect<bf>(ray);
        if constexpr (bf) { n_cull.fetch_add(1, std::memory_order_relaxed); if (h) n_cull_hit.fetch_add(1, std::memory_order_relaxed); }
        else { n_nocull.fetch_add(1, std::memory_order_relaxed); if (h) n_nocull_hit.fetch_add(1, std::memory_order_relaxed); }
        if (log) {
            query_record q{};
            q.o[0] = ray.origin.x; q.o[1] = ray.origin.y; q.o[2] = ray.origin.z;
            q.d[0] = ray.direction.x; q.d[1] = ray.direction.y; q.d[2] = ray.direction.z;
            q.cull = bf ? 1u : 0u;
            int64_t tri; F t, u, v;
            const bool got = trace_with_index<bf>(*inner, ray, t, u, v, tri);
            if (got != h.has_value() || (got && (t != h->distance || u != h->u || v != h->v))) {
                std::fprintf(stderr, "ref_harness: index re-run disagrees with reference intersect\n");
                std::abort();
            }
            q.t = t; q.u = u; q.v = v; q.tri = int32_t(tri);
            std::lock_guard g(log_mutex);
            if (log->size() < log_cap) log->push_back(q);
        }
        return h;
    }
};
static_assert(accelerator<tap_accel, F>);
static_assert(accelerator<A, F>);

struct ref_handle {
    std::shared_ptr<const scene<F>> sc;
    std::unique_ptr<A> accel;
    double build_seconds = 0;
};

}  // namespace

extern "C" {

void* ref_scene_load(const char* rtsc_path) {
    try {
        auto h = std::make_unique<ref_handle>();
        h->sc = std::make_shared<const scene<F>>(scene_from_rtsc(rtsc_path, h.get()));
        const auto t0 = std::chrono::steady_clock::now();
        h->accel = std::make_unique<A>(h->sc);
        h->build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        return h.release();
    } catch (const std::exception& e) {
        std::fprintf(stderr, "ref_scene_load: %s\n", e.what());
        return nullptr;
    }
}

void ref_scene_free(void* p) { delete static_cast<ref_handle*>(p); }

// info[0..11]: width, height, n_tris, n_nodes, n_packs, W, spp, max_ray_depth, gi_rays, kd_max_depth, kd_max_leaf, threads
void ref_info(void* p, uint64_t* info, double* build_seconds) {
    auto* h = static_cast<ref_handle*>(p);
    info[0] = h->sc->config.image_width; info[1] = h->sc->config.image_height;
    info[2] = h->accel->triangles.size(); info[3] = h->accel->tree.size(); info[4] = h->accel->triangle_packs.size();
    info[5] = A::simd_f::size(); info[6] = samples_per_pixel; info[7] = max_ray_depth;
    info[8] = diffuse_reflection_ray_count; info[9] = RT_KD_MAX_DEPTH; info[10] = RT_KD_MAX_LEAF;
    info[11] = std::thread::hardware_concurrency();
    if (build_seconds) *build_seconds = h->build_seconds;
}

// The reference's own frame entry point, all host threads, BUCKET_TILES as src/main.cpp:17.  Returns seconds
// around render_frame only (main.cpp:16-20).  rgb may be null (timing only).
double ref_render(void* p, float* rgb) {
    auto* h = static_cast<ref_handle*>(p);
    const auto t0 = std::chrono::steady_clock::now();
    auto img = render_frame<A, F>(*h->accel, scheduling_type::BUCKET_TILES);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rgb) {
        for (std::size_t y = 0; y < img.get_height(); ++y)
            for (std::size_t x = 0; x < img.get_width(); ++x) {
                const auto& c = img.get_pixel(y, x);
                float* o = rgb + (y * img.get_width() + x) * 3;
                o[0] = c.red; o[1] = c.green; o[2] = c.blue;
            }
    }
    return sec;
}

// Counting pass through the unmodified render_frame. counts: [cull queries, cull hits, no-cull queries, no-cull hits]
void ref_count(void* p, uint64_t* counts) {
    auto* h = static_cast<ref_handle*>(p);
    tap_accel tap; tap.scene_ptr = h->sc; tap.inner = h->accel.get();
    (void)render_frame<tap_accel, F>(tap, scheduling_type::BUCKET_TILES);
    counts[0] = tap.n_cull; counts[1] = tap.n_cull_hit; counts[2] = tap.n_nocull; counts[3] = tap.n_nocull_hit;
}

// Every closest-hit query render_frame issues, in issue order of ONE worker (SINGLE_TILE: one tile, so exactly one
// thread does all the work, in row-major pixel order), with its answer incl. triangle index.  Returns #records.
uint64_t ref_record(void* p, query_record* out, uint64_t cap, float* rgb) {
    auto* h = static_cast<ref_handle*>(p);
    std::vector<query_record> log;
    log.reserve(cap);
    tap_accel tap; tap.scene_ptr = h->sc; tap.inner = h->accel.get(); tap.log = &log; tap.log_cap = cap;
    auto img = render_frame<tap_accel, F>(tap, scheduling_type::SINGLE_TILE);
    std::memcpy(out, log.data(), log.size() * sizeof(query_record));
    if (rgb) {
        for (std::size_t y = 0; y < img.get_height(); ++y)
            for (std::size_t x = 0; x < img.get_width(); ++x) {
                const auto& c = img.get_pixel(y, x);
                float* o = rgb + (y * img.get_width() + x) * 3;
                o[0] = c.red; o[1] = c.green; o[2] = c.blue;
            }
    }
    return log.size();
}

// Batch closest-hit on caller-provided rays (o[3], d[3] per ray).  tri = -1 on miss.
void ref_trace(void* p, const float* rays, uint64_t n, int cull, float* tuv, int32_t* tri) {
    auto* h = static_cast<ref_handle*>(p);
    for (uint64_t i = 0; i < n; ++i) {
        const float* r = rays + 6 * i;
        const ray3<F> ray(vec3<F>{r[0], r[1], r[2]}, vec3<F>{r[3], r[4], r[5]});
        F t, u, v; int64_t id;
        if (cull) trace_with_index<true>(*h->accel, ray, t, u, v, id);
        else trace_with_index<false>(*h->accel, ray, t, u, v, id);
        tuv[3 * i] = t; tuv[3 * i + 1] = u; tuv[3 * i + 2] = v; tri[i] = int32_t(id);
    }
}

// Batch shadow query = the reference's is_occluded (render.hpp:110-131)
void ref_occluded(void* p, const float* rays, const float* max_t, uint64_t n, uint8_t* out) {
    auto* h = static_cast<ref_handle*>(p);
    for (uint64_t i = 0; i < n; ++i) {
        const float* r = rays + 6 * i;
        const ray3<F> ray(vec3<F>{r[0], r[1], r[2]}, vec3<F>{r[3], r[4], r[5]});
        out[i] = is_occluded<A, F>(*h->accel, ray, max_t[i]) ? 1 : 0;
    }
}

// Tree dump for builder parity.  Pass nulls to query sizes.  node6 per node: parent, child0, child1, start_idx,
// pack_count (all u64, EMPTY = 2^64-1); boxes 6 floats per node; pack_tris = W u64 indices per pack.
void ref_tree(void* p, uint64_t* node5, float* boxes, uint64_t* pack_tris) {
    auto* h = static_cast<ref_handle*>(p);
    const auto& a = *h->accel;
    for (std::size_t i = 0; i < a.tree.size(); ++i) {
        const auto& n = a.tree[i];
        if (node5) { node5[5 * i] = n.parent; node5[5 * i + 1] = n.child0; node5[5 * i + 2] = n.child1; node5[5 * i + 3] = n.start_idx; node5[5 * i + 4] = n.pack_count; }
        if (boxes) { float* b = boxes + 6 * i; b[0] = n.box.min.x; b[1] = n.box.min.y; b[2] = n.box.min.z; b[3] = n.box.max.x; b[4] = n.box.max.y; b[5] = n.box.max.z; }
    }
    if (pack_tris) {
        constexpr std::size_t W = A::simd_f::size();
        for (std::size_t i = 0; i < a.triangle_packs.size(); ++i)
            for (std::size_t l = 0; l < W; ++l) pack_tris[i * W + l] = a.triangle_packs[i].triangle_indices[l];
    }
}

}  // extern "C"
