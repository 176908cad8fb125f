// tests/helpers/kd8_host.cpp - TEST INFRASTRUCTURE: compiles simd-raytracer_b200/csrc/rt_bvh.cuh and rt_kd8.cuh (the accelerated traversals the
// CUDA kernels run) as plain C++ so that tests/test_kd8_host.py can check the algorithm and the flattened tree against the
// oracle on a machine without a GPU.  Built with -ffp-contract=off (the device TU is built with -fmad=false).
#include <cstdint>
static uint64_t g_kd8_nodes, g_kd8_tris, g_bvh_leaves;   // visit counters (rt_kd8.cuh / rt_bvh.cuh instrumentation hooks)
#define KD8_COUNT_NODE() (++g_kd8_nodes)
#define KD8_COUNT_TRI() (++g_kd8_tris)
#define BVH_COUNT_NODE() (++g_kd8_nodes)
#define BVH_COUNT_LEAF() (++g_bvh_leaves)
#include "../../simd-raytracer_b200/csrc/rt_bvh.cuh"          // includes rt_kd8.cuh

extern "C" void kd8_counters(uint64_t* nodes, uint64_t* tris, int reset) {
    if (nodes) *nodes = g_kd8_nodes;
    if (tris) *tris = g_kd8_tris;
    if (reset) g_kd8_nodes = g_kd8_tris = 0;
}
extern "C" uint64_t bvh_leaf_visits(int reset) {
    const uint64_t v = g_bvh_leaves;
    if (reset) g_bvh_leaves = 0;
    return v;
}

extern "C" void kd8_trace_batch(const uint32_t* nodes8, const float* tris, const float* root6, const float* rays6, uint64_t n,
                                int cull, int fast, float eps, const float* t_far, int any_hit, float* tuv, int32_t* tri, uint8_t* tie) {
    for (uint64_t i = 0; i < n; ++i) {
        const float* q = rays6 + 6 * i;
        const float far = t_far ? t_far[i] : FLT_MAX;
        rtb::KdHit h;
        if (cull) h = fast ? rtb::kd8_trace<true, true>(nodes8, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0)
                           : rtb::kd8_trace<true, false>(nodes8, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0);
        else h = fast ? rtb::kd8_trace<false, true>(nodes8, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0)
                      : rtb::kd8_trace<false, false>(nodes8, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0);
        tuv[3 * i] = h.tri >= 0 ? h.t : 0; tuv[3 * i + 1] = h.tri >= 0 ? h.u : 0; tuv[3 * i + 2] = h.tri >= 0 ? h.v : 0;
        tri[i] = h.tri;
        if (tie) tie[i] = (h.tri == rtb::KD_RERUN || (h.tri >= 0 && h.tie_t == h.t)) ? 1 : 0;     // the device re-runs these through the reference-order query
    }
}

// the same batch through the bounding-volume hierarchy (rt_scene_get_bvh_layout)
extern "C" void bvh_trace_batch(const float* nodes16, const float* tris, const float* root6, const float* rays6, uint64_t n,
                                int cull, int fast, float eps, const float* t_far, int any_hit, float* tuv, int32_t* tri, uint8_t* tie) {
    for (uint64_t i = 0; i < n; ++i) {
        const float* q = rays6 + 6 * i;
        const float far = t_far ? t_far[i] : FLT_MAX;
        rtb::KdHit h;
        if (cull) h = fast ? rtb::bvh_trace<true, true>(nodes16, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0)
                           : rtb::bvh_trace<true, false>(nodes16, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0);
        else h = fast ? rtb::bvh_trace<false, true>(nodes16, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0)
                      : rtb::bvh_trace<false, false>(nodes16, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0);
        tuv[3 * i] = h.tri >= 0 ? h.t : 0; tuv[3 * i + 1] = h.tri >= 0 ? h.u : 0; tuv[3 * i + 2] = h.tri >= 0 ? h.v : 0;
        tri[i] = h.tri;
        if (tie) tie[i] = (h.tri == rtb::KD_RERUN || (h.tri >= 0 && h.tie_t == h.t)) ? 1 : 0;
    }
}

// per-ray node-visit and triangle-test counts of the BVH traversal (developer sweeps: scripts/bvh_ray_lengths_cpu.py)
extern "C" void bvh_trace_batch_counts(const float* nodes16, const float* tris, const float* root6, const float* rays6, uint64_t n,
                                       int cull, float eps, const float* t_far, int any_hit, uint32_t* node_visits, uint32_t* tri_tests) {
    for (uint64_t i = 0; i < n; ++i) {
        const float* q = rays6 + 6 * i;
        const float far = t_far ? t_far[i] : FLT_MAX;
        const uint64_t n0 = g_kd8_nodes, t0 = g_kd8_tris;
        if (cull) rtb::bvh_trace<true, false>(nodes16, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0);
        else rtb::bvh_trace<false, false>(nodes16, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0);
        node_visits[i] = uint32_t(g_kd8_nodes - n0);
        tri_tests[i] = uint32_t(g_kd8_tris - t0);
    }
}

// the tile-culling predicate of k_tile_cull (csrc/rt_tilecull.cuh), one raster rectangle per call
#include "../../simd-raytracer_b200/csrc/rt_tilecull.cuh"
extern "C" int tile_misses_box_host(const float* m9, const float* pos3, float width, float height, float tan_half_fov, const float* lo3,
                                    const float* hi3, float x0, float y0, float x1, float y1) {
    rtb::TileCamera c;
    for (int k = 0; k < 9; ++k) c.m[k] = m9[k];
    for (int k = 0; k < 3; ++k) c.pos[k] = pos3[k];
    c.width = width; c.height = height; c.tan_half_fov = tan_half_fov;
    return rtb::tile_misses_box(c, lo3, hi3, x0, y0, x1, y1) ? 1 : 0;
}

// ---- four-wide hierarchy (csrc/rt_bvh4.cuh), collapsed on the fly from the two-wide nodes the scene exports ----------------
#include <map>
#include <vector>
#include "../../simd-raytracer_b200/csrc/rt_bvh4.cuh"
#include "../../simd-raytracer_b200/host/bvh4_collapse.hpp"
static const std::vector<uint32_t>& bvh4_of(const float* nodes16, uint64_t n_nodes2) {
    static std::map<std::pair<const float*, uint64_t>, std::vector<uint32_t>> cache;
    auto it = cache.find({nodes16, n_nodes2});
    if (it == cache.end()) it = cache.emplace(std::make_pair(nodes16, n_nodes2), rtb::bvh4_collapse(reinterpret_cast<const uint32_t*>(nodes16), n_nodes2)).first;
    return it->second;
}
extern "C" uint64_t bvh4_node_count(const float* nodes16, uint64_t n_nodes2) { return bvh4_of(nodes16, n_nodes2).size() / 32; }
extern "C" void bvh4_trace_batch(const float* nodes16, uint64_t n_nodes2, const float* tris, const float* root6, const float* rays6, uint64_t n,
                                 int cull, int fast, float eps, const float* t_far, int any_hit, float* tuv, int32_t* tri, uint8_t* tie,
                                 uint32_t* node_visits) {
    // a fresh collapse per batch unless the caller keeps the node array alive and unchanged (the cache is keyed by its address)
    const std::vector<uint32_t> nodes4v = rtb::bvh4_collapse(reinterpret_cast<const uint32_t*>(nodes16), n_nodes2);
    const float* nodes4 = reinterpret_cast<const float*>(nodes4v.data());
    for (uint64_t i = 0; i < n; ++i) {
        const float* q = rays6 + 6 * i;
        const float far = t_far ? t_far[i] : FLT_MAX;
        const uint64_t n0 = g_kd8_nodes;
        rtb::KdHit h;
        if (cull) h = fast ? rtb::bvh4_trace<true, true>(nodes4, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0)
                           : rtb::bvh4_trace<true, false>(nodes4, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0);
        else h = fast ? rtb::bvh4_trace<false, true>(nodes4, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0)
                      : rtb::bvh4_trace<false, false>(nodes4, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0);
        tuv[3 * i] = h.tri >= 0 ? h.t : 0; tuv[3 * i + 1] = h.tri >= 0 ? h.u : 0; tuv[3 * i + 2] = h.tri >= 0 ? h.v : 0;
        tri[i] = h.tri;
        if (tie) tie[i] = (h.tri == rtb::KD_RERUN || (h.tri >= 0 && h.tie_t == h.t)) ? 1 : 0;
        if (node_visits) node_visits[i] = uint32_t(g_kd8_nodes - n0);
    }
}
