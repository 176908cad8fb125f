// tests/helpers/kd8_host.cpp - TEST INFRASTRUCTURE: compiles simd-raytracer_b200/csrc/rt_bvh.cuh, rt_bvh4.cuh and rt_tri.cuh (the accelerated
// traversals the CUDA kernels run and their triangle test) as plain C++ so that tests/test_kd8_host.py can check the algorithm and the flattened tree against the
// oracle on a machine without a GPU.  Built with -ffp-contract=off (the device TU is built with -fmad=false).
#include <cstdint>
static uint64_t g_kd8_nodes, g_kd8_tris, g_bvh_leaves;   // visit counters (rt_tri.cuh / rt_bvh.cuh instrumentation hooks)
#define KD8_COUNT_TRI() (++g_kd8_tris)
#define BVH_COUNT_NODE() (++g_kd8_nodes)
#define BVH_COUNT_LEAF() (++g_bvh_leaves)
#include "../../simd-raytracer_b200/csrc/rt_bvh.cuh"          // includes rt_tri.cuh

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

// a batch of queries through the two-wide bounding-volume hierarchy (rt_scene_get_bvh_layout)
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
static int g_bvh4_max_sp;               // deepest stack of the current query (rt_bvh4.cuh instrumentation hook)
#define BVH4_TRACK_SP(sp) (g_bvh4_max_sp = (sp) > g_bvh4_max_sp ? (sp) : g_bvh4_max_sp)
#include "../../simd-raytracer_b200/csrc/rt_bvh4.cuh"
#include "../../simd-raytracer_b200/host/bvh4_collapse.hpp"
static const std::vector<uint32_t>& bvh4_of(const float* nodes16, uint64_t n_nodes2) {
    static std::map<std::pair<const float*, uint64_t>, std::vector<uint32_t>> cache;
    auto it = cache.find({nodes16, n_nodes2});
    if (it == cache.end()) it = cache.emplace(std::make_pair(nodes16, n_nodes2), rtb::bvh4_collapse(reinterpret_cast<const uint32_t*>(nodes16), n_nodes2)).first;
    return it->second;
}
extern "C" uint64_t bvh4_node_count(const float* nodes16, uint64_t n_nodes2) { return bvh4_of(nodes16, n_nodes2).size() / 32; }
// worst-case traversal stack entries of the collapsed hierarchy (what scene creation checks against BVH4_STACK), and that bound
extern "C" uint64_t bvh4_stack_need_host(const float* nodes16, uint64_t n_nodes2, uint64_t* capacity) {
    if (capacity) *capacity = uint64_t(rtb::BVH4_STACK);
    uint32_t in_the_same_walk = 0;                  // what scene creation uses: must be the figure of the separate pass
    const std::vector<uint32_t> nodes4 = rtb::bvh4_collapse(reinterpret_cast<const uint32_t*>(nodes16), n_nodes2, &in_the_same_walk);
    const uint32_t need = rtb::bvh4_stack_need(nodes4);
    return need == in_the_same_walk ? uint64_t(need) : ~uint64_t(0);
}
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
        g_bvh4_max_sp = 0;
        rtb::KdHit h;
        if (cull) h = fast ? rtb::bvh4_trace<true, true>(nodes4, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0)
                           : rtb::bvh4_trace<true, false>(nodes4, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0);
        else h = fast ? rtb::bvh4_trace<false, true>(nodes4, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0)
                      : rtb::bvh4_trace<false, false>(nodes4, tris, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], eps, far, any_hit != 0);
        tuv[3 * i] = h.tri >= 0 ? h.t : 0; tuv[3 * i + 1] = h.tri >= 0 ? h.u : 0; tuv[3 * i + 2] = h.tri >= 0 ? h.v : 0;
        tri[i] = h.tri;
        if (tie) tie[i] = (h.tri == rtb::KD_RERUN || (h.tri >= 0 && h.tie_t == h.t)) ? 1 : 0;
        if (node_visits) node_visits[i] = uint32_t(g_kd8_nodes - n0) | (uint32_t(g_bvh4_max_sp) << 16);     // low half: node visits, high half: deepest stack
    }
}

// ---- developer statistics: node visits of a W-wide hierarchy (W = 2..16) collapsed from the two-wide nodes the same way --------
// (scripts/bvh_ray_lengths_cpu.py; a plain recursive-list traversal with the two-wide traversal's box test, root test and leaf
// test - it answers "how many dependent node visits would a query need", nothing more)
namespace {
struct WideNode { int n; rtb::Bvh4Child c[16]; };
std::vector<WideNode> collapse_wide(const uint32_t* nodes16, uint64_t n_nodes2, int width) {
    constexpr uint32_t NO_CHILD = 0xFFFFFFFFu;
    std::vector<WideNode> out;
    if (!n_nodes2) return out;
    struct Todo { uint32_t node2; int parent, slot; };
    std::vector<Todo> todo{{0u, -1, -1}};
    while (!todo.empty()) {
        const Todo t = todo.back(); todo.pop_back();
        const int me = int(out.size());
        if (t.parent >= 0) out[size_t(t.parent)].c[t.slot].ref = uint32_t(me);
        WideNode w; w.n = 2;
        rtb::bvh2_children(nodes16, t.node2, w.c);
        for (int k = 0; k < w.n;) { if (w.c[k].cnt == NO_CHILD) { w.c[k] = w.c[w.n - 1]; --w.n; } else ++k; }
        while (w.n < width) {
            int best = -1;
            for (int k = 0; k < w.n; ++k) if (w.c[k].cnt == 0 && (best < 0 || rtb::bvh4_area(w.c[k]) > rtb::bvh4_area(w.c[best]))) best = k;
            if (best < 0) break;
            rtb::Bvh4Child g[2]; rtb::bvh2_children(nodes16, w.c[best].ref, g);
            int m = 0; rtb::Bvh4Child keep[2];
            for (int k = 0; k < 2; ++k) if (g[k].cnt != NO_CHILD) keep[m++] = g[k];
            if (m == 0) { w.c[best] = w.c[w.n - 1]; --w.n; continue; }
            w.c[best] = keep[0];
            if (m == 2) w.c[w.n++] = keep[1];
        }
        out.push_back(w);
        for (int k = w.n - 1; k >= 0; --k) if (w.c[k].cnt == 0) todo.push_back({w.c[k].ref, me, k});
    }
    return out;
}
}  // namespace
extern "C" void bvh_wide_visit_counts(const float* nodes16, uint64_t n_nodes2, int width, const float* tris, const float* root6, const float* rays6,
                                      uint64_t n, int cull, float eps, const float* t_far, int any_hit, uint32_t* node_visits) {
    const std::vector<WideNode> nodes = collapse_wide(reinterpret_cast<const uint32_t*>(nodes16), n_nodes2, width < 2 ? 2 : (width > 16 ? 16 : width));
    struct Entry { uint32_t ref, cnt; float t0; };
    std::vector<Entry> stack;
    for (uint64_t i = 0; i < n; ++i) {
        const float* q = rays6 + 6 * i;
        rtb::BvhState s;
        node_visits[i] = 0;
        if (!rtb::bvh_init(s, root6, root6 + 3, q[0], q[1], q[2], q[3], q[4], q[5], t_far ? t_far[i] : FLT_MAX, any_hit != 0) || s.phase == rtb::KD8_DONE) continue;
        stack.clear();
        stack.push_back({0u, 0u, 0.0f});
        const float ix = rtb::kd_rcp_estimate(s.dx), iy = rtb::kd_rcp_estimate(s.dy), iz = rtb::kd_rcp_estimate(s.dz);
        const float cx = -(s.ox * ix), cy = -(s.oy * iy), cz = -(s.oz * iz);
        while (!stack.empty()) {
            const Entry e = stack.back(); stack.pop_back();
            const float lim = rtb::kd_min(s.best.t, s.t_far);
            if (e.t0 > lim) continue;
            if (e.cnt) {
                if (cull) rtb::kd_test_leaf<true, false>(tris + size_t(e.ref) * rtb::KD8_TRI_FLOATS, e.cnt, s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, eps, s.best);
                else rtb::kd_test_leaf<false, false>(tris + size_t(e.ref) * rtb::KD8_TRI_FLOATS, e.cnt, s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, eps, s.best);
                if (s.any_hit && s.best.t <= s.t_far) break;
                continue;
            }
            ++node_visits[i];
            const WideNode& w = nodes[e.ref];
            Entry hit[16]; int m = 0;
            for (int k = 0; k < w.n; ++k) {
                float t_in, t_out;
                rtb::bvh_slab(ix, iy, iz, cx, cy, cz, w.c[k].lo[0], w.c[k].lo[1], w.c[k].lo[2], w.c[k].hi[0], w.c[k].hi[1], w.c[k].hi[2], t_in, t_out);
                if (!((t_in <= t_out) & (t_in <= lim))) continue;
                int j = m++;
                while (j > 0 && hit[j - 1].t0 < t_in) { hit[j] = hit[j - 1]; --j; }       // descending: the nearest is pushed last
                hit[j] = Entry{w.c[k].ref, w.c[k].cnt, t_in};
            }
            for (int j = 0; j < m; ++j) stack.push_back(hit[j]);
        }
    }
}
