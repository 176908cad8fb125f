// host/bvh_build.cpp - builder and flattener of the backend's bounding-volume hierarchy (the
// "wider-node accel" the reference's author lists as a TODO, README.md:118-124), the structure behind the accelerated
// query mode (csrc/rt_bvh.cuh).
//
// Why a BVH next to the two kd-trees: a ray that grazes a surface crosses many small kd leaves (config 2: 46 node visits and
// 14 triangle tests per shadow ray in the SAH kd-tree); a BVH has no empty cells to step through and references every
// triangle exactly once.  The answer of a closest-hit query does not depend on the structure (rt_kd8.cuh, header note): every
// triangle test is the reference's own arithmetic, ties between different triangles are re-run in reference order.
//
// Build: top-down, binned surface-area heuristic over triangle centroids (16 bins per axis), children bounds = union of the
// triangles' own boxes (TriGeom::bmin/bmax).  It is a policy of the task-parallel driver (kd_parallel.hpp), so the tree
// is the same for any thread count and comes out in DFS pre-order (child0 == index + 1), leaf lists in creation order.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kd_build.hpp"
#include "kd_parallel.hpp"

namespace rtb {

namespace {

constexpr int BVH_BINS = 16;
// cost of one node step relative to one triangle test in the leaf-or-split decision (RT_B200_BVH_COST: tuning sweeps only)
constexpr double BVH_COST_TRI = 1.0;
static double bvh_cost_step() {
    static const double v = [] { double x = 1.0; if (const char* e = std::getenv("RT_B200_BVH_COST")) std::sscanf(e, "%lf", &x); return x; }();
    return v;
}

inline double half_area(const float* lo, const float* hi) {
    const double dx = double(hi[0]) - lo[0], dy = double(hi[1]) - lo[1], dz = double(hi[2]) - lo[2];
    return dx * dy + dy * dz + dz * dx;
}
struct Box {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    void add(const float* a, const float* b) {
        for (int c = 0; c < 3; ++c) { lo[c] = std::min(lo[c], a[c]); hi[c] = std::max(hi[c], b[c]); }
    }
    void add(const Box& o) { add(o.lo, o.hi); }
    bool valid() const { return lo[0] <= hi[0]; }
};

struct BvhPolicy {
    struct Work {
        uint64_t depth = 0;
        float lo[3], hi[3];
        std::vector<uint32_t> tris;
        size_t size() const { return tris.size(); }
        bool empty() const { return tris.empty(); }
    };
    const Geometry& g;
    uint32_t max_leaf, max_depth;

    static float centroid(const TriGeom& t, int c) { return 0.5f * t.bmin[c] + 0.5f * t.bmax[c]; }

    bool expand(Work& w, uint32_t& axis_out, float& split_out, Work& c0, Work& c1) const {
        const size_t N = w.tris.size();
        if (N <= 1 || w.depth >= max_depth) return false;
        float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (uint32_t id : w.tris)
            for (int c = 0; c < 3; ++c) {
                const float x = centroid(g.tris[id], c);
                clo[c] = std::min(clo[c], x); chi[c] = std::max(chi[c], x);
            }
        const double area = half_area(w.lo, w.hi);
        int best_axis = -1, best_bin = 0;
        double best_cost = DBL_MAX;
        for (int axis = 0; axis < 3; ++axis) {
            const double width = double(chi[axis]) - clo[axis];
            if (!(width > 0)) continue;
            uint32_t cnt[BVH_BINS] = {0};
            Box bb[BVH_BINS];
            const double scale = BVH_BINS / width;
            for (uint32_t id : w.tris) {
                const TriGeom& t = g.tris[id];
                int b = int((double(centroid(t, axis)) - clo[axis]) * scale);
                b = std::min(std::max(b, 0), BVH_BINS - 1);
                ++cnt[b]; bb[b].add(t.bmin, t.bmax);
            }
            // sweep: right-to-left suffix boxes, then left-to-right
            Box suf[BVH_BINS]; uint32_t sufn[BVH_BINS];
            Box acc; uint32_t n = 0;
            for (int b = BVH_BINS - 1; b >= 0; --b) { if (cnt[b]) acc.add(bb[b]); n += cnt[b]; suf[b] = acc; sufn[b] = n; }
            Box left; uint32_t nl = 0;
            for (int b = 0; b + 1 < BVH_BINS; ++b) {
                if (cnt[b]) left.add(bb[b]);
                nl += cnt[b];
                const uint32_t nr = sufn[b + 1];
                if (nl == 0 || nr == 0) continue;
                const double cost = half_area(left.lo, left.hi) * nl + half_area(suf[b + 1].lo, suf[b + 1].hi) * nr;
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
            }
        }
        if (best_axis >= 0 && N <= max_leaf && area > 0) {
            const double split_cost = bvh_cost_step() + BVH_COST_TRI * best_cost / area;
            if (BVH_COST_TRI * double(N) <= split_cost) return false;        // cheaper as a leaf
        }
        c0.depth = c1.depth = w.depth + 1;
        c0.tris.clear(); c1.tris.clear();
        if (best_axis >= 0) {
            const double scale = BVH_BINS / (double(chi[best_axis]) - clo[best_axis]);
            for (uint32_t id : w.tris) {
                int b = int((double(centroid(g.tris[id], best_axis)) - clo[best_axis]) * scale);
                b = std::min(std::max(b, 0), BVH_BINS - 1);
                (b <= best_bin ? c0 : c1).tris.push_back(id);
            }
            axis_out = uint32_t(best_axis);
            split_out = float(clo[best_axis] + (best_bin + 1) / scale);
        } else {
            if (N <= max_leaf) return false;                                 // all centroids coincide: nothing to separate
            c0.tris.assign(w.tris.begin(), w.tris.begin() + N / 2);          // ... but too many for one leaf: halve the list
            c1.tris.assign(w.tris.begin() + N / 2, w.tris.end());
            axis_out = 0; split_out = clo[0];
        }
        for (Work* c : {&c0, &c1}) {
            Box b;
            for (uint32_t id : c->tris) b.add(g.tris[id].bmin, g.tris[id].bmax);
            std::memcpy(c->lo, b.lo, 12); std::memcpy(c->hi, b.hi, 12);
        }
        return true;
    }
    void emit(const Work& w, std::vector<uint32_t>& refs) const { refs.insert(refs.end(), w.tris.begin(), w.tris.end()); }
};

}  // namespace

KdTree build_bvh(const Geometry& g, uint32_t max_leaf) {
    BvhPolicy pol{g, std::max<uint32_t>(1, std::min<uint32_t>(max_leaf, 255)), 44};
    BvhPolicy::Work root;
    root.depth = 0;
    Box b;
    for (const TriGeom& t : g.tris) b.add(t.bmin, t.bmax);
    if (!b.valid()) { std::memcpy(b.lo, g.root_min, 12); std::memcpy(b.hi, g.root_max, 12); }
    std::memcpy(root.lo, b.lo, 12); std::memcpy(root.hi, b.hi, 12);
    root.tris.resize(g.tris.size());
    for (uint32_t i = 0; i < root.tris.size(); ++i) root.tris[i] = i;
    return kd_build_parallel(pol, std::move(root));
}

// 64-byte nodes, one per INNER node of the tree, each holding its two children's boxes and references:
//   float4 { c0.min.xyz, c0.max.x }  float4 { c0.max.yz, c1.min.xy }  float4 { c1.min.z, c1.max.xyz }  uint4 { ref0, ref1, cnt0, cnt1 }
//   cnt == 0: ref is the index of an inner node;  cnt > 0: the child is a leaf of cnt triangles starting at triangle record ref
//   cnt == 0xFFFFFFFF: no such child (only the synthetic root of a one-leaf tree has one)
// tris: the 48-byte records of the accelerated kd-tree's layout, one per triangle (a BVH references every triangle once).
BvhLayout flatten_bvh(const Geometry& g, const KdTree& t) {
    BvhLayout d;
    auto bits = [](float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; };
    const uint64_t n_nodes = t.nodes.size();
    d.n_refs = t.refs.size();
    if (d.n_refs >= (1ull << 31)) throw rt_error(RT_ERR_UNSUPPORTED, "too many triangles for the BVH layout");
    // inner nodes are renumbered densely in tree order
    std::vector<uint32_t> inner_index(n_nodes, 0);
    uint64_t n_inner = 0;
    for (uint64_t i = 0; i < n_nodes; ++i)
        if (t.nodes[i].first_ref == KD_NONE) inner_index[i] = uint32_t(n_inner++);
    const bool leaf_root = n_nodes > 0 && t.nodes[0].first_ref != KD_NONE;
    d.n_nodes = std::max<uint64_t>(n_inner, 1);
    d.nodes.assign(16 * d.n_nodes, 0u);
    // Every stored box is grown by an absolute pad (2e-5 of the largest coordinate magnitude of the scene).  The reference accepts
    // a hit when its ROUNDED u, v, t pass (kd_tree_simd.hpp:47,54,57), so an accepted hit point may lie a few ulps of the scene
    // scale outside the triangle's own box - e.g. a ray that starts on the ceiling plane and hits a wall 3e-7 above the wall's
    // top edge; the relative slack of the slab test (rt_bvh.cuh) vanishes at t ~ 0 and cannot cover that.
    float scale = 0.0f;
    if (n_nodes)
        for (int c = 0; c < 3; ++c) scale = std::max(scale, std::max(std::fabs(t.nodes[0].bmin[c]), std::fabs(t.nodes[0].bmax[c])));
    const float pad = 2e-5f * scale;
    auto put_child = [&](uint32_t* node, int slot, const KdNode* c, uint64_t ci) {
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};   // inverted: never hit
        uint32_t ref = 0, cnt = 0xFFFFFFFFu;
        if (c) {
            for (int a = 0; a < 3; ++a) { lo[a] = c->bmin[a] - pad; hi[a] = c->bmax[a] + pad; }
            if (c->first_ref != KD_NONE) { if (c->ref_count) { ref = uint32_t(c->first_ref); cnt = uint32_t(c->ref_count); } }   // an empty leaf is no child
            else { ref = inner_index[ci]; cnt = 0; }
        }
        const int b = slot ? 6 : 0;
        node[b + 0] = bits(lo[0]); node[b + 1] = bits(lo[1]); node[b + 2] = bits(lo[2]);
        node[b + 3] = bits(hi[0]); node[b + 4] = bits(hi[1]); node[b + 5] = bits(hi[2]);
        node[12 + slot] = ref; node[14 + slot] = cnt;
    };
    if (n_nodes == 0 || leaf_root) {
        put_child(d.nodes.data(), 0, n_nodes ? &t.nodes[0] : nullptr, 0);
        put_child(d.nodes.data(), 1, nullptr, 0);
    } else {
        parallel_for(n_nodes, 1 << 14, [&](uint64_t nb, uint64_t ne) {
            for (uint64_t i = nb; i < ne; ++i) {
                const KdNode& n = t.nodes[i];
                if (n.first_ref != KD_NONE) continue;
                uint32_t* node = d.nodes.data() + 16 * uint64_t(inner_index[i]);
                put_child(node, 0, n.child0 != KD_NONE ? &t.nodes[n.child0] : nullptr, n.child0);
                put_child(node, 1, n.child1 != KD_NONE ? &t.nodes[n.child1] : nullptr, n.child1);
            }
        });
    }
    d.tris.assign(12 * std::max<uint64_t>(d.n_refs, 1), 0u);
    parallel_for(d.n_refs, 1 << 16, [&](uint64_t rb, uint64_t re) {
        for (uint64_t r = rb; r < re; ++r) {
            const uint32_t id = t.refs[r];
            const TriGeom& tg = g.tris[id];
            uint32_t* p = d.tris.data() + 12 * r;
            p[0] = bits(tg.v0[0]); p[1] = bits(tg.v0[1]); p[2] = bits(tg.v0[2]); p[3] = id;
            p[4] = bits(tg.e1[0]); p[5] = bits(tg.e1[1]); p[6] = bits(tg.e1[2]);
            p[8] = bits(tg.e2[0]); p[9] = bits(tg.e2[1]); p[10] = bits(tg.e2[2]);
        }
    });
    if (n_nodes)
        for (int a = 0; a < 3; ++a) { d.root_min[a] = t.nodes[0].bmin[a] - pad; d.root_max[a] = t.nodes[0].bmax[a] + pad; }
    return d;
}

}  // namespace rtb
