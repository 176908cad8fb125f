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
#include <mutex>
#include <vector>

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

// The builder reads nothing of a triangle but its box, and a node's triangles are scattered all over the scene's arrays: gathered
// by triangle id, every access of a 10 M-triangle build is a cache miss (one thread: 200 ns per triangle and tree level).  So
// the work lists carry the boxes WITH the ids - 28 bytes per reference - and every pass of the build streams through memory.
struct TriBox { uint32_t id; float lo[3], hi[3]; };
inline float box_centroid(const TriBox& t, int c) { return 0.5f * t.lo[c] + 0.5f * t.hi[c]; }

struct BvhPolicy {
    struct Work {
        uint64_t depth = 0;
        float lo[3], hi[3];
        float clo[3], chi[3];            // bounds of the triangles' box centres (filled by the parent's partition pass)
        std::vector<TriBox> tris;
        size_t size() const { return tris.size(); }
        bool empty() const { return tris.empty(); }
    };
    uint32_t max_leaf, max_depth;

    // Two passes over the node's triangles: (1) bin the box centres on all three axes at once, (2) partition, which also
    // yields the children's boxes and centre bounds.  A big node (the top of a 10 M-triangle tree) splits both passes over the
    // build threads: counts and boxes merge by + / min / max and the partition keeps the list order (per-chunk counts, then
    // every chunk writes at its offsets), so the tree does not depend on the thread count.
    static constexpr size_t PARALLEL_NODE = size_t(1) << 18;

    // bins of the three axes; a bin's box is written when its first triangle arrives (most nodes of a tree are small: filling and
    // sweeping 48 empty boxes per node was most of a 10 M-triangle build)
    struct RawBox { float lo[3], hi[3]; };
    static void grow(RawBox& b, const float* lo, const float* hi) {
        for (int c = 0; c < 3; ++c) { b.lo[c] = std::min(b.lo[c], lo[c]); b.hi[c] = std::max(b.hi[c], hi[c]); }
    }
    struct Bins {
        uint32_t mask[3] = {0, 0, 0};                 // bit b: bin b of the axis holds something; cnt / bb of the other bins are not initialised
        uint32_t cnt[3][BVH_BINS];
        RawBox bb[3][BVH_BINS];
        void put(int a, int b, const float* lo, const float* hi) {
            if (!((mask[a] >> b) & 1u)) { mask[a] |= 1u << b; cnt[a][b] = 1; std::memcpy(bb[a][b].lo, lo, 12); std::memcpy(bb[a][b].hi, hi, 12); }
            else { ++cnt[a][b]; grow(bb[a][b], lo, hi); }
        }
        void merge(const Bins& o) {
            for (int a = 0; a < 3; ++a)
                for (uint32_t m = o.mask[a]; m; m &= m - 1u) {
                    const int b = __builtin_ctz(m);
                    if (!((mask[a] >> b) & 1u)) { mask[a] |= 1u << b; cnt[a][b] = o.cnt[a][b]; bb[a][b] = o.bb[a][b]; }
                    else { cnt[a][b] += o.cnt[a][b]; grow(bb[a][b], o.bb[a][b].lo, o.bb[a][b].hi); }
                }
        }
    };
    static int bin_of(float centre, double lo, double scale) {
        const int b = int((double(centre) - lo) * scale);
        return std::min(std::max(b, 0), BVH_BINS - 1);
    }
    void bin_range(const TriBox* refs, size_t n, const float* clo, const double* scale, Bins& out) const {
        for (size_t k = 0; k < n; ++k) {
            const TriBox& t = refs[k];
            for (int a = 0; a < 3; ++a) {
                if (!(scale[a] > 0)) continue;
                out.put(a, bin_of(box_centroid(t, a), clo[a], scale[a]), t.lo, t.hi);
            }
        }
    }
    struct Side { Box box, centres; };
    // the references [refs, refs + n) go left (bin <= best_bin) or right, appended at out0 / out1 in order; returns how many went left
    size_t partition_range(const TriBox* refs, size_t n, int axis, double lo, double scale, int best_bin, TriBox* out0, TriBox* out1,
                           Side& s0, Side& s1) const {
        size_t n0 = 0, n1 = 0;
        for (size_t k = 0; k < n; ++k) {
            const TriBox& t = refs[k];
            const float c[3] = {box_centroid(t, 0), box_centroid(t, 1), box_centroid(t, 2)};
            const bool left = bin_of(c[axis], lo, scale) <= best_bin;
            Side& sd = left ? s0 : s1;
            sd.box.add(t.lo, t.hi); sd.centres.add(c, c);
            if (left) out0[n0++] = t; else out1[n1++] = t;
        }
        return n0;
    }
    static void finish_child(Work& c, const Side& sd) {
        std::memcpy(c.lo, sd.box.lo, 12); std::memcpy(c.hi, sd.box.hi, 12);
        std::memcpy(c.clo, sd.centres.lo, 12); std::memcpy(c.chi, sd.centres.hi, 12);
    }

    bool expand(Work& w, uint32_t& axis_out, float& split_out, Work& c0, Work& c1) const {
        const size_t N = w.tris.size();
        if (N <= 1 || w.depth >= max_depth) return false;
        const float* clo = w.clo; const float* chi = w.chi;
        const double area = half_area(w.lo, w.hi);
        double scale[3];
        for (int a = 0; a < 3; ++a) { const double width = double(chi[a]) - clo[a]; scale[a] = width > 0 ? BVH_BINS / width : 0.0; }
        const bool wide_node = N >= PARALLEL_NODE && build_threads() > 1;
        Bins bins;
        if (wide_node) {
            std::mutex m;
            parallel_for(N, PARALLEL_NODE / 8, [&](uint64_t b, uint64_t e) {
                Bins local;
                bin_range(w.tris.data() + b, size_t(e - b), clo, scale, local);
                std::lock_guard<std::mutex> g(m);
                bins.merge(local);
            });
        } else bin_range(w.tris.data(), N, clo, scale, bins);
        int best_axis = -1, best_bin = 0;
        double best_cost = DBL_MAX;
        for (int axis = 0; axis < 3; ++axis) {
            if (!(scale[axis] > 0)) continue;
            const uint32_t* cnt = bins.cnt[axis];
            const RawBox* bb = bins.bb[axis];
            // the occupied bins, ascending.  A split behind an EMPTY bin costs exactly what the split behind the last occupied
            // bin before it costs (same two boxes, same counts), and the first of equal costs wins: only occupied bins compete
            int occ[BVH_BINS], n_occ = 0;
            for (uint32_t m = bins.mask[axis]; m; m &= m - 1u) occ[n_occ++] = __builtin_ctz(m);
            if (n_occ < 2) continue;
            RawBox suf[BVH_BINS]; uint32_t sufn[BVH_BINS];                  // indexed like occ: everything from occ[k] on
            suf[n_occ - 1] = bb[occ[n_occ - 1]]; sufn[n_occ - 1] = cnt[occ[n_occ - 1]];
            for (int k = n_occ - 2; k >= 1; --k) { suf[k] = suf[k + 1]; grow(suf[k], bb[occ[k]].lo, bb[occ[k]].hi); sufn[k] = sufn[k + 1] + cnt[occ[k]]; }
            RawBox left = bb[occ[0]]; uint32_t nl = 0;
            for (int k = 0; k + 1 < n_occ; ++k) {
                if (k) grow(left, bb[occ[k]].lo, bb[occ[k]].hi);
                nl += cnt[occ[k]];
                const double cost = half_area(left.lo, left.hi) * nl + half_area(suf[k + 1].lo, suf[k + 1].hi) * sufn[k + 1];
                if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = occ[k]; }
            }
        }
        if (best_axis >= 0 && N <= max_leaf && area > 0) {
            const double split_cost = bvh_cost_step() + BVH_COST_TRI * best_cost / area;
            if (BVH_COST_TRI * double(N) <= split_cost) return false;        // cheaper as a leaf
        }
        c0.depth = c1.depth = w.depth + 1;
        c0.tris.clear(); c1.tris.clear();
        if (best_axis >= 0) {
            const double lo = clo[best_axis], sc = scale[best_axis];
            Side s0, s1;
            if (wide_node) {
                // chunk-wise: count, place the chunks at their offsets, partition again into place (the list order is kept)
                const uint64_t threads = std::max<uint64_t>(1, std::min<uint64_t>(build_threads(), N / (PARALLEL_NODE / 8)));
                const uint64_t step = (N + threads - 1) / threads, chunks = (N + step - 1) / step;
                std::vector<size_t> left_of(chunks, 0);
                parallel_for(chunks, 1, [&](uint64_t cb, uint64_t ce) {
                    for (uint64_t c = cb; c < ce; ++c) {
                        const size_t b = size_t(c * step), e = std::min(N, size_t((c + 1) * step));
                        size_t n0 = 0;
                        for (size_t k = b; k < e; ++k) n0 += bin_of(box_centroid(w.tris[k], best_axis), lo, sc) <= best_bin;
                        left_of[c] = n0;
                    }
                });
                std::vector<size_t> off0(chunks + 1, 0), off1(chunks + 1, 0);
                for (uint64_t c = 0; c < chunks; ++c) {
                    const size_t b = size_t(c * step), e = std::min(N, size_t((c + 1) * step));
                    off0[c + 1] = off0[c] + left_of[c]; off1[c + 1] = off1[c] + (e - b - left_of[c]);
                }
                c0.tris.resize(off0[chunks]); c1.tris.resize(off1[chunks]);
                std::vector<Side> side0(chunks), side1(chunks);
                parallel_for(chunks, 1, [&](uint64_t cb, uint64_t ce) {
                    for (uint64_t c = cb; c < ce; ++c) {
                        const size_t b = size_t(c * step), e = std::min(N, size_t((c + 1) * step));
                        partition_range(w.tris.data() + b, e - b, best_axis, lo, sc, best_bin, c0.tris.data() + off0[c], c1.tris.data() + off1[c],
                                        side0[c], side1[c]);
                    }
                });
                for (uint64_t c = 0; c < chunks; ++c) {
                    if (side0[c].box.valid()) { s0.box.add(side0[c].box); s0.centres.add(side0[c].centres); }
                    if (side1[c].box.valid()) { s1.box.add(side1[c].box); s1.centres.add(side1[c].centres); }
                }
            } else {
                size_t n0 = 0;
                for (uint32_t m = bins.mask[best_axis] & ((2u << best_bin) - 1u); m; m &= m - 1u) n0 += bins.cnt[best_axis][__builtin_ctz(m)];
                c0.tris.resize(n0); c1.tris.resize(N - n0);                   // the bins know how many go left
                partition_range(w.tris.data(), N, best_axis, lo, sc, best_bin, c0.tris.data(), c1.tris.data(), s0, s1);
            }
            finish_child(c0, s0); finish_child(c1, s1);
            axis_out = uint32_t(best_axis);
            split_out = float(clo[best_axis] + (best_bin + 1) / sc);
        } else {
            if (N <= max_leaf) return false;                                 // all centroids coincide: nothing to separate
            c0.tris.assign(w.tris.begin(), w.tris.begin() + N / 2);          // ... but too many for one leaf: halve the list
            c1.tris.assign(w.tris.begin() + N / 2, w.tris.end());
            axis_out = 0; split_out = clo[0];
            for (Work* c : {&c0, &c1}) {
                Side sd;
                for (const TriBox& t : c->tris) {
                    const float cc[3] = {box_centroid(t, 0), box_centroid(t, 1), box_centroid(t, 2)};
                    sd.box.add(t.lo, t.hi); sd.centres.add(cc, cc);
                }
                finish_child(*c, sd);
            }
        }
        return true;
    }
    void emit(const Work& w, std::vector<uint32_t>& refs) const { for (const TriBox& t : w.tris) refs.push_back(t.id); }
};

}  // namespace

KdTree build_bvh(const Geometry& g, uint32_t max_leaf) {
    const size_t n = g.tris.size();
    BvhPolicy pol{std::max<uint32_t>(1, std::min<uint32_t>(max_leaf, 255)), 44};
    BvhPolicy::Work root;
    root.depth = 0;
    root.tris.resize(n);
    parallel_for(n, 1 << 16, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) { root.tris[i].id = uint32_t(i); std::memcpy(root.tris[i].lo, g.tris[i].bmin, 12); std::memcpy(root.tris[i].hi, g.tris[i].bmax, 12); }
    });
    Box b, centres;
    for (const TriBox& t : root.tris) {
        const float c[3] = {box_centroid(t, 0), box_centroid(t, 1), box_centroid(t, 2)};
        b.add(t.lo, t.hi); centres.add(c, c);
    }
    if (!b.valid()) { std::memcpy(b.lo, g.root_min, 12); std::memcpy(b.hi, g.root_max, 12); }
    std::memcpy(root.lo, b.lo, 12); std::memcpy(root.hi, b.hi, 12);
    std::memcpy(root.clo, centres.lo, 12); std::memcpy(root.chi, centres.hi, 12);
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
