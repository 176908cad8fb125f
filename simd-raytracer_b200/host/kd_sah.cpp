// host/kd_sah.cpp - builder of the backend's OWN kd-tree (the accelerated query mode, csrc/rt_kd8.cuh).
//
// The reference's builder (kd_tree_simd.hpp:146-185, restated in kd_build.cpp) must be kept for the parity-gated mode:
// it splits at the box midpoint on axis depth%3, assigns by triangle AABB overlap and stops at depth 8, which leaves
// hundreds of triangles per leaf.  The answer of a closest-hit query does not depend on the tree (rt_kd8.cuh), so the
// fast path is free to use a good one: surface-area heuristic over binned candidate planes, triangles clipped to the node
// box so that a big triangle is only referenced where it really is, empty space cut off, leaves of a few triangles.
//
// Conservative by construction: a triangle is referenced by every leaf its geometry touches - clipped bounds are grown by
// a relative epsilon, both children take a triangle that touches the plane (closed intervals, like aabb3.hpp:68-72), a
// reference is never dropped because clipping came out empty.  Output is the same KdTree as build_kd_tree: DFS pre-order
// (child0 == index + 1), leaf lists in creation order, so flatten_tree_only() applies unchanged.
#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>

#include "kd_build.hpp"
#include "kd_parallel.hpp"

namespace rtb {

namespace {

struct Ref { uint32_t tri; float lo[3], hi[3]; };
struct P3 { double x[3]; };

// bounds of (triangle ∩ box): Sutherland-Hodgman against the six planes, in double
bool clip_bounds(const TriGeom& t, const float* blo, const float* bhi, double grow, float* lo, float* hi) {
    P3 poly[16], tmp[16];
    int n = 3;
    for (int c = 0; c < 3; ++c) {
        poly[0].x[c] = t.v0[c];
        poly[1].x[c] = double(t.v0[c]) + double(t.e1[c]);
        poly[2].x[c] = double(t.v0[c]) + double(t.e2[c]);
    }
    for (int axis = 0; axis < 3 && n; ++axis)
        for (int side = 0; side < 2 && n; ++side) {
            const double plane = side ? double(bhi[axis]) + grow : double(blo[axis]) - grow;
            const double sgn = side ? -1.0 : 1.0;                 // inside: sgn * (x - plane) >= 0
            int m = 0;
            for (int i = 0; i < n; ++i) {
                const P3& a = poly[i];
                const P3& b = poly[(i + 1) % n];
                const double da = sgn * (a.x[axis] - plane), db = sgn * (b.x[axis] - plane);
                if (da >= 0) tmp[m++] = a;
                if ((da >= 0) != (db >= 0)) {
                    const double s = da / (da - db);
                    P3 p;
                    for (int c = 0; c < 3; ++c) p.x[c] = a.x[c] + s * (b.x[c] - a.x[c]);
                    p.x[axis] = plane;
                    tmp[m++] = p;
                }
            }
            n = m;
            std::memcpy(poly, tmp, sizeof(P3) * size_t(n));
        }
    if (n == 0) return false;
    for (int c = 0; c < 3; ++c) {
        double mn = poly[0].x[c], mx = poly[0].x[c];
        for (int i = 1; i < n; ++i) { mn = std::min(mn, poly[i].x[c]); mx = std::max(mx, poly[i].x[c]); }
        mn -= grow; mx += grow;
        // outward rounding to float, then clamp to the box
        float flo = float(mn), fhi = float(mx);
        if (double(flo) > mn) flo = std::nextafter(flo, -FLT_MAX);
        if (double(fhi) < mx) fhi = std::nextafter(fhi, FLT_MAX);
        lo[c] = std::max(flo, blo[c]);
        hi[c] = std::min(fhi, bhi[c]);
        if (lo[c] > hi[c]) { lo[c] = hi[c] = std::min(std::max(flo, blo[c]), bhi[c]); }
    }
    return true;
}

inline double half_area(const float* lo, const float* hi) {
    const double dx = double(hi[0]) - lo[0], dy = double(hi[1]) - lo[1], dz = double(hi[2]) - lo[2];
    return dx * dy + dy * dz + dz * dx;
}

constexpr int BINS = 32;
constexpr double COST_STEP = 1.0;
// cost of one triangle test relative to one node step, and the discount of a split that cuts off empty space; the
// RT_B200_SAH="tri_cost,empty_bonus" environment variable overrides them for tuning sweeps
struct SahCosts { double tri = 2.0, empty_bonus = 0.8; };
SahCosts sah_costs() {
    SahCosts c;
    if (const char* e = std::getenv("RT_B200_SAH")) {
        double a = 0, b = 0;
        if (std::sscanf(e, "%lf,%lf", &a, &b) == 2 && a > 0 && b > 0) { c.tri = a; c.empty_bonus = b; }
    }
    return c;
}

}  // namespace

namespace {

// the SAH builder as a policy of the parallel driver (kd_parallel.hpp)
struct SahPolicy {
    struct Work {
        uint64_t depth = 0;
        float lo[3], hi[3];
        std::vector<Ref> refs;
        size_t size() const { return refs.size(); }
        bool empty() const { return refs.empty(); }
    };
    const Geometry& g;
    uint32_t max_depth, max_leaf_size;
    double grow;
    double COST_TRI, EMPTY_BONUS;

    bool expand(Work& w, uint32_t& axis_out, float& split_out, Work& c0, Work& c1) const {
        const size_t N = w.refs.size();
        int best_axis = -1;
        float best_plane = 0;
        double best_cost = COST_TRI * double(N);                  // cost of making this a leaf
        if (w.depth < max_depth && N > std::max<uint32_t>(1, std::min<uint32_t>(max_leaf_size, 2))) {
            const double area = half_area(w.lo, w.hi);
            for (int axis = 0; axis < 3 && area > 0; ++axis) {
                const double lo = w.lo[axis], hi = w.hi[axis], width = hi - lo;
                if (!(width > 0)) continue;
                // candidate planes: BINS-1 uniform ones; counts use the closed-interval rule of the partition below
                uint32_t cnt_lo[BINS + 1] = {0}, cnt_hi[BINS + 1] = {0};
                for (const Ref& r : w.refs) {
                    int bl = int((double(r.lo[axis]) - lo) / width * BINS), bh = int((double(r.hi[axis]) - lo) / width * BINS);
                    bl = std::min(std::max(bl, 0), BINS); bh = std::min(std::max(bh, 0), BINS);
                    ++cnt_lo[bl]; ++cnt_hi[bh];
                }
                uint32_t nl = 0, nr = uint32_t(N);
                float other_lo[3], other_hi[3];
                std::memcpy(other_lo, w.lo, 12); std::memcpy(other_hi, w.hi, 12);
                for (int k = 1; k < BINS; ++k) {
                    nl += cnt_lo[k - 1];                           // refs starting before plane k
                    nr -= cnt_hi[k - 1];                           // refs ending before plane k are not on the right
                    const float plane = float(lo + width * (double(k) / BINS));
                    if (!(plane > w.lo[axis] && plane < w.hi[axis])) continue;
                    other_hi[axis] = plane;
                    const double al = half_area(w.lo, other_hi);
                    other_hi[axis] = w.hi[axis];
                    other_lo[axis] = plane;
                    const double ar = half_area(other_lo, w.hi);
                    other_lo[axis] = w.lo[axis];
                    const double NL = double(nl), NR = double(nr);     // estimates (bin resolution); the partition below is exact
                    double cost = COST_STEP + COST_TRI * (al * NL + ar * NR) / area;
                    if (NL == 0 || NR == 0) cost *= EMPTY_BONUS;
                    if (cost < best_cost) { best_cost = cost; best_axis = axis; best_plane = plane; }
                }
            }
        }
        if (best_axis < 0) return false;

        c0.refs.clear(); c1.refs.clear();
        c0.depth = c1.depth = w.depth + 1;
        std::memcpy(c0.lo, w.lo, 12); std::memcpy(c0.hi, w.hi, 12);
        std::memcpy(c1.lo, w.lo, 12); std::memcpy(c1.hi, w.hi, 12);
        c0.hi[best_axis] = best_plane;
        c1.lo[best_axis] = best_plane;
        for (const Ref& r : w.refs) {
            const bool left = r.lo[best_axis] <= best_plane, right = r.hi[best_axis] >= best_plane;
            if (left && right) {                                   // straddles (or touches) the plane: re-clip for tight bounds
                Ref a = r, b = r;
                if (!clip_bounds(g.tris[r.tri], c0.lo, c0.hi, grow, a.lo, a.hi)) {
                    a = r; a.hi[best_axis] = std::min(a.hi[best_axis], best_plane);
                }
                if (!clip_bounds(g.tris[r.tri], c1.lo, c1.hi, grow, b.lo, b.hi)) {
                    b = r; b.lo[best_axis] = std::max(b.lo[best_axis], best_plane);
                }
                c0.refs.push_back(a);
                c1.refs.push_back(b);
            } else if (left) c0.refs.push_back(r);
            else c1.refs.push_back(r);
        }
        // a split that separated nothing and cut no empty space would recurse for nothing
        if (c0.refs.size() == N && c1.refs.size() == N) { c0.refs.clear(); c1.refs.clear(); return false; }
        axis_out = uint32_t(best_axis); split_out = best_plane;
        return true;
    }
    void emit(const Work& w, std::vector<uint32_t>& refs) const {
        for (const Ref& r : w.refs) refs.push_back(r.tri);
    }
};

}  // namespace

KdTree build_kd_tree_sah(const Geometry& g, uint32_t max_depth, uint32_t max_leaf_size) {
    double extent = 0;
    for (int c = 0; c < 3; ++c) extent = std::max(extent, double(g.root_max[c]) - g.root_min[c]);
    const SahCosts costs = sah_costs();
    SahPolicy pol{g, max_depth, max_leaf_size, 1e-6 * (extent > 0 ? extent : 1.0), costs.tri, costs.empty_bonus};
    SahPolicy::Work root;
    root.depth = 0;
    std::memcpy(root.lo, g.root_min, 12); std::memcpy(root.hi, g.root_max, 12);
    root.refs.resize(g.tris.size());
    parallel_for(g.tris.size(), 1 << 16, [&](uint64_t b, uint64_t e) {
        for (uint64_t i = b; i < e; ++i) {
            Ref& r = root.refs[i];
            r.tri = uint32_t(i);
            std::memcpy(r.lo, g.tris[i].bmin, 12); std::memcpy(r.hi, g.tris[i].bmax, 12);
        }
    });
    return kd_build_parallel(pol, std::move(root));
}

}  // namespace rtb
