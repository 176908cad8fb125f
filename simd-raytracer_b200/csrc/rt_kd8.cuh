// csrc/rt_kd8.cuh - the accelerated closest-hit query (RT_FLAG_ORDERED): front-to-back split-plane traversal over the
// compact 8-byte nodes of the backend's OWN, deeper kd-tree, one ray per thread.
//
// Why a second tree: the reference's default template arguments <max_depth 8, max_leaf_size 64> (kd_tree_simd.hpp:65-66)
// leave up to 752 triangles in a leaf and its LIFO order is not front-to-back, so a ray that enters the dragon's box
// pays for hundreds of triangle tests (config 2: 237 per shadow ray).  The answer of a closest-hit query does not
// depend on the tree: every triangle is tested with the reference's own arithmetic (test_tri below = the same
// expressions as kd_tree_simd.hpp:25-60, no FMA in exact mode), so t/u/v of the winner are the reference's bits, and
// the minimum over "all triangles whose leaf the ray passes" is the same set minimum.  What the reference's visit
// order decides is only which of two DIFFERENT triangles with exactly equal t is reported (a shared edge; the cube
// standing on the floor in config 1 - coplanar faces).  That decision depends on the reference's leaf order, so this
// traversal does not guess: it records that a second triangle tied with the winner (KdHit::tie_t == t) and the
// caller re-runs exactly those rays (~0.02 %) through the reference-order query (trace_any in rt_device.cuh).
//
// The tree is built by the same host builder (host/kd_build.cpp) with more depth and small leaves, flattened to
//   nodes8  : 8 B per node.  inner: { f32 split, u32 axis | has0<<2 | has1<<3 | child1<<4 }, child0 = index+1
//                            leaf : { u32 first_packet, u32 3 | n_packets<<2 }
//   tris    : 48 B per leaf reference, array of structures { v0.xyz, id } { e1.xyz, - } { e2.xyz, - }
// Node boxes are implicit: the ray carries its parametric interval [t0,t1] down the tree.  Interval comparisons are
// widened by a relative slack so that rounding can only ADD a visit, never drop one; hits are never clipped to the
// leaf interval, pruning uses only "the node starts beyond the closest hit so far".
//
// Compiles as CUDA device code and as plain C++ (tests/helpers/kd8_host.cpp runs this very source on the CPU against the
// oracle), hence the small portability macros.
#pragma once

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

#ifndef KD8_COUNT_NODE
#define KD8_COUNT_NODE() ((void)0)      // instrumentation hooks for host-side experiments
#define KD8_COUNT_TRI() ((void)0)
#endif

namespace rtb {

struct KdHit { float t, u, v; int tri; float tie_t; };   // tie_t == t: another triangle has exactly the winner's t
// tri == KD_RERUN: the query has to be answered by the reference-order traversal (see kd8_init / bvh_init)
constexpr int KD_RERUN = -3;

RT_HD float kd_bits_to_float(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; std::memcpy(&f, &u, 4); return f;
#endif
}
RT_HD float kd_rcp_estimate(float x) {
#if defined(__CUDA_ARCH__)
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
    return 1.0f / x;
#endif
}
RT_HD float kd_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return fmaf(a, b, c);
#else
    return std::fmaf(a, b, c);
#endif
}
RT_HD bool kd_sign_bit(float f) {
#if defined(__CUDA_ARCH__)
    return (__float_as_uint(f) >> 31) != 0u;
#else
    uint32_t u; std::memcpy(&u, &f, 4); return (u >> 31) != 0u;
#endif
}
RT_HD float kd_min(float a, float b) { return (b < a) ? b : a; }      // NaN in b is ignored
RT_HD float kd_max(float a, float b) { return (a < b) ? b : a; }

// one ray vs one triangle; exact mode = kd_tree_simd.hpp:25-60 in order, behind the conservative rcp pre-filter
// (see rt_device.cuh test_lane); an exact-t tie with a different triangle is recorded, not resolved
template <bool CULL, bool FAST>
RT_HD void kd_test_tri(float v0x, float v0y, float v0z, float e1x, float e1y, float e1z, float e2x, float e2y, float e2z, int id,
                       float ox, float oy, float oz, float dx, float dy, float dz, float eps, KdHit& best) {
    float u, v, t;
    if (FAST) {
        const float pvx = kd_fma(dy, e2z, -(dz * e2y)), pvy = kd_fma(dz, e2x, -(dx * e2z)), pvz = kd_fma(dx, e2y, -(dy * e2x));
        const float det = kd_fma(e1z, pvz, kd_fma(e1y, pvy, e1x * pvx));
        if (!(CULL ? (eps <= det) : (eps <= fabsf(det)))) return;
        const float inv_det = 1.0f / det;
        const float tx = ox - v0x, ty = oy - v0y, tz = oz - v0z;
        u = kd_fma(tz, pvz, kd_fma(ty, pvy, tx * pvx)) * inv_det;
        if (!((0.0f <= u) & (u <= 1.0f))) return;
        const float qx = kd_fma(ty, e1z, -(tz * e1y)), qy = kd_fma(tz, e1x, -(tx * e1z)), qz = kd_fma(tx, e1y, -(ty * e1x));
        v = kd_fma(dz, qz, kd_fma(dy, qy, dx * qx)) * inv_det;
        if (!((0.0f <= v) & (u + v <= 1.0f))) return;
        t = kd_fma(e2z, qz, kd_fma(e2y, qy, e2x * qx)) * inv_det;
        if (!(eps < t)) return;
    } else {
        constexpr float M = 1e-4f;
        const float pvx = dy * e2z - dz * e2y;                                                           // :27
        const float pvy = dz * e2x - dx * e2z;                                                           // :28
        const float pvz = dx * e2y - dy * e2x;                                                           // :29
        const float det = e1x * pvx + e1y * pvy + e1z * pvz;                                             // :31
        const float tx = ox - v0x, ty = oy - v0y, tz = oz - v0z;                                         // :42-44
        const float un = tx * pvx + ty * pvy + tz * pvz;
        const float r = kd_rcp_estimate(det);
        const float ua = un * r;
        // the determinant test (:33-38) rarely rejects: it shares the branch of the first pre-filter, so that the three rows of
        // the triangle are loaded together instead of v0 waiting behind a branch of its own
        if (!(CULL ? (eps <= det) : (eps <= fabsf(det))) | (ua < -M) | (ua > 1.0f + M)) return;
        const float qx = ty * e1z - tz * e1y;                                                            // :49
        const float qy = tz * e1x - tx * e1z;                                                            // :50
        const float qz = tx * e1y - ty * e1x;                                                            // :51
        const float vn = dx * qx + dy * qy + dz * qz;
        const float va = vn * r;
        if ((va < -M) | (ua + va > 1.0f + 3.0f * M)) return;
        const float inv_det = 1.0f / det;                                                                // :40
        u = un * inv_det;                                                                                // :46
        v = vn * inv_det;                                                                                // :53
        t = (e2x * qx + e2y * qy + e2z * qz) * inv_det;                                                  // :56
        if (!((0.0f <= u) & (u <= 1.0f) & (0.0f <= v) & (u + v <= 1.0f) & (eps < t))) return;           // :47,:54,:57
    }
    if (t < best.t) { best.t = t; best.u = u; best.v = v; best.tri = id; }
    else if (t == best.t && id != best.tri) best.tie_t = t;
}

struct KdRow { float x, y, z, w; };
RT_HD KdRow kd_load_row(const float* p) {
#if defined(__CUDA_ARCH__)
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    return KdRow{q.x, q.y, q.z, q.w};
#else
    return KdRow{p[0], p[1], p[2], p[3]};
#endif
}
RT_HD int kd_as_int(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int i; std::memcpy(&i, &f, 4); return i;
#endif
}

// leaf triangles of the accelerated tree: 48 B per reference, three aligned 16-byte rows
//   { v0.xyz, id }  { e1.xyz, - }  { e2.xyz, - }
// (one ray per thread walks small leaves, so array-of-structures beats the 4-wide SoA packets of the reference-order path:
// no padding lanes, three LDG.128 per triangle)
constexpr uint32_t KD8_TRI_FLOATS = 12;

template <bool CULL, bool FAST>
RT_HD void kd_test_leaf(const float* tr, uint32_t count, float ox, float oy, float oz, float dx, float dy, float dz, float eps,
                        KdHit& best) {
    for (uint32_t k = 0; k < count; ++k, tr += KD8_TRI_FLOATS) {
        KD8_COUNT_TRI();
        const KdRow a = kd_load_row(tr), b = kd_load_row(tr + 4), c = kd_load_row(tr + 8);
        kd_test_tri<CULL, FAST>(a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z, kd_as_int(a.w), ox, oy, oz, dx, dy, dz, eps, best);
    }
}

constexpr int KD8_STACK = 32;     // tree depth is capped at 30 by the host

// one 16-byte vector store / load per push / pop (thread-local memory)
struct alignas(16) KdStackEntry { uint32_t node; float t0, t1; uint32_t pad; };
RT_HD void kd_stack_put(KdStackEntry* e, uint32_t node, float t0, float t1) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<float4*>(e) = make_float4(__uint_as_float(node), t0, t1, 0.0f);
#else
    e->node = node; e->t0 = t0; e->t1 = t1; e->pad = 0;
#endif
}
RT_HD void kd_stack_get(const KdStackEntry* e, uint32_t& node, float& t0, float& t1) {
#if defined(__CUDA_ARCH__)
    const float4 q = *reinterpret_cast<const float4*>(e);
    node = __float_as_uint(q.x); t0 = q.y; t1 = q.z;
#else
    node = e->node; t0 = e->t0; t1 = e->t1;
#endif
}
// x, y or z by a run-time axis without a branch (two predicates, two selects)
RT_HD float kd_sel3(uint32_t axis, float x, float y, float z) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("{\n\t.reg .pred p0, p1;\n\tsetp.eq.u32 p0, %1, 0;\n\tsetp.eq.u32 p1, %1, 1;\n\t"
        "selp.f32 %0, %3, %4, p1;\n\tselp.f32 %0, %2, %0, p0;\n\t}"
        : "=f"(r) : "r"(axis), "f"(x), "f"(y), "f"(z));
    return r;
#else
    return axis == 0u ? x : (axis == 1u ? y : z);
#endif
}

// Traversal state of one ray, so that a kernel can advance many rays in lock step and refill finished lanes
// (rt_stream.cuh).  phase: WALK = standing at `node` with interval [t0,t1]; LEAF = parked at the leaf `node`, which still
// has to be tested; DONE = query finished, `best` is the answer.
enum : int { KD8_WALK = 0, KD8_LEAF = 1, KD8_DONE = 2 };

struct Kd8State {
    float ox, oy, oz, dx, dy, dz;                // 1/d is re-derived per node visit (one MUFU) instead of living in three registers
    float t0, t1, t_far;
    uint32_t node;
    int sp, phase;
    bool any_hit;
    KdHit best;
};   // scalars only, so that it lives in registers; the stack is a separate (local-memory) array passed alongside

// returns false when the ray misses the root box (the query is then finished: a miss)
RT_HD bool kd8_init(Kd8State& s, const float* root_min, const float* root_max, float ox, float oy, float oz, float dx, float dy, float dz,
                    float t_far, bool any_hit) {
    s.ox = ox; s.oy = oy; s.oz = oz; s.dx = dx; s.dy = dy; s.dz = dz;
    const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
    s.t_far = t_far; s.any_hit = any_hit;
    s.best.t = FLT_MAX; s.best.u = 0.0f; s.best.v = 0.0f; s.best.tri = -1; s.best.tie_t = -1.0f;
    s.node = 0; s.sp = 0; s.phase = KD8_DONE;
    // parametric interval of the root box (aabb3.hpp:74-90 semantics: NaN from 0*inf leaves a bound unchanged)
    float t0 = 0.0f, t1 = FLT_MAX;
    const float ax = (root_min[0] - ox) * ix, bx = (root_max[0] - ox) * ix;
    const float ay = (root_min[1] - oy) * iy, by = (root_max[1] - oy) * iy;
    const float az = (root_min[2] - oz) * iz, bz = (root_max[2] - oz) * iz;
    t0 = kd_max(t0, (bx < ax) ? bx : ax); t1 = kd_min(t1, (bx < ax) ? ax : bx);
    t0 = kd_max(t0, (by < ay) ? by : ay); t1 = kd_min(t1, (by < ay) ? ay : by);
    t0 = kd_max(t0, (bz < az) ? bz : az); t1 = kd_min(t1, (bz < az) ? az : bz);
    // the root box is the reference's own root box and this is the reference's own arithmetic, so the decision "the ray
    // misses the scene" is taken exactly as the reference takes it (strict t_max < t_min, aabb3.hpp:81-88): a ray that starts
    // ON the scene boundary and leaves it is a miss there even where a triangle test alone would still accept a wall
    const float S = 2e-6f;                       // relative widening of every interval comparison below the root
    if (t1 < t0) return false;
    // t1 == 0: the ray starts ON the scene boundary and leaves the scene at once.  Which wall triangles the reference still
    // tests then depends on its own leaf boxes, so such rays (rare) are answered by the reference-order traversal: the lane
    // finishes at once with the KD_RERUN mark
    if (!(0.0f < t1)) { s.best.tri = KD_RERUN; return true; }
    s.t0 = kd_max(0.0f, t0 - S * fabsf(t0));
    s.t1 = t1 + S * fabsf(t1);
    if (!(s.t0 <= t_far)) return false;
    s.phase = KD8_WALK;
    return true;
}

RT_HD void kd8_pop(Kd8State& s, const KdStackEntry* stack) {
    if (!s.sp) { s.phase = KD8_DONE; return; }
    --s.sp;
    kd_stack_get(stack + s.sp, s.node, s.t0, s.t1);
    s.phase = KD8_WALK;
}

// One node visit (phase WALK): prune and pop, go down one level, or park at a leaf.
//
// Children are ordered along the ray: `near` is the half the ray is in BEFORE it crosses the plane (the lower half when the
// direction component is positive), `far` the half after it.  With ts the plane's parameter the ray is in near on
// [t0, ts] and in far on [ts, t1]; a crossing behind the node (ts < t0) leaves only far, one beyond it (ts > t1) only
// near.  A zero direction component gives ts = +-inf with the right sign for this rule (and NaN - both halves - when the
// origin also lies on the plane).  An origin exactly ON the plane with a non-zero component is in far for every t > 0
// (a camera at x = 0 and a binned plane at 0.0): near is skipped outright instead of being walked with a zero-length interval.
// Written without branches between the node fetch and the push: every lane of a warp runs the same instructions.
RT_HD void kd8_node_step(Kd8State& s, KdStackEntry* stack, const uint32_t* __restrict__ nodes8) {
    const float S = 2e-6f;
    bool pop = s.t0 > kd_min(s.best.t, s.t_far);                            // the node starts beyond the closest hit so far
    if (!pop) {
        KD8_COUNT_NODE();
#if defined(__CUDA_ARCH__)
        const uint2 nd = __ldg(reinterpret_cast<const uint2*>(nodes8) + s.node);
        const uint32_t first = nd.x, word = nd.y;
#else
        const uint32_t first = nodes8[2 * s.node], word = nodes8[2 * s.node + 1];
#endif
        const uint32_t axis = word & 3u;
        if (axis == 3u) { s.phase = KD8_LEAF; return; }                     // the leaf phase reads the node again
        const float split = kd_bits_to_float(first);
        const float oa = kd_sel3(axis, s.ox, s.oy, s.oz);
        const float da = kd_sel3(axis, s.dx, s.dy, s.dz);
        // the 1-ulp reciprocal is good enough here: the interval comparisons below carry a relative slack of 2e-6, and a
        // flushed subnormal component behaves like zero (its displacement over any finite t is below float resolution)
        const float ia = kd_rcp_estimate(da);
        const float ts = (split - oa) * ia;                                 // +-inf for a zero component; NaN if also oa == split
        const bool neg = kd_sign_bit(ia);                                   // direction component < 0 (or -0.0)
        const uint32_t lo_c = s.node + 1u, hi_c = word >> 4;                // lower half [min, split] at index + 1, upper at word >> 4
        const uint32_t near_c = neg ? hi_c : lo_c, far_c = neg ? lo_c : hi_c;
        const uint32_t has_near = neg ? (word & 8u) : (word & 4u), has_far = neg ? (word & 4u) : (word & 8u);
        const float w = S * kd_max(kd_min(fabsf(ts), FLT_MAX), s.t1);       // finite for ts = +-inf; 0 <= t0 <= t1 here
        const bool on_plane = (oa == split) & (ts == ts);
        const bool go_near = (has_near != 0u) & !(ts < s.t0 - w) & !on_plane;   // a NaN ts fails both compares: both halves
        const bool go_far = (has_far != 0u) & !(ts > s.t1 + w);
        const float far_t0 = kd_max(s.t0, ts - w), near_t1 = kd_min(s.t1, ts + w);
        if (go_near & go_far) { kd_stack_put(stack + s.sp, far_c, far_t0, s.t1); ++s.sp; }
        pop = !(go_near | go_far);
        s.node = go_near ? near_c : far_c;
        s.t1 = go_near ? near_t1 : s.t1;
        s.t0 = go_near ? s.t0 : far_t0;
    }
    if (pop) kd8_pop(s, stack);
}

// The parked leaf (phase LEAF): test its triangles, then pop the next node or finish.
template <bool CULL, bool FAST>
RT_HD void kd8_leaf_step(Kd8State& s, const KdStackEntry* stack, const uint32_t* __restrict__ nodes8, const float* __restrict__ tris, float eps) {
#if defined(__CUDA_ARCH__)
    const uint2 nd = __ldg(reinterpret_cast<const uint2*>(nodes8) + s.node);
    const uint32_t first = nd.x, count = nd.y >> 2;
#else
    const uint32_t first = nodes8[2 * s.node], count = nodes8[2 * s.node + 1] >> 2;
#endif
    kd_test_leaf<CULL, FAST>(tris + size_t(first) * KD8_TRI_FLOATS, count, s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, eps, s.best);
    if (s.any_hit && s.best.t <= s.t_far) { s.phase = KD8_DONE; return; }
    kd8_pop(s, stack);
}

// Closest hit with t <= t_far (t_far = FLT_MAX for a plain query).  any_hit: return at the first hit inside [.., t_far].
template <bool CULL, bool FAST>
RT_HD KdHit kd8_trace(const uint32_t* __restrict__ nodes8, const float* __restrict__ tris, const float* root_min, const float* root_max,
                      float ox, float oy, float oz, float dx, float dy, float dz, float eps, float t_far, bool any_hit) {
    Kd8State s;
    KdStackEntry stack[KD8_STACK];
    if (kd8_init(s, root_min, root_max, ox, oy, oz, dx, dy, dz, t_far, any_hit))
        while (s.phase != KD8_DONE) {
            if (s.phase == KD8_WALK) kd8_node_step(s, stack, nodes8);
            else kd8_leaf_step<CULL, FAST>(s, stack, nodes8, tris, eps);
        }
    return s.best;
}

}  // namespace rtb
