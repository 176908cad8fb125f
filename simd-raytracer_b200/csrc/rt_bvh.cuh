// csrc/rt_bvh.cuh - the accelerated closest-hit query over the backend's bounding-volume hierarchy (host/bvh_build.cpp),
// one ray per thread, near child first.  SURVEY.md section 8 row f4.
//
// The contract (rt_tri.cuh, whose triangle test kd_test_tri - the reference's own arithmetic, kd_tree_simd.hpp:25-60 - it
// uses): t/u/v of the winner are the reference's bits; the minimum is taken over
// every triangle whose box the ray touches; an exact-t tie between two different triangles is only recorded
// (KdHit::tie_t == t) and the caller re-runs those rays in reference order.  Box tests are conservative: both ends of a
// slab interval are widened by a relative slack far above the rounding of the slab arithmetic, a NaN (0 * inf) never
// rejects, and pruning only uses "the box starts beyond the closest hit so far".
//
// Node (64 B, four aligned 16-byte rows), one per inner node, holding its two children:
//   { c0.min.xyz, c0.max.x } { c0.max.yz, c1.min.xy } { c1.min.z, c1.max.xyz } { ref0, ref1, cnt0, cnt1 }
//   cnt == 0: ref = inner node index; cnt > 0: leaf of cnt triangle records starting at ref; cnt == ~0u: no child
//
// Compiles as CUDA device code and as plain C++ (tests/helpers/kd8_host.cpp runs this very source on the CPU).
#pragma once

#include "rt_tri.cuh"

#ifndef BVH_COUNT_NODE
#define BVH_COUNT_NODE() ((void)0)
#endif
#ifndef BVH_COUNT_LEAF
#define BVH_COUNT_LEAF() ((void)0)
#endif

namespace rtb {

constexpr int BVH_STACK = 48;                 // the builder caps the depth at 44
constexpr uint32_t BVH_NO_CHILD = 0xFFFFFFFFu;

struct alignas(16) BvhStackEntry { uint32_t ref, cnt; float t0; uint32_t pad; };
RT_HD void bvh_stack_put(BvhStackEntry* e, uint32_t ref, uint32_t cnt, float t0) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<float4*>(e) = make_float4(__uint_as_float(ref), __uint_as_float(cnt), t0, 0.0f);
#else
    e->ref = ref; e->cnt = cnt; e->t0 = t0; e->pad = 0;
#endif
}
RT_HD void bvh_stack_get(const BvhStackEntry* e, uint32_t& ref, uint32_t& cnt, float& t0) {
#if defined(__CUDA_ARCH__)
    const float4 q = *reinterpret_cast<const float4*>(e);
    ref = __float_as_uint(q.x); cnt = __float_as_uint(q.y); t0 = q.z;
#else
    ref = e->ref; cnt = e->cnt; t0 = e->t0;
#endif
}

// min / max that ignore a NaN operand (fminf / fmaxf semantics on the device)
RT_HD float bvh_min(float a, float b) {
#if defined(__CUDA_ARCH__)
    return fminf(a, b);
#else
    return (a != a) ? b : ((b != b) ? a : (b < a ? b : a));
#endif
}
RT_HD float bvh_max(float a, float b) {
#if defined(__CUDA_ARCH__)
    return fmaxf(a, b);
#else
    return (a != a) ? b : ((b != b) ? a : (a < b ? b : a));
#endif
}

// phase: WALK = standing at inner node `ref`; LEAF = standing at a leaf (cnt triangles from record `ref`); DONE
struct BvhState {
    float ox, oy, oz, dx, dy, dz;                 // 1/d is re-derived per node visit (three MUFU) instead of living in three registers
    float t0, t_far;
    uint32_t ref, cnt;
    int sp, phase;
    bool any_hit;
    KdHit best;
};

constexpr float BVH_SLACK = 1e-5f;

// RT_BVH_FMA_SLAB: slab test as one fused multiply-add per plane, t = l * (1/d) - o * (1/d), and no relative widening.
// The box tests only have to be conservative, and the builder pads every box by 2e-5 of the scene scale: the error of this
// form is <= |o| * 2^-23 (+ the 1-ulp reciprocal, a relative error of t that moves entry and exit together) in space units,
// far inside the padding as long as the origin is within BVH_FAR_ORIGIN scene extents of the scene box; rays from farther
// away are answered by the reference-order traversal (bvh_init marks them KD_RERUN).
#ifndef RT_BVH_FMA_SLAB
#define RT_BVH_FMA_SLAB 1
#endif
constexpr float BVH_FAR_ORIGIN = 8.0f;

// entry and exit parameter of the ray in a box; entry clamped to 0, NaNs (0 * inf) ignored
#if RT_BVH_FMA_SLAB
RT_HD void bvh_slab(float ix, float iy, float iz, float cx, float cy, float cz, float lx, float ly, float lz, float hx, float hy, float hz,
                    float& t_in, float& t_out) {
    const float ax = kd_fma(lx, ix, cx), bx = kd_fma(hx, ix, cx);
    const float ay = kd_fma(ly, iy, cy), by = kd_fma(hy, iy, cy);
    const float az = kd_fma(lz, iz, cz), bz = kd_fma(hz, iz, cz);
    t_in = bvh_max(bvh_max(bvh_min(ax, bx), bvh_min(ay, by)), bvh_max(bvh_min(az, bz), 0.0f));
    t_out = bvh_min(bvh_min(bvh_max(ax, bx), bvh_max(ay, by)), bvh_max(az, bz));
}
#else
RT_HD void bvh_slab(const BvhState& s, float ix, float iy, float iz, float lx, float ly, float lz, float hx, float hy, float hz,
                    float& t_in, float& t_out) {
    const float ax = (lx - s.ox) * ix, bx = (hx - s.ox) * ix;
    const float ay = (ly - s.oy) * iy, by = (hy - s.oy) * iy;
    const float az = (lz - s.oz) * iz, bz = (hz - s.oz) * iz;
    t_in = bvh_max(bvh_max(bvh_min(ax, bx), bvh_min(ay, by)), bvh_max(bvh_min(az, bz), 0.0f));
    t_out = bvh_min(bvh_min(bvh_max(ax, bx), bvh_max(ay, by)), bvh_max(az, bz));
}
#endif

// returns false when the ray misses the scene (the query is then finished: a miss).  ref_min / ref_max is the REFERENCE's
// root box (union of the mesh boxes, kd_tree_simd.hpp:101-104) and the test is the reference's own slab arithmetic with its
// strict t_max < t_min (aabb3.hpp:74-90): whether a ray that starts on the scene boundary and leaves it "misses the scene" is
// decided exactly as the reference decides it.  Below the root the BVH's own padded boxes take over.
RT_HD bool bvh_init(BvhState& s, const float* ref_min, const float* ref_max, float ox, float oy, float oz, float dx, float dy, float dz,
                    float t_far, bool any_hit) {
    s.ox = ox; s.oy = oy; s.oz = oz; s.dx = dx; s.dy = dy; s.dz = dz;
    const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;          // the reference's inv_direction (ray3.hpp:11-14)
    s.t_far = t_far; s.any_hit = any_hit;
    s.best.t = FLT_MAX; s.best.u = 0.0f; s.best.v = 0.0f; s.best.tri = -1; s.best.tie_t = -1.0f;
    s.ref = 0; s.cnt = 0; s.sp = 0; s.phase = KD8_DONE; s.t0 = 0.0f;
    float t0 = 0.0f, t1 = FLT_MAX;
    const float ax = (ref_min[0] - ox) * ix, bx = (ref_max[0] - ox) * ix;
    const float ay = (ref_min[1] - oy) * iy, by = (ref_max[1] - oy) * iy;
    const float az = (ref_min[2] - oz) * iz, bz = (ref_max[2] - oz) * iz;
    t0 = kd_max(t0, (bx < ax) ? bx : ax); t1 = kd_min(t1, (bx < ax) ? ax : bx);
    t0 = kd_max(t0, (by < ay) ? by : ay); t1 = kd_min(t1, (by < ay) ? ay : by);
    t0 = kd_max(t0, (bz < az) ? bz : az); t1 = kd_min(t1, (bz < az) ? az : bz);
    if (t1 < t0 || !(t0 * (1.0f - BVH_SLACK) <= t_far)) return false;
    // t1 == 0: the ray starts ON the scene boundary and leaves the scene at once.  Which wall triangles the reference still
    // tests then depends on its own leaf boxes, so such rays (rare) are answered by the reference-order traversal: the lane
    // finishes at once with the KD_RERUN mark
    if (!(0.0f < t1)) { s.best.tri = KD_RERUN; return true; }
#if RT_BVH_FMA_SLAB
    {   // l * inf - o * inf is NaN where (l - o) * inf is a signed infinity: a ray parallel to an axis (a direction component that
        // is zero or flushes to zero in rcp.approx.ftz) is answered by the reference-order traversal
        constexpr float TINY = 1e-30f;
        if (!(fabsf(dx) >= TINY) | !(fabsf(dy) >= TINY) | !(fabsf(dz) >= TINY)) { s.best.tri = KD_RERUN; return true; }
    }
    {   // an origin many scene extents away (or not finite): the padded boxes no longer cover the slab rounding
        const float ext = kd_max(kd_max(ref_max[0] - ref_min[0], ref_max[1] - ref_min[1]), ref_max[2] - ref_min[2]) * BVH_FAR_ORIGIN;
        const float far = kd_max(kd_max(kd_max(ref_min[0] - ox, ox - ref_max[0]), kd_max(ref_min[1] - oy, oy - ref_max[1])),
                                 kd_max(ref_min[2] - oz, oz - ref_max[2]));
        if (!(far <= ext)) { s.best.tri = KD_RERUN; return true; }
    }
#endif
    s.phase = KD8_WALK;
    return true;
}

// Next subtree from the stack.  An entry whose box starts beyond the closest hit so far is dropped HERE, so the entry
// parameter of the current node never has to live in the traversal state: a node reached by descending was tested against the
// same limit a moment ago.  At most RT_BVH_POP_CULL entries are looked at per call (in lock step a longer loop would stall the
// whole warp behind one lane that is emptying its stack); if all of them were dropped the lane stands at the pseudo node
// BVH_SKIP and goes on dropping at its next node step.  RT_BVH_POP_CULL = 0: the older form, s.t0 re-checked at every step.
#ifndef RT_BVH_POP_CULL
#define RT_BVH_POP_CULL 2
#endif
constexpr uint32_t BVH_SKIP = 0xFFFFFFFEu;
RT_HD void bvh_pop(BvhState& s, const BvhStackEntry* stack, float lim) {
#if RT_BVH_POP_CULL
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < RT_BVH_POP_CULL; ++k) {
        if (!s.sp) { s.phase = KD8_DONE; return; }
        --s.sp;
        float t0;
        bvh_stack_get(stack + s.sp, s.ref, s.cnt, t0);
        if (!(t0 > lim)) { s.phase = s.cnt ? KD8_LEAF : KD8_WALK; return; }
    }
    s.ref = BVH_SKIP; s.cnt = 0; s.phase = KD8_WALK;
#else
    (void)lim;
    if (!s.sp) { s.phase = KD8_DONE; return; }
    --s.sp;
    bvh_stack_get(stack + s.sp, s.ref, s.cnt, s.t0);
    s.phase = s.cnt ? KD8_LEAF : KD8_WALK;
#endif
}

// One inner-node visit (phase WALK): test both children's boxes, go to the nearer one, push the other.
RT_HD void bvh_node_step(BvhState& s, BvhStackEntry* stack, const float* __restrict__ nodes) {
    const float lim = kd_min(s.best.t, s.t_far);
#if RT_BVH_POP_CULL
    bool pop = s.ref == BVH_SKIP;                                            // still dropping stack entries (bvh_pop)
#else
    bool pop = s.t0 > lim;                                                   // the box starts beyond the closest hit so far
#endif
    if (!pop) {
        BVH_COUNT_NODE();
        const float* p = nodes + size_t(s.ref) * 16u;
        const KdRow q0 = kd_load_row(p), q1 = kd_load_row(p + 4), q2 = kd_load_row(p + 8), q3 = kd_load_row(p + 12);
        // the 1-ulp reciprocal is enough under the box slack; a flushed subnormal component behaves like zero (parallel ray)
        const float ix = kd_rcp_estimate(s.dx), iy = kd_rcp_estimate(s.dy), iz = kd_rcp_estimate(s.dz);
        float in0, out0, in1, out1;
#if RT_BVH_FMA_SLAB
        const float cx = -(s.ox * ix), cy = -(s.oy * iy), cz = -(s.oz * iz);
        bvh_slab(ix, iy, iz, cx, cy, cz, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, in0, out0);
        bvh_slab(ix, iy, iz, cx, cy, cz, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, in1, out1);
        const uint32_t ref0 = uint32_t(kd_as_int(q3.x)), ref1 = uint32_t(kd_as_int(q3.y));
        const uint32_t cnt0 = uint32_t(kd_as_int(q3.z)), cnt1 = uint32_t(kd_as_int(q3.w));
        const float e0 = in0, e1 = in1;
        const bool hit0 = (cnt0 != BVH_NO_CHILD) & (e0 <= out0) & (e0 <= lim);
        const bool hit1 = (cnt1 != BVH_NO_CHILD) & (e1 <= out1) & (e1 <= lim);
#else
        bvh_slab(s, ix, iy, iz, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, in0, out0);
        bvh_slab(s, ix, iy, iz, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, in1, out1);
        const uint32_t ref0 = uint32_t(kd_as_int(q3.x)), ref1 = uint32_t(kd_as_int(q3.y));
        const uint32_t cnt0 = uint32_t(kd_as_int(q3.z)), cnt1 = uint32_t(kd_as_int(q3.w));
        const float e0 = in0 * (1.0f - BVH_SLACK), e1 = in1 * (1.0f - BVH_SLACK);         // widened entry points (>= 0)
        const bool hit0 = (cnt0 != BVH_NO_CHILD) & (e0 <= out0 * (1.0f + BVH_SLACK)) & (e0 <= lim);
        const bool hit1 = (cnt1 != BVH_NO_CHILD) & (e1 <= out1 * (1.0f + BVH_SLACK)) & (e1 <= lim);
#endif
        const bool first1 = hit1 & (!hit0 | (e1 < e0));                     // child 1 is visited first
        if (hit0 & hit1) { bvh_stack_put(stack + s.sp, first1 ? ref0 : ref1, first1 ? cnt0 : cnt1, first1 ? e0 : e1); ++s.sp; }
        pop = !(hit0 | hit1);
        s.ref = first1 ? ref1 : ref0;
        s.cnt = first1 ? cnt1 : cnt0;
#if !RT_BVH_POP_CULL
        s.t0 = first1 ? e1 : e0;
#endif
        s.phase = s.cnt ? KD8_LEAF : KD8_WALK;
    }
    if (pop) bvh_pop(s, stack, lim);
}

// A leaf (phase LEAF): test its triangles, then pop the next subtree or finish.
template <bool CULL, bool FAST>
RT_HD void bvh_leaf_step(BvhState& s, const BvhStackEntry* stack, const float* __restrict__ tris, float eps) {
#if !RT_BVH_POP_CULL
    if (!(s.t0 > kd_min(s.best.t, s.t_far)))
#endif
    {
        BVH_COUNT_LEAF();
        kd_test_leaf<CULL, FAST>(tris + size_t(s.ref) * KD8_TRI_FLOATS, s.cnt, s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, eps, s.best);
        if (s.any_hit && s.best.t <= s.t_far) { s.phase = KD8_DONE; return; }
    }
    bvh_pop(s, stack, kd_min(s.best.t, s.t_far));
}

// Closest hit with t <= t_far (t_far = FLT_MAX for a plain query).  any_hit: return at the first hit inside [.., t_far].
template <bool CULL, bool FAST>
RT_HD KdHit bvh_trace(const float* __restrict__ nodes, const float* __restrict__ tris, const float* root_min, const float* root_max,
                      float ox, float oy, float oz, float dx, float dy, float dz, float eps, float t_far, bool any_hit) {
    BvhState s;
    BvhStackEntry stack[BVH_STACK];
    if (bvh_init(s, root_min, root_max, ox, oy, oz, dx, dy, dz, t_far, any_hit))
        while (s.phase != KD8_DONE) {
            if (s.phase == KD8_WALK) bvh_node_step(s, stack, nodes);
            else bvh_leaf_step<CULL, FAST>(s, stack, tris, eps);
        }
    return s.best;
}

}  // namespace rtb
