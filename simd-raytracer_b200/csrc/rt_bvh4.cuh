// csrc/rt_bvh4.cuh - the accelerated closest-hit query over a FOUR-wide hierarchy ("wider-node
// accel (BVH4/8)", the reference author's own TODO, README.md:118-124).  ROUND-1 STATE: the traversal and the collapse of the
// two-wide hierarchy (host/bvh4_collapse.hpp) are validated on the CPU against the oracle (tests/test_kd8_host.py runs this
// very source, structure "bvh4"); the CUDA kernels still walk the two-wide nodes of rt_bvh.cuh.  Why it is the next step:
// on a one-sample 1080p frame the trace kernels end 2-3x later than their median warp because single queries with 80-110 node
// visits are walked at one thread's latency; four-wide nodes halve the number of DEPENDENT node visits of
// a query (scripts/bvh_ray_lengths_cpu.py prints both).
//
// Same contract as rt_bvh.cuh, whose state, stack, root test (bvh_init), pop rule (bvh_pop) and leaf step (bvh_leaf_step) it
// shares: every triangle test is the reference's own arithmetic (kd_test_tri), the minimum is taken over every triangle whose
// box the ray touches, an exact-t tie between two different triangles is only recorded (KdHit::tie_t == t) and the caller
// re-runs those rays in reference order.  Box tests are conservative (the boxes are the padded boxes of the two-wide tree).
//
// Node (128 B = one cache line, eight aligned 16-byte rows), holding its up to four children in structure-of-arrays form:
//   { min.x[4] } { min.y[4] } { min.z[4] } { max.x[4] } { max.y[4] } { max.z[4] } { ref[4] } { cnt[4] }
//   cnt == 0: ref = inner node index; cnt > 0: leaf of cnt triangle records starting at ref (the SAME 48-byte records as the
//   two-wide tree); cnt == ~0u: no child
#pragma once

#include "rt_bvh.cuh"

namespace rtb {

constexpr int BVH4_STACK = 64;
constexpr uint32_t BVH4_NODE_FLOATS = 32;

// compare-and-swap of two (entry distance, ref, cnt) triples: predicate selects only, nothing indexed at run time
RT_HD void bvh4_order(float& ta, uint32_t& ra, uint32_t& ca, float& tb, uint32_t& rb, uint32_t& cb) {
    const bool sw = tb < ta;
    const float t = sw ? tb : ta; tb = sw ? ta : tb; ta = t;
    const uint32_t r = sw ? rb : ra; rb = sw ? ra : rb; ra = r;
    const uint32_t c = sw ? cb : ca; cb = sw ? ca : cb; ca = c;
}

// One inner-node visit (phase WALK): test the four children's boxes, go to the nearest one, push the others so that the
// nearer one is popped first.  Children the ray does not touch get the key +inf; a five-comparator network sorts the four
// keys; everything lives in scalars (registers on the device).
RT_HD void bvh4_node_step(BvhState& s, BvhStackEntry* stack, const float* __restrict__ nodes) {
    const float lim = kd_min(s.best.t, s.t_far);
    bool pop = s.ref == BVH_SKIP;                                            // still dropping stack entries (bvh_pop)
    if (!pop) {
        BVH_COUNT_NODE();
        const float* p = nodes + size_t(s.ref) * BVH4_NODE_FLOATS;
        const KdRow lox = kd_load_row(p), loy = kd_load_row(p + 4), loz = kd_load_row(p + 8);
        const KdRow hix = kd_load_row(p + 12), hiy = kd_load_row(p + 16), hiz = kd_load_row(p + 20);
        const KdRow rr = kd_load_row(p + 24), cc = kd_load_row(p + 28);
        const float ix = kd_rcp_estimate(s.dx), iy = kd_rcp_estimate(s.dy), iz = kd_rcp_estimate(s.dz);
        const float cx = -(s.ox * ix), cy = -(s.oy * iy), cz = -(s.oz * iz);
        uint32_t r0 = uint32_t(kd_as_int(rr.x)), r1 = uint32_t(kd_as_int(rr.y)), r2 = uint32_t(kd_as_int(rr.z)), r3 = uint32_t(kd_as_int(rr.w));
        uint32_t c0 = uint32_t(kd_as_int(cc.x)), c1 = uint32_t(kd_as_int(cc.y)), c2 = uint32_t(kd_as_int(cc.z)), c3 = uint32_t(kd_as_int(cc.w));
        const float MISS = FLT_MAX;             // entry distances are <= lim <= FLT_MAX; a touched child at exactly FLT_MAX cannot be closer than a hit
        float in, out, t0, t1, t2, t3;
        bvh_slab(ix, iy, iz, cx, cy, cz, lox.x, loy.x, loz.x, hix.x, hiy.x, hiz.x, in, out);
        const bool h0 = (c0 != BVH_NO_CHILD) & (in <= out) & (in <= lim); t0 = h0 ? in : MISS;
        bvh_slab(ix, iy, iz, cx, cy, cz, lox.y, loy.y, loz.y, hix.y, hiy.y, hiz.y, in, out);
        const bool h1 = (c1 != BVH_NO_CHILD) & (in <= out) & (in <= lim); t1 = h1 ? in : MISS;
        bvh_slab(ix, iy, iz, cx, cy, cz, lox.z, loy.z, loz.z, hix.z, hiy.z, hiz.z, in, out);
        const bool h2 = (c2 != BVH_NO_CHILD) & (in <= out) & (in <= lim); t2 = h2 ? in : MISS;
        bvh_slab(ix, iy, iz, cx, cy, cz, lox.w, loy.w, loz.w, hix.w, hiy.w, hiz.w, in, out);
        const bool h3 = (c3 != BVH_NO_CHILD) & (in <= out) & (in <= lim); t3 = h3 ? in : MISS;
        // untouched children: the key alone would tie with a touched child at FLT_MAX, so they also lose their identity
        c0 = h0 ? c0 : BVH_NO_CHILD; c1 = h1 ? c1 : BVH_NO_CHILD; c2 = h2 ? c2 : BVH_NO_CHILD; c3 = h3 ? c3 : BVH_NO_CHILD;
        bvh4_order(t0, r0, c0, t1, r1, c1); bvh4_order(t2, r2, c2, t3, r3, c3);
        bvh4_order(t0, r0, c0, t2, r2, c2); bvh4_order(t1, r1, c1, t3, r3, c3);
        bvh4_order(t1, r1, c1, t2, r2, c2);
        // an untouched child (cnt == NO_CHILD) may sort in front of a touched one only when both keys are FLT_MAX: skip by identity
        if (c3 != BVH_NO_CHILD) { bvh_stack_put(stack + s.sp, r3, c3, t3); ++s.sp; }                     // farthest first
        if (c2 != BVH_NO_CHILD) { bvh_stack_put(stack + s.sp, r2, c2, t2); ++s.sp; }
        if (c0 != BVH_NO_CHILD) {
            if (c1 != BVH_NO_CHILD) { bvh_stack_put(stack + s.sp, r1, c1, t1); ++s.sp; }
            s.ref = r0; s.cnt = c0;
        } else { s.ref = r1; s.cnt = c1; }       // only reachable when every touched key is FLT_MAX
        pop = (c0 == BVH_NO_CHILD) & (c1 == BVH_NO_CHILD);
        if (!pop) s.phase = s.cnt ? KD8_LEAF : KD8_WALK;
    }
    if (pop) bvh_pop(s, stack, lim);
}

// Closest hit with t <= t_far (t_far = FLT_MAX for a plain query).  any_hit: return at the first hit inside [.., t_far].
template <bool CULL, bool FAST>
RT_HD KdHit bvh4_trace(const float* __restrict__ nodes4, const float* __restrict__ tris, const float* root_min, const float* root_max,
                       float ox, float oy, float oz, float dx, float dy, float dz, float eps, float t_far, bool any_hit) {
    BvhState s;
    BvhStackEntry stack[BVH4_STACK];
    if (bvh_init(s, root_min, root_max, ox, oy, oz, dx, dy, dz, t_far, any_hit))
        while (s.phase != KD8_DONE) {
            if (s.phase == KD8_WALK) bvh4_node_step(s, stack, nodes4);
            else bvh_leaf_step<CULL, FAST>(s, stack, tris, eps);
        }
    return s.best;
}

}  // namespace rtb
