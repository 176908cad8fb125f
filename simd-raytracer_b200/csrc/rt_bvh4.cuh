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

// One inner-node visit (phase WALK): test the four children's boxes, go to the nearest one, push the others so that the
// nearer one is popped first.
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
        const float lx[4] = {lox.x, lox.y, lox.z, lox.w}, ly[4] = {loy.x, loy.y, loy.z, loy.w}, lz[4] = {loz.x, loz.y, loz.z, loz.w};
        const float hx[4] = {hix.x, hix.y, hix.z, hix.w}, hy[4] = {hiy.x, hiy.y, hiy.z, hiy.w}, hz[4] = {hiz.x, hiz.y, hiz.z, hiz.w};
        const uint32_t ref[4] = {uint32_t(kd_as_int(rr.x)), uint32_t(kd_as_int(rr.y)), uint32_t(kd_as_int(rr.z)), uint32_t(kd_as_int(rr.w))};
        const uint32_t cnt[4] = {uint32_t(kd_as_int(cc.x)), uint32_t(kd_as_int(cc.y)), uint32_t(kd_as_int(cc.z)), uint32_t(kd_as_int(cc.w))};
        // the children the ray touches, ordered by entry distance (insertion into a list of at most four)
        float et[4];
        uint32_t er[4], ec[4];
        int n = 0;
        for (int k = 0; k < 4; ++k) {
            float t_in, t_out;
            bvh_slab(ix, iy, iz, cx, cy, cz, lx[k], ly[k], lz[k], hx[k], hy[k], hz[k], t_in, t_out);
            if (!((cnt[k] != BVH_NO_CHILD) & (t_in <= t_out) & (t_in <= lim))) continue;
            int j = n++;
            while (j > 0 && et[j - 1] > t_in) { et[j] = et[j - 1]; er[j] = er[j - 1]; ec[j] = ec[j - 1]; --j; }
            et[j] = t_in; er[j] = ref[k]; ec[j] = cnt[k];
        }
        for (int j = n - 1; j >= 1; --j) { bvh_stack_put(stack + s.sp, er[j], ec[j], et[j]); ++s.sp; }    // farthest first
        pop = n == 0;
        if (n) { s.ref = er[0]; s.cnt = ec[0]; s.phase = s.cnt ? KD8_LEAF : KD8_WALK; }
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
