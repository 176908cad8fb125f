// csrc/rt_bvh4.cuh - the accelerated closest-hit query over a FOUR-wide hierarchy ("wider-node
// accel (BVH4/8)", the reference author's own TODO, README.md:118-124).  The stream kernels (rt_stream.cuh) walk these nodes
// when the scene was built with accel_width 4: four-wide nodes halve the number of DEPENDENT node visits of a query, which is
// what bounds a launch whose longest queries are walked at one thread's latency.
//
// Same contract as rt_bvh.cuh, whose state (BvhState) and root test (bvh_init) it shares: every triangle test is the
// reference's own arithmetic (kd_test_tri = kd_tree_simd.hpp:25-60), the minimum is taken over every triangle whose box the ray
// touches, an exact-t tie between two different triangles is only recorded (KdHit::tie_t == t) and the caller re-runs those
// rays in reference order.  Box tests are conservative (the boxes are the padded boxes of the two-wide tree,
// host/bvh4_collapse.hpp).
//
// Node (128 B = one cache line, eight aligned 16-byte rows), its up to four children in structure-of-arrays form:
//   { min.x[4] } { min.y[4] } { min.z[4] } { max.x[4] } { max.y[4] } { max.z[4] } { child[4] } { unused }
//   child = ref << 3 | cnt.  cnt == 0: ref = inner node index; cnt in 1..7: leaf of cnt triangle records starting at ref (the
//   SAME 48-byte records as the two-wide tree); child == ~0u: no child
//
// Traversal stack: 8-byte entries { entry distance, child }.  The stack is a template parameter - put(pos, t0, child) /
// put_if / get(pos, t0, child) / overflows(entries) - because its home differs: shared memory with a local-memory tail in the stream kernels
// (rt_stream.cuh StreamStack4), a plain array in the one-ray-per-thread query below and on the CPU.
//
// Compiles as CUDA device code and as plain C++ (tests/helpers/kd8_host.cpp runs this very source on the CPU).
#pragma once

#include "rt_bvh.cuh"

#ifndef BVH4_TRACK_SP
#define BVH4_TRACK_SP(sp) ((void)0)     // instrumentation hook for host-side experiments: stack entries after a visit
#endif

#if !RT_BVH_FMA_SLAB || !RT_BVH_POP_CULL
#error "rt_bvh4.cuh is written against the fused slab test (RT_BVH_FMA_SLAB=1) and the pop-time cull (RT_BVH_POP_CULL>=1) of rt_bvh.cuh"
#endif

namespace rtb {

constexpr uint32_t BVH4_NODE_FLOATS = 32;
constexpr uint32_t BVH4_NONE = 0xFFFFFFFFu;
constexpr uint32_t BVH4_MAX_LEAF = 7;           // cnt lives in three bits
constexpr uint32_t BVH4_MAX_REF = 0x1FFFFFFDu;  // ref lives in 29 bits; BVH_SKIP >> 3 and BVH4_NONE >> 3 stay free
// A visit pushes at most three entries and descends at least one level of the two-wide tree, whose depth the builder caps at
// 44 (host/bvh_build.cpp), so 3 * 44 + 1 entries always suffice; the collapse computes the real need of a tree
// (bvh4_stack_need, a few dozen) and scene creation refuses a tree that would need more.
constexpr int BVH4_STACK = 3 * 44 + 4;

#ifndef RT_BVH4_LOAD256
#define RT_BVH4_LOAD256 1                // fetch a node with four 256-bit loads (rt_tri.cuh kd_load_row_pair) instead of seven 128-bit ones
#endif

RT_HD uint32_t bvh4_child(uint32_t ref, uint32_t cnt) { return (ref << 3) | cnt; }

struct Bvh4ArrayStack {
    struct Entry { float t0; uint32_t child; } e[BVH4_STACK];
    RT_HD void put(int pos, float t0, uint32_t child) { e[pos].t0 = t0; e[pos].child = child; }
    RT_HD void put_if(bool on, int pos, float t0, uint32_t child) { if (on) put(pos, t0, child); }
    RT_HD bool overflows(int) const { return false; }          // BVH4_STACK entries cover every tree the builder can produce
    RT_HD void get(int pos, float& t0, uint32_t& child) const { t0 = e[pos].t0; child = e[pos].child; }
};

// Next subtree from the stack (the rule of bvh_pop: an entry that starts beyond the closest hit so far is dropped here, at most
// RT_BVH_POP_CULL of them per call; BVH_SKIP = still dropping).
template <class Stack>
RT_HD void bvh4_pop(BvhState& s, const Stack& stack, float lim) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < RT_BVH_POP_CULL; ++k) {
        if (!s.sp) { s.phase = KD8_DONE; return; }
        --s.sp;
        float t0; uint32_t child;
        stack.get(s.sp, t0, child);
        if (!(t0 > lim)) { s.ref = child >> 3; s.cnt = child & 7u; s.phase = s.cnt ? KD8_LEAF : KD8_WALK; return; }
    }
    s.ref = BVH_SKIP; s.cnt = 0; s.phase = KD8_WALK;
}

// One inner-node visit (phase WALK): test the four children's boxes, go to the nearest one, put the others on the stack so
// that the nearer one is popped first.  No sorting network and no data movement: every touched child learns its RANK among the
// four entry distances (six compares; an untouched child carries +inf, ties go to the lower slot) and is stored straight at
// its stack position sp + (touched - 1 - rank); rank 0 becomes the current node.
template <class Stack>
RT_HD void bvh4_node_step(BvhState& s, Stack& stack, const float* __restrict__ nodes) {
    const float lim = kd_min(s.best.t, s.t_far);
    bool pop = s.ref == BVH_SKIP;                                            // still dropping stack entries (bvh4_pop)
    if (!pop) {
        BVH_COUNT_NODE();
        const float* p = nodes + size_t(s.ref) * BVH4_NODE_FLOATS;
#if RT_BVH4_LOAD256
        KdRow lox, loy, loz, hix, hiy, hiz, cc, unused;
        kd_load_row_pair(p, lox, loy); kd_load_row_pair(p + 8, loz, hix); kd_load_row_pair(p + 16, hiy, hiz); kd_load_row_pair(p + 24, cc, unused);
        (void)unused;
#else
        const KdRow lox = kd_load_row(p), loy = kd_load_row(p + 4), loz = kd_load_row(p + 8);
        const KdRow hix = kd_load_row(p + 12), hiy = kd_load_row(p + 16), hiz = kd_load_row(p + 20);
        const KdRow cc = kd_load_row(p + 24);
#endif
        const float ix = kd_rcp_estimate(s.dx), iy = kd_rcp_estimate(s.dy), iz = kd_rcp_estimate(s.dz);
        const float cx = -(s.ox * ix), cy = -(s.oy * iy), cz = -(s.oz * iz);
        const uint32_t c0 = uint32_t(kd_as_int(cc.x)), c1 = uint32_t(kd_as_int(cc.y)), c2 = uint32_t(kd_as_int(cc.z)), c3 = uint32_t(kd_as_int(cc.w));
        const float MISS = kd_bits_to_float(0x7F800000u);                    // +inf: behind every touched child (entry <= lim <= FLT_MAX)
        float in, out;
        bvh_slab(ix, iy, iz, cx, cy, cz, lox.x, loy.x, loz.x, hix.x, hiy.x, hiz.x, in, out);
        const bool h0 = (c0 != BVH4_NONE) & (in <= out) & (in <= lim); const float t0 = h0 ? in : MISS;
        bvh_slab(ix, iy, iz, cx, cy, cz, lox.y, loy.y, loz.y, hix.y, hiy.y, hiz.y, in, out);
        const bool h1 = (c1 != BVH4_NONE) & (in <= out) & (in <= lim); const float t1 = h1 ? in : MISS;
        bvh_slab(ix, iy, iz, cx, cy, cz, lox.z, loy.z, loz.z, hix.z, hiy.z, hiz.z, in, out);
        const bool h2 = (c2 != BVH4_NONE) & (in <= out) & (in <= lim); const float t2 = h2 ? in : MISS;
        bvh_slab(ix, iy, iz, cx, cy, cz, lox.w, loy.w, loz.w, hix.w, hiy.w, hiz.w, in, out);
        const bool h3 = (c3 != BVH4_NONE) & (in <= out) & (in <= lim); const float t3 = h3 ? in : MISS;
        const int touched = int(h0) + int(h1) + int(h2) + int(h3);
        pop = touched == 0;
        if (!pop) {
            // a < b for slots a < b: "b is strictly nearer"; equal keys keep slot order, so the four ranks are a permutation
            const int n10 = t1 < t0, n20 = t2 < t0, n30 = t3 < t0, n21 = t2 < t1, n31 = t3 < t1, n32 = t3 < t2;
            const int r0 = n10 + n20 + n30;
            const int r1 = (1 - n10) + n21 + n31;
            const int r2 = (2 - n20 - n21) + n32;
            const int r3 = 3 - n30 - n31 - n32;
            const int top = s.sp + touched - 1;                             // rank r >= 1 goes to top - r: rank 1 is popped first
            // a stack with fewer entries than the tree's worst case (the shared-memory stack of the stream kernels, sized for
            // what rays really need): a query that would outgrow it is not continued with a subtree missing - it ends here with
            // the KD_OVERFLOW mark and is answered by the reference-order traversal, like a query with an exact-t tie
            if (stack.overflows(top)) { s.best.tri = KD_OVERFLOW; s.phase = KD8_DONE; return; }
            // exactly one touched child has rank 0 (an untouched one ranks behind every touched one): it becomes the current node;
            // the others store themselves, no branch between them
            stack.put_if(h0 & (r0 != 0), top - r0, t0, c0);
            stack.put_if(h1 & (r1 != 0), top - r1, t1, c1);
            stack.put_if(h2 & (r2 != 0), top - r2, t2, c2);
            stack.put_if(h3 & (r3 != 0), top - r3, t3, c3);
            const uint32_t cur = (r0 == 0) ? c0 : (r1 == 0) ? c1 : (r2 == 0) ? c2 : c3;
            s.sp = top;
            BVH4_TRACK_SP(top);
            s.ref = cur >> 3; s.cnt = cur & 7u;
            s.phase = s.cnt ? KD8_LEAF : KD8_WALK;
        }
    }
    if (pop) bvh4_pop(s, stack, lim);
}

// A leaf (phase LEAF): test its triangles, then pop the next subtree or finish.
template <bool CULL, bool FAST, class Stack>
RT_HD void bvh4_leaf_step(BvhState& s, const Stack& stack, const float* __restrict__ tris, float eps) {
    BVH_COUNT_LEAF();
    kd_test_leaf<CULL, FAST>(tris + size_t(s.ref) * KD8_TRI_FLOATS, s.cnt, s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, eps, s.best);
    if (s.any_hit && s.best.t <= s.t_far) { s.phase = KD8_DONE; return; }
    bvh4_pop(s, stack, kd_min(s.best.t, s.t_far));
}

// Closest hit with t <= t_far (t_far = FLT_MAX for a plain query).  any_hit: return at the first hit inside [.., t_far].
template <bool CULL, bool FAST>
RT_HD KdHit bvh4_trace(const float* __restrict__ nodes4, const float* __restrict__ tris, const float* root_min, const float* root_max,
                       float ox, float oy, float oz, float dx, float dy, float dz, float eps, float t_far, bool any_hit) {
    BvhState s;
    Bvh4ArrayStack stack;
    if (bvh_init(s, root_min, root_max, ox, oy, oz, dx, dy, dz, t_far, any_hit))
        while (s.phase != KD8_DONE) {
            if (s.phase == KD8_WALK) bvh4_node_step(s, stack, nodes4);
            else bvh4_leaf_step<CULL, FAST>(s, stack, tris, eps);
        }
    return s.best;
}

}  // namespace rtb
