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
//   packets : 160 B, 4 triangles SoA (v0, e1, e2 rows + ids)
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

namespace rtb {

struct KdHit { float t, u, v; int tri; float tie_t; };   // tie_t == t: another triangle has exactly the winner's t

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
        if (!(CULL ? (eps <= det) : (eps <= fabsf(det)))) return;                                        // :33-38
        const float tx = ox - v0x, ty = oy - v0y, tz = oz - v0z;                                         // :42-44
        const float un = tx * pvx + ty * pvy + tz * pvz;
        const float r = kd_rcp_estimate(det);
        const float ua = un * r;
        if ((ua < -M) | (ua > 1.0f + M)) return;
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

template <bool CULL, bool FAST>
RT_HD void kd_test_packets(const float* pk, uint32_t count, float ox, float oy, float oz, float dx, float dy, float dz, float eps,
                           KdHit& best) {
    for (uint32_t p = 0; p < count; ++p, pk += 40) {
        const KdRow v0x = kd_load_row(pk), v0y = kd_load_row(pk + 4), v0z = kd_load_row(pk + 8);
        const KdRow e1x = kd_load_row(pk + 12), e1y = kd_load_row(pk + 16), e1z = kd_load_row(pk + 20);
        const KdRow e2x = kd_load_row(pk + 24), e2y = kd_load_row(pk + 28), e2z = kd_load_row(pk + 32);
        const KdRow id = kd_load_row(pk + 36);
        kd_test_tri<CULL, FAST>(v0x.x, v0y.x, v0z.x, e1x.x, e1y.x, e1z.x, e2x.x, e2y.x, e2z.x, kd_as_int(id.x), ox, oy, oz, dx, dy, dz, eps, best);
        kd_test_tri<CULL, FAST>(v0x.y, v0y.y, v0z.y, e1x.y, e1y.y, e1z.y, e2x.y, e2y.y, e2z.y, kd_as_int(id.y), ox, oy, oz, dx, dy, dz, eps, best);
        kd_test_tri<CULL, FAST>(v0x.z, v0y.z, v0z.z, e1x.z, e1y.z, e1z.z, e2x.z, e2y.z, e2z.z, kd_as_int(id.z), ox, oy, oz, dx, dy, dz, eps, best);
        kd_test_tri<CULL, FAST>(v0x.w, v0y.w, v0z.w, e1x.w, e1y.w, e1z.w, e2x.w, e2y.w, e2z.w, kd_as_int(id.w), ox, oy, oz, dx, dy, dz, eps, best);
    }
}

constexpr int KD8_STACK = 32;     // tree depth is capped at 30 by the host

struct KdStackEntry { uint32_t node; float t0, t1; };

// Closest hit with t <= t_far (t_far = FLT_MAX for a plain query).  any_hit: return at the first hit inside [.., t_far].
template <bool CULL, bool FAST>
RT_HD KdHit kd8_trace(const uint32_t* __restrict__ nodes8, const float* __restrict__ packets, const float* root_min, const float* root_max,
                      float ox, float oy, float oz, float dx, float dy, float dz, float eps, float t_far, bool any_hit) {
    KdHit best; best.t = FLT_MAX; best.u = 0.0f; best.v = 0.0f; best.tri = -1; best.tie_t = -1.0f;
    const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
    // parametric interval of the root box (aabb3.hpp:74-90 semantics: NaN from 0*inf leaves a bound unchanged)
    float t0 = 0.0f, t1 = FLT_MAX;
    {
        const float ax = (root_min[0] - ox) * ix, bx = (root_max[0] - ox) * ix;
        const float ay = (root_min[1] - oy) * iy, by = (root_max[1] - oy) * iy;
        const float az = (root_min[2] - oz) * iz, bz = (root_max[2] - oz) * iz;
        t0 = kd_max(t0, (bx < ax) ? bx : ax); t1 = kd_min(t1, (bx < ax) ? ax : bx);
        t0 = kd_max(t0, (by < ay) ? by : ay); t1 = kd_min(t1, (by < ay) ? ay : by);
        t0 = kd_max(t0, (bz < az) ? bz : az); t1 = kd_min(t1, (bz < az) ? az : bz);
    }
    const float S = 2e-6f;                       // relative widening of every interval comparison
    if (t1 + S * fabsf(t1) < t0) return best;
    t0 = kd_max(0.0f, t0 - S * fabsf(t0));
    t1 = t1 + S * fabsf(t1);

    KdStackEntry stack[KD8_STACK];
    int sp = 0;
    uint32_t node = 0;
    for (;;) {
        float limit = kd_min(best.t, t_far);
        bool pop = t0 > limit;
        if (!pop) {
            const uint32_t first = nodes8[2 * node], word = nodes8[2 * node + 1];
            const uint32_t axis = word & 3u;
            if (axis == 3u) {
                kd_test_packets<CULL, FAST>(packets + size_t(first) * 40u, word >> 2, ox, oy, oz, dx, dy, dz, eps, best);
                if (any_hit && best.t <= t_far) return best;
                pop = true;
            } else {
                const float split = kd_bits_to_float(first);
                const float oa = axis == 0u ? ox : (axis == 1u ? oy : oz);
                const float da = axis == 0u ? dx : (axis == 1u ? dy : dz);
                const float ia = axis == 0u ? ix : (axis == 1u ? iy : iz);
                const uint32_t c0 = (word & 4u) ? node + 1u : 0xFFFFFFFFu;       // lower half  [min, split]
                const uint32_t c1 = (word & 8u) ? (word >> 4) : 0xFFFFFFFFu;     // upper half  [split, max]
                const bool below = oa < split;
                const uint32_t near_c = below ? c0 : c1, far_c = below ? c1 : c0;
                const float ts = (split - oa) * ia;                               // exact sign; +-inf for da == 0; NaN if also oa == split
                bool go_near = true, go_far = true;
                float near_t1 = t1, far_t0 = t0;
                if (oa == split || ts != ts) {
                    // origin on the plane: both halves, intervals kept
                } else if (ts < 0.0f || da == 0.0f) {
                    go_far = false;                                               // moving away from / parallel to the plane
                } else {
                    const float w = S * kd_max(fabsf(ts), kd_max(fabsf(t0), fabsf(t1)));
                    if (ts > t1 + w) go_far = false;                              // leaves the node before the plane
                    else if (ts < t0 - w) go_near = false;                        // crossed the plane before entering the node
                    else { near_t1 = kd_min(t1, ts + w); far_t0 = kd_max(t0, ts - w); }
                }
                go_near = go_near && near_c != 0xFFFFFFFFu;
                go_far = go_far && far_c != 0xFFFFFFFFu;
                if (go_near) {
                    if (go_far) { stack[sp].node = far_c; stack[sp].t0 = far_t0; stack[sp].t1 = t1; ++sp; }
                    node = near_c; t1 = near_t1;
                } else if (go_far) {
                    node = far_c; t0 = far_t0;
                } else pop = true;
            }
        }
        if (pop) {
            if (!sp) break;
            --sp;
            node = stack[sp].node; t0 = stack[sp].t0; t1 = stack[sp].t1;
        }
    }
    return best;
}

}  // namespace rtb
