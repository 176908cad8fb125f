// csrc/rt_tri.cuh - the triangle test and the small helpers the accelerated traversals share (rt_bvh.cuh, rt_bvh4.cuh).
//
// The answer of a closest-hit query does not depend on the structure that is walked: every triangle is tested with the
// reference's own arithmetic (kd_test_tri below = the expressions of kd_tree_simd.hpp:25-60 in order, no FMA in exact mode), so
// t/u/v of the winner are the reference's bits, and the minimum over "all triangles whose box the ray touches" is the same set
// minimum.  What the reference's visit order decides is only which of two DIFFERENT triangles with exactly equal t is reported
// (a shared edge; the cube standing on the floor in config 1 - coplanar faces).  That decision depends on the reference's leaf
// order, so the accelerated traversals do not guess: they record that a second triangle tied with the winner
// (KdHit::tie_t == t) and the caller re-runs exactly those rays (~0.02 %) through the reference-order query (trace_any in
// rt_device.cuh).
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

#ifndef KD8_COUNT_TRI
#define KD8_COUNT_TRI() ((void)0)       // instrumentation hook for host-side experiments
#endif

namespace rtb {

struct KdHit { float t, u, v; int tri; float tie_t; };   // tie_t == t: another triangle has exactly the winner's t
// tri == KD_RERUN: the query has to be answered by the reference-order traversal (see bvh_init)
constexpr int KD_RERUN = -3;
// tri == KD_OVERFLOW: the same, because the query outgrew its traversal stack (rt_bvh4.cuh); counted, so that the host can react
constexpr int KD_OVERFLOW = -4;

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
// two adjacent 16-byte rows with ONE 256-bit load (sm_100: LDG.E.256; p must be 32-byte aligned).  A lane's rows of a node sit
// in one cache line, but every load instruction of a warp whose lanes stand at different nodes costs one L1 wavefront per lane:
// fetching a 128-byte node as four 256-bit loads instead of seven 128-bit ones nearly halves the node traffic through L1, which
// is what bounds the traversal of scenes beyond L2.
RT_HD void kd_load_row_pair(const float* p, KdRow& a, KdRow& b) {
#if defined(__CUDA_ARCH__)
    asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#else
    a = KdRow{p[0], p[1], p[2], p[3]}; b = KdRow{p[4], p[5], p[6], p[7]};
#endif
}
RT_HD int kd_as_int(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int i; std::memcpy(&i, &f, 4); return i;
#endif
}

// leaf triangles of the bounding-volume hierarchy: 48 B per triangle, three aligned 16-byte rows
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

// Phases of a lane's traversal state (BvhState in rt_bvh.cuh), so that a kernel can advance many rays in lock step and refill
// finished lanes (rt_stream.cuh): WALK = standing at an inner node; LEAF = parked at a leaf that still has to be tested;
// DONE = query finished, `best` is the answer.
enum : int { KD8_WALK = 0, KD8_LEAF = 1, KD8_DONE = 2 };

}  // namespace rtb
