// csrc/rt_device.cuh - device-side data model and the traversal / intersection core (sm_100a).
//
// This translation unit is compiled with -fmad=false: nvcc never contracts a*b+c into FMA here, so every
// expression below rounds exactly where it is written, like the canonical (-ffp-contract=off) reference build that
// reproduces the published golden image.  Division, reciprocal and square root are
// the IEEE round-to-nearest forms (-prec-div / -prec-sqrt defaults), subnormals are kept (-ftz=false default).
// The FAST template variants use explicit fmaf() and are never bit-exact by contract.
//
// Reference lines cited as file:line under /root/reference/include/raytracer/.
#pragma once

#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace rtb {

// ---- resident scene (all pointers into HBM) -------------------------------------------------------------------
struct DMaterial { uint32_t kind; float albedo[3]; float ior; uint32_t smooth; int32_t texture; };       // 28 B
struct DTexture { uint32_t kind; float c0[3], c1[3], scalar; uint32_t w, h, off; };                      // 44 B
struct DLight { float pos[3], intensity; };                                                              // 16 B

struct DScene {
    const uint2* __restrict__ nodes8;       // 8-byte kd nodes (ordered traversal)
    const float4* __restrict__ nodes32;     // 2 x float4 per node: box + the same two words (reference-order traversal)
    const float4* __restrict__ packets;     // 10 x float4 per 4-triangle SoA packet
    const uint4* __restrict__ tri_index;    // vi0, vi1, vi2, material
    const float4* __restrict__ tri_normal;  // face normal
    const float4* __restrict__ tri_uv;      // 2 x float4 per triangle: uv0.xy uv1.xy | uv2.xy - -
    const float4* __restrict__ vnormals;
    const DMaterial* __restrict__ materials;
    const DTexture* __restrict__ textures;
    const DLight* __restrict__ lights;
    const uint8_t* __restrict__ texels;
    uint32_t n_lights, n_nodes;
    uint32_t width, height;
    float bg[3];
    float cam_pos[3];
    float cam_m[9];
    float root_min[3], root_max[3];
    int has_transmissive;
};

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }      // vec3.hpp:76-78
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }      // vec3.hpp:80-82
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }                           // vec3.hpp:43-45
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }         // vec3.hpp:94-97
__device__ __forceinline__ float dot(V3 a, V3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }    // vec3.hpp:119-122
__device__ __forceinline__ V3 cross(V3 a, V3 b) {                                                        // vec3.hpp:124-131
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float len2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }               // vec3.hpp:84-86
__device__ __forceinline__ float len(V3 a) { return __fsqrt_rn(len2(a)); }                               // vec3.hpp:88-90
__device__ __forceinline__ V3 normalized(V3 a) {                                                         // vec3.hpp:104-108
    const float inv = __fdiv_rn(1.0f, len(a));
    return mk(a.x * inv, a.y * inv, a.z * inv);
}
__device__ __forceinline__ float max_std(float a, float b) { return (a < b) ? b : a; }                   // std::max(a, b)

struct Hit { float t, u, v; int tri; };

// ---- ray vs. 4-triangle SoA packet --------------------------------------------------------------------------------
// triangle_packet<F,W>::intersect (kd_tree_simd.hpp:25-60), one lane, in the reference's operation order, folded
// into intersect_leaf's running closest (kd_tree_simd.hpp:266-302).  Using the traversal-wide best_t directly with
// a strict `<` is equivalent to the reference's leaf-local candidate followed by `c->t < best_t` (:224): in both,
// the first occurrence of the smallest t strictly below the incoming best wins.
template <bool CULL, bool FAST>
__device__ __forceinline__ void test_lane(float v0x, float v0y, float v0z, float e1x, float e1y, float e1z,
                                          float e2x, float e2y, float e2z, int id, V3 o, V3 d, float eps, Hit& best) {
    float pvx, pvy, pvz, det;
    if (FAST) {
        pvx = fmaf(d.y, e2z, -(d.z * e2y)); pvy = fmaf(d.z, e2x, -(d.x * e2z)); pvz = fmaf(d.x, e2y, -(d.y * e2x));
        det = fmaf(e1z, pvz, fmaf(e1y, pvy, e1x * pvx));
    } else {
        pvx = d.y * e2z - d.z * e2y;                                                                     // :27
        pvy = d.z * e2x - d.x * e2z;                                                                     // :28
        pvz = d.x * e2y - d.y * e2x;                                                                     // :29
        det = e1x * pvx + e1y * pvy + e1z * pvz;                                                         // :31
    }
    const bool ok_det = CULL ? (eps <= det) : (eps <= fabsf(det));                                       // :33-38
    if (!ok_det) return;
    const float inv_det = FAST ? __frcp_rn(det) : __fdiv_rn(1.0f, det);                                  // :40
    const float tx = o.x - v0x, ty = o.y - v0y, tz = o.z - v0z;                                          // :42-44
    float u, v, t, qx, qy, qz;
    if (FAST) {
        u = fmaf(tz, pvz, fmaf(ty, pvy, tx * pvx)) * inv_det;
        qx = fmaf(ty, e1z, -(tz * e1y)); qy = fmaf(tz, e1x, -(tx * e1z)); qz = fmaf(tx, e1y, -(ty * e1x));
        v = fmaf(d.z, qz, fmaf(d.y, qy, d.x * qx)) * inv_det;
        t = fmaf(e2z, qz, fmaf(e2y, qy, e2x * qx)) * inv_det;
    } else {
        u = (tx * pvx + ty * pvy + tz * pvz) * inv_det;                                                  // :46
        qx = ty * e1z - tz * e1y;                                                                        // :49
        qy = tz * e1x - tx * e1z;                                                                        // :50
        qz = tx * e1y - ty * e1x;                                                                        // :51
        v = (d.x * qx + d.y * qy + d.z * qz) * inv_det;                                                  // :53
        t = (e2x * qx + e2y * qy + e2z * qz) * inv_det;                                                  // :56
    }
    const bool ok = (0.0f <= u) & (u <= 1.0f) & (0.0f <= v) & (u + v <= 1.0f) & (eps < t);               // :47,:54,:57
    if (ok && t < best.t) { best.t = t; best.u = u; best.v = v; best.tri = id; }                         // :284-298, :224
}

template <bool CULL, bool FAST>
__device__ __forceinline__ void test_packets(const float4* __restrict__ pk, uint32_t count, V3 o, V3 d, float eps,
                                             Hit& best, float t_stop = -1.0f) {
    for (uint32_t p = 0; p < count; ++p, pk += 10) {
        if (best.t <= t_stop) return;     // any-hit early out (see occluded_query); never taken for t_stop < 0
        const float4 v0x = __ldg(pk + 0), v0y = __ldg(pk + 1), v0z = __ldg(pk + 2);
        const float4 e1x = __ldg(pk + 3), e1y = __ldg(pk + 4), e1z = __ldg(pk + 5);
        const float4 e2x = __ldg(pk + 6), e2y = __ldg(pk + 7), e2z = __ldg(pk + 8);
        const float4 idf = __ldg(pk + 9);
        test_lane<CULL, FAST>(v0x.x, v0y.x, v0z.x, e1x.x, e1y.x, e1z.x, e2x.x, e2y.x, e2z.x, __float_as_int(idf.x), o, d, eps, best);
        test_lane<CULL, FAST>(v0x.y, v0y.y, v0z.y, e1x.y, e1y.y, e1z.y, e2x.y, e2y.y, e2z.y, __float_as_int(idf.y), o, d, eps, best);
        test_lane<CULL, FAST>(v0x.z, v0y.z, v0z.z, e1x.z, e1y.z, e1z.z, e2x.z, e2y.z, e2z.z, __float_as_int(idf.z), o, d, eps, best);
        test_lane<CULL, FAST>(v0x.w, v0y.w, v0z.w, e1x.w, e1y.w, e1z.w, e2x.w, e2y.w, e2z.w, __float_as_int(idf.w), o, d, eps, best);
    }
}

// ---- slab test ----------------------------------------------------------------------------------------------------
// aabb3<F>::intersect(ray), core/math/aabb3.hpp:74-90.  std::minmax(a,b) = (b<a) ? (b,a) : (a,b); the running
// max / min ignore a NaN operand (std::max(t_min,t1) = (t_min<t1) ? t1 : t_min), which fmaxf / fminf also do.  The
// reference's per-axis early return equals one test at the end because t_min only grows and t_max only shrinks.
__device__ __forceinline__ bool slab(float4 lo, float4 hi, V3 o, V3 inv, float& t_min_out) {
    float t_min = 0.0f, t_max = FLT_MAX;
    {
        const float a = (lo.x - o.x) * inv.x, b = (hi.x - o.x) * inv.x;
        const bool sw = b < a;
        t_min = fmaxf(t_min, sw ? b : a); t_max = fminf(t_max, sw ? a : b);
    }
    {
        const float a = (lo.y - o.y) * inv.y, b = (hi.y - o.y) * inv.y;
        const bool sw = b < a;
        t_min = fmaxf(t_min, sw ? b : a); t_max = fminf(t_max, sw ? a : b);
    }
    {
        const float a = (lo.z - o.z) * inv.z, b = (hi.z - o.z) * inv.z;
        const bool sw = b < a;
        t_min = fmaxf(t_min, sw ? b : a); t_max = fminf(t_max, sw ? a : b);
    }
    t_min_out = t_min;
    return !(t_max < t_min);
}

constexpr int KD_STACK = 32;   // kd_max_depth is capped at 30 by the host

// ---- closest hit, reference visit order ---------------------------------------------------------------------------
// kd_tree_simd_accel::intersect<bf> (kd_tree_simd.hpp:187-229): LIFO stack, child0 pushed before child1 (child1 is
// visited first), every popped node slab-tested against its own box, pruned when best_t < box.t_min (strict).
// Returns the closest hit with t < t_limit (FLT_MAX for a plain query).
template <bool CULL, bool FAST>
__device__ __forceinline__ Hit trace_reference_order(const DScene& sc, V3 o, V3 d, float eps, float t_stop = -1.0f) {
    const V3 inv = mk(__fdiv_rn(1.0f, d.x), __fdiv_rn(1.0f, d.y), __fdiv_rn(1.0f, d.z));                 // ray3.hpp:11-14
    Hit best; best.t = FLT_MAX; best.u = 0.0f; best.v = 0.0f; best.tri = -1;
    uint32_t stack[KD_STACK];
    int sp = 0;
    stack[sp++] = 0u;
    while (sp) {
        const uint32_t idx = stack[--sp];
        const float4 lo = __ldg(sc.nodes32 + 2 * idx), hi = __ldg(sc.nodes32 + 2 * idx + 1);
        float t_min;
        if (!slab(lo, hi, o, inv, t_min) || best.t < t_min) continue;                                    // :202-205
        const uint32_t word = __float_as_uint(hi.w);
        if ((word & 3u) != 3u) {                                                                         // inner, :207-214
            if (word & 4u) stack[sp++] = idx + 1u;
            if (word & 8u) stack[sp++] = word >> 4;
        } else {                                                                                         // leaf, :216-226
            test_packets<CULL, FAST>(sc.packets + 10ull * __float_as_uint(lo.w), word >> 2, o, d, eps, best, t_stop);
            if (best.t <= t_stop) return best;
        }
    }
    return best;
}

// ---- closest hit, front-to-back over the 8-byte nodes ----------------------------------------------------------------
// Same tree, same leaf test, different visit order: the child on the ray origin's side of the split plane first,
// the subtree skipped when the ray's parametric interval inside the parent box does not reach it or starts beyond
// the closest hit found so far.  The closest t is order independent; exact-t ties between DIFFERENT triangles are
// resolved towards the lower triangle id (the rule that matched the reference on every tie observed, SURVEY.md
// section 7).  Not the parity-gated mode; see DESIGN.md.
template <bool CULL, bool FAST>
__device__ __forceinline__ void test_packets_ordered(const float4* __restrict__ pk, uint32_t count, V3 o, V3 d,
                                                     float eps, Hit& best) {
    Hit local; local.t = FLT_MAX; local.u = 0.0f; local.v = 0.0f; local.tri = -1;
    test_packets<CULL, FAST>(pk, count, o, d, eps, local);
    if (local.tri >= 0 && (local.t < best.t || (local.t == best.t && local.tri < best.tri))) best = local;
}

template <bool CULL, bool FAST>
__device__ __forceinline__ Hit trace_ordered(const DScene& sc, V3 o, V3 d, float eps, float t_stop = -1.0f) {
    const V3 inv = mk(__fdiv_rn(1.0f, d.x), __fdiv_rn(1.0f, d.y), __fdiv_rn(1.0f, d.z));
    Hit best; best.t = FLT_MAX; best.u = 0.0f; best.v = 0.0f; best.tri = -1;
    float t0;
    {
        const float4 lo = make_float4(sc.root_min[0], sc.root_min[1], sc.root_min[2], 0.0f);
        const float4 hi = make_float4(sc.root_max[0], sc.root_max[1], sc.root_max[2], 0.0f);
        if (!slab(lo, hi, o, inv, t0)) return best;
    }
    // exit distance of the root box (slab() only reports the entry)
    float t1 = FLT_MAX;
    {
        const float ax = (sc.root_min[0] - o.x) * inv.x, bx = (sc.root_max[0] - o.x) * inv.x;
        const float ay = (sc.root_min[1] - o.y) * inv.y, by = (sc.root_max[1] - o.y) * inv.y;
        const float az = (sc.root_min[2] - o.z) * inv.z, bz = (sc.root_max[2] - o.z) * inv.z;
        t1 = fminf(t1, fmaxf(ax, bx)); t1 = fminf(t1, fmaxf(ay, by)); t1 = fminf(t1, fmaxf(az, bz));
    }
    struct Entry { uint32_t node; float t0, t1; };
    Entry stack[KD_STACK];
    int sp = 0;
    uint32_t idx = 0;
    const float oa[3] = {o.x, o.y, o.z}, ia[3] = {inv.x, inv.y, inv.z}, da[3] = {d.x, d.y, d.z};
    for (;;) {
        bool pop = false;
        if (best.t < t0) pop = true;
        else {
            const uint2 n = __ldg(sc.nodes8 + idx);
            const uint32_t axis = n.y & 3u;
            if (axis == 3u) {
                test_packets_ordered<CULL, FAST>(sc.packets + 10ull * n.x, n.y >> 2, o, d, eps, best);
                if (best.t <= t_stop) return best;
                pop = true;
            } else {
                const float split = __uint_as_float(n.x);
                const float oc = axis == 0 ? oa[0] : (axis == 1 ? oa[1] : oa[2]);
                const float ic = axis == 0 ? ia[0] : (axis == 1 ? ia[1] : ia[2]);
                const float dc = axis == 0 ? da[0] : (axis == 1 ? da[1] : da[2]);
                const float ts = (split - oc) * ic;
                // child on the origin's side first; on the plane, the side the ray is heading to
                const bool below = (oc < split) || (oc == split && dc <= 0.0f);
                const uint32_t c0 = (n.y & 4u) ? idx + 1u : 0xFFFFFFFFu;
                const uint32_t c1 = (n.y & 8u) ? (n.y >> 4) : 0xFFFFFFFFu;
                const uint32_t near_c = below ? c0 : c1, far_c = below ? c1 : c0;
                // every comparison is widened by `slack`, so rounding in ts / t0 / t1 can only ADD a visit
                const float slack = 1e-6f * fmaxf(fabsf(ts), 1.0f);
                bool go_near = true, go_far = true;
                float near_t1 = t1, far_t0 = t0;
                if (ts == ts) {                       // NaN: the ray lies in the split plane -> both, intervals kept
                    if (ts > t1 + slack || ts < -slack) go_far = false;          // plane beyond the exit / behind the origin
                    else if (ts < t0 - slack) go_near = false;                   // plane before the entry
                    else { near_t1 = fminf(t1, ts + slack); far_t0 = fmaxf(t0, ts - slack); }
                }
                go_near = go_near && near_c != 0xFFFFFFFFu;
                go_far = go_far && far_c != 0xFFFFFFFFu;
                if (go_near && go_far) {
                    stack[sp].node = far_c; stack[sp].t0 = far_t0; stack[sp].t1 = t1; ++sp;
                    idx = near_c; t1 = near_t1;
                } else if (go_near) {
                    idx = near_c; t1 = near_t1;
                } else if (go_far) {
                    idx = far_c; t0 = far_t0;
                } else pop = true;
            }
        }
        if (pop) {
            if (!sp) break;
            --sp;
            idx = stack[sp].node; t0 = stack[sp].t0; t1 = stack[sp].t1;
        }
    }
    return best;
}

template <bool CULL, bool FAST, bool ORDERED>
__device__ __forceinline__ Hit trace_closest(const DScene& sc, V3 o, V3 d, float eps) {
    if (ORDERED) return trace_ordered<CULL, FAST>(sc, o, d, eps);
    return trace_reference_order<CULL, FAST>(sc, o, d, eps);
}
// same query, but the traversal may return as soon as it holds a candidate with t <= t_stop (the running closest
// only decreases, so "closest.t <= t_stop" is already decided)
template <bool CULL, bool FAST, bool ORDERED>
__device__ __forceinline__ Hit trace_closest_stop(const DScene& sc, V3 o, V3 d, float eps, float t_stop) {
    if (ORDERED) return trace_ordered<CULL, FAST>(sc, o, d, eps, t_stop);
    return trace_reference_order<CULL, FAST>(sc, o, d, eps, t_stop);
}

// ---- Philox4x32-10 (Salmon et al., SC'11) ---------------------------------------------------------------------------
// Counter-based, so the wavefront can draw a path node's numbers from (pixel, sample, path) alone.  The keying
// scheme (tags, slots) is specified in DESIGN.md "Random numbers"; the oracle implements the same specification.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ float u01(uint32_t x) { return float(x >> 8) * 0x1p-24f; }
constexpr uint32_t TAG_ROOT = 0x52544230u, TAG_CHILD = 0x4348494Cu, TAG_GI = 0x47495F5Fu;
constexpr uint32_t SLOT_REFRACT = 0, SLOT_REFLECT = 1, SLOT_GI0 = 2;
__device__ __forceinline__ uint2 child_key(uint2 key, uint32_t slot) {
    const uint4 r = philox4x32_10(make_uint4(slot, 0u, 0u, TAG_CHILD), key);
    return make_uint2(r.x, r.y);
}

}  // namespace rtb
