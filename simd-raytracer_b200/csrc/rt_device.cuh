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

#include "rt_bvh4.cuh"

namespace rtb {

// Programmatic dependent launch (rt_api.cu launch_k): first statement of every kernel of a pass.  Waits until the kernel in
// front of this one in the stream has completed and its writes are visible, then lets the next kernel of the stream be set up.
// A no-op when the kernel was launched without the attribute.
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}


// ---- resident scene (all pointers into HBM) -------------------------------------------------------------------
struct DMaterial { uint32_t kind; float albedo[3]; float ior; uint32_t smooth; int32_t texture; };       // 28 B
struct DTexture { uint32_t kind; float c0[3], c1[3], scalar; uint32_t w, h, off; };                      // 44 B
struct DLight { float pos[3], intensity; };                                                              // 16 B

struct DScene {
    const float* __restrict__ b_nodes;      // accelerated mode: 64-byte two-child nodes of the bounding-volume hierarchy (rt_bvh.cuh)
    const float* __restrict__ b_tris;       //                   and its 48-byte triangle records (one per triangle)
    const float* __restrict__ w_nodes;      // accelerated mode, four-wide variant (rt_bvh4.cuh): 128-byte nodes over the same records; null = two-wide
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
    uint32_t w_stack_rows;                  // four-wide stream kernels: stack entries per lane in shared memory (rt_stream.cuh)
    uint32_t width, height;
    float bg[3];
    float cam_pos[3];
    float cam_m[9];
    float root_min[3], root_max[3];
    float b_root_min[3], b_root_max[3];     // root box of the BVH (tight bounds of the triangles)
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

// ---- the structure behind the accelerated query mode -----------------------------------------------------------------------
// The backend's own bounding-volume hierarchy (host/bvh_build.cpp), two-wide (rt_bvh.cuh) or four-wide (rt_bvh4.cuh); both give
// the same hits (same triangle arithmetic, same tie rule).
using AccelState = BvhState;
using AccelStackEntry = BvhStackEntry;
constexpr int ACCEL_STACK = BVH_STACK;
__device__ __forceinline__ bool accel_init(AccelState& st, const DScene& sc, float ox, float oy, float oz, float dx, float dy, float dz,
                                           float t_far, bool any_hit) {
    return bvh_init(st, sc.root_min, sc.root_max, ox, oy, oz, dx, dy, dz, t_far, any_hit);
}
__device__ __forceinline__ void accel_node_step(AccelState& st, AccelStackEntry* stack, const DScene& sc) { bvh_node_step(st, stack, sc.b_nodes); }
template <bool CULL, bool FAST>
__device__ __forceinline__ void accel_leaf_step(AccelState& st, const AccelStackEntry* stack, const DScene& sc, float eps) {
    bvh_leaf_step<CULL, FAST>(st, stack, sc.b_tris, eps);
}
// the one-ray-per-thread query of the batch entry points (rt_trace_closest / rt_trace_occluded / rt_trace_primary) walks the
// hierarchy the scene was built with; the stream kernels of the frame path pick theirs at compile time (rt_stream.cuh)
template <bool CULL, bool FAST>
__device__ __forceinline__ KdHit accel_trace(const DScene& sc, float ox, float oy, float oz, float dx, float dy, float dz, float eps,
                                             float t_far, bool any_hit) {
    if (sc.w_nodes) return bvh4_trace<CULL, FAST>(sc.w_nodes, sc.b_tris, sc.root_min, sc.root_max, ox, oy, oz, dx, dy, dz, eps, t_far, any_hit);
    return bvh_trace<CULL, FAST>(sc.b_nodes, sc.b_tris, sc.root_min, sc.root_max, ox, oy, oz, dx, dy, dz, eps, t_far, any_hit);
}

// ---- ray vs. 4-triangle SoA packet --------------------------------------------------------------------------------
// triangle_packet<F,W>::intersect (kd_tree_simd.hpp:25-60), one lane, in the reference's operation order, folded
// into intersect_leaf's running closest (kd_tree_simd.hpp:266-302).  Using the traversal-wide best_t directly with
// a strict `<` is equivalent to the reference's leaf-local candidate followed by `c->t < best_t` (:224): in both,
// the first occurrence of the smallest t strictly below the incoming best wins.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Exact mode computes the reference's expressions in the reference's order, but only for triangles that can still
// pass: u and v are first estimated with the 1-ulp MUFU reciprocal (relative error of the estimate < 1e-6) and a
// triangle is dropped when the estimate is outside [0,1] by more than 1e-4 - the exactly rounded value is then
// outside too, so the outcome is the reference's.  NaN estimates never drop anything (every compare is false).
template <bool CULL, bool FAST>
__device__ __forceinline__ void test_lane(float v0x, float v0y, float v0z, float e1x, float e1y, float e1z,
                                          float e2x, float e2y, float e2z, int id, V3 o, V3 d, float eps, Hit& best) {
    constexpr float M = 1e-4f;
    if (FAST) {
        const float pvx = fmaf(d.y, e2z, -(d.z * e2y)), pvy = fmaf(d.z, e2x, -(d.x * e2z)), pvz = fmaf(d.x, e2y, -(d.y * e2x));
        const float det = fmaf(e1z, pvz, fmaf(e1y, pvy, e1x * pvx));
        if (!(CULL ? (eps <= det) : (eps <= fabsf(det)))) return;
        const float inv_det = __frcp_rn(det);
        const float tx = o.x - v0x, ty = o.y - v0y, tz = o.z - v0z;
        const float u = fmaf(tz, pvz, fmaf(ty, pvy, tx * pvx)) * inv_det;
        if (!((0.0f <= u) & (u <= 1.0f))) return;
        const float qx = fmaf(ty, e1z, -(tz * e1y)), qy = fmaf(tz, e1x, -(tx * e1z)), qz = fmaf(tx, e1y, -(ty * e1x));
        const float v = fmaf(d.z, qz, fmaf(d.y, qy, d.x * qx)) * inv_det;
        if (!((0.0f <= v) & (u + v <= 1.0f))) return;
        const float t = fmaf(e2z, qz, fmaf(e2y, qy, e2x * qx)) * inv_det;
        if ((eps < t) & (t < best.t)) { best.t = t; best.u = u; best.v = v; best.tri = id; }
        return;
    }
    const float pvx = d.y * e2z - d.z * e2y;                                                             // :27
    const float pvy = d.z * e2x - d.x * e2z;                                                             // :28
    const float pvz = d.x * e2y - d.y * e2x;                                                             // :29
    const float det = e1x * pvx + e1y * pvy + e1z * pvz;                                                 // :31
    if (!(CULL ? (eps <= det) : (eps <= fabsf(det)))) return;                                            // :33-38
    const float tx = o.x - v0x, ty = o.y - v0y, tz = o.z - v0z;                                          // :42-44
    const float un = tx * pvx + ty * pvy + tz * pvz;                                                     // :46 (numerator)
    const float r = rcp_approx(det);
    const float ua = un * r;
    if ((ua < -M) | (ua > 1.0f + M)) return;
    const float qx = ty * e1z - tz * e1y;                                                                // :49
    const float qy = tz * e1x - tx * e1z;                                                                // :50
    const float qz = tx * e1y - ty * e1x;                                                                // :51
    const float vn = d.x * qx + d.y * qy + d.z * qz;                                                     // :53 (numerator)
    const float va = vn * r;
    if ((va < -M) | (ua + va > 1.0f + 3.0f * M)) return;
    const float inv_det = __fdiv_rn(1.0f, det);                                                          // :40
    const float u = un * inv_det;                                                                        // :46
    const float v = vn * inv_det;                                                                        // :53
    const float t = (e2x * qx + e2y * qy + e2z * qz) * inv_det;                                          // :56
    const bool ok = (0.0f <= u) & (u <= 1.0f) & (0.0f <= v) & (u + v <= 1.0f) & (eps < t);               // :47,:54,:57
    if (ok && t < best.t) { best.t = t; best.u = u; best.v = v; best.tri = id; }                         // :284-298, :224
}

template <bool CULL, bool FAST>
__device__ __forceinline__ void test_packets(const float4* __restrict__ pk, uint32_t count, V3 o, V3 d, float eps,
                                             Hit& best) {
    for (uint32_t p = 0; p < count; ++p, pk += 10) {
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

// ---- leaf test, one ray per WARP ------------------------------------------------------------------------------------
// The direct analogue of the reference's W-wide packet (kd_tree_simd.hpp:25-60): one ray against 32 triangles of the
// leaf at a time, lane k testing triangles k, k+32, ... of the leaf's list, then a (t, position) min-reduction over
// the lanes - "first triangle in list order among the smallest t", exactly what intersect_leaf's pack loop keeps
// (:276-298: first pack wins ties, lowest lane inside a pack).  Padding lanes of the last packet repeat the leaf's
// last triangle at a higher position and can only tie with it.
// `limit`: only candidates with t < limit are of interest (the caller's running closest).
template <bool CULL, bool FAST>
__device__ __forceinline__ bool leaf_warp(const float* __restrict__ packets, uint32_t first_packet, uint32_t n_packets, V3 o, V3 d,
                                          float eps, float limit, Hit& out) {
    const uint32_t lane = threadIdx.x & 31u;
    Hit c; c.t = limit; c.u = 0.0f; c.v = 0.0f; c.tri = -1;
    uint32_t pos = 0xFFFFFFFFu;
    const uint32_t n = n_packets * 4u;
    for (uint32_t k = lane; k < n; k += 32u) {
        const float* p = packets + (size_t(first_packet) + (k >> 2)) * 40u + (k & 3u);
        const float before = c.t;
        test_lane<CULL, FAST>(__ldg(p), __ldg(p + 4), __ldg(p + 8), __ldg(p + 12), __ldg(p + 16), __ldg(p + 20), __ldg(p + 24),
                              __ldg(p + 28), __ldg(p + 32), __float_as_int(__ldg(p + 36)), o, d, eps, c);
        if (c.t < before) pos = k;
    }
    if (!__ballot_sync(0xFFFFFFFFu, pos != 0xFFFFFFFFu)) return false;
    float t = c.t;
    uint32_t q = pos;
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        const float ot = __shfl_xor_sync(0xFFFFFFFFu, t, off);
        const uint32_t oq = __shfl_xor_sync(0xFFFFFFFFu, q, off);
        if (ot < t || (ot == t && oq < q)) { t = ot; q = oq; }
    }
    const int src = __ffs(__ballot_sync(0xFFFFFFFFu, pos == q)) - 1;       // positions are unique
    out.t = t;
    out.u = __shfl_sync(0xFFFFFFFFu, c.u, src);
    out.v = __shfl_sync(0xFFFFFFFFu, c.v, src);
    out.tri = __shfl_sync(0xFFFFFFFFu, c.tri, src);
    return true;
}

// ---- closest hit ---------------------------------------------------------------------------------------------------------
// kd_tree_simd_accel::intersect<bf> (kd_tree_simd.hpp:187-229) for the 32 rays of a warp.  Must be called by all 32
// lanes (`active` = this lane has a ray).
//
// Inner nodes are walked per lane, exactly as the reference does per ray: LIFO stack, child0 pushed before child1
// (so child1 is visited first), every popped node slab-tested against its own box, pruned when best_t < box.t_min
// (strict).
//
// Leaves are where the time goes (up to 752 triangles per leaf with the reference's default <8,64> tree) and where a
// one-ray-per-lane loop diverges, so they are handled by the warp: lanes that arrived at the SAME leaf (coherent
// rays) and form a large enough group walk it one ray per lane, sharing every triangle fetch; any other leaf visit is
// executed by all 32 lanes for its one ray (leaf_warp).
//
// t_stop >= 0: the caller only needs to know whether closest.t <= t_stop; the running closest only decreases, so a
// lane stops as soon as it holds such a candidate (occluded_query, non-transmissive scenes).
constexpr int COHERENT_GROUP = 20;

template <bool CULL, bool FAST>
__device__ __forceinline__ Hit trace_warp(const DScene& sc, bool active, V3 o, V3 d, float eps, float t_stop = -1.0f) {
    const uint32_t lane = threadIdx.x & 31u;
    const V3 inv = mk(__fdiv_rn(1.0f, d.x), __fdiv_rn(1.0f, d.y), __fdiv_rn(1.0f, d.z));                 // ray3.hpp:11-14
    Hit best; best.t = FLT_MAX; best.u = 0.0f; best.v = 0.0f; best.tri = -1;
    uint32_t stack[KD_STACK];
    int sp = 0;
    if (active) stack[sp++] = 0u;
    const float* packets = reinterpret_cast<const float*>(sc.packets);
    for (;;) {
        // ---- walk inner nodes until this lane stands at a leaf it has to test ----
        uint32_t leaf_first = 0, leaf_count = 0;
        bool want = false;
        while (sp) {
            const uint32_t idx = stack[--sp];
            const float4 lo = __ldg(sc.nodes32 + 2 * idx), hi = __ldg(sc.nodes32 + 2 * idx + 1);
            float t_min;
            if (!slab(lo, hi, o, inv, t_min) || best.t < t_min) continue;                                // :202-205
            const uint32_t word = __float_as_uint(hi.w);
            const uint32_t axis = word & 3u;
            if (axis != 3u) {                                                                            // inner, :207-214
                if (word & 4u) stack[sp++] = idx + 1u;
                if (word & 8u) stack[sp++] = word >> 4;
            } else {                                                                                     // leaf, :216-226
                leaf_first = __float_as_uint(lo.w); leaf_count = word >> 2;
                want = true;
                break;
            }
        }
        uint32_t pending = __ballot_sync(0xFFFFFFFFu, want);
        if (!pending) break;
        // ---- leaf tests ----
        while (pending) {
            const int leader = __ffs(pending) - 1;
            const uint32_t lf = __shfl_sync(0xFFFFFFFFu, leaf_first, leader);
            const uint32_t lc = __shfl_sync(0xFFFFFFFFu, leaf_count, leader);
            const bool same = want && leaf_first == lf;
            uint32_t group = __ballot_sync(0xFFFFFFFFu, same);
            Hit cand; cand.t = FLT_MAX; cand.u = 0.0f; cand.v = 0.0f; cand.tri = -1;
            bool have = false;
            const float my_limit = best.t;                  // a later leaf needs a strictly smaller t (:224)
            if (__popc(group) >= COHERENT_GROUP) {
                if (same) {
                    cand.t = my_limit;
                    test_packets<CULL, FAST>(sc.packets + 10ull * lf, lc, o, d, eps, cand);
                    have = cand.tri >= 0;
                }
            } else {
                group = 1u << leader;
                const V3 ro = mk(__shfl_sync(0xFFFFFFFFu, o.x, leader), __shfl_sync(0xFFFFFFFFu, o.y, leader), __shfl_sync(0xFFFFFFFFu, o.z, leader));
                const V3 rd = mk(__shfl_sync(0xFFFFFFFFu, d.x, leader), __shfl_sync(0xFFFFFFFFu, d.y, leader), __shfl_sync(0xFFFFFFFFu, d.z, leader));
                const float limit = __shfl_sync(0xFFFFFFFFu, my_limit, leader);
                Hit w;
                const bool found = leaf_warp<CULL, FAST>(packets, lf, lc, ro, rd, eps, limit, w);
                if (found && lane == uint32_t(leader)) { cand = w; have = true; }
            }
            if (have) {
                if (cand.t < best.t) best = cand;                                                        // :224
                if (best.t <= t_stop) sp = 0;
            }
            if (group & (1u << lane)) want = false;
            pending &= ~group;
        }
    }
    return best;
}

// ---- one entry point for the kernels -----------------------------------------------------------------------------------
// ACCEL == false: the reference's tree in the reference's order (bit-exact by construction, statistics included).
//                 any_hit lets a lane stop once closest.t <= t_far is decided; t_far is otherwise ignored.
// ACCEL == true : rt_bvh.cuh / rt_bvh4.cuh - the backend's own hierarchy, near child first, one ray per thread; nodes beyond t_far are
//                 not visited at all (a hit beyond t_far and a miss mean the same to every caller that passes t_far).
//                 Must be called by all 32 lanes (tie re-runs use the warp-cooperative query).
template <bool CULL, bool FAST, bool ACCEL>
__device__ __forceinline__ Hit trace_any(const DScene& sc, bool active, V3 o, V3 d, float eps, float t_far = FLT_MAX, bool any_hit = false) {
    if (ACCEL) {
        Hit h; h.t = FLT_MAX; h.u = 0.0f; h.v = 0.0f; h.tri = -1;
        bool tie = false;
        if (active) {
            const KdHit k = accel_trace<CULL, FAST>(sc, o.x, o.y, o.z, d.x, d.y, d.z, eps, t_far, any_hit);
            h.t = k.t; h.u = k.u; h.v = k.v; h.tri = k.tri;
            tie = (k.tri == KD_RERUN) || (!any_hit && k.tri >= 0 && k.tie_t == k.t);
        }
        // two different triangles at exactly the winner's t: which one the reference reports depends on its leaf order,
        // so those rays (rare) take the reference-order query
        if (__ballot_sync(0xFFFFFFFFu, tie)) {
            const Hit e = trace_warp<CULL, FAST>(sc, tie, o, d, eps);
            if (tie) h = e;
        }
        return h;
    }
    return trace_warp<CULL, FAST>(sc, active, o, d, eps, (any_hit && t_far < FLT_MAX) ? t_far : -1.0f);
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
