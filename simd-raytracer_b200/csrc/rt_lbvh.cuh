// csrc/rt_lbvh.cuh - device-side builder of the backend's bounding-volume hierarchy ("device-side
// or parallel host builder").  The host builder (host/bvh_build.cpp, binned SAH on the task-parallel driver) needs ~6 s for
// 10 M triangles on 16 host threads; this one builds the same KIND of structure - two-wide 64-byte nodes over 48-byte triangle
// records with at most `leaf` (1..4) triangles per leaf, plus the four-wide collapse - in tens of milliseconds on the GPU.  Only the backend's
// OWN hierarchy can be built this way: the reference's kd-tree must stay the reference's (host/kd_build.cpp).  The answer of a
// query does not depend on the hierarchy (rt_tri.cuh), so the parity tests apply unchanged; what a different tree changes is
// the number of node visits.
//
// Method: a linear BVH.  (1) 63-bit Morton code of every triangle's box centre inside the scene's box; (2) radix sort of
// (code, triangle) pairs (cub::DeviceRadixSort - library plumbing, not a hot path); (3) the binary radix tree over the sorted
// codes, every inner node found independently from the common prefixes of its neighbours (Karras 2012), codes made unique by
// the position in the sorted order; (4) boxes bottom-up, the second thread to arrive at a node merges its children; (5) every
// radix-tree node that spans more than `leaf` triangles becomes a two-wide node (numbered densely by a prefix sum), one that
// spans <= `leaf` a leaf over a contiguous run of the SORTED triangle records; (6) the four-wide collapse level by level from the
// root (the greedy rule of host/bvh4_collapse.hpp), which also yields the tree depth and the worst-case stack need that scene
// creation checks.  Boxes carry the same absolute padding as the host builder's (2e-5 of the scene's largest coordinate).
#pragma once

#include <cfloat>
#include <cstdint>
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#endif                                          // tests/helpers/lbvh_host.cpp runs these kernels thread by thread on the CPU over its own shims

namespace rtb {

constexpr uint32_t LBVH_MAX_LEAF = 4;           // most triangles a leaf can hold (the leaf size is a build parameter, 1..LBVH_MAX_LEAF)
constexpr uint32_t LBVH_NONE = 0xFFFFFFFFu;

struct LbvhBox { float lo[3], hi[3]; };

__device__ __forceinline__ uint64_t lbvh_spread21(uint32_t v) {         // 21 bits -> every third bit of 63
    uint64_t x = v & 0x1FFFFFull;
    x = (x | x << 32) & 0x1F00000000FFFFull;
    x = (x | x << 16) & 0x1F0000FF0000FFull;
    x = (x | x << 8) & 0x100F00F00F00F00Full;
    x = (x | x << 4) & 0x10C30C30C30C30C3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// tri9: v0, e1, e2 per triangle (the numbers every triangle test reads).  The box is that of v0, v0 + e1, v0 + e2.
__device__ __forceinline__ LbvhBox lbvh_tri_box(const float* __restrict__ tri9, uint32_t id) {
    const float* p = tri9 + size_t(id) * 9;
    LbvhBox b;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float a = p[c], u = p[c] + p[3 + c], v = p[c] + p[6 + c];
        b.lo[c] = fminf(a, fminf(u, v)); b.hi[c] = fmaxf(a, fmaxf(u, v));
    }
    return b;
}

__global__ void k_lbvh_morton(const float* __restrict__ tri9, uint32_t n, const float* __restrict__ root6, uint64_t* __restrict__ keys,
                              uint32_t* __restrict__ ids) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const LbvhBox b = lbvh_tri_box(tri9, i);
    uint64_t code = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float ext = root6[3 + c] - root6[c];
        float x = ext > 0.0f ? ((0.5f * b.lo[c] + 0.5f * b.hi[c]) - root6[c]) / ext : 0.0f;
        x = fminf(fmaxf(x, 0.0f), 1.0f);
        const uint32_t q = min(uint32_t(x * 2097152.0f), 2097151u);
        code |= lbvh_spread21(q) << (2 - c);
    }
    keys[i] = code; ids[i] = i;
}

// common-prefix length of the codes at sorted positions i and j (j may be out of range: -1); equal codes are told apart by i, j
__device__ __forceinline__ int lbvh_delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    return a == b ? 64 + __clz(uint32_t(i) ^ uint32_t(j)) : __clzll((long long)(a ^ b));
}

// inner node i of the binary radix tree (0 <= i < n - 1): its range [first, last] of sorted positions and its two children.
// child encoding: bit 31 set = leaf (a single sorted position), else inner node index.
__global__ void k_lbvh_tree(const uint64_t* __restrict__ keys, int n, uint32_t* __restrict__ left, uint32_t* __restrict__ right,
                            uint32_t* __restrict__ first_out, uint32_t* __restrict__ last_out, uint32_t* __restrict__ parent_inner,
                            uint32_t* __restrict__ parent_leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const uint32_t cl = (lo == gamma) ? (0x80000000u | uint32_t(gamma)) : uint32_t(gamma);
    const uint32_t cr = (hi == gamma + 1) ? (0x80000000u | uint32_t(gamma + 1)) : uint32_t(gamma + 1);
    left[i] = cl; right[i] = cr; first_out[i] = uint32_t(lo); last_out[i] = uint32_t(hi);
    if (cl & 0x80000000u) parent_leaf[gamma] = uint32_t(i); else parent_inner[gamma] = uint32_t(i);
    if (cr & 0x80000000u) parent_leaf[gamma + 1] = uint32_t(i); else parent_inner[gamma + 1] = uint32_t(i);
    if (i == 0) parent_inner[0] = LBVH_NONE;
}

// boxes bottom-up: one thread per leaf walks towards the root; at every inner node the first thread to arrive stops, the second
// (its sibling's subtree is complete and visible after the fence) merges the two children's boxes and goes on
__global__ void k_lbvh_boxes(const float* __restrict__ tri9, const uint32_t* __restrict__ ids, int n, const uint32_t* __restrict__ left,
                             const uint32_t* __restrict__ right, const uint32_t* __restrict__ parent_inner, const uint32_t* __restrict__ parent_leaf,
                             LbvhBox* __restrict__ leaf_box, LbvhBox* __restrict__ inner_box, uint32_t* __restrict__ arrived) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    leaf_box[k] = lbvh_tri_box(tri9, ids[k]);
    __threadfence();
    uint32_t node = parent_leaf[k];
    while (node != LBVH_NONE) {
        if (atomicAdd(&arrived[node], 1u) == 0u) return;
        __threadfence();
        const uint32_t cl = left[node], cr = right[node];
        const volatile LbvhBox* a = (cl & 0x80000000u) ? leaf_box + (cl & 0x7FFFFFFFu) : inner_box + cl;
        const volatile LbvhBox* b = (cr & 0x80000000u) ? leaf_box + (cr & 0x7FFFFFFFu) : inner_box + cr;
        LbvhBox m;
#pragma unroll
        for (int c = 0; c < 3; ++c) { m.lo[c] = fminf(a->lo[c], b->lo[c]); m.hi[c] = fmaxf(a->hi[c], b->hi[c]); }
        inner_box[node] = m;
        __threadfence();
        node = parent_inner[node];
    }
}

// kept[i] = 1 when radix-tree node i spans more than `leaf` triangles: it becomes a two-wide node (dense index = exclusive
// prefix sum of kept, computed by the host glue with cub::DeviceScan)
__global__ void k_lbvh_mark(const uint32_t* __restrict__ first, const uint32_t* __restrict__ last, int n_inner, uint32_t leaf, uint32_t* __restrict__ kept) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_inner) kept[i] = (last[i] - first[i] + 1u > leaf) ? 1u : 0u;
}

// the 64-byte two-wide nodes of csrc/rt_bvh.cuh: { c0.min.xyz, c0.max.xyz, c1.min.xyz, c1.max.xyz, ref0, ref1, cnt0, cnt1 }
__global__ void k_lbvh_emit_nodes(const uint32_t* __restrict__ left, const uint32_t* __restrict__ right, const uint32_t* __restrict__ first,
                                  const uint32_t* __restrict__ last, const uint32_t* __restrict__ kept, const uint32_t* __restrict__ dense,
                                  const LbvhBox* __restrict__ leaf_box, const LbvhBox* __restrict__ inner_box, int n_inner, float pad,
                                  uint32_t* __restrict__ nodes16) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_inner || !kept[i]) return;
    uint32_t* node = nodes16 + size_t(dense[i]) * 16;
    const uint32_t ch[2] = {left[i], right[i]};
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const uint32_t c = ch[s];
        LbvhBox b; uint32_t ref, cnt;
        if (c & 0x80000000u) { b = leaf_box[c & 0x7FFFFFFFu]; ref = c & 0x7FFFFFFFu; cnt = 1u; }
        else {
            b = inner_box[c];
            if (kept[c]) { ref = dense[c]; cnt = 0u; }
            else { ref = first[c]; cnt = last[c] - first[c] + 1u; }           // a whole small subtree = one leaf over a run of sorted records
        }
        const int o = s ? 6 : 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) { node[o + a] = __float_as_uint(b.lo[a] - pad); node[o + 3 + a] = __float_as_uint(b.hi[a] + pad); }
        node[12 + s] = ref; node[14 + s] = cnt;
    }
}

// the 48-byte triangle records, in sorted order: { v0.xyz, id } { e1.xyz, 0 } { e2.xyz, 0 }
__global__ void k_lbvh_emit_tris(const float* __restrict__ tri9, const uint32_t* __restrict__ ids, uint32_t n, uint32_t* __restrict__ tris12) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t id = ids[k];
    const float* p = tri9 + size_t(id) * 9;
    uint4* o = reinterpret_cast<uint4*>(tris12 + size_t(k) * 12);
    o[0] = make_uint4(__float_as_uint(p[0]), __float_as_uint(p[1]), __float_as_uint(p[2]), id);
    o[1] = make_uint4(__float_as_uint(p[3]), __float_as_uint(p[4]), __float_as_uint(p[5]), 0u);
    o[2] = make_uint4(__float_as_uint(p[6]), __float_as_uint(p[7]), __float_as_uint(p[8]), 0u);
}

// ---- four-wide collapse, one tree level per launch ------------------------------------------------------------------------------
// A frontier entry is a two-wide node that becomes a four-wide node: { two-wide index, four-wide index, stack entries below it,
// depth of the two-wide node }.  Every thread collapses one entry by the rule of host/bvh4_collapse.hpp (while fewer than four children, replace the
// inner child with the largest surface area by its two children), writes the 128-byte node of csrc/rt_bvh4.cuh, and appends its
// inner children to the next frontier with four-wide indices drawn from a global counter.
struct LbvhFrontier { uint32_t node2, node4, sp, depth2; };
struct LbvhCounters { uint32_t n_nodes4, next_count, stack_need, depth2; };   // depth2: deepest leaf of the TWO-wide tree (root's children = 1)

struct LbvhChild { float lo[3], hi[3]; uint32_t ref, cnt; };
__device__ __forceinline__ void lbvh_children2(const uint32_t* __restrict__ nodes16, uint32_t node, LbvhChild out[2]) {
    const uint4* q = reinterpret_cast<const uint4*>(nodes16 + size_t(node) * 16);
    const uint4 a = q[0], b = q[1], c = q[2], d = q[3];
    out[0].lo[0] = __uint_as_float(a.x); out[0].lo[1] = __uint_as_float(a.y); out[0].lo[2] = __uint_as_float(a.z);
    out[0].hi[0] = __uint_as_float(a.w); out[0].hi[1] = __uint_as_float(b.x); out[0].hi[2] = __uint_as_float(b.y);
    out[1].lo[0] = __uint_as_float(b.z); out[1].lo[1] = __uint_as_float(b.w); out[1].lo[2] = __uint_as_float(c.x);
    out[1].hi[0] = __uint_as_float(c.y); out[1].hi[1] = __uint_as_float(c.z); out[1].hi[2] = __uint_as_float(c.w);
    out[0].ref = d.x; out[1].ref = d.y; out[0].cnt = d.z; out[1].cnt = d.w;
}
__device__ __forceinline__ float lbvh_area(const LbvhChild& c) {
    const float x = c.hi[0] - c.lo[0], y = c.hi[1] - c.lo[1], z = c.hi[2] - c.lo[2];
    return x * y + y * z + z * x;
}

__global__ void k_lbvh_collapse_level(const uint32_t* __restrict__ nodes16, const LbvhFrontier* __restrict__ in, uint32_t n_in,
                                      LbvhFrontier* __restrict__ out, uint32_t out_cap, LbvhCounters* __restrict__ ctr, uint32_t* __restrict__ nodes32,
                                      uint32_t nodes4_cap) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_in) return;
    const LbvhFrontier f = in[t];
    LbvhChild c[4];
    int n = 2;
    lbvh_children2(nodes16, f.node2, c);
    int lvl[4] = {1, 1, 0, 0};                        // two-wide depth of every child below f.node2
    for (int k = 0; k < n;) { if (c[k].cnt == LBVH_NONE) { c[k] = c[n - 1]; lvl[k] = lvl[n - 1]; --n; } else ++k; }
    while (n < 4) {
        int best = -1;
        for (int k = 0; k < n; ++k)
            if (c[k].cnt == 0u && (best < 0 || lbvh_area(c[k]) > lbvh_area(c[best]))) best = k;
        if (best < 0) break;
        LbvhChild g[2];
        lbvh_children2(nodes16, c[best].ref, g);
        const int l = lvl[best] + 1;
        int m = 0;
        LbvhChild keep[2];
        for (int k = 0; k < 2; ++k) if (g[k].cnt != LBVH_NONE) keep[m++] = g[k];
        if (m == 0) { c[best] = c[n - 1]; lvl[best] = lvl[n - 1]; --n; continue; }
        c[best] = keep[0]; lvl[best] = l;
        if (m == 2) { c[n] = keep[1]; lvl[n] = l; ++n; }
    }
    const uint32_t below = f.sp + uint32_t(n ? n - 1 : 0);
    atomicMax(&ctr->stack_need, below);
    uint32_t* node = nodes32 + size_t(f.node4) * 32;
    uint32_t child[4];
    for (int k = 0; k < 4; ++k) {
        const bool have = k < n;
        for (int a = 0; a < 3; ++a) {
            node[a * 4 + k] = __float_as_uint(have ? c[k].lo[a] : 0.0f);
            node[12 + a * 4 + k] = __float_as_uint(have ? c[k].hi[a] : 0.0f);
        }
        child[k] = LBVH_NONE;
        if (have) {
            if (c[k].cnt == 0u) {
                const uint32_t idx4 = atomicAdd(&ctr->n_nodes4, 1u);
                const uint32_t slot = atomicAdd(&ctr->next_count, 1u);
                if (idx4 < nodes4_cap && slot < out_cap) out[slot] = LbvhFrontier{c[k].ref, idx4, below, f.depth2 + uint32_t(lvl[k])};
                child[k] = idx4 << 3;
            } else {
                child[k] = (c[k].ref << 3) | c[k].cnt;
                atomicMax(&ctr->depth2, f.depth2 + uint32_t(lvl[k]));
            }
        }
        node[24 + k] = child[k];
        node[28 + k] = 0u;
    }
}

}  // namespace rtb
