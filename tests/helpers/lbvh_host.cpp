// tests/helpers/lbvh_host.cpp - TEST INFRASTRUCTURE: runs the kernels of simd-raytracer_b200/csrc/rt_lbvh.cuh (the device-side
// hierarchy builder) thread by thread on the CPU, in the order the host glue (csrc/rt_api.cu device_build_bvh) launches them, so
// that tests/test_lbvh_host.py can check the builder on a machine without a GPU: the tree covers every triangle once, boxes
// enclose, the level-by-level four-wide collapse equals host/bvh4_collapse.hpp on the same two-wide tree, and the traversals of
// rt_bvh.cuh / rt_bvh4.cuh over the result give the oracle's hits.  A std::stable_sort stands in for the radix sort.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

// ---- the few CUDA built-ins the kernels use --------------------------------------------------------------------------------
#define __global__
#define __device__
#define __forceinline__ inline
struct Dim { unsigned x; };
static Dim blockIdx, blockDim, threadIdx;
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p = o + v; return o; }
static inline uint32_t atomicMax(uint32_t* p, uint32_t v) { const uint32_t o = *p; *p = std::max(o, v); return o; }
static inline void __threadfence() {}
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
using std::max;
using std::min;
#include "../../simd-raytracer_b200/csrc/rt_lbvh.cuh"
#include "../../simd-raytracer_b200/host/bvh4_collapse.hpp"

#define LAUNCH(n_threads, CALL)                                              \
    do {                                                                     \
        blockDim.x = 128;                                                    \
        for (uint64_t g_ = 0; g_ < uint64_t(n_threads); ++g_) {              \
            blockIdx.x = unsigned(g_ / 128); threadIdx.x = unsigned(g_ % 128); \
            CALL;                                                            \
        }                                                                    \
    } while (0)

namespace {
struct Built {
    std::vector<uint32_t> nodes16, tris12, nodes32;
    uint32_t n2 = 0, n4 = 0, stack_need = 0, depth2 = 0;
    float root6[6];
};
Built g_built;

// do the two four-wide trees agree node by node (same child order, boxes, leaf words), whatever their numbering?
bool same_tree4(const uint32_t* a, uint32_t na, const uint32_t* b, uint32_t nb) {
    struct P { uint32_t x, y; };
    std::vector<P> todo{{0u, 0u}};
    uint64_t seen = 0;
    while (!todo.empty()) {
        const P p = todo.back();
        todo.pop_back();
        if (p.x >= na || p.y >= nb) return false;
        ++seen;
        const uint32_t *u = a + size_t(p.x) * 32, *v = b + size_t(p.y) * 32;
        if (std::memcmp(u, v, 24 * 4) != 0) return false;
        for (int k = 0; k < 4; ++k) {
            const uint32_t cu = u[24 + k], cv = v[24 + k];
            if ((cu == 0xFFFFFFFFu) != (cv == 0xFFFFFFFFu)) return false;
            if (cu == 0xFFFFFFFFu) continue;
            if ((cu & 7u) != (cv & 7u)) return false;
            if (cu & 7u) { if (cu != cv) return false; }
            else todo.push_back({cu >> 3, cv >> 3});
        }
    }
    return seen == na && seen == nb;
}
}  // namespace

// returns 0, or a negative code naming the first step that went wrong; sizes through out6 = { n2, n4, stack_need, depth2, same-as-host-collapse, host stack need }
extern "C" int lbvh_build_host(const float* tri9, uint32_t n, const float* root6, uint32_t leaf, uint32_t* out6) {
    using namespace rtb;
    Built b;
    if (n <= 4 * LBVH_MAX_LEAF || leaf < 1 || leaf > LBVH_MAX_LEAF) return -1;
    const uint32_t n_inner = n - 1;
    std::vector<uint64_t> keys(n);
    std::vector<uint32_t> ids(n);
    LAUNCH(n, k_lbvh_morton(tri9, n, root6, keys.data(), ids.data()));
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return keys[x] < keys[y]; });
    std::vector<uint64_t> skeys(n);
    std::vector<uint32_t> sids(n);
    for (uint32_t k = 0; k < n; ++k) { skeys[k] = keys[order[k]]; sids[k] = ids[order[k]]; }
    std::vector<uint32_t> left(n), right(n), first(n), last(n), parent_inner(n, 0xDEADBEEFu), parent_leaf(n, 0xDEADBEEFu), arrived(n, 0u);
    std::vector<LbvhBox> leaf_box(n), inner_box(n);
    LAUNCH(n, k_lbvh_tree(skeys.data(), int(n), left.data(), right.data(), first.data(), last.data(), parent_inner.data(), parent_leaf.data()));
    LAUNCH(n, k_lbvh_boxes(tri9, sids.data(), int(n), left.data(), right.data(), parent_inner.data(), parent_leaf.data(), leaf_box.data(),
                           inner_box.data(), arrived.data()));
    for (uint32_t i = 0; i < n_inner; ++i) if (arrived[i] != 2u) return -2;          // every inner node was completed exactly once
    std::vector<uint32_t> kept(n, 0u), dense(n, 0u);
    LAUNCH(n, k_lbvh_mark(first.data(), last.data(), int(n_inner), leaf, kept.data()));
    uint32_t run = 0;
    for (uint32_t i = 0; i < n_inner; ++i) { dense[i] = run; run += kept[i]; }
    const uint32_t n2 = run;
    if (!n2) return -3;
    float scale = 0.0f;
    for (int c = 0; c < 3; ++c) scale = std::max(scale, std::max(std::fabs(inner_box[0].lo[c]), std::fabs(inner_box[0].hi[c])));
    const float pad = 2e-5f * scale;
    b.nodes16.assign(size_t(n2) * 16, 0u);
    b.tris12.assign(size_t(n) * 12, 0u);
    LAUNCH(n, k_lbvh_emit_nodes(left.data(), right.data(), first.data(), last.data(), kept.data(), dense.data(), leaf_box.data(), inner_box.data(),
                                int(n_inner), pad, b.nodes16.data()));
    LAUNCH(n, k_lbvh_emit_tris(tri9, sids.data(), n, b.tris12.data()));
    b.nodes32.assign(size_t(n2) * 32, 0u);
    std::vector<LbvhFrontier> fa(n2), fb(n2);
    LbvhCounters ctr{1u, 0u, 0u, 0u};
    fa[0] = LbvhFrontier{0u, 0u, 0u, 0u};
    uint32_t n_in = 1;
    for (int level = 0; n_in; ++level) {
        if (level > 256) return -4;
        LAUNCH(n_in, k_lbvh_collapse_level(b.nodes16.data(), fa.data(), n_in, fb.data(), n2, &ctr, b.nodes32.data(), n2));
        n_in = ctr.next_count;
        ctr.next_count = 0;
        std::swap(fa, fb);
    }
    if (ctr.n_nodes4 > n2) return -5;
    b.nodes32.resize(size_t(ctr.n_nodes4) * 32);
    b.n2 = n2; b.n4 = ctr.n_nodes4; b.stack_need = ctr.stack_need; b.depth2 = ctr.depth2;
    for (int c = 0; c < 3; ++c) { b.root6[c] = inner_box[0].lo[c] - pad; b.root6[3 + c] = inner_box[0].hi[c] + pad; }
    const std::vector<uint32_t> ref4 = rtb::bvh4_collapse(b.nodes16.data(), n2);
    out6[0] = n2; out6[1] = b.n4; out6[2] = b.stack_need; out6[3] = b.depth2;
    out6[4] = same_tree4(b.nodes32.data(), b.n4, ref4.data(), uint32_t(ref4.size() / 32)) ? 1u : 0u;
    out6[5] = rtb::bvh4_stack_need(ref4);
    g_built = std::move(b);
    return 0;
}
extern "C" void lbvh_fetch_host(uint32_t* nodes16, uint32_t* tris12, uint32_t* nodes32, float* root6) {
    std::memcpy(nodes16, g_built.nodes16.data(), g_built.nodes16.size() * 4);
    std::memcpy(tris12, g_built.tris12.data(), g_built.tris12.size() * 4);
    std::memcpy(nodes32, g_built.nodes32.data(), g_built.nodes32.size() * 4);
    std::memcpy(root6, g_built.root6, 24);
}
