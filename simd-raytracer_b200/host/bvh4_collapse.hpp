// host/bvh4_collapse.hpp - the four-wide hierarchy of csrc/rt_bvh4.cuh from the flattened two-wide one (BvhLayout::nodes,
// host/bvh_build.cpp): a node's child list starts as its two children; while it holds fewer than four entries, the inner child
// with the largest surface area is replaced by its own two children.  The child boxes are taken over unchanged (they carry the
// builder's padding), leaves keep pointing into the same triangle records, nodes are emitted in DFS pre-order.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace rtb {

struct Bvh4Child { float lo[3], hi[3]; uint32_t ref, cnt; };

inline void bvh2_children(const uint32_t* nodes16, uint32_t node, Bvh4Child out[2]) {
    float f[12];
    std::memcpy(f, nodes16 + size_t(node) * 16, sizeof f);
    const uint32_t* w = nodes16 + size_t(node) * 16 + 12;
    out[0] = Bvh4Child{{f[0], f[1], f[2]}, {f[3], f[4], f[5]}, w[0], w[2]};
    out[1] = Bvh4Child{{f[6], f[7], f[8]}, {f[9], f[10], f[11]}, w[1], w[3]};
}

inline float bvh4_area(const Bvh4Child& c) {
    const float x = c.hi[0] - c.lo[0], y = c.hi[1] - c.lo[1], z = c.hi[2] - c.lo[2];
    return x * y + y * z + z * x;
}

// nodes16: the two-wide nodes (16 words each), node 0 the root; returns the four-wide nodes (32 words each), node 0 the root.
// Throws std::length_error when a leaf or an index does not fit the packed child word (ref << 3 | cnt).
// stack_need (optional): the worst-case traversal stack of the result, as bvh4_stack_need below computes it - the collapse visits
// every node anyway, so scene creation gets the figure without a second pass over a 10 M-triangle tree.
inline std::vector<uint32_t> bvh4_collapse(const uint32_t* nodes16, uint64_t n_nodes2, uint32_t* stack_need = nullptr) {
    constexpr uint32_t NO_CHILD = 0xFFFFFFFFu, MAX_LEAF = 7u, MAX_REF = 0x1FFFFFFDu;       // csrc/rt_bvh4.cuh
    std::vector<uint32_t> out;
    if (stack_need) *stack_need = 0;
    if (!n_nodes2) return out;
    out.reserve(size_t(n_nodes2) * 16);               // a four-wide tree has about half the nodes of the two-wide one
    struct Todo { uint32_t node2, slot, sp; };        // two-wide subtree root -> where its four-wide node index has to be written; stack entries below it
    std::vector<Todo> todo{{0u, NO_CHILD, 0u}};
    uint32_t need = 0;
    while (!todo.empty()) {
        const Todo t = todo.back();
        todo.pop_back();
        const uint32_t me = uint32_t(out.size() / 32);
        if (me > MAX_REF) throw std::length_error("bvh4: more than 2^29 nodes");
        if (t.slot != NO_CHILD) out[t.slot] = me << 3;
        Bvh4Child c[4];
        int n = 2;
        bvh2_children(nodes16, t.node2, c);
        // a missing child of the two-wide node is dropped from the list
        for (int k = 0; k < n;) { if (c[k].cnt == NO_CHILD) { c[k] = c[n - 1]; --n; } else ++k; }
        while (n < 4) {
            int best = -1;
            for (int k = 0; k < n; ++k)
                if (c[k].cnt == 0 && (best < 0 || bvh4_area(c[k]) > bvh4_area(c[best]))) best = k;
            if (best < 0) break;
            Bvh4Child g[2];
            bvh2_children(nodes16, c[best].ref, g);
            int m = 0;
            Bvh4Child keep[2];
            for (int k = 0; k < 2; ++k) if (g[k].cnt != NO_CHILD) keep[m++] = g[k];
            if (m == 0) { c[best] = c[n - 1]; --n; continue; }
            c[best] = keep[0];
            if (m == 2) c[n++] = keep[1];
        }
        const size_t base = out.size();
        out.resize(base + 32, 0u);
        float f[24];
        uint32_t child[4];
        for (int k = 0; k < 4; ++k) {
            const bool have = k < n;
            for (int a = 0; a < 3; ++a) {
                f[a * 4 + k] = have ? c[k].lo[a] : 0.0f;
                f[12 + a * 4 + k] = have ? c[k].hi[a] : 0.0f;
            }
            if (have && (c[k].cnt > MAX_LEAF || c[k].ref > MAX_REF)) throw std::length_error("bvh4: leaf larger than 7 triangles or index beyond 2^29");
            child[k] = have ? ((c[k].ref << 3) | c[k].cnt) : NO_CHILD;          // an inner child's index is patched in when it is emitted
        }
        std::memcpy(out.data() + base, f, sizeof f);
        std::memcpy(out.data() + base + 24, child, sizeof child);
        // a visit of a node with n children leaves n - 1 entries below the child it descends into (bvh4_stack_need)
        const uint32_t below = t.sp + uint32_t(n ? n - 1 : 0);
        need = std::max(need, below);
        // inner children: pushed in reverse so that child 0 comes next
        for (int k = n - 1; k >= 0; --k)
            if (c[k].cnt == 0) todo.push_back({c[k].ref, uint32_t(base + 24 + k), below});
    }
    if (stack_need) *stack_need = need;
    return out;
}

// Stack entries the traversal of csrc/rt_bvh4.cuh can need on this tree, whatever the ray: a visit of a node with n children
// leaves n - 1 entries below the child it descends into, and any child may be the one visited first.
inline uint32_t bvh4_stack_need(const std::vector<uint32_t>& nodes4) {
    if (nodes4.empty()) return 0;
    struct Item { uint32_t node, sp; };
    std::vector<Item> todo{{0u, 0u}};
    uint32_t need = 0;
    while (!todo.empty()) {
        const Item it = todo.back();
        todo.pop_back();
        const uint32_t* child = nodes4.data() + size_t(it.node) * 32 + 24;
        uint32_t n = 0;
        for (int k = 0; k < 4; ++k) n += child[k] != 0xFFFFFFFFu;
        const uint32_t below = it.sp + (n ? n - 1 : 0);
        need = std::max(need, below);
        for (int k = 0; k < 4; ++k)
            if (child[k] != 0xFFFFFFFFu && (child[k] & 7u) == 0u) todo.push_back({child[k] >> 3, below});
    }
    return need;
}

}  // namespace rtb
