// host/kd_parallel.hpp - task-parallel driver shared by the top-down builders.
//
// Both builders (kd_build.cpp: the reference's median kd-tree, kd_tree_simd.hpp:146-185; bvh_build.cpp: the backend's own
// bounding-volume hierarchy) decide a node's split from the node's own triangle list only, so subtrees are independent and the tree is a pure
// function of the input.  The driver exploits that without changing a single node:
//
//   1. the top of the tree is expanded by recursive tasks (child1 on a new thread, child0 inline) until a subtree holds
//      at most `cutoff` references; every such subtree is built by the sequential loop into a LOCAL KdTree;
//   2. the pieces are then laid out in DFS pre-order - a subtree occupies one contiguous index range, child0 == index+1,
//      leaf lists in leaf creation order, exactly the numbering of the reference's recursion (kd_tree_simd.hpp:172-184)
//      and of the sequential loop - and copied into place in parallel with their index offsets.
//
// With one thread the driver degenerates to the sequential loop; tests/test_host.py checks that 1 and N threads give
// byte-identical trees and that both equal the oracle's restatement of the reference builder.
//
// Policy concept:
//   struct Work { uint64_t depth; float lo[3], hi[3]; size_t size() const; bool empty() const; ... };
//   bool expand(Work& w, uint32_t& axis, float& split, Work& c0, Work& c1) const;   false: w becomes a leaf (w untouched)
//   void emit(const Work& w, std::vector<uint32_t>& refs) const;                     append the leaf's triangle ids
#pragma once

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#include "kd_build.hpp"

namespace rtb {

inline unsigned build_threads() {
    if (const char* e = std::getenv("RT_B200_BUILD_THREADS")) {
        const int v = std::atoi(e);
        if (v > 0) return unsigned(v);
    }
    const unsigned hc = std::thread::hardware_concurrency();
    return hc ? hc : 1u;
}

// fn(begin, end) over [0, n) in contiguous chunks, one per thread
template <class F>
void parallel_for(uint64_t n, uint64_t min_chunk, F&& fn) {
    const uint64_t threads = std::max<uint64_t>(1, std::min<uint64_t>(build_threads(), n / std::max<uint64_t>(min_chunk, 1)));
    if (threads <= 1) { fn(uint64_t(0), n); return; }
    // an exception in a worker (bad_alloc on a huge scene) must reach the caller's guarded() block, not std::terminate
    std::vector<std::thread> pool;
    std::exception_ptr failed;
    std::mutex failed_mutex;
    auto guarded_fn = [&](uint64_t b, uint64_t e) {
        try { fn(b, e); }
        catch (...) { std::lock_guard<std::mutex> g(failed_mutex); if (!failed) failed = std::current_exception(); }
    };
    const uint64_t step = (n + threads - 1) / threads;
    for (uint64_t b = step; b < n; b += step) pool.emplace_back([&guarded_fn, b, step, n] { guarded_fn(b, std::min(n, b + step)); });
    guarded_fn(uint64_t(0), std::min(n, step));
    for (auto& th : pool) th.join();
    if (failed) std::rethrow_exception(failed);
}

// the sequential loop: builds the subtree rooted at `root` into t (local indices, root parent = KD_NONE)
template <class Policy>
void kd_build_subtree(const Policy& pol, typename Policy::Work&& root, KdTree& t) {
    using Work = typename Policy::Work;
    struct Item { Work w; uint64_t parent; int which; };
    std::vector<Item> todo;
    todo.push_back(Item{std::move(root), KD_NONE, 0});
    // A node gets its index when it is taken off the stack; child1 is stacked below child0, so the whole child0 subtree
    // is numbered first: the reference's recursion order (kd_tree_simd.hpp:172-184).
    while (!todo.empty()) {
        Item it = std::move(todo.back());
        todo.pop_back();
        const uint64_t idx = t.nodes.size();
        KdNode n{};
        n.parent = it.parent; n.child0 = n.child1 = n.first_ref = KD_NONE; n.ref_count = 0;
        std::memcpy(n.bmin, it.w.lo, 12); std::memcpy(n.bmax, it.w.hi, 12);
        n.axis = 3; n.split = 0.0f;
        t.nodes.push_back(n);
        if (it.parent != KD_NONE) (it.which ? t.nodes[it.parent].child1 : t.nodes[it.parent].child0) = idx;
        if (it.w.depth > t.depth) t.depth = it.w.depth;
        Work c0, c1;
        uint32_t axis = 3; float split = 0.0f;
        if (!pol.expand(it.w, axis, split, c0, c1)) {
            t.nodes[idx].first_ref = t.refs.size();
            t.nodes[idx].ref_count = it.w.size();
            pol.emit(it.w, t.refs);
            ++t.n_leaves;
            if (it.w.size() > t.max_leaf_refs) t.max_leaf_refs = it.w.size();
            continue;
        }
        t.nodes[idx].axis = axis; t.nodes[idx].split = split;
        { Work drop = std::move(it.w); (void)drop; }                       // the parent's list is no longer needed
        if (!c1.empty()) todo.push_back(Item{std::move(c1), idx, 1});      // empty children are never created
        if (!c0.empty()) todo.push_back(Item{std::move(c0), idx, 0});
    }
}

template <class Policy>
struct KdPiece {
    bool is_sub = false;
    KdTree sub;                                   // is_sub: a finished local subtree
    KdNode node{};                                // otherwise: one inner node of the top of the tree
    std::unique_ptr<KdPiece> c[2];
    uint64_t n_nodes = 0, n_refs = 0;
};

template <class Policy>
void kd_build_piece(const Policy& pol, typename Policy::Work&& w, KdPiece<Policy>& p, size_t cutoff, int spawn_levels) {
    using Work = typename Policy::Work;
    Work c0, c1;
    uint32_t axis = 3; float split = 0.0f;
    if (w.size() <= cutoff || !pol.expand(w, axis, split, c0, c1)) {
        p.is_sub = true;
        kd_build_subtree(pol, std::move(w), p.sub);
        p.n_nodes = p.sub.nodes.size(); p.n_refs = p.sub.refs.size();
        return;
    }
    p.node.parent = KD_NONE; p.node.child0 = p.node.child1 = p.node.first_ref = KD_NONE; p.node.ref_count = 0;
    std::memcpy(p.node.bmin, w.lo, 12); std::memcpy(p.node.bmax, w.hi, 12);
    p.node.axis = axis; p.node.split = split;
    const uint64_t depth = w.depth;
    { Work drop = std::move(w); (void)drop; }
    std::thread other;
    std::exception_ptr other_failed;                  // an exception on the spawned thread is rethrown here, after the join
    if (!c1.empty()) {
        p.c[1] = std::make_unique<KdPiece<Policy>>();
        if (spawn_levels > 0 && !c0.empty())
            other = std::thread([&pol, &p, &c1, &other_failed, cutoff, spawn_levels] {
                try { kd_build_piece(pol, std::move(c1), *p.c[1], cutoff, spawn_levels - 1); }
                catch (...) { other_failed = std::current_exception(); }
            });
    }
    try {
        if (!c0.empty()) {
            p.c[0] = std::make_unique<KdPiece<Policy>>();
            kd_build_piece(pol, std::move(c0), *p.c[0], cutoff, spawn_levels - 1);
        }
    } catch (...) {
        if (other.joinable()) other.join();
        throw;
    }
    if (other.joinable()) { other.join(); if (other_failed) std::rethrow_exception(other_failed); }
    else if (p.c[1]) kd_build_piece(pol, std::move(c1), *p.c[1], cutoff, spawn_levels - 1);
    p.n_nodes = 1 + (p.c[0] ? p.c[0]->n_nodes : 0) + (p.c[1] ? p.c[1]->n_nodes : 0);
    p.n_refs = (p.c[0] ? p.c[0]->n_refs : 0) + (p.c[1] ? p.c[1]->n_refs : 0);
    p.sub.depth = depth;
}

template <class Policy>
KdTree kd_build_parallel(const Policy& pol, typename Policy::Work&& root) {
    const unsigned threads = build_threads();
    const size_t total = root.size();
    if (threads <= 1 || total < 8192) {
        KdTree t;
        kd_build_subtree(pol, std::move(root), t);
        return t;
    }
    int spawn_levels = 0;
    while ((1u << spawn_levels) < threads) ++spawn_levels;
    spawn_levels += 2;                                                     // about four tasks per thread: SAH subtrees are uneven
    const size_t cutoff = std::max<size_t>(total / (size_t(threads) * 16), 2048);
    KdPiece<Policy> top;
    kd_build_piece(pol, std::move(root), top, cutoff, spawn_levels);

    KdTree t;
    t.nodes.resize(top.n_nodes);
    t.refs.resize(top.n_refs);
    struct Place { KdPiece<Policy>* p; uint64_t node_off, ref_off, parent; };
    std::vector<Place> subs;
    // DFS pre-order placement: node, then the whole child0 subtree, then child1
    struct Frame { KdPiece<Policy>* p; uint64_t parent; int which; };
    std::vector<Frame> stack;
    stack.push_back(Frame{&top, KD_NONE, 0});
    uint64_t next_node = 0, next_ref = 0;
    while (!stack.empty()) {
        const Frame f = stack.back();
        stack.pop_back();
        const uint64_t idx = next_node;
        if (f.parent != KD_NONE) (f.which ? t.nodes[f.parent].child1 : t.nodes[f.parent].child0) = idx;
        if (f.p->is_sub) {
            subs.push_back(Place{f.p, next_node, next_ref, f.parent});
            next_node += f.p->n_nodes; next_ref += f.p->n_refs;
            t.n_leaves += f.p->sub.n_leaves;
            t.max_leaf_refs = std::max(t.max_leaf_refs, f.p->sub.max_leaf_refs);
            t.depth = std::max(t.depth, f.p->sub.depth);
            continue;
        }
        t.nodes[idx] = f.p->node;
        t.nodes[idx].parent = f.parent;
        t.depth = std::max(t.depth, f.p->sub.depth);
        ++next_node;
        if (f.p->c[1]) stack.push_back(Frame{f.p->c[1].get(), idx, 1});
        if (f.p->c[0]) stack.push_back(Frame{f.p->c[0].get(), idx, 0});
    }
    parallel_for(subs.size(), 1, [&](uint64_t b, uint64_t e) {
        for (uint64_t k = b; k < e; ++k) {
            const Place& pl = subs[k];
            KdTree& s = pl.p->sub;
            for (uint64_t i = 0; i < s.nodes.size(); ++i) {
                KdNode n = s.nodes[i];
                n.parent = (n.parent == KD_NONE) ? pl.parent : n.parent + pl.node_off;
                if (n.child0 != KD_NONE) n.child0 += pl.node_off;
                if (n.child1 != KD_NONE) n.child1 += pl.node_off;
                if (n.first_ref != KD_NONE) n.first_ref += pl.ref_off;
                t.nodes[pl.node_off + i] = n;
            }
            if (!s.refs.empty()) std::memcpy(t.refs.data() + pl.ref_off, s.refs.data(), s.refs.size() * sizeof(uint32_t));
            std::vector<KdNode>().swap(s.nodes);
            std::vector<uint32_t>().swap(s.refs);
        }
    });
    return t;
}

}  // namespace rtb
