// ORACLE — TEST INFRASTRUCTURE ONLY (see sf_math.hpp header).
// Restates /root/reference/src/physics/collision/bvh.rs: incremental binary AABB tree
// (insert :105-169, test_aabb :171-190 + :299-336, test_point :192-209 + :355-391,
// sweep_aabb :211-235 + :418-513). Needed so the oracle reproduces the reference's candidate
// pair ORDER (which fixes the DFS / Gauss-Seidel order of the sequential reference mode) and
// its best-first ray traversal, not just the pair set.
#pragma once
#include <queue>
#include <vector>
#include "sf_narrow.hpp"

namespace sfo {

template <class R>
class Bvh {
    struct Node { AABB<R> aabb; bool leaf; uint32_t left, right, coll; };
    std::vector<Node> nodes_;

public:
    void clear() { nodes_.clear(); }
    bool empty() const { return nodes_.empty(); }

    void insert(uint32_t coll, const AABB<R>& aabb) {
        Node nn{aabb, true, 0, 0, coll};
        if (nodes_.empty()) { nodes_.push_back(nn); return; }
        size_t cur = 0;
        for (;;) {
            Node cn = nodes_[cur];
            if (!cn.leaf) {
                AABB<R> lu = aabb.union_(nodes_[cn.left].aabb);
                AABB<R> ru = aabb.union_(nodes_[cn.right].aabb);
                if (lu.area() <= ru.area()) {
                    nodes_[cur].aabb = lu.union_(cn.aabb);
                    if (!nodes_[cn.left].leaf) nodes_[cn.left].aabb = lu;
                    cur = cn.left;
                } else {
                    nodes_[cur].aabb = ru.union_(cn.aabb);
                    if (!nodes_[cn.right].leaf) nodes_[cn.right].aabb = ru;
                    cur = cn.right;
                }
                if (cur == 0) nodes_[cur].aabb = cn.aabb.union_(nn.aabb);  // unreachable in practice; kept
            } else {
                nodes_.push_back(cn);
                nodes_.push_back(nn);
                nodes_[cur] = Node{cn.aabb.union_(nn.aabb), false, uint32_t(nodes_.size() - 2),
                                   uint32_t(nodes_.size() - 1), 0};
                return;
            }
        }
    }

    // Calls f(coll) for every leaf reached, in the reference's traversal order.
    template <class F> void test_aabb(const AABB<R>& q, F&& f) const {
        if (nodes_.empty()) return;
        if (nodes_.size() == 1) { if (nodes_[0].aabb.intersects(q)) f(nodes_[0].coll); return; }
        std::vector<uint32_t> stack;
        long next = 0;
        while (next >= 0) {
            const Node& n = nodes_[size_t(next)];
            if (!n.leaf) {
                bool l = q.intersects(nodes_[n.left].aabb), r = q.intersects(nodes_[n.right].aabb);
                if (l && r) { stack.push_back(n.right); next = n.left; }
                else if (l) next = n.left;
                else if (r) next = n.right;
                else if (stack.empty()) next = -1;
                else { next = stack.back(); stack.pop_back(); }
            } else {
                uint32_t c = n.coll;
                if (stack.empty()) next = -1; else { next = stack.back(); stack.pop_back(); }
                f(c);
            }
        }
    }

    template <class F> void test_point(Vec2<R> p, F&& f) const {
        if (nodes_.empty()) return;
        if (nodes_.size() == 1) { if (nodes_[0].aabb.contains_point(p)) f(nodes_[0].coll); return; }
        std::vector<uint32_t> stack;
        long next = 0;
        while (next >= 0) {
            const Node& n = nodes_[size_t(next)];
            if (!n.leaf) {
                bool l = nodes_[n.left].aabb.contains_point(p), r = nodes_[n.right].aabb.contains_point(p);
                if (l && r) { stack.push_back(n.right); next = n.left; }
                else if (l) next = n.left;
                else if (r) next = n.right;
                else if (stack.empty()) next = -1;
                else { next = stack.back(); stack.pop_back(); }
            } else {
                uint32_t c = n.coll;
                if (stack.empty()) next = -1; else { next = stack.back(); stack.pop_back(); }
                f(c);
            }
        }
    }

    // Best-first sweep. f(t, coll) returns false to stop (spherecast's early return).
    template <class F> void sweep_aabb(R half, const Ray<R>& ray, R max_t, F&& f) const {
        struct Ent { uint32_t node; R t; uint64_t seq; };
        // BinaryHeap with reversed ordering on t = min-heap on t. Tie order inside Rust's
        // BinaryHeap is unspecified; we break ties by insertion sequence (deterministic).
        auto cmp = [](const Ent& a, const Ent& b) { return a.t > b.t || (a.t == b.t && a.seq > b.seq); };
        std::priority_queue<Ent, std::vector<Ent>, decltype(cmp)> heap(cmp);
        uint64_t seq = 0;
        bool have = false;
        Ent next{};
        if (nodes_.empty()) return;
        if (nodes_.size() == 1) {
            R t;
            if (ray_aabb(ray, nodes_[0].aabb.padded(half), &t) && t <= max_t) { next = {0, t, 0}; have = true; }
        } else { next = {0, R(0), 0}; have = true; }
        auto pop = [&]() { if (heap.empty()) have = false; else { next = heap.top(); heap.pop(); have = true; } };
        auto choose = [&](uint32_t idx, R t) {
            if (!heap.empty() && heap.top().t < t) { Ent top = heap.top(); heap.pop(); heap.push({idx, t, seq++}); next = top; }
            else next = {idx, t, 0};
            have = true;
        };
        while (have) {
            const Node& n = nodes_[next.node];
            if (!n.leaf) {
                R tl = R(0), tr = R(0);      // set by ray_aabb whenever it reports a hit
                bool hl = ray_aabb(ray, nodes_[n.left].aabb.padded(half), &tl);
                bool hr = ray_aabb(ray, nodes_[n.right].aabb.padded(half), &tr);
                if (hl && hr) {
                    R ct, ft; uint32_t ci, fi;
                    if (tl <= tr) { ct = tl; ci = n.left; ft = tr; fi = n.right; }
                    else { ct = tr; ci = n.right; ft = tl; fi = n.left; }
                    if (ct >= max_t) { pop(); continue; }
                    if (ft < max_t) heap.push({fi, ft, seq++});
                    choose(ci, ct);
                } else if (hl) {
                    if (tl >= max_t) { pop(); continue; }
                    choose(n.left, tl);
                } else if (hr) {
                    if (tr >= max_t) { pop(); continue; }
                    choose(n.right, tr);
                } else pop();
            } else {
                R t = next.t; uint32_t c = n.coll;
                pop();
                if (!f(t, c)) return;
            }
        }
    }
};

}  // namespace sfo
