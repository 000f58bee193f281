// voronoids.hpp -- header-only C++ mirror of the reference's Rust API for the insertion path, over the C ABI of
// voronoids_b200.h.  Same names and argument meaning as /root/reference/src/delaunay_tree.rs so that code (and tests)
// written against the crate read the same:
//
//   Rust (reference)                                         C++ (here)
//   DelaunayTree::<3,4>::new(vertices)            :390       voronoids::DelaunayTree<3,4>::make(vertices)   ("new" is a C++ keyword)
//   tree.add_points_to_tree(vertices)             :336       tree.add_points_to_tree(vertices)
//   let u = TreeUpdate::new(id, p, &tree);
//   tree.insert_point(&u)                         :710,:125  tree.insert_point(p)
//   tree.insert_points_parallel(&updates)         :213       tree.insert_points_parallel(points)
//   tree.check_delaunay()                         :512       tree.check_delaunay()
//   tree.max_simplex_id / vertices.len()          :26-29     tree.max_simplex_id() / tree.n_vertices()
//   geometry::circumsphere / in_sphere / bounding_sphere   geometry.rs:58,91,99    voronoids::geometry::*
// A Rust panic is a voronoids::Error exception carrying the vor_status.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "voronoids_b200.h"

namespace voronoids {

struct Error : std::runtime_error {
    vor_status status;
    Error(vor_status s, const char *what) : std::runtime_error(what), status(s) {}
};
inline void check(vor_status s) {
    if (s != VOR_OK && s != VOR_ERR_DUPLICATE_POINT) throw Error(s, vor_last_error());
}

template <size_t N, size_t M> class DelaunayTree {
    static_assert(M == N + 1 && (N == 2 || N == 3), "DelaunayTree<2,3> or DelaunayTree<3,4>");
    vor_tree *h_ = nullptr;

  public:
    using Point = std::array<double, N>;
    // DelaunayTree::new: bounding sphere of `vertices`, 10x super simplex, nothing inserted (delaunay_tree.rs:390 / :545)
    static DelaunayTree make(const std::vector<Point> &vertices, int device = 0) {
        DelaunayTree t;
        check(vor_tree_create((int)N, vertices.empty() ? nullptr : vertices[0].data(), vertices.size(), device, &t.h_));
        return t;
    }
    DelaunayTree() = default;
    DelaunayTree(DelaunayTree &&o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    DelaunayTree &operator=(DelaunayTree &&o) noexcept { std::swap(h_, o.h_); return *this; }
    DelaunayTree(const DelaunayTree &) = delete;
    ~DelaunayTree() { vor_tree_destroy(h_); }

    void add_points_to_tree(const std::vector<Point> &vertices) {
        check(vor_tree_insert(h_, vertices.empty() ? nullptr : vertices[0].data(), vertices.size(), VOR_INSERT_PARALLEL));
    }
    void insert_points_parallel(const std::vector<Point> &vertices) { add_points_to_tree(vertices); }
    void insert_point(const Point &p) { check(vor_tree_insert(h_, p.data(), 1, VOR_INSERT_SINGLE)); }
    bool check_delaunay() {
        int ok = 0;
        check(vor_tree_check_delaunay(h_, &ok, nullptr));
        return ok != 0;
    }
    uint64_t max_simplex_id() {
        uint64_t m = 0;
        check(vor_tree_counts(h_, nullptr, nullptr, &m));
        return m;
    }
    uint64_t n_vertices() {
        uint64_t v = 0;
        check(vor_tree_counts(h_, &v, nullptr, nullptr));
        return v;
    }
    // canonical Delaunay graph: sorted unique (lo, hi) input-index pairs
    std::vector<std::array<uint32_t, 2>> edges() {
        size_t n = 0;
        check(vor_tree_edges(h_, nullptr, 0, &n));
        std::vector<std::array<uint32_t, 2>> e(n);
        check(vor_tree_edges(h_, n ? e[0].data() : nullptr, n, &n));
        return e;
    }
    vor_tree *handle() { return h_; }
};

namespace geometry {
template <size_t N, size_t M> std::pair<std::array<double, N>, double> circumsphere(const std::array<std::array<double, N>, M> &v, int device = 0) {
    std::array<double, N> c{};
    double r = 0;
    check(vor_circumsphere((int)N, v[0].data(), 1, c.data(), &r, device));
    return {c, r};
}
template <size_t N> bool in_sphere(const std::array<double, N> &vertex, const std::array<double, N> &center, double radius, int device = 0) {
    int32_t out = 0;
    check(vor_in_sphere((int)N, vertex.data(), center.data(), &radius, 1, &out, device));
    return out != 0;
}
template <size_t N> std::pair<std::array<double, N>, double> bounding_sphere(const std::vector<std::array<double, N>> &pts, int device = 0) {
    std::array<double, N> c{};
    double r = 0;
    check(vor_bounding_sphere((int)N, pts[0].data(), pts.size(), c.data(), &r, device));
    return {c, r};
}
} // namespace geometry
} // namespace voronoids
