// A program written against include/voronoids.hpp the way one is written against the crate
// (/root/reference/examples/parallel_insert.rs:7-33, tests/test_delaunay_tree.rs:6-38): DelaunayTree::new on all points,
// a few single inserts, the rest through add_points_to_tree, check_delaunay, then the edge list.
// usage: hpp_smoke points.bin n  ->  prints "max0 <id> vertices <n> ok <0|1> edges <m> sum <lo-sum> <hi-sum>"
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "voronoids.hpp"

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    const size_t n = (size_t)atol(argv[2]);
    std::vector<std::array<double, 3>> pts(n);
    FILE *f = fopen(argv[1], "rb");
    if (!f || fread(pts.data(), sizeof(double) * 3, n, f) != n) return 3;
    fclose(f);
    try {
        auto tree = voronoids::DelaunayTree<3, 4>::make(pts);
        const unsigned long long max0 = tree.max_simplex_id();          // == 4, tests/test_delaunay_tree.rs:20
        const size_t n_seq = n < 100 ? n : 100;
        for (size_t i = 0; i < n_seq; i++) tree.insert_point(pts[i]);   // tests/test_delaunay_tree.rs:23-26
        tree.add_points_to_tree(std::vector<std::array<double, 3>>(pts.begin() + n_seq, pts.end()));
        const bool ok = tree.check_delaunay();
        const auto e = tree.edges();
        unsigned long long slo = 0, shi = 0;
        for (const auto &x : e) { slo += x[0]; shi += x[1]; }
        printf("max0 %llu vertices %llu ok %d edges %zu sum %llu %llu\n", max0, (unsigned long long)tree.n_vertices(), ok ? 1 : 0, e.size(), slo, shi);
        const auto cs = voronoids::geometry::circumsphere<3, 4>({{{1, 0, 0}, {0, 0, 0}, {0, 1, 0}, {0, 0, 1}}});
        printf("circumsphere %.17g %.17g %.17g %.17g\n", cs.first[0], cs.first[1], cs.first[2], cs.second);   // tests/test_geometry.rs:5-15
    } catch (const voronoids::Error &err) {
        printf("error %d %s\n", (int)err.status, err.what());
        return 1;
    }
    return 0;
}
