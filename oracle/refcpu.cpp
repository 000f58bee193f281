/*
 * oracle/refcpu.cpp -- TEST INFRASTRUCTURE and the timed CPU baseline
 * ("port" of kazewong/Voronoids; NOT the upstream Rust binary, which cannot be
 * built in this image: no cargo/rustc).
 *
 * Restates, step for step and with the reference's own floating-point tests:
 *   DelaunayTree::new            /root/reference/src/delaunay_tree.rs:390-510 (3D), :545-640 (2D)
 *   locate + find_all_neighbors  /root/reference/src/delaunay_tree.rs:33-75
 *   get_new_simplices            /root/reference/src/delaunay_tree.rs:77-123
 *   pair_simplices               /root/reference/src/delaunay_tree.rs:674-695
 *   TreeUpdate::new              /root/reference/src/delaunay_tree.rs:710-739
 *   insert_point                 /root/reference/src/delaunay_tree.rs:125-211
 *   insert_points_parallel       /root/reference/src/delaunay_tree.rs:213-334
 *   add_points_to_tree           /root/reference/src/delaunay_tree.rs:336-386
 *   make_queue, find_placement   /root/reference/src/scheduler.rs:6-55
 *   delaunay (python entry)      /root/reference/src/lib.rs:104-125
 *   check_delaunay               /root/reference/src/delaunay_tree.rs:512-541, :642-671
 *
 * Deliberate differences (all make this baseline FASTER than the original, so
 * a speed-up quoted against it is conservative):
 *   - DashMap<usize,_> (SipHash + shard RwLock per access) -> dense std::vector
 *     indexed by id (ids are never reused in the reference either), per-vertex
 *     spin locks only where threads really share a record (Vertex.simplex).
 *   - kiddo 4.2.0 kd-tree (not under /root/reference) -> a bucketed kd-tree
 *     (bucket 32, median split on overflow), nearest_one with squared
 *     Euclidean distance, ties resolved to the earliest-added item.
 *   - rayon par_iter -> OpenMP parallel for (dynamic schedule).
 * A Rust panic (`unwrap` on a missing simplex, "No simplex found", singular LU)
 * becomes err != 0 and the run stops.
 */
#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <unordered_map>
#include <vector>
#include <omp.h>

#include "ref_geometry.h"

namespace {

template <int N> struct KdTree {
    // bucketed kd-tree, incremental add; items keep insertion order inside a bucket
    struct Node {
        int dim = -1; // -1 leaf
        double split = 0;
        int left = -1, right = -1;
        std::vector<std::pair<std::array<double, N>, uint64_t>> items;
    };
    std::vector<Node> nodes;
    static constexpr size_t B = 32;
    KdTree() { nodes.emplace_back(); }
    void add(const std::array<double, N> &p, uint64_t item) {
        int cur = 0;
        while (nodes[cur].dim >= 0) cur = p[nodes[cur].dim] < nodes[cur].split ? nodes[cur].left : nodes[cur].right;
        nodes[cur].items.emplace_back(p, item);
        if (nodes[cur].items.size() > B) split(cur);
    }
    void split(int cur) {
        auto &it = nodes[cur].items;
        int best = -1;
        double bestext = 0;
        std::array<double, N> lo, hi;
        for (int d = 0; d < N; d++) { lo[d] = INFINITY; hi[d] = -INFINITY; }
        for (auto &e : it)
            for (int d = 0; d < N; d++) { lo[d] = std::min(lo[d], e.first[d]); hi[d] = std::max(hi[d], e.first[d]); }
        for (int d = 0; d < N; d++)
            if (hi[d] - lo[d] > bestext) { bestext = hi[d] - lo[d]; best = d; }
        if (best < 0) return; // all coincident: keep an oversize bucket
        std::vector<double> vals;
        for (auto &e : it) vals.push_back(e.first[best]);
        std::nth_element(vals.begin(), vals.begin() + vals.size() / 2, vals.end());
        double s = vals[vals.size() / 2];
        if (s == lo[best]) { // make sure the left side is non-empty
            double nxt = INFINITY;
            for (double v : vals) if (v > s) nxt = std::min(nxt, v);
            s = nxt;
        }
        Node l, r;
        for (auto &e : it) (e.first[best] < s ? l : r).items.push_back(e);
        int li = (int)nodes.size();
        nodes.push_back(std::move(l));
        nodes.push_back(std::move(r));
        nodes[cur].dim = best;
        nodes[cur].split = s;
        nodes[cur].left = li;
        nodes[cur].right = li + 1;
        nodes[cur].items.clear();
        nodes[cur].items.shrink_to_fit();
    }
    void search(int cur, const std::array<double, N> &p, double &bd, uint64_t &bi) const {
        const Node &nd = nodes[cur];
        if (nd.dim < 0) {
            for (auto &e : nd.items) {
                double d = 0;
                for (int k = 0; k < N; k++) d += (e.first[k] - p[k]) * (e.first[k] - p[k]);
                if (d < bd || (d == bd && e.second < bi)) { bd = d; bi = e.second; }
            }
            return;
        }
        double diff = p[nd.dim] - nd.split;
        int near = diff < 0 ? nd.left : nd.right, far = diff < 0 ? nd.right : nd.left;
        search(near, p, bd, bi);
        if (diff * diff <= bd) search(far, p, bd, bi);
    }
    uint64_t nearest_one(const std::array<double, N> &p) const {
        double bd = INFINITY;
        uint64_t bi = ~0ull;
        search(0, p, bd, bi);
        return bi;
    }
};

struct SpinLock {
    std::atomic_flag f = ATOMIC_FLAG_INIT;
    void lock() { while (f.test_and_set(std::memory_order_acquire)) {} }
    void unlock() { f.clear(std::memory_order_release); }
};

template <int N> struct Tree {
    static constexpr int M = N + 1;
    using Pt = std::array<double, N>;
    struct Simplex {
        std::array<size_t, M> vertices;
        Pt center;
        double radius;
        std::vector<size_t> neighbors;
        bool alive = false;
    };
    struct Vertex {
        Pt coordinates;
        std::vector<size_t> simplex;
    };
    struct Update {
        Pt vertex;
        std::vector<size_t> killed_sites;
        std::vector<std::array<size_t, M>> simplices;
        std::vector<size_t> simplices_id;
        std::vector<Pt> centers;
        std::vector<double> radii;
        std::vector<std::pair<size_t, size_t>> neighbors;
        std::vector<std::pair<size_t, size_t>> new_neighbors;
    };

    KdTree<N> kdtree;
    std::vector<Vertex> vertices;       // key = index (ids are dense: vertices.len())
    std::vector<Simplex> simplices;     // key = id; !alive == removed from the map
    std::unique_ptr<SpinLock[]> vlocks;
    size_t vlock_n = 0;
    size_t max_simplex_id = 0;
    std::vector<int64_t> input_index;   // per vertex id: index of the input point, -1 for super/ghost
    std::atomic<int> err{0};            // 1 no simplex found, 2 singular LU, 3 missing simplex
    size_t rounds = 0;
    size_t n_live = 0;

    void ensure_simplices(size_t n) {
        if (simplices.size() < n) simplices.resize(std::max(n, simplices.size() * 2));
    }
    void ensure_vlocks(size_t n) {
        if (vlock_n >= n) return;
        size_t nn = std::max(n, vlock_n * 2);
        vlocks.reset(new SpinLock[nn]);
        vlock_n = nn;
    }

    // ---- DelaunayTree::new (delaunay_tree.rs:390-510 / :545-640)
    explicit Tree(const double *pts, long n) {
        double sup[M * N], c[N], r;
        ref_super_simplex(N, pts, n, sup, c, &r);
        const int nsv = 2 * M; // super + ghost vertices
        std::vector<Pt> vs(nsv);
        for (int i = 0; i < M; i++)
            for (int k = 0; k < N; k++) vs[i][k] = sup[i * N + k];
        if (N == 3) {
            vs[4] = vs[0]; vs[5] = vs[0]; vs[6] = vs[0]; vs[7] = vs[1]; // delaunay_tree.rs:407-412
        } else {
            vs[3] = vs[0]; vs[4] = vs[1]; vs[5] = vs[2];               // delaunay_tree.rs:559-566
        }
        for (int i = 0; i < nsv; i++) kdtree.add(vs[i], (uint64_t)i);
        static const std::vector<std::vector<size_t>> vs3 = {{0, 1, 2, 3}, {0, 1, 3, 4}, {0, 1, 2, 4}, {0, 2, 3, 4}, {1}, {2}, {3}, {4}};
        static const std::vector<std::vector<size_t>> vs2 = {{0, 1, 2}, {0, 1, 3}, {0, 2, 3}, {1}, {2}, {3}};
        vertices.resize(nsv);
        for (int i = 0; i < nsv; i++) {
            vertices[i].coordinates = vs[i];
            vertices[i].simplex = N == 3 ? vs3[i] : vs2[i];
        }
        input_index.assign(nsv, -1);
        ensure_simplices(64);
        Pt zero{};
        Pt cc;
        for (int k = 0; k < N; k++) cc[k] = c[k];
        auto put = [&](size_t id, std::array<size_t, M> v, Pt ce, double ra, std::vector<size_t> nb) {
            simplices[id].vertices = v; simplices[id].center = ce; simplices[id].radius = ra;
            simplices[id].neighbors = nb; simplices[id].alive = true;
        };
        if constexpr (N == 3) {
            put(0, {0, 1, 2, 3}, cc, r, {1, 2, 3, 4});
            put(1, {4, 0, 1, 2}, zero, 0., {0});
            put(2, {5, 0, 2, 3}, zero, 0., {0});
            put(3, {6, 0, 3, 1}, zero, 0., {0});
            put(4, {7, 1, 2, 3}, zero, 0., {0});
            max_simplex_id = 4;
        } else {
            put(0, {0, 1, 2}, cc, r, {1, 2, 3});
            put(1, {3, 0, 1}, zero, 0., {0});
            put(2, {4, 0, 2}, zero, 0., {0});
            put(3, {5, 1, 2}, zero, 0., {0});
            max_simplex_id = 3;
        }
        n_live = M + 1;
    }

    bool in_sphere(const Pt &v, const Simplex &s) const { return ref_in_sphere(N, v.data(), s.center.data(), s.radius); }

    // ---- locate (delaunay_tree.rs:33-58)
    void find_all_neighbors(std::vector<size_t> &out, size_t node, const Pt &v) const {
        // recursion of delaunay_tree.rs:60-75 (the to_vec() clones are not restated)
        const std::vector<size_t> &nb = simplices[node].neighbors;
        for (size_t k = 0; k < nb.size(); k++) {
            size_t id = nb[k];
            const Simplex &s = simplices[id];
            if (std::find(out.begin(), out.end(), id) == out.end() && in_sphere(v, s)) {
                out.push_back(id);
                find_all_neighbors(out, id, v);
            }
        }
    }
    std::vector<size_t> locate(const Pt &v) {
        std::vector<size_t> out;
        size_t nn = (size_t)kdtree.nearest_one(v);
        const std::vector<size_t> &inc = vertices[nn].simplex;
        for (size_t id : inc) {
            if (!simplices[id].alive) { err = 3; return out; }
            if (in_sphere(v, simplices[id])) {
                out.push_back(id);
                find_all_neighbors(out, id, v);
            }
        }
        if (out.empty()) { err = 1; return out; } // panic!("No simplex found ...")
        std::sort(out.begin(), out.end());
        out.erase(std::unique(out.begin(), out.end()), out.end());
        return out;
    }

    // ---- get_new_simplices (delaunay_tree.rs:77-123)
    void get_new_simplices(size_t killed_id, const Pt &v, size_t vid, Update &u) {
        const Simplex &ks = simplices[killed_id];
        for (size_t nbid : ks.neighbors) {
            const Simplex &ns = simplices[nbid];
            if (!in_sphere(v, ns)) {
                std::array<size_t, M> nsx{};
                nsx[0] = vid;
                int count = 1;
                for (int i = 0; i < M; i++) {
                    bool has = false;
                    for (int k = 0; k < M; k++) has |= ks.vertices[k] == ns.vertices[i];
                    if (has && count < M) nsx[count++] = ns.vertices[i];
                }
                double vv[M * N];
                for (int k = 0; k < N; k++) vv[k] = v[k];
                for (int i = 1; i < M; i++)
                    for (int k = 0; k < N; k++) vv[i * N + k] = vertices[nsx[i]].coordinates[k];
                Pt c;
                double r;
                if (N == 2) ref_circumsphere_2d(vv, c.data(), &r);
                else if (ref_circumsphere_3d(vv, c.data(), &r)) err = 2;
                u.simplices.push_back(nsx);
                u.centers.push_back(c);
                u.radii.push_back(r);
                u.neighbors.emplace_back(nbid, killed_id);
            }
        }
    }

    // ---- pair_simplices (delaunay_tree.rs:674-695)
    static void pair_simplices(const std::vector<std::array<size_t, M>> &s, const std::vector<size_t> &ids,
                               std::vector<std::pair<size_t, size_t>> &out) {
        size_t n = s.size();
        for (size_t i = 0; i < n; i++)
            for (size_t j = i + 1; j < n; j++) {
                int count = 0;
                for (int k = 0; k < M; k++)
                    for (int m = 0; m < M; m++)
                        if (s[i][m] == s[j][k]) { count++; break; }
                if (count == N) {
                    out.emplace_back(ids[i], ids[j]);
                    out.emplace_back(ids[j], ids[i]);
                }
            }
    }

    // ---- TreeUpdate::new (delaunay_tree.rs:710-739)
    Update make_update(size_t id, const Pt &v) {
        Update u;
        u.vertex = v;
        u.killed_sites = locate(v);
        for (size_t k : u.killed_sites) get_new_simplices(k, v, id, u);
        u.simplices_id.resize(u.simplices.size());
        for (size_t i = 0; i < u.simplices.size(); i++) u.simplices_id[i] = i + 1;
        pair_simplices(u.simplices, u.simplices_id, u.new_neighbors);
        return u;
    }

    // the six mutation steps shared by insert_point (:133-208) and insert_points_parallel (:240-330)
    void apply(const Update &u, size_t base, size_t vid, bool locked) {
        for (size_t i = 0; i < u.simplices.size(); i++) {
            Simplex &s = simplices[base + u.simplices_id[i]];
            s.vertices = u.simplices[i];
            s.center = u.centers[i];
            s.radius = u.radii[i];
            s.neighbors.assign(1, u.neighbors[i].first);
            s.alive = true;
        }
        for (size_t i = 0; i < u.neighbors.size(); i++) {
            Simplex &nb = simplices[u.neighbors[i].first];
            for (size_t j = 0; j < nb.neighbors.size(); j++)
                if (nb.neighbors[j] == u.neighbors[i].second) nb.neighbors[j] = base + u.simplices_id[i];
        }
        for (auto &pr : u.new_neighbors) simplices[base + pr.first].neighbors.push_back(base + pr.second);
        vertices[vid].coordinates = u.vertex;
        vertices[vid].simplex.clear();
        for (size_t i = 0; i < u.simplices.size(); i++)
            for (int j = 0; j < M; j++) {
                size_t v = u.simplices[i][j];
                if (locked) vlocks[v].lock();
                vertices[v].simplex.push_back(base + u.simplices_id[i]);
                if (locked) vlocks[v].unlock();
            }
        for (size_t k : u.killed_sites) {
            if (!simplices[k].alive) { err = 3; return; }
            for (int i = 0; i < M; i++) {
                size_t v = simplices[k].vertices[i];
                if (locked) vlocks[v].lock();
                auto &lst = vertices[v].simplex;
                lst.erase(std::remove(lst.begin(), lst.end(), k), lst.end());
                if (locked) vlocks[v].unlock();
            }
        }
        for (size_t k : u.killed_sites) {
            simplices[k].alive = false;
            std::vector<size_t>().swap(simplices[k].neighbors);
        }
    }

    // ---- insert_point (delaunay_tree.rs:125-211)
    void insert_point(const Update &u, int64_t input_idx) {
        size_t vid = vertices.size();
        kdtree.add(u.vertex, (uint64_t)vid);
        ensure_simplices(max_simplex_id + u.simplices.size() + 1);
        vertices.emplace_back();
        input_index.push_back(input_idx);
        apply(u, max_simplex_id, vid, false);
        max_simplex_id += u.simplices.size();
        n_live += u.simplices.size();
        n_live -= u.killed_sites.size();
    }

    // ---- insert_points_parallel (delaunay_tree.rs:213-334)
    void insert_points_parallel(const std::vector<Update> &ups, const std::vector<int64_t> &input_idx) {
        std::vector<size_t> off(ups.size() + 1, 0);
        for (size_t i = 0; i < ups.size(); i++) off[i + 1] = off[i] + ups[i].simplices.size(); // serial fold :221-225
        size_t length = vertices.size();
        for (size_t i = 0; i < ups.size(); i++) kdtree.add(ups[i].vertex, (uint64_t)(length + i)); // serial :228-231
        ensure_simplices(max_simplex_id + off.back() + 1);
        vertices.resize(length + ups.size());
        input_index.resize(length + ups.size());
        ensure_vlocks(vertices.size());
        long nk = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : nk)
        for (long i = 0; i < (long)ups.size(); i++) {
            input_index[length + i] = input_idx[i];
            apply(ups[i], max_simplex_id + off[i], length + i, true);
            nk += (long)ups[i].killed_sites.size();
        }
        max_simplex_id += off.back();
        n_live += off.back();
        n_live -= (size_t)nk;
    }

    // ---- make_queue (scheduler.rs:6-28)
    struct QItem { size_t id; Pt v; std::vector<size_t> fp; };
    std::vector<QItem> make_queue(const double *pts, long n) {
        std::vector<QItem> q((size_t)n);
#pragma omp parallel for schedule(dynamic, 64)
        for (long i = 0; i < n; i++) {
            Pt v;
            for (int k = 0; k < N; k++) v[k] = pts[i * N + k];
            std::vector<size_t> killed = locate(v);
            std::vector<size_t> fp;
            for (size_t s : killed)
                for (size_t a : simplices[s].neighbors)
                    for (size_t b : simplices[a].neighbors) fp.push_back(b);
            std::sort(fp.begin(), fp.end());
            fp.erase(std::unique(fp.begin(), fp.end()), fp.end());
            q[i].id = (size_t)i;
            q[i].v = v;
            q[i].fp = std::move(fp);
        }
        return q;
    }
    // ---- find_placement (scheduler.rs:30-55) -- serial, as in the reference
    static std::vector<size_t> find_placement(const std::vector<QItem> &q) {
        std::unordered_map<size_t, std::vector<size_t>> occupancy;
        std::vector<size_t> placement(q.size(), 0);
        for (const QItem &it : q) {
            for (size_t s : it.fp) occupancy[s].push_back(it.id);
            size_t best = 0;
            for (size_t s : it.fp) {
                const std::vector<size_t> &o = occupancy[s];
                size_t r = o.size() == 1 ? 1 : placement[o[o.size() - 2]] + 1;
                best = std::max(best, r);
            }
            placement[it.id] = best; // .max().unwrap(): an empty footprint cannot happen (killed is non-empty)
        }
        return placement;
    }

    // ---- add_points_to_tree (delaunay_tree.rs:336-386)
    void add_points_to_tree(const double *pts, long n, int64_t first_input) {
        if (n == 0) return;
        std::vector<QItem> queue = make_queue(pts, n);
        if (err) return;
        std::vector<size_t> placement = find_placement(queue);
        size_t maxr = *std::max_element(placement.begin(), placement.end());
        std::vector<std::vector<size_t>> batches(maxr + 1);
        for (size_t i = 0; i < queue.size(); i++) batches[placement[i]].push_back(i);
        for (size_t r = 1; r <= maxr; r++) {
            const std::vector<size_t> &b = batches[r];
            if (b.empty()) continue;
            size_t n_points = vertices.size();
            std::vector<Update> ups(b.size());
            std::vector<int64_t> iidx(b.size());
#pragma omp parallel for schedule(dynamic, 16)
            for (long k = 0; k < (long)b.size(); k++) {
                ups[k] = make_update(n_points + k, queue[b[k]].v);
                iidx[k] = first_input + (int64_t)b[k];
            }
            if (err) return;
            insert_points_parallel(ups, iidx);
            if (err) return;
            rounds++;
        }
    }

    void insert_sequential(const double *pts, long n, int64_t first_input) {
        for (long i = 0; i < n; i++) {
            Pt v;
            for (int k = 0; k < N; k++) v[k] = pts[i * N + k];
            Update u = make_update(vertices.size(), v);
            if (err) return;
            insert_point(u, first_input + i);
            if (err) return;
        }
    }

    // ---- check_delaunay (delaunay_tree.rs:512-541) brute force
    bool check_delaunay() const {
        bool result = true;
        const size_t lim = 2 * M - 1; // ids > 7 (3D) / > 5 (2D)
        for (size_t s = 0; s <= max_simplex_id; s++) {
            if (!simplices[s].alive) continue;
            bool allreal = true;
            for (int k = 0; k < M; k++) allreal &= simplices[s].vertices[k] > lim;
            if (!allreal) continue;
            for (size_t v = 0; v < vertices.size(); v++) {
                bool isv = false;
                for (int k = 0; k < M; k++) isv |= simplices[s].vertices[k] == v;
                if (!isv && in_sphere(vertices[v].coordinates, simplices[s])) result = false;
            }
        }
        return result;
    }

    std::vector<uint32_t> edges() const {
        std::vector<uint64_t> keys;
        for (size_t s = 0; s <= max_simplex_id && s < simplices.size(); s++) {
            if (!simplices[s].alive) continue;
            for (int i = 0; i < M; i++)
                for (int j = i + 1; j < M; j++) {
                    int64_t a = input_index[simplices[s].vertices[i]], b = input_index[simplices[s].vertices[j]];
                    if (a < 0 || b < 0) continue;
                    uint64_t lo = (uint64_t)std::min(a, b), hi = (uint64_t)std::max(a, b);
                    keys.push_back(lo << 32 | hi);
                }
        }
        std::sort(keys.begin(), keys.end());
        keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
        std::vector<uint32_t> out(keys.size() * 2);
        for (size_t i = 0; i < keys.size(); i++) { out[2 * i] = (uint32_t)(keys[i] >> 32); out[2 * i + 1] = (uint32_t)keys[i]; }
        return out;
    }
};

struct Handle {
    int dim;
    Tree<2> *t2 = nullptr;
    Tree<3> *t3 = nullptr;
    std::vector<uint32_t> edge_cache;
    bool have_edges = false;
};

} // namespace

#define DISPATCH(h, expr2, expr3) ((h)->dim == 2 ? (expr2) : (expr3))

extern "C" {

void *vo_ref_create(int dim, const double *pts, long n) {
    Handle *h = new Handle;
    h->dim = dim;
    if (dim == 2) h->t2 = new Tree<2>(pts, n);
    else h->t3 = new Tree<3>(pts, n);
    return h;
}
void vo_ref_destroy(void *hv) {
    Handle *h = (Handle *)hv;
    if (!h) return;
    delete h->t2;
    delete h->t3;
    delete h;
}
int vo_ref_insert_sequential(void *hv, const double *pts, long n, long first_input) {
    Handle *h = (Handle *)hv;
    h->have_edges = false;
    if (h->dim == 2) { h->t2->insert_sequential(pts, n, first_input); return h->t2->err; }
    h->t3->insert_sequential(pts, n, first_input);
    return h->t3->err;
}
int vo_ref_add_points_to_tree(void *hv, const double *pts, long n, long first_input, int nthreads) {
    Handle *h = (Handle *)hv;
    h->have_edges = false;
    if (nthreads > 0) omp_set_num_threads(nthreads);
    if (h->dim == 2) { h->t2->add_points_to_tree(pts, n, first_input); return h->t2->err; }
    h->t3->add_points_to_tree(pts, n, first_input);
    return h->t3->err;
}
/* lib.rs:104-125 : <=1e5 points sequential, else first 1e5 sequential + add_points_to_tree(rest) */
void *vo_ref_delaunay(int dim, const double *pts, long n, int nthreads, int *err) {
    void *h = vo_ref_create(dim, pts, n);
    int e;
    if (n > 100000) {
        e = vo_ref_insert_sequential(h, pts, 100000, 0);
        if (!e) e = vo_ref_add_points_to_tree(h, pts + (size_t)100000 * dim, n - 100000, 100000, nthreads);
    } else {
        e = vo_ref_insert_sequential(h, pts, n, 0);
    }
    if (err) *err = e;
    return h;
}
/* out: [n_vertices, n_live_simplices, max_simplex_id, rounds] */
void vo_ref_counts(void *hv, uint64_t *out) {
    Handle *h = (Handle *)hv;
    out[0] = DISPATCH(h, h->t2->vertices.size(), h->t3->vertices.size());
    out[1] = DISPATCH(h, h->t2->n_live, h->t3->n_live);
    out[2] = DISPATCH(h, h->t2->max_simplex_id, h->t3->max_simplex_id);
    out[3] = DISPATCH(h, h->t2->rounds, h->t3->rounds);
}
int vo_ref_check_delaunay(void *hv) {
    Handle *h = (Handle *)hv;
    return DISPATCH(h, h->t2->check_delaunay(), h->t3->check_delaunay()) ? 1 : 0;
}
uint64_t vo_ref_edges(void *hv, uint32_t *out, uint64_t cap) {
    Handle *h = (Handle *)hv;
    if (!h->have_edges) {
        h->edge_cache = DISPATCH(h, h->t2->edges(), h->t3->edges());
        h->have_edges = true;
    }
    uint64_t m = h->edge_cache.size() / 2;
    if (out) memcpy(out, h->edge_cache.data(), sizeof(uint32_t) * 2 * std::min(m, cap));
    return m;
}
/* scheduler test hook (tests/test_scheduler.rs): placement of n points against the current tree */
int vo_ref_placement(void *hv, const double *pts, long n, uint64_t *placement) {
    Handle *h = (Handle *)hv;
    if (h->dim == 2) {
        auto q = h->t2->make_queue(pts, n);
        auto p = Tree<2>::find_placement(q);
        for (long i = 0; i < n; i++) placement[i] = p[i];
        return h->t2->err;
    }
    auto q = h->t3->make_queue(pts, n);
    auto p = Tree<3>::find_placement(q);
    for (long i = 0; i < n; i++) placement[i] = p[i];
    return h->t3->err;
}
int vo_ref_max_threads(void) { return omp_get_max_threads(); }

} // extern "C"
