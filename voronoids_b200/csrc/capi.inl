// capi.inl -- implementation of include/voronoids_b200.h on top of Engine<D>.
// Included by vor_lib.cu (product, CUDA backend) and by tests/emu/emu_lib.cpp (kernel-logic unit tests).
#include <memory>
#include <mutex>
#include <new>
#include <thread>

#include "../../include/voronoids_b200.h"

namespace {

thread_local std::string g_err;
vor::EngineOptions g_opts;
std::once_flag g_opts_once;
std::mutex g_opts_mu;   // vor_delaunay_batch runs one host thread per device: options are read under this lock

vor::EngineOptions current_options() {
    std::call_once(g_opts_once, [] { vor::options_from_env(g_opts); });
    std::lock_guard<std::mutex> lk(g_opts_mu);
    return g_opts;
}

template <class F> vor_status guarded(F &&f) {
    try {
        return f();
    } catch (const vor::EngineError &e) {
        g_err = e.msg;
        return (vor_status)(e.code == vor::ERR_WALK ? VOR_ERR_INTERNAL : e.code);
    } catch (const vor::be::CudaError &e) {
        g_err = e.what();
        return (vor_status)e.code;
    } catch (const std::bad_alloc &) {
        g_err = "host allocation failed";
        return VOR_ERR_OOM;
    } catch (const std::exception &e) {
        g_err = e.what();
        return VOR_ERR_INTERNAL;
    }
}

struct DevBuf {
    void *p = nullptr;
    explicit DevBuf(size_t bytes) { p = vor::be::dmalloc(bytes); }
    ~DevBuf() { vor::be::dfree(p); }
    DevBuf(const DevBuf &) = delete;
};

std::vector<int> offsets32(const int64_t *off, size_t n_sets) {
    std::vector<int> o(n_sets + 1);
    for (size_t i = 0; i <= n_sets; i++) {
        if (off[i] < 0 || off[i] > 0x7fffffff || (i && off[i] < off[i - 1])) throw vor::EngineError{vor::ERR_ARG, "bad set offsets"};
        o[i] = (int)off[i];
    }
    return o;
}

} // namespace

struct vor_tree {
    int dim = 3;
    int device = 0;
    size_t n_sets = 1;
    vor::be::Stream stream{};
    std::unique_ptr<vor::Engine<2>> e2;
    std::unique_ptr<vor::Engine<3>> e3;
    template <class F> auto visit(F &&f) { return dim == 2 ? f(*e2) : f(*e3); }
};

extern "C" {

const char *vor_last_error(void) { return g_err.c_str(); }
uint64_t vor_kernel_launches(void) { return vor::be::g_launches; }
void vor_release_memory(void) { vor::be::release_cached(); }

int vor_set_option(const char *name, double value) {
    current_options();
    std::lock_guard<std::mutex> lk(g_opts_mu);
    const std::string n(name ? name : "");
    if (n == "slot_cap") g_opts.slot_cap = (int)value;
    else if (n == "min_attempt") g_opts.min_attempt = (int)value;
    else if (n == "attempt_div") g_opts.attempt_div = value;
    else if (n == "stage0") g_opts.stage0 = (int)value;
    else if (n == "stats") g_opts.stats = (int)value;
    else if (n == "verbose") g_opts.verbose = (int)value;
    else if (n == "profile") g_opts.profile = (int)value;
    else if (n == "coop") g_opts.coop = (int)value;
    else if (n == "rounds_per_sync") g_opts.rounds_per_sync = (int)value;
    else if (n == "select_mode") g_opts.select_mode = (int)value;
    else if (n == "red") g_opts.red = (int)value;
    else if (n == "commit_smem") g_opts.commit_smem = (int)value;
    else if (n == "split_exact") g_opts.split_exact = (int)value;
    else if (n == "mid_twin") g_opts.mid_twin = (int)value;
    else if (n == "edge_wedge") g_opts.edge_wedge = (int)value;
    else if (n == "edge_dir_x") g_opts.edge_dir[0] = value;
    else if (n == "edge_dir_y") g_opts.edge_dir[1] = value;
    else if (n == "edge_dir_z") g_opts.edge_dir[2] = value;
    else if (n == "tet_factor") g_opts.tet_factor = value;
    else if (n == "compact_frac") g_opts.compact_frac = value;
    else if (n == "stage_log") g_opts.stage_log = (int)value;
    else if (n == "tiled") g_opts.tiled = (int)value;
    else if (n == "subround") g_opts.subround = (int)value;
    else if (n == "pdl") g_opts.pdl = (int)value;
    else if (n == "carry_frac") g_opts.carry_frac = value;
    else if (n == "capk") { g_opts.capk = (int)value; g_opts.capb = 2 * g_opts.capk + 4; }
    else if (n == "big_slots") g_opts.big_slots = (int)value;
    else if (n == "big_capk") g_opts.big_capk = (int)value;
    else return -1;
    return 0;
}

vor_status vor_tree_create_batch_device(int dim, const double *d_points, const int64_t *set_offsets, size_t n_sets, int device,
                                        void *cuda_stream, vor_tree **out) {
    return guarded([&]() -> vor_status {
        if (!out || (dim != 2 && dim != 3) || n_sets < 1 || !set_offsets) { g_err = "bad argument"; return VOR_ERR_ARG; }
        vor::be::set_device(device);
        std::unique_ptr<vor_tree> t(new vor_tree);
        t->dim = dim;
        t->device = device;
        t->n_sets = n_sets;
        t->stream = (vor::be::Stream)(uintptr_t)cuda_stream;
        const std::vector<int> off = offsets32(set_offsets, n_sets);
        if (dim == 2) { t->e2.reset(new vor::Engine<2>(t->stream, current_options())); t->e2->create(d_points, off.back(), off.data(), (int)n_sets); }
        else { t->e3.reset(new vor::Engine<3>(t->stream, current_options())); t->e3->create(d_points, off.back(), off.data(), (int)n_sets); }
        *out = t.release();
        return VOR_OK;
    });
}

vor_status vor_tree_create_device(int dim, const double *d_points, size_t n, int device, void *cuda_stream, vor_tree **out) {
    const int64_t off[2] = {0, (int64_t)n};
    return vor_tree_create_batch_device(dim, d_points, off, 1, device, cuda_stream, out);
}

vor_status vor_tree_create_batch(int dim, const double *points, const int64_t *set_offsets, size_t n_sets, int device, vor_tree **out) {
    return guarded([&]() -> vor_status {
        if (!set_offsets || n_sets < 1 || (dim != 2 && dim != 3)) { g_err = "bad argument"; return VOR_ERR_ARG; }
        vor::be::set_device(device);
        const size_t n = (size_t)set_offsets[n_sets];
        DevBuf d(sizeof(double) * n * dim);
        vor::be::h2d_big(d.p, points, sizeof(double) * n * dim, vor::be::Stream{});
        return vor_tree_create_batch_device(dim, (const double *)d.p, set_offsets, n_sets, device, nullptr, out);
    });
}

// E1 of SURVEY.md §8(e): sets are independent units.  Contiguous blocks of ceil(n_sets / n_dev) sets per device, one
// host thread per device (create + insert of that block as ONE batch tree), no exchange between devices.
vor_status vor_delaunay_batch(int dim, const double *points, const int64_t *set_offsets, size_t n_sets, const int *devices, size_t n_dev,
                              vor_tree **trees, int64_t *shard) {
    if (!points || !set_offsets || !devices || !trees || !shard || n_sets < 1 || n_dev < 1 || (dim != 2 && dim != 3)) {
        g_err = "bad argument";
        return VOR_ERR_ARG;
    }
    const size_t per = (n_sets + n_dev - 1) / n_dev;
    for (size_t d = 0; d <= n_dev; d++) shard[d] = (int64_t)std::min(n_sets, d * per);
    std::vector<vor_status> st(n_dev, VOR_OK);
    std::vector<std::string> msg(n_dev);
    std::vector<std::thread> th;
    for (size_t d = 0; d < n_dev; d++) {
        trees[d] = nullptr;
        th.emplace_back([&, d]() {
            // nothing may escape a thread body (std::terminate): allocation failures become a status like everything else
            const vor_status r = guarded([&]() -> vor_status {
                const size_t lo = (size_t)shard[d], hi = (size_t)shard[d + 1];
                if (hi <= lo) return VOR_OK;   // more devices than sets: nothing for this one
                std::vector<int64_t> off(hi - lo + 1);
                for (size_t s = lo; s <= hi; s++) off[s - lo] = set_offsets[s] - set_offsets[lo];
                const double *p = points + (size_t)set_offsets[lo] * dim;
                vor_status q = vor_tree_create_batch(dim, p, off.data(), hi - lo, devices[d], &trees[d]);
                if (q == VOR_OK) q = vor_tree_insert_batch(trees[d], p, off.data());
                return q;
            });
            if (r != VOR_OK && r != VOR_ERR_DUPLICATE_POINT) { st[d] = r; msg[d] = vor_last_error(); }   // g_err is thread-local
        });
    }
    for (auto &t : th) t.join();
    for (size_t d = 0; d < n_dev; d++)
        if (st[d] != VOR_OK) {
            g_err = "device " + std::to_string(devices[d]) + ": " + msg[d];
            for (size_t k = 0; k < n_dev; k++) { vor_tree_destroy(trees[k]); trees[k] = nullptr; }
            return st[d];
        }
    return VOR_OK;
}

// Streaming driver for batches far beyond one device store (BASELINE.json configs[4]: 8,192 sets x 100k points; the caller
// pattern of examples/parallel_insert.rs:56-78, a loop over batches).  The sets of this call are cut into chunks of at
// most `chunk_sets` sets / `chunk_points` points; every chunk is ONE batch tree (create + insert + edge list) whose store
// goes back to the caching allocator for the next chunk; with host input the copy of chunk k+1 (pinned staging, its own
// stream, a helper thread) overlaps the rounds of chunk k.  Results per set: edge count and checksum64 of the set-local
// canonical edge list; `cb` (optional) receives every chunk's edge list (global input indices of this call, on the host).
vor_status vor_delaunay_batch_stream(int dim, const double *points, int points_on_device, const int64_t *set_offsets, size_t n_sets, int device,
                                     size_t chunk_sets, size_t chunk_points, uint64_t *n_edges, uint64_t *checksums, vor_chunk_cb cb, void *user) {
    return guarded([&]() -> vor_status {
        if (!points || !set_offsets || n_sets < 1 || (dim != 2 && dim != 3) || (!n_edges && !checksums && !cb)) { g_err = "bad argument"; return VOR_ERR_ARG; }
        vor::be::set_device(device);
        if (chunk_sets == 0) chunk_sets = 128;
        if (chunk_points == 0) chunk_points = dim == 3 ? (size_t)13 << 20 : (size_t)48 << 20;   // 2^29 simplex slots per store
        // chunk table
        std::vector<size_t> cut{0};
        for (size_t s = 0; s < n_sets;) {
            size_t e = s;
            while (e < n_sets && e - s < chunk_sets && (e == s || (size_t)(set_offsets[e + 1] - set_offsets[s]) <= chunk_points)) e++;
            cut.push_back(e);
            s = e;
        }
        const size_t nchunks = cut.size() - 1;
        size_t maxPts = 0;
        for (size_t c = 0; c < nchunks; c++) maxPts = std::max(maxPts, (size_t)(set_offsets[cut[c + 1]] - set_offsets[cut[c]]));
        // host input: two device buffers, chunk c+1 is copied by a helper thread while chunk c is triangulated
        std::unique_ptr<DevBuf> buf[2];
        vor::be::Stream copyStream{};
        bool haveCopyStream = false;
#if VOR_GPU
        if (!points_on_device) {
            buf[0].reset(new DevBuf(sizeof(double) * maxPts * dim));
            if (nchunks > 1) buf[1].reset(new DevBuf(sizeof(double) * maxPts * dim));
            VOR_CUDA(cudaStreamCreateWithFlags(&copyStream, cudaStreamNonBlocking));
            haveCopyStream = true;
        }
#else
        if (!points_on_device) { buf[0].reset(new DevBuf(sizeof(double) * maxPts * dim)); if (nchunks > 1) buf[1].reset(new DevBuf(sizeof(double) * maxPts * dim)); }
#endif
        auto chunk_src = [&](size_t c) { return points + (size_t)set_offsets[cut[c]] * dim; };
        auto chunk_n = [&](size_t c) { return (size_t)(set_offsets[cut[c + 1]] - set_offsets[cut[c]]); };
        std::string copyErr;
        auto copy_chunk = [&](size_t c) {
            try {
                vor::be::set_device(device);
                vor::be::h2d_big(buf[c & 1]->p, chunk_src(c), sizeof(double) * chunk_n(c) * dim, copyStream);
                vor::be::sync(copyStream);
            } catch (const std::exception &e) { copyErr = e.what(); }
        };
        vor_status result = VOR_OK;
        bool anyDup = false;
        std::thread copier;
        if (!points_on_device) copy_chunk(0);
        for (size_t c = 0; c < nchunks && result == VOR_OK; c++) {
            if (copier.joinable()) copier.join();
            if (!copyErr.empty()) { g_err = "host to device copy: " + copyErr; result = VOR_ERR_CUDA; break; }
            if (!points_on_device && c + 1 < nchunks) copier = std::thread(copy_chunk, c + 1);
            const double *d_pts = points_on_device ? chunk_src(c) : (const double *)buf[c & 1]->p;
            const size_t ns = cut[c + 1] - cut[c];
            std::vector<int64_t> off(ns + 1);
            for (size_t s = 0; s <= ns; s++) off[s] = set_offsets[cut[c] + s] - set_offsets[cut[c]];
            vor_tree *t = nullptr;
            vor_status r = vor_tree_create_batch_device(dim, d_pts, off.data(), ns, device, nullptr, &t);
            if (r == VOR_OK) r = vor_tree_insert_batch_device(t, d_pts, off.data());
            if (r == VOR_ERR_DUPLICATE_POINT) { anyDup = true; r = VOR_OK; }
            if (r == VOR_OK) {
                r = guarded([&]() -> vor_status {
                    const std::vector<int> off32 = offsets32(off.data(), ns);
                    std::vector<unsigned long long> cnt(ns), sum(ns);
                    t->visit([&](auto &e) { e.per_set_edge_stats(off32.data(), cnt.data(), sum.data()); return 0; });
                    for (size_t s = 0; s < ns; s++) {
                        if (n_edges) n_edges[cut[c] + s] = cnt[s];
                        if (checksums) checksums[cut[c] + s] = sum[s];
                    }
                    if (cb) {
                        uint32_t *h = nullptr;
                        long long m = 0;
                        t->visit([&](auto &e) { h = e.edges_to_host_block(&m); return 0; });
                        cb(user, cut[c], ns, (int64_t)set_offsets[cut[c]], h, (size_t)m);
                        vor::be::g_hostpool.free(h);
                    }
                    return VOR_OK;
                });
            }
            vor_tree_destroy(t);
            if (r != VOR_OK) { g_err = "chunk " + std::to_string(c) + " (sets " + std::to_string(cut[c]) + ".." + std::to_string(cut[c + 1]) + "): " + g_err; result = r; }
        }
        if (copier.joinable()) copier.join();
#if VOR_GPU
        if (haveCopyStream) cudaStreamDestroy(copyStream);
#endif
        (void)haveCopyStream;
        if (result == VOR_OK && anyDup) { g_err = "duplicate point(s) dropped"; return VOR_ERR_DUPLICATE_POINT; }
        return result;
    });
}

vor_status vor_tree_create(int dim, const double *points, size_t n, int device, vor_tree **out) {
    const int64_t off[2] = {0, (int64_t)n};
    return vor_tree_create_batch(dim, points, off, 1, device, out);
}

void vor_tree_destroy(vor_tree *t) {
    if (!t) return;
    try { vor::be::set_device(t->device); } catch (...) {}
    delete t;
}

void vor_tree_set_stream(vor_tree *t, void *cuda_stream) {
    if (!t) return;
    t->stream = (vor::be::Stream)(uintptr_t)cuda_stream;
    if (t->e2) t->e2->stream = t->stream;
    if (t->e3) t->e3->stream = t->stream;
}

vor_status vor_tree_insert_batch_device(vor_tree *t, const double *d_points, const int64_t *set_offsets) {
    return guarded([&]() -> vor_status {
        if (!t || !set_offsets) { g_err = "bad argument"; return VOR_ERR_ARG; }
        vor::be::set_device(t->device);
        const std::vector<int> off = offsets32(set_offsets, t->n_sets);
        const int dup0 = t->visit([](auto &e) { return e.hcnt->ndup; });
        t->visit([&](auto &e) { e.insert(d_points, off.back(), off.data()); return 0; });
        const int dup1 = t->visit([](auto &e) { return e.hcnt->ndup; });
        if (dup1 > dup0) { g_err = std::to_string(dup1 - dup0) + " duplicate point(s) dropped"; return VOR_ERR_DUPLICATE_POINT; }
        return VOR_OK;
    });
}

vor_status vor_tree_insert_device(vor_tree *t, const double *d_points, size_t n, vor_insert_mode) {
    if (!t || t->n_sets != 1) { g_err = "single-set insert on a batch tree"; return VOR_ERR_ARG; }
    const int64_t off[2] = {0, (int64_t)n};
    return vor_tree_insert_batch_device(t, d_points, off);
}

vor_status vor_tree_insert_batch(vor_tree *t, const double *points, const int64_t *set_offsets) {
    return guarded([&]() -> vor_status {
        if (!t || !set_offsets) { g_err = "bad argument"; return VOR_ERR_ARG; }
        vor::be::set_device(t->device);
        const size_t n = (size_t)set_offsets[t->n_sets];
        DevBuf d(sizeof(double) * n * t->dim);
        vor::be::h2d_big(d.p, points, sizeof(double) * n * t->dim, t->stream);
        return vor_tree_insert_batch_device(t, (const double *)d.p, set_offsets);
    });
}

vor_status vor_tree_insert(vor_tree *t, const double *points, size_t n, vor_insert_mode) {
    if (!t || t->n_sets != 1) { g_err = "single-set insert on a batch tree"; return VOR_ERR_ARG; }
    const int64_t off[2] = {0, (int64_t)n};
    return vor_tree_insert_batch(t, points, off);
}

vor_status vor_delaunay(int dim, const double *points, size_t n, int device, vor_tree **out) {
    return guarded([&]() -> vor_status {
        if (!out || (dim != 2 && dim != 3)) { g_err = "bad argument"; return VOR_ERR_ARG; }
        vor::be::set_device(device);
        // warm the host pool for the edge list the caller is about to ask for (~7.8 edges per 3D point, ~3 per 2D point):
        // the page-locking of a fresh block runs on a side thread under the insertion
        std::thread warm;
#if VOR_GPU
        const size_t est = (size_t)((dim == 3 ? 8.2 : 3.2) * (double)n) * 8;
        if (n >= 100000 && vor::be::HostPool::want_pinned() && !vor::be::g_hostpool.has_block(est))
            warm = std::thread([est, device] {
                try {
                    vor::be::set_device(device);
                    bool pinned = false;
                    void *p = vor::be::g_hostpool.alloc(est, &pinned);
                    vor::be::g_hostpool.free(p);
                } catch (...) {}
            });
#endif
        struct Joiner { std::thread &t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{warm};
        DevBuf d(sizeof(double) * n * dim);
        vor::be::h2d_big(d.p, points, sizeof(double) * n * dim, vor::be::Stream{});
        vor_tree *t = nullptr;
        vor_status s = vor_tree_create_device(dim, (const double *)d.p, n, device, nullptr, &t);
        if (s != VOR_OK) return s;
        s = vor_tree_insert_device(t, (const double *)d.p, n, VOR_INSERT_PARALLEL);
        if (s != VOR_OK && s != VOR_ERR_DUPLICATE_POINT) { vor_tree_destroy(t); return s; }
        *out = t;
        return s;
    });
}

vor_status vor_tree_counts(vor_tree *t, uint64_t *n_vertices, uint64_t *n_simplices, uint64_t *max_simplex_id) {
    return guarded([&]() -> vor_status {
        if (!t) { g_err = "null tree"; return VOR_ERR_ARG; }
        vor::be::set_device(t->device);
        return t->visit([&](auto &e) -> vor_status {
            if (n_vertices) *n_vertices = (uint64_t)e.nsuper * 2 + (uint64_t)e.insertedTotal;
            if (max_simplex_id) *max_simplex_id = (uint64_t)(e.M * e.nsets) + (uint64_t)e.created_all();
            if (n_simplices) *n_simplices = (uint64_t)e.export_simplices(nullptr, nullptr, nullptr, nullptr, 0);
            return VOR_OK;
        });
    });
}

vor_status vor_tree_edges(vor_tree *t, uint32_t *edges, size_t cap, size_t *n_edges) {
    return guarded([&]() -> vor_status {
        if (!t) { g_err = "null tree"; return VOR_ERR_ARG; }
        vor::be::set_device(t->device);
        return t->visit([&](auto &e) -> vor_status {
            const long long m = e.edges();
            if (n_edges) *n_edges = (size_t)m;
            if (edges && (long long)cap < m) { g_err = "edge buffer too small: " + std::to_string(cap) + " < " + std::to_string(m); return VOR_ERR_ARG; }
            if (edges) e.copy_edges(edges, (long long)cap);
            return VOR_OK;
        });
    });
}

vor_status vor_tree_edges_host(vor_tree *t, uint32_t **edges, size_t *n_edges) {
    return guarded([&]() -> vor_status {
        if (!t || !edges || !n_edges) { g_err = "null argument"; return VOR_ERR_ARG; }
        vor::be::set_device(t->device);
        return t->visit([&](auto &e) -> vor_status {
            long long m = 0;
            *edges = e.edges_to_host_block(&m);
            *n_edges = (size_t)m;
            return VOR_OK;
        });
    });
}

vor_status vor_host_free(void *block) {
    if (!vor::be::g_hostpool.free(block)) { g_err = "not a block handed out by this library"; return VOR_ERR_ARG; }
    return VOR_OK;
}

vor_status vor_tree_edges_device(vor_tree *t, const uint32_t **d_edges, size_t *n_edges, uint64_t *checksum) {
    return guarded([&]() -> vor_status {
        if (!t) { g_err = "null tree"; return VOR_ERR_ARG; }
        vor::be::set_device(t->device);
        return t->visit([&](auto &e) -> vor_status {
            const long long m = e.edges();
            if (n_edges) *n_edges = (size_t)m;
            if (d_edges) *d_edges = e.d_edges;
            if (checksum) *checksum = e.edge_checksum();
            return VOR_OK;
        });
    });
}

vor_status vor_tree_export_simplices(vor_tree *t, int32_t *vertices, int32_t *neighbors, double *centers, double *radii, size_t cap,
                                     size_t *n_simplices) {
    return guarded([&]() -> vor_status {
        if (!t) { g_err = "null tree"; return VOR_ERR_ARG; }
        vor::be::set_device(t->device);
        return t->visit([&](auto &e) -> vor_status {
            const int n = e.export_simplices(nullptr, nullptr, nullptr, nullptr, 0);
            if (n_simplices) *n_simplices = (size_t)n;
            if (!vertices && !neighbors && !centers && !radii) return VOR_OK;
            if (cap < (size_t)n) { g_err = "export buffer too small"; return VOR_ERR_ARG; }
            e.export_simplices(vertices, neighbors, centers, radii, 2 * e.M * e.nsets);
            return VOR_OK;
        });
    });
}

vor_status vor_tree_locate(vor_tree *t, const double *points, size_t n, int32_t *out_ids, size_t cap, int32_t *counts) {
    return guarded([&]() -> vor_status {
        if (!t || !points || !out_ids || !counts || cap == 0 || t->n_sets != 1) { g_err = "bad argument"; return VOR_ERR_ARG; }
        vor::be::set_device(t->device);
        return t->visit([&](auto &e) -> vor_status {
            e.locate(points, (int)n, (int)cap, out_ids, counts);
            for (size_t i = 0; i < n; i++)
                if (counts[i] == -2) { g_err = "query point outside the super simplex"; return VOR_ERR_OUTSIDE; }
            return VOR_OK;
        });
    });
}

vor_status vor_tree_export_vertices(vor_tree *t, double *coords, int64_t *simp_off, int32_t *simps, size_t cap, size_t *n_vertices,
                                    size_t *n_incidences) {
    return guarded([&]() -> vor_status {
        if (!t) { g_err = "null tree"; return VOR_ERR_ARG; }
        vor::be::set_device(t->device);
        return t->visit([&](auto &e) -> vor_status {
            long long ninc = 0;
            static_assert(sizeof(long long) == sizeof(int64_t), "offset width");
            const long long n = e.export_vertices(coords, (long long *)simp_off, simps, (long long)cap, &ninc);
            if (n_vertices) *n_vertices = (size_t)n;
            if (n_incidences) *n_incidences = (size_t)ninc;
            return VOR_OK;
        });
    });
}

vor_status vor_make_queue(vor_tree *t, const double *points, size_t n, int64_t *offsets, int32_t *ids, size_t cap, size_t *total) {
    return guarded([&]() -> vor_status {
        if (!t || (!points && n) || !offsets) { g_err = "null argument"; return VOR_ERR_ARG; }
        vor::be::set_device(t->device);
        return t->visit([&](auto &e) -> vor_status {
            // conflict regions and footprints grow until every query fits (uniform 3D: ~20 killed, ~150 footprint ids)
            int kcap = 64, fcap = 512;
            std::vector<int> fp, cnt(n);
            for (;;) {
                fp.assign(n * (size_t)fcap, 0);
                e.make_queue(points, (int)n, kcap, fcap, fp.data(), cnt.data());
                bool again = false;
                for (size_t i = 0; i < n; i++) {
                    if (cnt[i] == -2) { g_err = "query point outside the super simplex"; return VOR_ERR_OUTSIDE; }
                    if (cnt[i] == -1) again = true;
                }
                if (!again) break;
                if (kcap >= (1 << 14)) { g_err = "conflict region exceeds the query capacity"; return VOR_ERR_CAPACITY; }
                kcap *= 4; fcap *= 4;
            }
            int64_t run = 0;
            for (size_t i = 0; i < n; i++) { offsets[i] = run; run += cnt[i]; }
            offsets[n] = run;
            if (total) *total = (size_t)run;
            if (!ids) return VOR_OK;
            if (cap < (size_t)run) { g_err = "footprint buffer too small"; return VOR_ERR_ARG; }
            for (size_t i = 0; i < n; i++) std::copy(fp.begin() + i * (size_t)fcap, fp.begin() + i * (size_t)fcap + cnt[i], ids + offsets[i]);
            return VOR_OK;
        });
    });
}

vor_status vor_find_placement(const int64_t *offsets, const int32_t *ids, size_t n, uint64_t *placement, int device) {
    return guarded([&]() -> vor_status {
        if (!offsets || !placement || (!ids && n && offsets[n] > 0)) { g_err = "null argument"; return VOR_ERR_ARG; }
        if (n == 0) return VOR_OK;
        vor::be::set_device(device);
        const long long total = offsets[n];
        int maxid = -1;
        for (long long x = 0; x < total; x++) {
            if (ids[x] < 0) { g_err = "negative simplex id in the queue"; return VOR_ERR_ARG; }
            maxid = std::max(maxid, ids[x]);
        }
        for (size_t i = 0; i < n; i++)
            if (offsets[i + 1] <= offsets[i]) { g_err = "empty footprint (the reference panics: scheduler.rs:52)"; return VOR_ERR_NO_CONFLICT; }
        vor::be::Stream s = 0;
        DevBuf b_off(sizeof(long long) * (n + 1)), b_ids(sizeof(int) * (size_t)std::max(total, 1LL)), b_last(sizeof(int) * (size_t)(maxid + 2)),
            b_round(sizeof(unsigned long long) * n);   // RAII: nothing leaks when a copy or the launch throws
        long long *d_off = (long long *)b_off.p;
        int *d_ids = (int *)b_ids.p;
        int *d_last = (int *)b_last.p;
        unsigned long long *d_round = (unsigned long long *)b_round.p;
        static_assert(sizeof(long long) == sizeof(int64_t), "offset width");
        vor::be::h2d(d_off, offsets, sizeof(long long) * (n + 1), s);
        vor::be::h2d(d_ids, ids, sizeof(int) * (size_t)total, s);
        vor::be::dmemset(d_last, 0, sizeof(int) * (size_t)(maxid + 2), s);
        vor::PlacementArgs pa{d_off, d_ids, (int)n, d_last, d_round};
        VOR_LAUNCH(vor::PlacementArgs, vor::placement_body, 1, pa, s);
        vor::be::d2h(placement, d_round, sizeof(unsigned long long) * n, s);
        vor::be::sync(s);
        return VOR_OK;
    });
}

vor_status vor_tree_check_delaunay(vor_tree *t, int *ok, int32_t *fail_counts) {
    return guarded([&]() -> vor_status {
        if (!t) { g_err = "null tree"; return VOR_ERR_ARG; }
        vor::be::set_device(t->device);
        return t->visit([&](auto &e) -> vor_status {
            int f[8];
            e.validate(f);
            if (ok) *ok = (f[0] | f[1] | f[2] | f[3] | f[4] | f[5]) == 0;
            if (fail_counts) for (int i = 0; i < 6; i++) fail_counts[i] = f[i];
            return VOR_OK;
        });
    });
}

// ---- slab mode: one triangulation over several GPUs, see voronoids_b200/slab.py for the driver
vor_status vor_slab_local_bounds(int dim, const double *d_points, size_t n, int device, double *lo, double *hi) {
    return guarded([&]() -> vor_status {
        if ((dim != 2 && dim != 3) || !lo || !hi || n > 0x7fffffff) { g_err = "bad argument"; return VOR_ERR_ARG; }
        vor::be::set_device(device);
        if (dim == 2) vor::Engine<2>::local_bounds(d_points, (int)n, lo, hi, vor::be::Stream{});
        else vor::Engine<3>::local_bounds(d_points, (int)n, lo, hi, vor::be::Stream{});
        return VOR_OK;
    });
}
vor_status vor_slab_count_outside(int dim, const double *d_points, size_t n, int device, const double *lo, const double *hi, uint64_t *count) {
    return guarded([&]() -> vor_status {
        if ((dim != 2 && dim != 3) || !lo || !hi || !count || n > 0x7fffffff) { g_err = "bad argument"; return VOR_ERR_ARG; }
        vor::be::set_device(device);
        *count = (uint64_t)(dim == 2 ? vor::Engine<2>::count_outside(d_points, (int)n, lo, hi, vor::be::Stream{})
                                     : vor::Engine<3>::count_outside(d_points, (int)n, lo, hi, vor::be::Stream{}));
        return VOR_OK;
    });
}
vor_status vor_tree_create_bounds(int dim, const double *lo, const double *hi, uint64_t outside, size_t capacity_hint, int device, void *cuda_stream,
                                  vor_tree **out) {
    return guarded([&]() -> vor_status {
        if (!out || (dim != 2 && dim != 3) || !lo || !hi || capacity_hint > 0x7fffffff) { g_err = "bad argument"; return VOR_ERR_ARG; }
        vor::be::set_device(device);
        std::unique_ptr<vor_tree> t(new vor_tree);
        t->dim = dim;
        t->device = device;
        t->n_sets = 1;
        t->stream = (vor::be::Stream)(uintptr_t)cuda_stream;
        const int off[2] = {0, 0};
        const int hint = (int)capacity_hint;
        if (dim == 2) { t->e2.reset(new vor::Engine<2>(t->stream, current_options())); t->e2->create(nullptr, hint, off, 1, lo, hi, outside ? 1 : 0); }
        else { t->e3.reset(new vor::Engine<3>(t->stream, current_options())); t->e3->create(nullptr, hint, off, 1, lo, hi, outside ? 1 : 0); }
        *out = t.release();
        return VOR_OK;
    });
}
vor_status vor_tree_certify_slab(vor_tree *t, const uint8_t *owned, size_t n_owned, int axis, double range_lo, double range_hi, double shell,
                                 uint64_t *n_uncertified, double *need) {
    return guarded([&]() -> vor_status {
        if (!t || !owned || !n_uncertified || !need || axis < 0 || axis >= t->dim) { g_err = "bad argument"; return VOR_ERR_ARG; }
        vor::be::set_device(t->device);
        return t->visit([&](auto &e) -> vor_status {
            if ((size_t)e.ninput != n_owned) { g_err = "owned flags must cover every inserted point"; return VOR_ERR_ARG; }
            *n_uncertified = (uint64_t)e.certify_slab(owned, axis, range_lo, range_hi, shell, need);
            return VOR_OK;
        });
    });
}

vor_status vor_tree_uncertified_slab(vor_tree *t, const uint8_t *owned, size_t n_owned, int axis, double range_lo, double range_hi, double shell,
                                     double *verts, double *reach, size_t cap, uint64_t *n_uncertified, double *need) {
    return guarded([&]() -> vor_status {
        if (!t || !owned || !n_uncertified || !need || !verts || !reach || cap < 1 || cap > (1u << 20) || axis < 0 || axis >= t->dim) { g_err = "bad argument"; return VOR_ERR_ARG; }
        vor::be::set_device(t->device);
        return t->visit([&](auto &e) -> vor_status {
            if ((size_t)e.ninput != n_owned) { g_err = "owned flags must cover every inserted point"; return VOR_ERR_ARG; }
            *n_uncertified = (uint64_t)e.certify_slab(owned, axis, range_lo, range_hi, shell, need, verts, reach, (int)cap);
            return VOR_OK;
        });
    });
}

vor_status vor_points_in_spheres(int dim, const double *d_points, size_t n, const double *simplices, size_t k, int device, uint64_t *inside) {
    return guarded([&]() -> vor_status {
        if ((dim != 2 && dim != 3) || !simplices || !inside || k < 1 || k > (1u << 20) || (n && !d_points)) { g_err = "bad argument"; return VOR_ERR_ARG; }
        vor::be::set_device(device);
        if (n == 0) { for (size_t j = 0; j < k; j++) inside[j] = 0; return VOR_OK; }
        auto run = [&](auto tag) {
            constexpr int D = decltype(tag)::value;
            constexpr int M = D + 1;
            vor::be::Stream st{};
            DevBuf dsimp(sizeof(double) * k * M * D), dins(sizeof(unsigned long long) * k), dcnt(sizeof(vor::Counters));
            vor::be::h2d(dsimp.p, simplices, sizeof(double) * k * M * D, st);
            vor::be::dmemset(dins.p, 0, sizeof(unsigned long long) * k, st);
            vor::be::dmemset(dcnt.p, 0, sizeof(vor::Counters), st);
            const vor::InSpheresArgs<D> a{d_points, (const double *)dsimp.p, (int)k, (unsigned long long *)dins.p, (vor::Counters *)dcnt.p};
            for (size_t done = 0; done < n; done += (size_t)1 << 30) {
                vor::InSpheresArgs<D> b = a;
                b.pts = d_points + done * D;
                VOR_LAUNCH(vor::InSpheresArgs<D>, vor::in_spheres_body<D>, (long long)std::min(n - done, (size_t)1 << 30), b, st);
            }
            std::vector<unsigned long long> h(k);
            vor::be::d2h(h.data(), dins.p, sizeof(unsigned long long) * k, st);
            vor::Counters hc;
            vor::be::d2h(&hc, dcnt.p, sizeof(hc), st);
            vor::be::sync(st);
            if (hc.err) throw vor::EngineError{hc.err, "points_in_spheres: coordinate range exceeds the exact arithmetic"};
            for (size_t j = 0; j < k; j++) inside[j] = h[j];
        };
        if (dim == 3) run(std::integral_constant<int, 3>{});
        else run(std::integral_constant<int, 2>{});
        return VOR_OK;
    });
}

vor_status vor_tree_edges_slab(vor_tree *t, const int64_t *global_index, const uint8_t *owned, size_t n, uint32_t **edges, size_t *n_edges) {
    return guarded([&]() -> vor_status {
        if (!t || !global_index || !owned || !edges || !n_edges) { g_err = "bad argument"; return VOR_ERR_ARG; }
        vor::be::set_device(t->device);
        return t->visit([&](auto &e) -> vor_status {
            if ((size_t)e.ninput != n) { g_err = "global_index / owned must cover every inserted point"; return VOR_ERR_ARG; }
            long long m = 0;
            static_assert(sizeof(long long) == sizeof(int64_t), "index width");
            *edges = e.slab_edges_to_host_block(reinterpret_cast<const long long *>(global_index), owned, &m);
            *n_edges = (size_t)m;
            return VOR_OK;
        });
    });
}

vor_status vor_debug_corrupt(vor_tree *t, int kind) {
    return guarded([&]() -> vor_status {
        if (!t) { g_err = "null tree"; return VOR_ERR_ARG; }
        vor::be::set_device(t->device);
        return t->visit([&](auto &e) -> vor_status { e.debug_corrupt(kind); return VOR_OK; });
    });
}

vor_status vor_tree_super_simplex(vor_tree *t, size_t set, double *super_vertices, double *center, double *radius) {
    return guarded([&]() -> vor_status {
        if (!t || set >= t->n_sets) { g_err = "bad argument"; return VOR_ERR_ARG; }
        return t->visit([&](auto &e) -> vor_status {
            const int D = e.M - 1;
            if (super_vertices) memcpy(super_vertices, &e.superXYZ[set * e.M * D], sizeof(double) * e.M * D);
            if (center) memcpy(center, &e.center[set * D], sizeof(double) * D);
            if (radius) *radius = e.radius[set];
            return VOR_OK;
        });
    });
}

vor_status vor_tree_stats(vor_tree *t, uint64_t *s) {
    return guarded([&]() -> vor_status {
        if (!t || !s) { g_err = "bad argument"; return VOR_ERR_ARG; }
        vor::be::set_device(t->device);
        return t->visit([&](auto &e) -> vor_status {
            e.pull_counters();
            const vor::Counters &c = *e.hcnt;
            s[0] = e.rs.rounds; s[1] = e.rs.attempts; s[2] = e.rs.winners; s[3] = e.rs.owner_resets; s[4] = e.rs.compactions; s[5] = e.rs.stages;
            s[6] = c.walk_steps; s[7] = c.tests; s[8] = c.killed; s[9] = c.created; s[10] = c.exact_calls; s[11] = c.exact_zero;
            s[12] = (uint64_t)c.ndup; s[13] = (uint64_t)c.ntets; s[14] = c.aborted; s[15] = c.tests_ok; s[16] = c.sph_undecided;
            s[17] = (uint64_t)c.nflag_set; s[18] = e.rs.slots;
            return VOR_OK;
        });
    });
}

vor_status vor_tree_profile(vor_tree *t, double *out) {
    return guarded([&]() -> vor_status {
        if (!t || !out) { g_err = "bad argument"; return VOR_ERR_ARG; }
        return t->visit([&](auto &e) -> vor_status {
            for (int i = 0; i < 4; i++) { out[i] = e.prof.ms[i]; out[4 + i] = e.prof.cnt[i]; }
            return VOR_OK;
        });
    });
}

// ---- geometry batches
vor_status vor_circumsphere(int dim, const double *verts, size_t n, double *centers, double *radii, int device) {
    return guarded([&]() -> vor_status {
        if (dim != 2 && dim != 3) { g_err = "bad dim"; return VOR_ERR_ARG; }
        vor::be::set_device(device);
        const vor::be::Stream st{};
        const size_t M = dim + 1;
        DevBuf dv(sizeof(double) * n * M * dim), dc(sizeof(double) * n * dim), dr(sizeof(double) * n);
        vor::be::h2d(dv.p, verts, sizeof(double) * n * M * dim, st);
        if (dim == 3) {
            vor::CircumBatchArgs<3> a{(const double *)dv.p, (double *)dc.p, (double *)dr.p};
            VOR_LAUNCH(vor::CircumBatchArgs<3>, vor::circum_batch_body, n, a, st);
        } else {
            vor::CircumBatchArgs<2> a{(const double *)dv.p, (double *)dc.p, (double *)dr.p};
            VOR_LAUNCH(vor::CircumBatchArgs<2>, vor::circum_batch_body, n, a, st);
        }
        vor::be::d2h(centers, dc.p, sizeof(double) * n * dim, st);
        vor::be::d2h(radii, dr.p, sizeof(double) * n, st);
        vor::be::sync(st);
        return VOR_OK;
    });
}

vor_status vor_in_sphere(int dim, const double *p, const double *c, const double *r, size_t n, int32_t *out, int device) {
    return guarded([&]() -> vor_status {
        if (dim != 2 && dim != 3) { g_err = "bad dim"; return VOR_ERR_ARG; }
        vor::be::set_device(device);
        const vor::be::Stream st{};
        DevBuf dp(sizeof(double) * n * dim), dc(sizeof(double) * n * dim), dr(sizeof(double) * n), dout(sizeof(int) * n);
        vor::be::h2d(dp.p, p, sizeof(double) * n * dim, st);
        vor::be::h2d(dc.p, c, sizeof(double) * n * dim, st);
        vor::be::h2d(dr.p, r, sizeof(double) * n, st);
        vor::InSphereBatchArgs a{(const double *)dp.p, (const double *)dc.p, (const double *)dr.p, (int *)dout.p, dim};
        VOR_LAUNCH(vor::InSphereBatchArgs, vor::in_sphere_batch_body, n, a, st);
        vor::be::d2h(out, dout.p, sizeof(int) * n, st);
        vor::be::sync(st);
        return VOR_OK;
    });
}

vor_status vor_bounding_sphere(int dim, const double *points, size_t n, double *center, double *radius, int device) {
    vor_tree *t = nullptr;
    vor_status s = vor_tree_create(dim, points, n, device, &t);
    if (s != VOR_OK) return s;
    s = vor_tree_super_simplex(t, 0, nullptr, center, nullptr);
    if (radius) *radius = t->visit([&](auto &e) { return e.radiusBase[0]; }); // before the 10x of delaunay_tree.rs:393
    vor_tree_destroy(t);
    return s;
}

vor_status vor_predicates(int kind, const double *rows, size_t n, int32_t *out, uint64_t *n_exact, int device) {
    return guarded([&]() -> vor_status {
        static const int width[4] = {6, 12, 8, 15};
        if (kind < 0 || kind > 3) { g_err = "bad kind"; return VOR_ERR_ARG; }
        vor::be::set_device(device);
        const vor::be::Stream st{};
        DevBuf dr(sizeof(double) * n * width[kind]), dout(sizeof(int) * n), dc(sizeof(vor::Counters));
        vor::be::h2d(dr.p, rows, sizeof(double) * n * width[kind], st);
        vor::be::dmemset(dc.p, 0, sizeof(vor::Counters), st);
        vor::PredBatchArgs a{(const double *)dr.p, (int *)dout.p, (vor::Counters *)dc.p, kind};
        VOR_LAUNCH(vor::PredBatchArgs, vor::pred_batch_body, n, a, st);
        vor::Counters hc;
        vor::be::d2h(out, dout.p, sizeof(int) * n, st);
        vor::be::d2h(&hc, dc.p, sizeof(hc), st);
        vor::be::sync(st);
        if (n_exact) *n_exact = hc.exact_calls;
        if (hc.err == vor::ERR_RANGE) { g_err = "coordinate range exceeds the exact-arithmetic capacity"; return VOR_ERR_RANGE; }
        return VOR_OK;
    });
}

vor_status vor_sphere_filter(int dim, const double *origin, double reach, const double *rows, size_t n, int32_t *out, float *blocks, int device) {
    return guarded([&]() -> vor_status {
        if ((dim != 2 && dim != 3) || !origin || !rows || !out) { g_err = "bad argument"; return VOR_ERR_ARG; }
        vor::be::set_device(device);
        const vor::be::Stream st{};
        const size_t w = dim == 3 ? 15 : 8;
        DevBuf dr(sizeof(double) * n * w), dout(sizeof(int) * n), db(sizeof(float) * n * 5);
        vor::be::h2d(dr.p, rows, sizeof(double) * n * w, st);
        vor::SphereRef ref{origin[0], origin[1], dim == 3 ? origin[2] : 0.0, 8.0 * vor::SPH_EPS * reach};
        vor::SphereBatchArgs a{(const double *)dr.p, (int *)dout.p, blocks ? (float *)db.p : nullptr, ref, dim};
        VOR_LAUNCH(vor::SphereBatchArgs, vor::sphere_batch_body, n, a, st);
        vor::be::d2h(out, dout.p, sizeof(int) * n, st);
        if (blocks) vor::be::d2h(blocks, db.p, sizeof(float) * n * 5, st);
        vor::be::sync(st);
        return VOR_OK;
    });
}

} // extern "C"
