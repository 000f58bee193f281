// engine.cuh -- host side of the device engine: store management, bootstrap, BRIO ordering, the round loop and
// the output passes.  Backend-agnostic (backend_cuda.cuh in the product, tests/emu/backend_emu.h in unit tests).
//
// Reference counterparts:
//   Engine::create   DelaunayTree::new          /root/reference/src/delaunay_tree.rs:390-510 (3D), :545-640 (2D)
//   Engine::insert   add_points_to_tree         /root/reference/src/delaunay_tree.rs:336-386
//                    (and the sequential loop of lib.rs:110-120: same result, the triangulation is unique)
//   Engine::edges    Delaunay-graph extraction  SURVEY.md §8a row G (implicit in lib.rs:73-101)
//   Engine::validate check_delaunay             /root/reference/src/delaunay_tree.rs:512-541
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "coop_kernels.cuh"
#include "output_kernels.cuh"
#include "setup_kernels.cuh"

namespace vor {

struct EngineOptions {
    int slot_cap = 1 << 19;      // attempt slots per round (scratch size)
    int big_slots = 256;         // overflow slots per round
    int big_capk = 8192;         // killed capacity of an overflow slot
    int capk = 64;               // killed capacity of a regular slot
    int capb = 132;              // boundary capacity of a regular slot (2*capk + 4)
    int min_attempt = 8192;      // attempt at least this many points per round (keeps the SMs busy)
    double attempt_div = 64.0;   // otherwise attempt about (inserted vertices)/attempt_div points per round: about n/60
                                 // disjoint footprints fit into a mesh of n vertices, so denser attempts only lose (measured)
    int stage0 = 256;            // size of the first stage
    int stage_log = 1;           // stage sizes grow by 2^stage_log
    int stats = 0;               // accumulate W/E/K/C counters (atomics; keep off when timing)
    int verbose = 0;
    int profile = 0;             // CUDA-event time per kernel class (attempt / check / retri / setup)
    int coop = 1;                // lane-group cooperative kernels (GPU build); 0 = thread-per-point bodies
    int bulk_locate = 1;         // locate every point of a stage at its start (thread per point) instead of in its first attempt
    int select_mode = 1;         // 1 = stratified selection along the Morton-ordered active list, 0 = random subset
    int rounds_per_sync = 8;     // rounds launched back to back between two host read-backs of the counters
    int commit_smem = 1;         // commit retriangulates cavities staged in shared memory (0 = through the global store)
    int edge_wedge = 0;          // 3D edge list: 1 = the simplex whose wedge at the edge contains a fixed direction emits it (edges_wedge_body),
                                 // 0 = pivot around every edge.  Parity-green with provoked ties, measured no faster: count pass 16.7 vs 15.9 ms
                                 // per 10M points (4 random vertex gathers and 128 registers against the pivots' L2-local record gathers)
    double edge_dir[3] = {0.0, 0.0, 0.0};   // test hook: direction of the wedge test instead of the generic one (ties can be provoked with it)
    int mid_twin = 1;            // allow the switch to the hot kernel's twin with the FP64 determinant stage (see run_stage_pipelined)
    int split_exact = 1;         // attempt kernel as a hot twin without exact predicates + an exact twin for the points it flags
    int red = 1;                 // kill reservation as a fire-and-forget reduction (match.any dedup), see k_attempt_coop
    int pdl = 131072;            // rounds of at most this many slots launch their kernels with programmatic dependent launch (the next
                                 // kernel's blocks are set up while the current one drains: -15..20 % on 100k..1M points; on large
                                 // rounds the early-resident blocks cost more than the launch gap: 10M points 128 ms with every round, 117 without,
                                 // 112.6 with this threshold)
    int subround = 0;            // > 0: rounds larger than this many slots run as spatially contiguous sub-rounds of about this size (measured
                                 // slower at every size: 16k 144 ms, 32k 129, 64k 120, off 114 -- launch tails outweigh the L2 reuse)
    int tiled = 0;               // round kernels as resident blocks pulling tiles of slots from a device queue (measured slower: off)
    int smem_pad = 0;            // diagnostics: extra dynamic shared memory per block of the round kernels (caps resident warps)
    int persist_waves = 1000000; // grid of the round kernels = resident blocks x this (1 = persistent warps; large = one block per slot pair)
    double carry_frac = 0.125;   // a stage that is not the last one of its insert call stops once fewer than this fraction of its points is
                                 // pending: the stragglers join the active list of the next stage instead of costing the stage a tail of
                                 // nearly empty rounds at full round latency (the triangulation does not depend on the insertion order)
    double compact_frac = 0.85;  // the active list is compacted (at a host read-back) once fewer than this fraction of it is pending
    double tet_factor = 0.0;     // simplex slots per vertex (0 = default: 31 in 3D, 7.5 in 2D)
};

inline void options_from_env(EngineOptions &o) {
    if (const char *e = getenv("VOR_SLOT_CAP")) o.slot_cap = atoi(e);
    if (const char *e = getenv("VOR_MIN_ATTEMPT")) o.min_attempt = atoi(e);
    if (const char *e = getenv("VOR_ATTEMPT_DIV")) o.attempt_div = atof(e);
    if (const char *e = getenv("VOR_STAGE0")) o.stage0 = atoi(e);
    if (const char *e = getenv("VOR_STAGE_LOG")) o.stage_log = atoi(e);
    if (const char *e = getenv("VOR_STATS")) o.stats = atoi(e);
    if (const char *e = getenv("VOR_VERBOSE")) o.verbose = atoi(e);
    if (const char *e = getenv("VOR_TET_FACTOR")) o.tet_factor = atof(e);
    if (const char *e = getenv("VOR_COMPACT_FRAC")) o.compact_frac = atof(e);
    if (const char *e = getenv("VOR_CARRY_FRAC")) o.carry_frac = atof(e);
    if (const char *e = getenv("VOR_PERSIST_WAVES")) o.persist_waves = std::max(1, atoi(e));
    if (const char *e = getenv("VOR_SMEM_PAD")) o.smem_pad = atoi(e);
    if (const char *e = getenv("VOR_TILED")) o.tiled = atoi(e);
    if (const char *e = getenv("VOR_SUBROUND")) o.subround = atoi(e);
    if (const char *e = getenv("VOR_PDL")) o.pdl = atoi(e);
    if (const char *e = getenv("VOR_COOP")) o.coop = atoi(e);
    if (const char *e = getenv("VOR_ROUNDS_PER_SYNC")) o.rounds_per_sync = atoi(e);
    if (const char *e = getenv("VOR_SELECT_MODE")) o.select_mode = atoi(e);
    if (const char *e = getenv("VOR_BULK_LOCATE")) o.bulk_locate = atoi(e);
    if (const char *e = getenv("VOR_RED")) o.red = atoi(e);
    if (const char *e = getenv("VOR_COMMIT_SMEM")) o.commit_smem = atoi(e);
    if (const char *e = getenv("VOR_SPLIT_EXACT")) o.split_exact = atoi(e);
    if (const char *e = getenv("VOR_MID_TWIN")) o.mid_twin = atoi(e);
    if (const char *e = getenv("VOR_EDGE_WEDGE")) o.edge_wedge = atoi(e);
    if (const char *e = getenv("VOR_CAPK")) { o.capk = atoi(e); o.capb = 2 * o.capk + 4; }
}

struct EngineError {
    int code;
    std::string msg;
};

struct RunStats {
    unsigned long long rounds = 0, attempts = 0, winners = 0, owner_resets = 0, compactions = 0, stages = 0, slots = 0;
};

template <int D> class Engine {
  public:
    using Pt = typename Dim<D>::Pt;
    static constexpr int M = D + 1;

    EngineOptions opt;
    be::Stream stream;
    int nsets = 1;
    int nsuper = 0;
    // host copies of the bootstrap data (per set)
    std::vector<double> boxLo, boxHi, center, radius, radiusBase, superXYZ;
    std::vector<int> setInserted;
    // device store
    Mesh<D> mesh{};
    int vcap = 0, nv = 0;
    int *inputIdx = nullptr, *vidOfInput = nullptr;
    uint64_t *keysAll = nullptr;
    int ninput = 0, incap = 0;
    double *d_boxLo = nullptr, *d_boxHi = nullptr;
    Counters *hcnt = nullptr; // pinned host mirror
    // scratch
    Scratch scr{};
    bool slowPending = false, splitDisabled = false, midTwin = false;
    int flagPending = 0;
    int occHot = -1, occCommit = -1, occTiledHot = 1, occTiledCommit = 1;   // resident blocks of the round kernels on this device
    int tileNow = 4;
    int *act = nullptr, *act2 = nullptr, *blockCnt = nullptr, *slowFlag = nullptr;
    long long *d_misc = nullptr;
    long long insertedTotal = 0;
    long long remainingInCall = 0;   // points of the current insert call not inserted yet
    int maxStageCall = 0;            // largest stage of the current insert call
    int carried = 0;                 // pending entries at the front of `act` handed over by the previous stage (pipelined path)
    int actcap = 0;
    // key layout
    int setBits = 0, axisBits = 0;
    // priority / epoch state
    int bits = 1, epoch = 0, epochMax = 0;
    // seeding reference stage (vertex range)
    int refLo = 0, refHi = 0;
    uint64_t callSalt = 0x5851F42D4C957F2DULL;
    RunStats rs;
    be::Prof prof;
    // cached outputs
    uint32_t *d_edges = nullptr;
    long long nedges = -1;

    explicit Engine(be::Stream s, const EngineOptions &o) : opt(o), stream(s) {
        hcnt = (Counters *)be::hmalloc_pinned(sizeof(Counters));
        memset(hcnt, 0, sizeof(Counters));
        d_misc = (long long *)be::dmalloc(sizeof(long long) * 8);
        prof.on = opt.profile != 0;
    }
    ~Engine() {
        be::dfree(mesh.pts); be::dfree(mesh.tet); if (!VOR_INTERLEAVE) be::dfree(mesh.owner); be::dfree(mesh.seed);
        be::dfree(mesh.ptTet); be::dfree(mesh.cnt); be::dfree(inputIdx); be::dfree(vidOfInput); be::dfree(keysAll);
        be::dfree(d_boxLo); be::dfree(d_boxHi); be::dfree(act); be::dfree(act2); be::dfree(blockCnt); be::dfree(d_edges); be::dfree(slowFlag);
        free_scratch();
        be::dfree(d_misc);
        be::hfree_pinned(hcnt);
    }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;

    [[noreturn]] static void fail(int code, const std::string &m) { throw EngineError{code, m}; }

    // ------------------------------------------------------------------ memory helpers
    template <class T> void grow(T *&p, size_t oldn, size_t newn) {
        T *q = (T *)be::dmalloc(sizeof(T) * newn);
        if (p && oldn) be::d2d(q, p, sizeof(T) * oldn, stream);
        be::sync(stream);
        be::dfree(p);
        p = q;
    }
    void fill_i(int *p, int val, size_t n) {
        size_t done = 0;
        while (done < n) { // launches are int-indexed
            const size_t c = std::min(n - done, (size_t)1 << 30);
            FillArgs a{p + done, val};
            VOR_LAUNCH(FillArgs, fill_body, c, a, stream);
            done += c;
        }
    }
    void ensure_vertices(int need) {
        if (need <= vcap) return;
        const int nc = std::max(need, vcap + vcap / 2);
        grow(mesh.pts, (size_t)nv, (size_t)nc);
        grow(mesh.seed, (size_t)nv, (size_t)nc);
        grow(mesh.ptTet, (size_t)nv, (size_t)nc);
        grow(inputIdx, (size_t)nv, (size_t)nc);
        grow(keysAll, (size_t)std::max(nv - nsuper, 0), (size_t)nc);
        {   // flags of the exact twin: zero for every vertex that is not flagged (caching allocator: blocks come back dirty)
            int *q = (int *)be::dmalloc(sizeof(int) * (size_t)nc);
            be::dmemset(q, 0, sizeof(int) * (size_t)nc, stream);
            if (slowFlag && nv) be::d2d(q, slowFlag, sizeof(int) * (size_t)nv, stream);
            be::sync(stream);
            be::dfree(slowFlag);
            slowFlag = q;
        }
        vcap = nc;
    }
    void ensure_inputs(int need) {
        if (need <= incap) return;
        const int nc = std::max(need, incap + incap / 2);
        grow(vidOfInput, (size_t)ninput, (size_t)nc);
        incap = nc;
    }
    void ensure_simplices(long long need) {
        if (need <= mesh.cap) return;
        if (need > (1LL << 29) - 1) fail(ERR_OOM, "more than 2^29 simplex slots needed");
        long long nc = std::max(need, (long long)mesh.cap + mesh.cap / 2);
        nc = std::min(nc, (1LL << 29) - 1);
        const int old = mesh.cap;
        // record + ownership / sphere block per simplex (sphere.cuh: one 64 B line when interleaved); a slot is written whole
        // by the round that creates the simplex (store_rec, store_blk) and nothing reads it before, so new capacity needs
        // no initialisation
        grow(mesh.tet, REC4 * (size_t)old, REC4 * (size_t)nc);
        if (VOR_INTERLEAVE) mesh.owner = reinterpret_cast<int *>(mesh.tet);
        else grow(mesh.owner, OWS * (size_t)old, OWS * (size_t)nc);
        mesh.cap = (int)nc;
        if (opt.verbose) fprintf(stderr, "[vor] simplex capacity -> %lld\n", nc);
    }
    void free_scratch() {
        be::dfree(scr.killed); be::dfree(scr.bfacet); be::dfree(scr.bouter); be::dfree(scr.slotAct); be::dfree(scr.slotNk);
        be::dfree(scr.slotNb); be::dfree(scr.slotStatus); be::dfree(scr.slotBig); be::dfree(scr.bigK); be::dfree(scr.bigF);
        be::dfree(scr.bigO); be::dfree(scr.winners); be::dfree(scr.wbase); be::dfree(scr.slowSlots); be::dfree(scr.slotInfo);
        scr = Scratch{};
    }
    void ensure_scratch(int nslots) {
        if (nslots <= scr.nslots) return;
        free_scratch();
        scr.nslots = nslots;
        scr.capk = opt.capk;
        scr.capb = opt.capb;
        scr.killed = (int *)be::dmalloc(sizeof(int) * (size_t)scr.capk * nslots);
        scr.bfacet = (int *)be::dmalloc(sizeof(int) * (size_t)scr.capb * nslots);
        scr.bouter = (int *)be::dmalloc(sizeof(int) * (size_t)scr.capb * nslots);
        scr.slotAct = (int *)be::dmalloc(sizeof(int) * (size_t)nslots);
        scr.slotNk = (int *)be::dmalloc(sizeof(int) * (size_t)nslots);
        scr.slotNb = (int *)be::dmalloc(sizeof(int) * (size_t)nslots);
        scr.slotStatus = (int *)be::dmalloc(sizeof(int) * (size_t)nslots);
        scr.slotBig = (int *)be::dmalloc(sizeof(int) * (size_t)nslots);
        scr.winners = (int *)be::dmalloc(sizeof(int) * (size_t)nslots);
        scr.wbase = (int *)be::dmalloc(sizeof(int) * (size_t)nslots);
        scr.slowSlots = (int *)be::dmalloc(sizeof(int) * (size_t)nslots);
        scr.slotInfo = (int4 *)be::dmalloc(sizeof(int4) * (size_t)nslots);
        scr.nbig = opt.big_slots;
        scr.bigCapK = opt.big_capk;
        scr.bigCapB = 2 * opt.big_capk + 4;
        scr.bigK = (int *)be::dmalloc(sizeof(int) * (size_t)scr.nbig * scr.bigCapK);
        scr.bigF = (int *)be::dmalloc(sizeof(int) * (size_t)scr.nbig * scr.bigCapB);
        scr.bigO = (int *)be::dmalloc(sizeof(int) * (size_t)scr.nbig * scr.bigCapB);
    }
    void pull_counters() {
        be::d2h(hcnt, mesh.cnt, sizeof(Counters), stream);
        be::sync(stream);
    }
    // (win_total, created_all) = direct counters of the thread-per-point path + the privatised parts of k_commit_coop
    unsigned long long win_total() const {
        unsigned long long w = hcnt->win_total;
        for (int p = 0; p < NPART; p++) w += hcnt->part[p][0] >> 40;
        return w;
    }
    unsigned long long created_all() const {
        unsigned long long c = hcnt->created_all;
        for (int p = 0; p < NPART; p++) c += hcnt->part[p][0] & ((1ULL << 40) - 1ULL);
        return c;
    }
    void push_counters() { be::h2d(mesh.cnt, hcnt, sizeof(Counters), stream); }
    void check_device_error(const char *where) {
        if (hcnt->err) {
            static const char *names[] = {"ok", "no conflict", "degenerate", "duplicate point", "cuda", "out of memory",
                                          "cavity exceeds overflow scratch", "coordinate range exceeds exact arithmetic capacity",
                                          "point outside the super simplex", "walk did not terminate", "bad argument"};
            const int c = hcnt->err;
            fail(c, std::string(where) + ": " + (c >= 0 && c <= 10 ? names[c] : "error"));
        }
    }

    // ------------------------------------------------------------------ bootstrap  (DelaunayTree::new)
    // d_in: n x D points on the device (all the points the tree will ever see, as in the reference);
    // h_setOff: nsets+1 offsets (nullptr => one set)
    // forcedLo/forcedHi/forcedOutside (single set): the bounding box and the "some point fails the strict in_sphere test of
    // the half-diagonal sphere" count come from the caller instead of from d_in -- a slab of a point set that is spread
    // over several GPUs bootstraps from the GLOBAL bounds so that every slab builds the same super simplex (d_in may be
    // null, n is a capacity hint).
    void create(const double *d_in, int n, const int *h_setOff, int nsets_, const double *forcedLo = nullptr, const double *forcedHi = nullptr,
                int forcedOutside = 0) {
        nsets = nsets_ < 1 ? 1 : nsets_;
        if (forcedLo && nsets != 1) fail(ERR_ARG, "forced bounds address single-set trees");
        nsuper = nsets * M;
        std::vector<int> off(nsets + 1);
        if (h_setOff) off.assign(h_setOff, h_setOff + nsets + 1);
        else { off[0] = 0; off[1] = n; }
        // chunk table
        std::vector<ChunkDesc> chunks;
        for (int s = 0; s < nsets; s++)
            for (int lo = off[s]; lo < off[s + 1]; lo += BBOX_CHUNK) chunks.push_back(ChunkDesc{s, lo, std::min(lo + BBOX_CHUNK, off[s + 1])});
        const int nch = (int)chunks.size();
        ChunkDesc *d_chunks = (ChunkDesc *)be::dmalloc(sizeof(ChunkDesc) * (size_t)std::max(nch, 1));
        double *d_partial = (double *)be::dmalloc(sizeof(double) * 2 * D * (size_t)std::max(nch, 1));
        be::h2d(d_chunks, chunks.data(), sizeof(ChunkDesc) * (size_t)nch, stream);
        std::vector<double> partial((size_t)nch * 2 * D);
        if (!forcedLo) {
            BboxArgs<D> ba{d_in, d_chunks, d_partial};
            VOR_LAUNCH(BboxArgs<D>, bbox_chunk_body<D>, nch, ba, stream);
            be::d2h(partial.data(), d_partial, sizeof(double) * partial.size(), stream);
            be::sync(stream);
        }
        boxLo.assign((size_t)nsets * D, INFINITY);
        boxHi.assign((size_t)nsets * D, -INFINITY);
        if (forcedLo)
            for (int k = 0; k < D; k++) { boxLo[k] = forcedLo[k]; boxHi[k] = forcedHi[k]; }
        for (int c = 0; c < nch && !forcedLo; c++)
            for (int k = 0; k < D; k++) {
                const int s = chunks[c].set;
                boxLo[(size_t)s * D + k] = std::fmin(boxLo[(size_t)s * D + k], partial[(size_t)c * 2 * D + k]);
                boxHi[(size_t)s * D + k] = std::fmax(boxHi[(size_t)s * D + k], partial[(size_t)c * 2 * D + D + k]);
            }
        // bounding sphere (geometry.rs:99-142): centre, half diagonal, then the 1.5x rule
        center.assign((size_t)nsets * D, 0.0);
        radius.assign((size_t)nsets, 0.0);
        radiusBase.assign((size_t)nsets, 0.0);
        for (int s = 0; s < nsets; s++) {
            double ud = 0.0, ld = 0.0;
            for (int k = 0; k < D; k++) center[(size_t)s * D + k] = (boxHi[(size_t)s * D + k] + boxLo[(size_t)s * D + k]) / 2.0;
            for (int k = 0; k < D; k++) { const double d = boxHi[(size_t)s * D + k] - center[(size_t)s * D + k]; ud += d * d; }
            for (int k = 0; k < D; k++) { const double d = boxLo[(size_t)s * D + k] - center[(size_t)s * D + k]; ld += d * d; }
            ud = std::sqrt(ud);
            ld = std::sqrt(ld);
            radius[s] = ud > ld ? ud : ld;
        }
        double *d_center = (double *)be::dmalloc(sizeof(double) * (size_t)nsets * D);
        double *d_radius = (double *)be::dmalloc(sizeof(double) * (size_t)nsets);
        int *d_outside = (int *)be::dmalloc(sizeof(int) * (size_t)nsets);
        be::h2d(d_center, center.data(), sizeof(double) * (size_t)nsets * D, stream);
        be::h2d(d_radius, radius.data(), sizeof(double) * (size_t)nsets, stream);
        be::dmemset(d_outside, 0, sizeof(int) * (size_t)nsets, stream);
        std::vector<int> outside(nsets);
        if (!forcedLo) {
            OutsideArgs<D> oa{d_in, d_chunks, d_center, d_radius, d_outside};
            VOR_LAUNCH(OutsideArgs<D>, count_outside_body<D>, nch, oa, stream);
            be::d2h(outside.data(), d_outside, sizeof(int) * (size_t)nsets, stream);
        } else outside[0] = forcedOutside;
        be::sync(stream);
        be::dfree(d_chunks); be::dfree(d_partial); be::dfree(d_center); be::dfree(d_radius); be::dfree(d_outside);
        // super simplex (delaunay_tree.rs:392-406 / :547-558); host libm cos/sin as rustc's f64::cos/sin
        superXYZ.assign((size_t)nsets * M * D, 0.0);
        const double PI = 3.14159265358979323846264338327950288;
        const double a1 = 2. * PI / 3., a2 = 4. * PI / 3.;
        for (int s = 0; s < nsets; s++) {
            if (outside[s] > 0) radius[s] = radius[s] * 1.5;
            radiusBase[s] = radius[s];
            radius[s] *= 10.0;
            // a single point / coincident points / empty set: zero or non-finite radius, the super simplex collapses
            // (the reference panics there with "No simplex found", delaunay_tree.rs:47-54)
            if (!(radius[s] > 0.0) || !std::isfinite(radius[s])) fail(ERR_DEGENERATE, "bounding sphere of the point set has zero or non-finite radius");
            const double r = radius[s];
            const double *c = &center[(size_t)s * D];
            double *sv = &superXYZ[(size_t)s * M * D];
            if (D == 3) {
                sv[0] = c[0]; sv[1] = c[1]; sv[2] = c[2] + r;
                sv[3] = c[0] + r; sv[4] = c[1]; sv[5] = c[2] - r;
                sv[6] = c[0] + r * std::cos(a1); sv[7] = c[1] + r * std::sin(a1); sv[8] = c[2] - r;
                sv[9] = c[0] + r * std::cos(a2); sv[10] = c[1] + r * std::sin(a2); sv[11] = c[2] - r;
            } else {
                sv[0] = c[0] + r; sv[1] = c[1];
                sv[2] = c[0] + r * std::cos(a1); sv[3] = c[1] + r * std::sin(a1);
                sv[4] = c[0] + r * std::cos(a2); sv[5] = c[1] + r * std::sin(a2);
            }
        }
        // store
        setInserted.assign(nsets, 0);
        setBits = 0;
        while ((1 << setBits) < nsets) setBits++;
        axisBits = std::min(D == 3 ? 19 : 28, (STAGE_SHIFT - setBits) / D);
        mesh.nsuper = nsuper;
        mesh.cnt = (Counters *)be::dmalloc(sizeof(Counters));
        ensure_vertices(nsuper + n + 16);
        ensure_inputs(n + 16);
        const double tf = opt.tet_factor > 0 ? opt.tet_factor : (D == 3 ? 31.0 : 7.5);
        ensure_simplices((long long)(tf * (double)(n + 64)) + 4LL * nsets + 1024);
        d_boxLo = (double *)be::dmalloc(sizeof(double) * (size_t)nsets * D);
        d_boxHi = (double *)be::dmalloc(sizeof(double) * (size_t)nsets * D);
        be::h2d(d_boxLo, boxLo.data(), sizeof(double) * (size_t)nsets * D, stream);
        be::h2d(d_boxHi, boxHi.data(), sizeof(double) * (size_t)nsets * D, stream);
        // super vertices + root simplices (simplex s = root of set s), positively oriented
        std::vector<Pt> sp((size_t)nsuper);
        std::vector<int4> rtet(2 * (size_t)nsets);
        std::vector<int> rseed((size_t)nsuper, -1);
        Counters dummy;
        memset(&dummy, 0, sizeof(dummy));
        PredCtx cx{&dummy};
        for (int s = 0; s < nsets; s++) {
            for (int k = 0; k < M; k++) set_host_pt(sp[(size_t)s * M + k], &superXYZ[((size_t)s * M + k) * D]);
            int4 v;
            v.x = s * M; v.y = s * M + 1; v.z = s * M + 2; v.w = (D == 3) ? s * M + 3 : -1;
            if (orient_host(cx, &sp[(size_t)s * M]) < 0) std::swap(v.x, v.y);
            rtet[2 * (size_t)s] = v;
            rtet[2 * (size_t)s + 1] = int4{-1, -1, -1, -1};
        }
        // origin of the float sphere centres: centre of the union of the sets' boxes; every legal query point lies
        // inside a super simplex, so |p - origin| <= reach and fl(p - origin) is off by at most EPS * reach per axis
        {
            double ulo[3] = {INFINITY, INFINITY, INFINITY}, uhi[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int s = 0; s < nsets; s++)
                for (int k = 0; k < D; k++) {
                    ulo[k] = std::fmin(ulo[k], boxLo[(size_t)s * D + k]);
                    uhi[k] = std::fmax(uhi[k], boxHi[(size_t)s * D + k]);
                }
            double org[3] = {0.0, 0.0, 0.0}, reach = 0.0;
            for (int k = 0; k < D; k++) org[k] = 0.5 * (ulo[k] + uhi[k]);
            for (int s = 0; s < nsets; s++)
                for (int j = 0; j < M; j++) {
                    double d1 = 0.0;
                    for (int k = 0; k < D; k++) d1 += std::fabs(superXYZ[((size_t)s * M + j) * D + k] - org[k]);
                    reach = std::fmax(reach, d1);
                }
            mesh.sref.ox = org[0]; mesh.sref.oy = org[1]; mesh.sref.oz = org[2];
            mesh.sref.qerr = 8.0 * SPH_EPS * reach;
        }
        std::vector<int> rblk((size_t)8 * nsets);
        for (int s = 0; s < nsets; s++) {
            const SphereBlk sb = sphere_host(&sp[(size_t)s * M]);
            int *w = &rblk[(size_t)8 * s];
            w[0] = OWNER_FREE; w[1] = OWNER_FREE; w[2] = f2i(sb.cx); w[3] = f2i(sb.cy); w[4] = f2i(sb.cz); w[5] = f2i(sb.rin2); w[6] = f2i(sb.rout2); w[7] = 0;
        }
        std::vector<int4> rline;   // must outlive the asynchronous copy: synchronised at the end of create()
        if (VOR_INTERLEAVE) {
            rline.resize((size_t)REC4 * nsets);
            for (int s = 0; s < nsets; s++) {
                memcpy(&rline[(size_t)REC4 * s], &rblk[(size_t)8 * s], sizeof(int) * 8);
                rline[(size_t)REC4 * s + TVO4] = rtet[2 * (size_t)s];
                rline[(size_t)REC4 * s + TVO4 + 1] = rtet[2 * (size_t)s + 1];
            }
            be::h2d(mesh.tet, rline.data(), sizeof(int4) * rline.size(), stream);
        } else {
            be::h2d(mesh.owner, rblk.data(), sizeof(int) * rblk.size(), stream);
            be::h2d(mesh.tet, rtet.data(), sizeof(int4) * 2 * (size_t)nsets, stream);
        }
        be::h2d(mesh.pts, sp.data(), sizeof(Pt) * (size_t)nsuper, stream);
        be::h2d(mesh.seed, rseed.data(), sizeof(int) * (size_t)nsuper, stream);
        fill_i(mesh.ptTet, 0, (size_t)nsuper);
        fill_i(inputIdx, -1, (size_t)nsuper);
        nv = nsuper;
        memset(hcnt, 0, sizeof(Counters));
        hcnt->ntets = nsets;
        hcnt->sph_lo = nsets;
        push_counters();
        be::sync(stream);
    }
    static void set_host_pt(double4 &p, const double *s) { p.x = s[0]; p.y = s[1]; p.z = s[2]; p.w = 0.0; }
    static void set_host_pt(double2 &p, const double *s) { p.x = s[0]; p.y = s[1]; }
    static int orient_host(PredCtx &cx, const double4 *p) { return orient3d(cx, p[0], p[1], p[2], p[3]); }
    static int orient_host(PredCtx &cx, const double2 *p) { return orient2d(cx, p[0], p[1], p[2]); }
    SphereBlk sphere_host(const double4 *p) const { return sphere_make(p[0], p[1], p[2], p[3], mesh.sref); }
    SphereBlk sphere_host(const double2 *p) const { return sphere_make(p[0], p[1], p[2], mesh.sref); }

    // ------------------------------------------------------------------ insertion  (add_points_to_tree)
    // d_in: n x D points on the device; h_setOff: nsets+1 offsets into d_in (nullptr => one set)
    void insert(const double *d_in, int n, const int *h_setOff) {
        if (n <= 0) return;
        const auto tins0 = std::chrono::steady_clock::now();
        invalidate_outputs();
        std::vector<int> off(nsets + 1);
        if (h_setOff) off.assign(h_setOff, h_setOff + nsets + 1);
        else { off[0] = 0; off[1] = n; }
        const double tf = opt.tet_factor > 0 ? opt.tet_factor : (D == 3 ? 31.0 : 7.5);
        // Simplex slots are never recycled and a neighbour code packs (slot << 2 | facet) into an int: a tree holds at
        // most 2^29 - 1 slots, dead ones included (~27 slots are created per 3D point, ~6 per 2D point, over ALL sets and
        // ALL inserts of the tree).  Refuse BEFORE any state is advanced; bigger batches go through
        // vor_delaunay_batch_stream, which cuts them into stores of this size.
        if ((long long)hcnt->ntets + (long long)((tf - (D == 3 ? 3.5 : 1.2)) * (double)n) > (1LL << 29) - 1)
            fail(ERR_CAPACITY, "tree would exceed 2^29 simplex slots (about 19M 3D / 85M 2D points per tree, all sets and inserts together): "
                               "use vor_delaunay_batch_stream for larger batches");
        ensure_vertices(nv + n);
        ensure_inputs(ninput + n);
        ensure_simplices(std::min((long long)hcnt->ntets + (long long)(tf * (double)(n + 64)), (1LL << 29) - 1));

        // ---- keys, sort, gather
        std::vector<int> s0(nsets);
        for (int s = 0; s < nsets; s++) s0[s] = std::max(opt.stage0, setInserted[s]);
        int *d_off = (int *)be::dmalloc(sizeof(int) * (size_t)(nsets + 1));
        int *d_s0 = (int *)be::dmalloc(sizeof(int) * (size_t)nsets);
        uint64_t *k0 = (uint64_t *)be::dmalloc(sizeof(uint64_t) * (size_t)n);
        uint32_t *v0 = (uint32_t *)be::dmalloc(sizeof(uint32_t) * (size_t)n);
        uint32_t *v1 = (uint32_t *)be::dmalloc(sizeof(uint32_t) * (size_t)n);
        be::h2d(d_off, off.data(), sizeof(int) * (size_t)(nsets + 1), stream);
        be::h2d(d_s0, s0.data(), sizeof(int) * (size_t)nsets, stream);
        callSalt = mix64(callSalt + (uint64_t)ninput);
        prof.start(3, stream);
        KeyArgs<D> ka{d_in, d_off, d_boxLo, d_boxHi, d_s0, k0, v0, nsets, setBits, axisBits, callSalt, std::max(1, opt.stage_log)};
        VOR_LAUNCH(KeyArgs<D>, keys_body<D>, n, ka, stream);
        uint64_t *kdst = keysAll + (nv - nsuper);
        be::sort_pairs(k0, kdst, v0, v1, (size_t)n, stream);
        GatherArgs<D> ga{d_in, v1, mesh.pts, inputIdx, vidOfInput, mesh.seed, mesh.ptTet, nv, ninput};
        VOR_LAUNCH(GatherArgs<D>, gather_body<D>, n, ga, stream);
        int *d_stageLo = (int *)be::dmalloc(sizeof(int) * 64);
        be::dmemset(d_stageLo, 0xff, sizeof(int) * 64, stream);
        StageBoundsArgs sa{kdst, d_stageLo};
        VOR_LAUNCH(StageBoundsArgs, stage_bounds_body, n, sa, stream);
        prof.stop(stream);
        int stageLo[65];
        be::d2h(stageLo, d_stageLo, sizeof(int) * 64, stream);
        be::sync(stream);
        be::dfree(d_off); be::dfree(d_s0); be::dfree(k0); be::dfree(v0); be::dfree(v1); be::dfree(d_stageLo);
        stageLo[64] = n;
        for (int st = 63; st >= 0; st--)
            if (stageLo[st] < 0) stageLo[st] = stageLo[st + 1]; // empty stage

        const int vbase = nv;
        nv += n;
        const int inputBase = ninput;
        ninput += n;
        (void)inputBase;

        // ---- priority bits: wide enough for the largest stage of this call
        int maxStage = 1;
        for (int st = 0; st < 64; st++) maxStage = std::max(maxStage, stageLo[st + 1] - stageLo[st]);
        bits = 1;
        ensure_scratch(std::min(maxStage, opt.slot_cap));
#if VOR_GPU
        // cooperative path: the priority is a hash of the SLOT index (unique within a round), so the key needs
        // log2(slots) bits and the epoch field is wide: owner[] is reset about once per 1000 rounds
        const int prioRange = opt.coop ? std::min(maxStage, scr.nslots) : maxStage;
#else
        const int prioRange = maxStage;
#endif
        while ((1 << bits) < prioRange) bits++;
        if (bits + 1 > 28) fail(ERR_ARG, "stage too large for the priority key");
        epochMax = (1 << (30 - (bits + 1))) - 1;
        reset_owners();
        ensure_scratch(std::min(maxStage, opt.slot_cap));
        // active list: a stage plus the stragglers the stage before it handed over (at most carry_frac of that stage)
        maxStageCall = maxStage;
        const int actNeed = maxStage + (int)std::min<long long>((long long)(std::max(0.0, std::min(opt.carry_frac, 0.5)) * (double)maxStage) + 1024, maxStage);
        if (actNeed > actcap) {
            be::dfree(act); be::dfree(act2); be::dfree(blockCnt);
            actcap = actNeed;
            act = (int *)be::dmalloc(sizeof(int) * (size_t)actcap);
            act2 = (int *)be::dmalloc(sizeof(int) * (size_t)actcap);
            blockCnt = (int *)be::dmalloc(sizeof(int) * (size_t)(actcap / 256 + 2));
        }

        remainingInCall = n;
        if (opt.verbose) {
            be::sync(stream);
            fprintf(stderr, "[vor] insert: %d points, setup (alloc + keys + sort + gather) %.3f ms\n", n,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tins0).count());
        }
        // ---- stages
        int lastStage = 0;
        for (int st = 0; st < 64; st++)
            if (stageLo[st + 1] > stageLo[st]) lastStage = st;
        carried = 0;
        for (int st = 0; st < 64; st++) {
            const int lo = vbase + stageLo[st], hi = vbase + stageLo[st + 1];
            if (hi <= lo) continue;
            const auto t0 = std::chrono::steady_clock::now();
            const unsigned long long r0 = rs.rounds;
            run_stage(lo, hi, st == lastStage);
            if (opt.verbose) {
                be::sync(stream);
                const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
                fprintf(stderr, "[vor] stage %d: %d points, %llu rounds, %.3f ms (%.1f us/round, %.1f ns/point)\n", st, hi - lo, rs.rounds - r0, ms,
                        1e3 * ms / (double)std::max<unsigned long long>(rs.rounds - r0, 1), 1e6 * ms / (double)(hi - lo));
            }
            refLo = lo;
            refHi = hi;
            rs.stages++;
        }
        for (int s = 0; s < nsets; s++) setInserted[s] += off[s + 1] - off[s];
        pull_counters();
        prof.resolve(stream);
        check_device_error("insert");
    }

    void reset_owners() {
        ResetOwnerArgs ra{mesh.owner, mesh.cnt};
        VOR_LAUNCH(ResetOwnerArgs, reset_owner_body, mesh.cap, ra, stream);
        epoch = epochMax;
        rs.owner_resets++;
    }

    // ------------------------------------------------------------------ pipelined stage (GPU, cooperative kernels)
    // `rounds_per_sync` rounds are launched back to back; every kernel sizes itself from device-side counters, the
    // host reads the counters once per batch (pending count, allocator, errors) and grows / compacts there.
#if VOR_GPU
    // simplex slots to keep free for a batch of R rounds: every attempted point could win, but never more than the
    // points of this insert call that are still pending
    long long batch_margin(int R, int nsel, double newPerPoint) const {
        const double byRounds = (double)R * (double)nsel * newPerPoint;
        const double byRemaining = (double)remainingInCall * (newPerPoint * 0.85) + 65536.0;
        return (long long)std::min(byRounds, byRemaining) + 4096;
    }
    template <int G> void launch_round(const AttemptArgs<D> &aa, const CheckArgs<D> &ca, const RoundSel &sel, bool slowNow) {
        constexpr int HG = HotCfg<D>::G, CG = CommitCfg<D>::G;     // lanes per attempted point in the hot / commit kernels
        // resident warps: as many blocks as fit on the machine (occupancy queried once), each warp strides over the slots
        if (occHot < 0) {
            int dev = 0, nsm = 148, o1 = 1, o2 = 1;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o1, k_attempt_hot<D, HG, 0>, VOR_HOT_BLOCK, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o2, k_commit_coop<D, CG>, VOR_COOP_BLOCK, 0);
            occHot = std::max(1, o1) * nsm;
            occCommit = std::max(1, o2) * nsm;
            int o3 = 1, o4 = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o3, k_attempt_hot_tiled<D>, VOR_TILE_BLOCK, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o4, k_commit_tiled<D>, VOR_TILE_BLOCK, 0);
            occTiledHot = std::max(1, o3) * nsm;
            occTiledCommit = std::max(1, o4) * nsm;
        }
        const long long nslots = (long long)sel.last - sel.first;
        const bool pdlNow = nslots <= (long long)opt.pdl;
        const unsigned grid = (unsigned)std::min<long long>((nslots * CG + VOR_COOP_BLOCK - 1) / VOR_COOP_BLOCK, (long long)occCommit * opt.persist_waves);
        const unsigned agrid = (unsigned)((((long long)sel.last - sel.first) * G + VOR_ATTEMPT_BLOCK - 1) / VOR_ATTEMPT_BLOCK);
        prof.start(0, stream);
        if (opt.red && aa.slowFlag) {
            // hot twin without the exact predicates in its call tree; while flagged points are pending (host
            // knowledge, one read-back old) the exact twin follows and attempts only those
            if (opt.tiled) {
                // tile = slots per block visit: large rounds 128, small rounds down to one slot per warp
                const long long per = (long long)sel.nsel / std::max(1, occTiledHot);
                tileNow = 4;
                while (tileNow < VOR_TILE_BLOCK && tileNow < per) tileNow <<= 1;
                const long long ntiles = ((long long)sel.nsel + tileNow - 1) / tileNow;
                k_attempt_hot_tiled<D><<<(unsigned)std::min<long long>(ntiles, occTiledHot), VOR_TILE_BLOCK, (size_t)opt.smem_pad, stream>>>(aa, sel, tileNow);
            } else {
                const unsigned hgrid = (unsigned)std::min<long long>((nslots * HG + VOR_HOT_BLOCK - 1) / VOR_HOT_BLOCK, (long long)occHot * opt.persist_waves);
                // the twin with the FP64 determinant stage inside for input that keeps leaving the sphere filter (midTwin, decided at a read-back)
                if (midTwin) be::launch_pdl(pdlNow, k_attempt_hot<D, HG, 1>, hgrid, VOR_HOT_BLOCK, (size_t)opt.smem_pad, stream, aa, sel);
                else be::launch_pdl(pdlNow, k_attempt_hot<D, HG, 0>, hgrid, VOR_HOT_BLOCK, (size_t)opt.smem_pad, stream, aa, sel);
            }
            if (slowNow) {
                // the slots the hot kernel queued (points it flagged in earlier rounds); a small grid-stride launch
                AttemptArgs<D> as = aa;
                as.thr = 2u;
                be::launch_pdl(pdlNow, k_attempt_slow<D, 1>, (unsigned)std::min<long long>(agrid, 148LL * 8), VOR_ATTEMPT_BLOCK, (size_t)0, stream, as, sel);
                be::g_launches++;
            }
        } else if (opt.red) {
            k_attempt_coop<D, G, 1, 1><<<agrid, VOR_ATTEMPT_BLOCK, 0, stream>>>(aa, sel);
        } else {
            AttemptArgs<D> af = aa;
            af.slowFlag = nullptr;
            k_attempt_coop<D, G, 0, 1><<<agrid, VOR_ATTEMPT_BLOCK, 0, stream>>>(af, sel);
        }
        prof.stop(stream);
        prof.start(2, stream);
        if (opt.tiled && opt.red && aa.slowFlag) {
            const long long ntiles = ((long long)sel.nsel + tileNow - 1) / tileNow;
            k_commit_tiled<D><<<(unsigned)std::min<long long>(ntiles, occTiledCommit), VOR_TILE_BLOCK, (size_t)opt.smem_pad, stream>>>(
                ca, act, sel, (opt.stats ? 1 : 0) | (opt.commit_smem ? 0 : 2), tileNow);
        } else
            be::launch_pdl(pdlNow, k_commit_coop<D, CG>, grid, VOR_COOP_BLOCK, (size_t)opt.smem_pad, stream, ca, (const int *)act, sel,
                           (opt.stats ? 1 : 0) | (opt.commit_smem ? 0 : 2));
        prof.stop(stream);
        prof.start(1, stream);
        {
            // new simplices of the round: at most ~36 per attempted point; grid-stride over whatever the allocator handed out
            const long long want = (nslots * (D == 3 ? 36 : 9) + 255) / 256;
            be::launch_pdl(pdlNow, k_spheres<D>, (unsigned)std::max(1LL, std::min(want, 148LL * 16)), 256, (size_t)0, stream, mesh);
        }
        prof.stop(stream);
        be::g_launches += 3;
        be::check_launch("round kernels");
    }
    void run_stage_pipelined(int lo, int hi, bool last) {
        // the active list: the stragglers of the stage before (already at the front of `act`, with seeds), then this stage's points
        const int fresh = hi - lo;
        const int total = carried + fresh;
        int nact = total, pending = total;
        // a stage that is not the last one hands its last few pending points over to the next stage
        const int carryLimit = last ? 0 : std::min((int)(std::max(0.0, std::min(opt.carry_frac, 0.5)) * (double)fresh), actcap - maxStageCall);
        SeedArgs<D> sd{keysAll, mesh.pts, mesh.ptTet, mesh.owner, mesh.seed, nsuper, lo, refLo, refHi, D * axisBits};
        VOR_LAUNCH(SeedArgs<D>, init_seeds_body<D>, fresh, sd, stream);
        if (opt.bulk_locate) {
            LocateArgs<D> la{mesh, lo, opt.stats};
            VOR_LAUNCH(LocateArgs<D>, locate_body<D>, fresh, la, stream);
        }
        IotaArgs ia{act + carried, lo};
        VOR_LAUNCH(IotaArgs, iota_body, fresh, ia, stream);
        carried = 0;
        pull_counters();
        const unsigned long long win0 = win_total();
        const int dup0 = hcnt->ndup;
        int stall = 0;
        uint32_t roundSalt = (uint32_t)mix64((uint64_t)lo * 0x9E37u + rs.rounds);
        const double newPerPoint = D == 3 ? 36.0 : 9.0;   // allocator margin per attempted point (mean is 27 / 6)
        while (pending > std::max(0, carryLimit)) {
            const int R = opt.verbose > 2 ? 1 : std::max(1, opt.rounds_per_sync);
            // attempt about max(min_attempt, inserted/attempt_div) of the pending points per round, one per run of
            // `stride` consecutive entries of the active list (which also holds the entries inserted since the last
            // compaction)
            const double target = std::max((double)opt.min_attempt, (double)insertedTotal / opt.attempt_div);
            int stride = 1;
            if ((double)pending > target) stride = std::max(1, (int)std::floor((double)pending / target));
            stride = std::max(stride, (nact + scr.nslots - 1) / scr.nslots);
            const int nsel = (nact + stride - 1) / stride;
            ensure_simplices((long long)hcnt->ntets + batch_margin(R, std::min(nsel, pending), newPerPoint));
            for (int r = 0; r < R; r++) {
                if (epoch <= 0) reset_owners();
                roundSalt = roundSalt * 1664525u + 1013904223u;
                const int keybase = epoch << (bits + 1);
                const RoundSel sel{nact, stride, (int)((roundSalt >> 8) % (uint32_t)stride), nsel, 0, nsel};
                // hot / exact twins of the attempt kernel -- unless this input keeps leaving the FP64 filter (near-degenerate:
                // thousands of flagged points), where one kernel with the exact path inside is the better deal
                const bool split = opt.split_exact && !splitDisabled;
                AttemptArgs<D> aa{mesh, scr, act, bits, roundSalt, 0u, stride, sel.offset, keybase, opt.stats, split ? slowFlag : nullptr};
                CheckArgs<D> ca{mesh, scr, bits, roundSalt, keybase, split ? slowFlag : nullptr};
                // the exact twin is pure latency for a handful of points (~26 us per launch): it runs once per batch of
                // rounds, and in every round once the flagged points are most of what is left of the stage
                const bool slowNow = slowPending && (r == R - 1 || 4LL * flagPending >= (long long)pending);
                if (opt.subround > 0 && nsel > opt.subround + opt.subround / 2 && !opt.tiled) {
                    // a large round as a sequence of spatially contiguous sub-rounds (slots are in Morton order): what the
                    // attempt kernel of a sub-round pulled into L2 is still there for its commit and sphere kernels
                    const int parts = (nsel + opt.subround - 1) / opt.subround;
                    for (int q = 0; q < parts; q++) {
                        RoundSel sub = sel;
                        sub.first = (int)((long long)nsel * q / parts);
                        sub.last = (int)((long long)nsel * (q + 1) / parts);
                        launch_round<32>(aa, ca, sub, slowNow && q == parts - 1);
                    }
                } else
                    launch_round<32>(aa, ca, sel, slowNow);
                epoch--;
                rs.rounds++;
                rs.slots += (unsigned long long)nsel;
            }
            pull_counters();
            check_device_error("round");
            slowPending = hcnt->nflag_set > hcnt->nflag_done;
            flagPending = hcnt->nflag_set - hcnt->nflag_done;
            // an input that keeps leaving the filters (near-degenerate: the jittered lattice) is better off with the one
            // kernel that has the exact path inside
            if ((long long)hcnt->nflag_set > 512 + insertedTotal / 32) splitDisabled = true;
            // ... and one that leaves the SPHERE filter more often than uniform input does (~6e-4 of the points) first gets the hot
            // kernel's twin with the FP64 determinant stage inside
            if (opt.mid_twin && (long long)hcnt->nflag_set > 256 + insertedTotal / 256) midTwin = true;
            const long long done = (long long)(win_total() - win0) + (long long)(hcnt->ndup - dup0);
            const int newPending = total - (int)done;
            insertedTotal += (long long)(pending - newPending);
            remainingInCall -= (long long)(pending - newPending);
            if (newPending == pending) { if (++stall > 64) fail(ERR_WALK, "no progress in 64 consecutive batches of rounds"); }
            else stall = 0;
            pending = newPending;
            if (hcnt->oom_soft) {
                // some winners found no room: retire the slots handed out beyond the old capacity and grow
                const int oldcap = mesh.cap;
                ensure_simplices((long long)hcnt->ntets + batch_margin(R, nsel, newPerPoint));
                if (hcnt->ntets > oldcap) {
                    MarkDeadArgs md{mesh.owner, oldcap};
                    VOR_LAUNCH(MarkDeadArgs, mark_dead_body, std::min(hcnt->ntets, mesh.cap) - oldcap, md, stream);
                }
                hcnt->oom_soft = 0;
                be::h2d(&mesh.cnt->oom_soft, &hcnt->oom_soft, sizeof(int), stream);
            }
            if (opt.verbose > 1)
                fprintf(stderr, "[vor] stage [%d,%d) rounds %llu: nact %d pending %d stride %d nsel %d simplices %d\n", lo, hi, rs.rounds, nact,
                        pending, stride, nsel, hcnt->ntets);
            if (opt.verbose > 2) {
                int f[8];
                validate(f);
                if (f[0] | f[1] | f[2] | f[3] | f[4] | f[5]) {
                    fprintf(stderr, "[vor] VALIDATION FAILED after round %llu: %d %d %d %d %d\n", rs.rounds, f[0], f[1], f[2], f[3], f[4]);
                    fail(ERR_CUDA, "debug validation failed");
                }
            }
            if (pending > 0 && nact > 4096 && (double)pending < opt.compact_frac * (double)nact) {
                nact = compact_active(nact);
                if (nact != pending) fail(ERR_CUDA, "active list compaction lost points");
            }
        }
        if (pending > 0) {
            // hand the stragglers over: they keep their seeds and lead the active list of the next stage
            nact = compact_active(nact);
            if (nact != pending) fail(ERR_CUDA, "active list compaction lost points");
            carried = nact;
        }
        rs.attempts = hcnt->attempts;
        rs.winners = win_total();
        if (opt.verbose)
            fprintf(stderr, "[vor] stage [%d,%d) done: rounds so far %llu, simplices %d, %d point(s) handed to the next stage\n", lo, hi, rs.rounds,
                    hcnt->ntets, carried);
    }
#endif

    void run_stage(int lo, int hi, bool last) {
        (void)last;
#if VOR_GPU
        if (opt.coop) { run_stage_pipelined(lo, hi, last); return; }
#endif
        int nact = hi - lo;
        int pending = nact;
        SeedArgs<D> sd{keysAll, mesh.pts, mesh.ptTet, mesh.owner, mesh.seed, nsuper, lo, refLo, refHi, D * axisBits};
        VOR_LAUNCH(SeedArgs<D>, init_seeds_body<D>, nact, sd, stream);
        IotaArgs ia{act, lo};
        VOR_LAUNCH(IotaArgs, iota_body, nact, ia, stream);
        int stall = 0;
        uint32_t roundSalt = (uint32_t)mix64((uint64_t)lo * 0x9E37u + rs.rounds);
        while (pending > 0) {
            if (epoch <= 0) reset_owners();
            const double target = std::max((double)opt.min_attempt, (double)insertedTotal / opt.attempt_div);
            uint32_t thr = 1u << bits;
            if ((double)pending > target) {
                // priorities are uniform over [0, 2^bits); `nact` of them are in use
                const double f = target / (double)pending;
                thr = (uint32_t)std::max(1.0, f * (double)(1u << bits));
            }
            roundSalt = roundSalt * 1664525u + 1013904223u;
            const int keybase = epoch << (bits + 1);
            hcnt->nslots = 0; hcnt->nwinners = 0; hcnt->nbig = 0;
            const int ndup0 = hcnt->ndup;
            // only the round-local counters are rewritten (the allocator and statistics live on the device)
            be::h2d(&mesh.cnt->nslots, &hcnt->nslots, sizeof(int) * 3, stream);
            const double fsel0 = (double)thr / (double)(1u << bits);
            const int stride = (opt.select_mode == 1 && thr < (1u << bits)) ? std::max(1, (int)std::lround(1.0 / fsel0)) : 0;
            const int offset = stride > 0 ? (int)(roundSalt % (uint32_t)stride) : 0;
            AttemptArgs<D> aa{mesh, scr, act, bits, roundSalt, thr, stride, offset, keybase, opt.stats, nullptr};
            CheckArgs<D> ca{mesh, scr, bits, roundSalt, keybase, nullptr};
            int nw = 0, used = 0;
            {
                prof.start(0, stream);
                VOR_LAUNCH(AttemptArgs<D>, attempt_body<D>, nact, aa, stream);
                prof.stop(stream);
                const int maxSlots = std::min(nact, scr.nslots);
                prof.start(1, stream);
                VOR_LAUNCH_FULL(CheckArgs<D>, check_body<D>, maxSlots, ca, stream);
                prof.stop(stream);
                pull_counters();
                check_device_error("round");
                nw = hcnt->nwinners;
                used = std::min(hcnt->nslots, scr.nslots);
                ensure_simplices((long long)hcnt->ntets);
                RetriArgs<D> ra{mesh, scr, act, opt.stats};
                prof.start(2, stream);
                VOR_LAUNCH(RetriArgs<D>, retri_body<D>, nw, ra, stream);
                prof.stop(stream);
            }
            const int dropped = hcnt->ndup - ndup0;
            pending -= nw + dropped;
            insertedTotal += nw;
            rs.rounds++;
            rs.attempts += (unsigned long long)used;
            rs.winners += (unsigned long long)nw;
            epoch--;
            if (opt.verbose > 1)
                fprintf(stderr, "[vor] stage [%d,%d) round %llu: nact %d pending %d attempts %d winners %d ntets %d\n", lo, hi, rs.rounds, nact,
                        pending, used, nw, hcnt->ntets);
            if (opt.verbose > 2) {
                int f[8];
                validate(f);
                if (f[0] | f[1] | f[2] | f[3] | f[4] | f[5]) {
                    fprintf(stderr, "[vor] VALIDATION FAILED after round %llu: %d %d %d %d %d\n", rs.rounds, f[0], f[1], f[2], f[3], f[4]);
                    fail(ERR_CUDA, "debug validation failed");
                }
            }
            if (nw + dropped == 0) {
                if (++stall > 64) fail(ERR_WALK, "no progress in 64 consecutive rounds");
            } else stall = 0;
            if (pending > 0 && nact > 4096 && (double)pending < opt.compact_frac * (double)nact) {
                nact = compact_active(nact);
                if (nact != pending) fail(ERR_CUDA, "active list compaction lost points");
            }
        }
        if (opt.verbose)
            fprintf(stderr, "[vor] stage [%d,%d) done: rounds so far %llu, simplices %d\n", lo, hi, rs.rounds, hcnt->ntets);
    }

    // debug: winners' footprints must be pairwise disjoint (host check, small cases only)
    void debug_footprints(int nw) {
        std::vector<int> win(nw), wb(nw);
        be::d2h(win.data(), scr.winners, sizeof(int) * nw, stream);
        be::sync(stream);
        std::vector<std::pair<int, int>> seen; // (tet, winner)
        for (int w = 0; w < nw; w++) {
            int slot = win[w], nk, nb, big;
            be::d2h(&nk, scr.slotNk + slot, 4, stream); be::d2h(&nb, scr.slotNb + slot, 4, stream); be::d2h(&big, scr.slotBig + slot, 4, stream);
            be::sync(stream);
            ScrView sv = scr_view(scr, slot, big);
            { int a; be::d2h(&a, scr.slotAct + slot, 4, stream); be::sync(stream); int v; be::d2h(&v, act + a, 4, stream); be::sync(stream);
              fprintf(stderr, "[vor] winner %d slot %d a %d v %d nk %d nb %d big %d\n", w, slot, a, v, nk, nb, big); }
            for (int j = 0; j < nk; j++) { int t; be::d2h(&t, sv.k + (size_t)j * sv.stride, 4, stream); be::sync(stream); seen.push_back({t, w}); }
            for (int j = 0; j < nb; j++) { int c; be::d2h(&c, sv.o + (size_t)j * sv.stride, 4, stream); be::sync(stream); if (c >= 0) seen.push_back({c >> 2, w}); }
        }
        std::sort(seen.begin(), seen.end());
        for (size_t i = 1; i < seen.size(); i++)
            if (seen[i].first == seen[i - 1].first && seen[i].second != seen[i - 1].second)
                fprintf(stderr, "[vor] OVERLAP: simplex %d in footprints of winners %d and %d\n", seen[i].first, seen[i - 1].second, seen[i].second);
    }

    int compact_active(int nact) {
        long long total = 0;
#if VOR_GPU
        if (opt.coop) {
            const int nb = (nact + 255) / 256;
            k_compact_count<<<nb, 256, 0, stream>>>(act, mesh.seed, blockCnt, nact);
            k_compact_scan<<<1, 1024, 0, stream>>>(blockCnt, nb, d_misc);
            k_compact_scatter<<<nb, 256, 0, stream>>>(act, mesh.seed, blockCnt, act2, nact);
            be::g_launches += 3;
        } else
#endif
        {
            const int chunk = 256;
            const int nb = (nact + chunk - 1) / chunk;
            CompactArgs ca{act, mesh.seed, act2, blockCnt, nact, chunk};
            VOR_LAUNCH(CompactArgs, compact_count_body, nb, ca, stream);
            ScanArgs sa{nullptr, blockCnt, 0, 0, nb, d_misc};
            VOR_LAUNCH(ScanArgs, scan_serial_body, 1, sa, stream);
            VOR_LAUNCH(CompactArgs, compact_scatter_body, nb, ca, stream);
        }
        be::d2h(&total, d_misc, sizeof(long long), stream);
        be::sync(stream);
        std::swap(act, act2);
        rs.compactions++;
        return (int)total;
    }

    // ------------------------------------------------------------------ outputs
    void invalidate_outputs() {
        be::dfree(d_edges);
        d_edges = nullptr;
        nedges = -1;
    }

    // exclusive scan of a[0..n) in place; returns the total
    long long scan_exclusive(int *a, int n) {
        const int chunk = 1024;
        const int nch = (n + chunk - 1) / chunk;
        int *sums = (int *)be::dmalloc(sizeof(int) * (size_t)std::max(nch, 1));
        long long *d_total = (long long *)be::dmalloc(sizeof(long long));
        ScanArgs sa{a, sums, n, chunk, nch, d_total};
        VOR_LAUNCH(ScanArgs, scan_sum_body, nch, sa, stream);
        VOR_LAUNCH(ScanArgs, scan_serial_body, 1, sa, stream);
        VOR_LAUNCH(ScanArgs, scan_apply_body, nch, sa, stream);
        long long total = 0;
        be::d2h(&total, d_total, sizeof(long long), stream);
        be::sync(stream);
        be::dfree(sums);
        be::dfree(d_total);
        return total;
    }

    // canonical edge list on the device; returns the number of edges
    long long edges() {
        if (nedges >= 0) return nedges;
        const auto t0 = std::chrono::steady_clock::now();
        auto lap = [&](const char *what) {
            if (!opt.verbose) return;
            be::sync(stream);
            fprintf(stderr, "[vor] edges: %s at %.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        };
        const int nt = hcnt->ntets;
        int *deg = (int *)be::dmalloc(sizeof(int) * (size_t)(ninput + 1));
        int *cursor = (int *)be::dmalloc(sizeof(int) * (size_t)(ninput + 1));
        be::dmemset(deg, 0, sizeof(int) * (size_t)(ninput + 1), stream);
        be::dmemset(cursor, 0, sizeof(int) * (size_t)(ninput + 1), stream);
        unsigned char *emask = D == 3 ? (unsigned char *)be::dmalloc((size_t)nt + 16) : nullptr;
        EdgeArgs<D> ea{mesh, inputIdx, deg, cursor, nullptr, 0, emask};
        bool wedge = false;
        if constexpr (D == 3) {
            if (opt.edge_wedge) {
                // direction of the wedge test: generic, of the size of the data (any non-zero vector is correct)
                double r = 0.0;
                for (int s = 0; s < nsets; s++) r = std::max(r, radiusBase[s]);
                if (!(r > 0.0) || !std::isfinite(r)) r = 1.0;
                EdgeWedgeArgs wa{mesh, inputIdx, deg, emask, 0.7548776662466927 * r, 0.5698402909980532 * r, 0.3247179572447461 * r};
                if (opt.edge_dir[0] != 0.0 || opt.edge_dir[1] != 0.0 || opt.edge_dir[2] != 0.0) { wa.dx = opt.edge_dir[0]; wa.dy = opt.edge_dir[1]; wa.dz = opt.edge_dir[2]; }
                VOR_LAUNCH(EdgeWedgeArgs, edges_wedge_body, nt, wa, stream);
                wedge = true;
            }
        }
        if (!wedge) VOR_LAUNCH(EdgeArgs<D>, edges_body<D>, nt, ea, stream);
        lap("count pass");
        const long long total = scan_exclusive(deg, ninput + 1);
        lap("scan");
        if (total > 0x7fffffffLL) fail(ERR_OOM, "more than 2^31 edges");
        uint32_t *hi = (uint32_t *)be::dmalloc(sizeof(uint32_t) * (size_t)std::max(total, 1LL));
        d_edges = (uint32_t *)be::dmalloc(sizeof(uint32_t) * 2 * (size_t)std::max(total, 1LL));
        ea.hi = hi;
        ea.pass = 1;
        VOR_LAUNCH(EdgeArgs<D>, edges_body<D>, nt, ea, stream);
        lap("fill pass");
        RowSortArgs ra{deg, hi, d_edges, ninput};
        VOR_LAUNCH(RowSortArgs, row_sort_body, ninput, ra, stream);
        be::sync(stream);
        lap("row sort");
        be::dfree(deg); be::dfree(cursor); be::dfree(hi); be::dfree(emask);
        nedges = total;
        return nedges;
    }
    void copy_edges(uint32_t *h_out, long long cap) {
        const long long m = edges();
        be::d2h_big(h_out, d_edges, sizeof(uint32_t) * 2 * (size_t)std::min(m, cap), stream);
    }
    // edge list in a host block of the caching host allocator; the caller owns it afterwards (vor_host_free)
    uint32_t *edges_to_host_block(long long *m_out) {
        const long long m = edges();
        bool pinned = false;
        uint32_t *h = (uint32_t *)be::g_hostpool.alloc(sizeof(uint32_t) * 2 * (size_t)std::max(m, 1LL), &pinned);
        if (pinned) { be::d2h(h, d_edges, sizeof(uint32_t) * 2 * (size_t)m, stream); be::sync(stream); }
        else be::d2h_big(h, d_edges, sizeof(uint32_t) * 2 * (size_t)m, stream);
        *m_out = m;
        return h;
    }
    unsigned long long edge_checksum() {
        const long long m = edges();
        unsigned long long *d_sum = (unsigned long long *)be::dmalloc(sizeof(unsigned long long));
        be::dmemset(d_sum, 0, sizeof(unsigned long long), stream);
        EdgeSumArgs ea{d_edges, d_sum};
        VOR_LAUNCH(EdgeSumArgs, edge_sum_body, m, ea, stream);
        unsigned long long h = 0;
        be::d2h(&h, d_sum, sizeof(h), stream);
        be::sync(stream);
        be::dfree(d_sum);
        return h;
    }

    // slab mode: local bounds of n device points, and how many of them fail the strict in_sphere test of
    // the bounding sphere of the box [lo, hi] (geometry.rs:99-142) -- the two reductions a group of slabs has to combine
    // (min / max / sum) before every slab can build the SAME super simplex.  Static: no tree needed.
    static void local_bounds(const double *d_in, int n, double *lo, double *hi, be::Stream stream) {
        std::vector<ChunkDesc> chunks;
        for (int a = 0; a < n; a += BBOX_CHUNK) chunks.push_back(ChunkDesc{0, a, std::min(a + BBOX_CHUNK, n)});
        const int nch = (int)chunks.size();
        for (int k = 0; k < D; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; }
        if (!nch) return;
        DevTmp<ChunkDesc> dch((size_t)nch);
        DevTmp<double> dp((size_t)nch * 2 * D);
        be::h2d(dch.p, chunks.data(), sizeof(ChunkDesc) * (size_t)nch, stream);
        BboxArgs<D> ba{d_in, dch.p, dp.p};
        VOR_LAUNCH(BboxArgs<D>, bbox_chunk_body<D>, nch, ba, stream);
        std::vector<double> partial((size_t)nch * 2 * D);
        be::d2h(partial.data(), dp.p, sizeof(double) * partial.size(), stream);
        be::sync(stream);
        for (int c = 0; c < nch; c++)
            for (int k = 0; k < D; k++) {
                lo[k] = std::fmin(lo[k], partial[(size_t)c * 2 * D + k]);
                hi[k] = std::fmax(hi[k], partial[(size_t)c * 2 * D + D + k]);
            }
    }
    static int count_outside(const double *d_in, int n, const double *lo, const double *hi, be::Stream stream) {
        double c[D], ud = 0.0, ld = 0.0;
        for (int k = 0; k < D; k++) c[k] = (hi[k] + lo[k]) / 2.0;
        for (int k = 0; k < D; k++) { const double d = hi[k] - c[k]; ud += d * d; }
        for (int k = 0; k < D; k++) { const double d = lo[k] - c[k]; ld += d * d; }
        ud = std::sqrt(ud); ld = std::sqrt(ld);
        const double r = ud > ld ? ud : ld;
        std::vector<ChunkDesc> chunks;
        for (int a = 0; a < n; a += BBOX_CHUNK) chunks.push_back(ChunkDesc{0, a, std::min(a + BBOX_CHUNK, n)});
        const int nch = (int)chunks.size();
        if (!nch) return 0;
        DevTmp<ChunkDesc> dch((size_t)nch);
        DevTmp<double> dc((size_t)D), dr(1);
        DevTmp<int> dout(1);
        be::h2d(dch.p, chunks.data(), sizeof(ChunkDesc) * (size_t)nch, stream);
        be::h2d(dc.p, c, sizeof(double) * D, stream);
        be::h2d(dr.p, &r, sizeof(double), stream);
        be::dmemset(dout.p, 0, sizeof(int), stream);
        OutsideArgs<D> oa{d_in, dch.p, dc.p, dr.p, dout.p};
        VOR_LAUNCH(OutsideArgs<D>, count_outside_body<D>, nch, oa, stream);
        int out = 0;
        be::d2h(&out, dout.p, sizeof(int), stream);
        be::sync(stream);
        return out;
    }
    // certification pass of a slab: every live simplex with a vertex flagged in `h_owned` (per input index) must have the
    // part of its circumsphere that lies inside the data box within [range_lo, range_hi] along `axis` -- the range in
    // which this tree holds EVERY point of the global set.  Returns the number of simplices that do not, and the extent
    // they need (need[0] <= range_lo, need[1] >= range_hi).
    // h_verts / h_reach (optional): the first `cap` uncertified simplices as M x D vertex coordinates + their reach along the axis
    long long certify_slab(const unsigned char *h_owned, int axis, double range_lo, double range_hi, double shell, double *need,
                           double *h_verts = nullptr, double *h_reach = nullptr, int cap = 0) {
        DevTmp<unsigned char> down((size_t)std::max(ninput, 1));
        DevTmp<unsigned long long> dcount(1);
        DevTmp<double> dneed(2);
        const bool list = h_verts && h_reach && cap > 0;
        DevTmp<double> dverts(list ? (size_t)cap * M * D : 1), dreach(list ? (size_t)cap * 2 : 1);
        be::h2d(down.p, h_owned, (size_t)ninput, stream);
        be::dmemset(dcount.p, 0, sizeof(unsigned long long), stream);
        const double init[2] = {range_lo, range_hi};
        be::h2d(dneed.p, init, sizeof(init), stream);
        CertifyArgs<D> ca{mesh, inputIdx, down.p, dcount.p, dneed.p, axis, range_lo, range_hi, shell};
        for (int k = 0; k < 3; k++) { ca.boxLo[k] = k < D ? boxLo[k] : 0.0; ca.boxHi[k] = k < D ? boxHi[k] : 0.0; }
        if (list) { ca.listVerts = dverts.p; ca.listReach = dreach.p; ca.listCap = cap; }
        VOR_LAUNCH(CertifyArgs<D>, certify_body<D>, hcnt->ntets, ca, stream);
        unsigned long long cnt = 0;
        be::d2h(&cnt, dcount.p, sizeof(cnt), stream);
        be::d2h(need, dneed.p, sizeof(double) * 2, stream);
        be::sync(stream);
        if (list && cnt > 0) {
            const size_t k = (size_t)std::min<unsigned long long>(cnt, (unsigned long long)cap);
            be::d2h(h_verts, dverts.p, sizeof(double) * k * M * D, stream);
            be::d2h(h_reach, dreach.p, sizeof(double) * k * 2, stream);
            be::sync(stream);
        }
        return (long long)cnt;
    }

    // slab mode: this slab's part of the global canonical edge list in a host block of the caching host allocator
    // (h_gmap / h_owned: per local input index); the caller owns the block (vor_host_free)
    uint32_t *slab_edges_to_host_block(const long long *h_gmap, const unsigned char *h_owned, long long *m_out) {
        const long long m = edges();
        if (m > 0x7ffffff0LL) fail(ERR_OOM, "too many edges");
        DevTmp<long long> dg((size_t)std::max(ninput, 1));
        DevTmp<unsigned char> down((size_t)std::max(ninput, 1));
        DevTmp<unsigned long long> k0((size_t)m + 2), k1((size_t)m + 2), dcount(1);
        be::h2d(dg.p, h_gmap, sizeof(long long) * (size_t)ninput, stream);
        be::h2d(down.p, h_owned, (size_t)ninput, stream);
        SlabEdgeArgs ka{d_edges, dg.p, down.p, k0.p};
        VOR_LAUNCH(SlabEdgeArgs, slab_edge_key_body, m, ka, stream);
        be::dmemset(k0.p + m, 0xff, sizeof(unsigned long long), stream);   // sentinel ~0 behind the last key
        be::sort_keys(reinterpret_cast<uint64_t *>(k0.p), reinterpret_cast<uint64_t *>(k1.p), (size_t)m + 1, stream);
        be::dmemset(dcount.p, 0, sizeof(unsigned long long), stream);
        KeyCountArgs ca{k1.p, dcount.p};
        VOR_LAUNCH(KeyCountArgs, key_count_body, m, ca, stream);
        unsigned long long cnt = 0;
        be::d2h(&cnt, dcount.p, sizeof(cnt), stream);
        be::sync(stream);
        DevTmp<uint32_t> dout(2 * (size_t)std::max<unsigned long long>(cnt, 1));
        KeyUnpackArgs ua{k1.p, dout.p};
        VOR_LAUNCH(KeyUnpackArgs, key_unpack_body, (long long)cnt, ua, stream);
        bool pinned = false;
        uint32_t *h = (uint32_t *)be::g_hostpool.alloc(sizeof(uint32_t) * 2 * (size_t)std::max<unsigned long long>(cnt, 1), &pinned);
        if (pinned) { be::d2h(h, dout.p, sizeof(uint32_t) * 2 * (size_t)cnt, stream); be::sync(stream); }
        else be::d2h_big(h, dout.p, sizeof(uint32_t) * 2 * (size_t)cnt, stream);
        *m_out = (long long)cnt;
        return h;
    }

    // per-set (edge count, checksum64 of the set-local edge list) of a batch tree; h_setOff = nsets + 1 input offsets
    void per_set_edge_stats(const int *h_setOff, unsigned long long *h_cnt, unsigned long long *h_sum) {
        const long long m = edges();
        DevTmp<int> doff((size_t)nsets + 1);
        DevTmp<unsigned long long> dc((size_t)nsets), ds((size_t)nsets);
        be::h2d(doff.p, h_setOff, sizeof(int) * (size_t)(nsets + 1), stream);
        be::dmemset(dc.p, 0, sizeof(unsigned long long) * (size_t)nsets, stream);
        be::dmemset(ds.p, 0, sizeof(unsigned long long) * (size_t)nsets, stream);
        const int run = 128;
        SetEdgeArgs a{d_edges, m, doff.p, nsets, dc.p, ds.p, run};
        VOR_LAUNCH(SetEdgeArgs, set_edge_stats_body, (int)((m + run - 1) / run), a, stream);
        be::d2h(h_cnt, dc.p, sizeof(unsigned long long) * (size_t)nsets, stream);
        be::d2h(h_sum, ds.p, sizeof(unsigned long long) * (size_t)nsets, stream);
        be::sync(stream);
    }

    // fail[0..4] as in ValidateArgs; returns number of live simplices
    unsigned long long validate(int *fail_out) {
        int *d_fail = (int *)be::dmalloc(sizeof(int) * 8);
        unsigned long long *d_live = (unsigned long long *)be::dmalloc(sizeof(unsigned long long));
        be::dmemset(d_fail, 0, sizeof(int) * 8, stream);
        be::dmemset(d_live, 0, sizeof(unsigned long long), stream);
        ValidateArgs<D> va{mesh, d_fail, d_live};
        VOR_LAUNCH(ValidateArgs<D>, validate_body<D>, hcnt->ntets, va, stream);
        unsigned long long live = 0;
        be::d2h(fail_out, d_fail, sizeof(int) * 8, stream);
        be::d2h(&live, d_live, sizeof(live), stream);
        be::sync(stream);
        be::dfree(d_fail);
        be::dfree(d_live);
        pull_counters();
        return live;
    }

    // TEST HOOK (vor_debug_corrupt): damage the finished mesh in one of six ways so that check_delaunay has something to
    // reject -- kind k makes fail counter k fire: 0 orientation, 1 dead neighbour, 2 asymmetric adjacency, 3 facet
    // mismatch, 4 not Delaunay (a vertex moved into a neighbour's circumsphere), 5 sphere filter certifying nonsense.
    void debug_corrupt(int kind) {
        const int nt = hcnt->ntets;
        if (nt > (1 << 24)) fail(ERR_ARG, "debug_corrupt is a test hook for small meshes");
        std::vector<int4> rec((size_t)REC4 * nt);
        std::vector<int> own;
        be::d2h(rec.data(), mesh.tet, sizeof(int4) * rec.size(), stream);
        if (!VOR_INTERLEAVE) { own.resize((size_t)OWS * nt); be::d2h(own.data(), mesh.owner, sizeof(int) * own.size(), stream); }
        be::sync(stream);
        auto killw = [&](int t) { return VOR_INTERLEAVE ? reinterpret_cast<const int *>(rec.data())[(size_t)OWS * t] : own[(size_t)OWS * t]; };
        auto tv = [&](int t) -> int4 & { return rec[(size_t)REC4 * t + TVO4]; };
        auto tn = [&](int t) -> int4 & { return rec[(size_t)REC4 * t + TVO4 + 1]; };
        auto real = [&](int t) {
            if (killw(t) < 0) return false;
            for (int k = 0; k < M; k++) if (get4(tv(t), k) < nsuper || get4(tn(t), k) < 0) return false;
            return true;
        };
        int t = -1;
        for (int c = nt / 2; c < nt && t < 0; c++) {
            if (!real(c)) continue;
            bool ok = true;
            for (int k = 0; k < M; k++) ok = ok && real(get4(tn(c), k) >> 2);
            if (ok) t = c;
        }
        if (t < 0) fail(ERR_ARG, "debug_corrupt: no interior simplex found");
        int4 v = tv(t), nb = tn(t);
        auto put_rec = [&]() {
            be::h2d(mesh.tet + REC4 * (size_t)t + TVO4, &v, sizeof(int4), stream);
            be::h2d(mesh.tet + REC4 * (size_t)t + TVO4 + 1, &nb, sizeof(int4), stream);
        };
        if (kind == 0) { std::swap(v.x, v.y); put_rec(); }
        else if (kind == 1) {
            int d = -1;
            for (int c = 0; c < nt && d < 0; c++) if (killw(c) < 0) d = c;
            if (d < 0) fail(ERR_ARG, "debug_corrupt: no dead simplex");
            nb.x = d * 4; put_rec();
        } else if (kind == 2) { nb.x = (nb.x & ~3) | (((nb.x & 3) + 1) % M); put_rec(); }
        else if (kind == 3) {
            int other = -1;
            for (int c = nt / 4; c < nt && other < 0; c++) {
                if (!real(c) || c == t) continue;
                const int cand = tv(c).x;
                bool used = false;
                for (int k = 0; k < M; k++) used = used || get4(v, k) == cand || get4(tv(nb.x >> 2), k) == cand;
                if (!used) other = cand;
            }
            if (other < 0) fail(ERR_ARG, "debug_corrupt: no foreign vertex");
            v.y = other; put_rec();
        } else if (kind == 4) {
            const int n0 = nb.x >> 2, j0 = nb.x & 3;
            const int w = get4(tv(n0), j0);
            Pt c{};
            std::vector<Pt> pv(M);
            for (int k = 0; k < M; k++) be::d2h(&pv[k], mesh.pts + get4(v, k), sizeof(Pt), stream);
            be::sync(stream);
            for (int k = 0; k < M; k++) { c.x += pv[k].x / M; c.y += pv[k].y / M; if constexpr (D == 3) c.z += pv[k].z / M; }
            be::h2d(mesh.pts + w, &c, sizeof(Pt), stream);
        } else if (kind == 5) {
            const float big = 1e30f;
            be::h2d(mesh.owner + OWS * (size_t)t + 5, &big, sizeof(float), stream);
        } else fail(ERR_ARG, "debug_corrupt: kind must be 0..5");
        be::sync(stream);
        invalidate_outputs();
    }

    // live simplices in slot order: liveId[c] = slot, compactOf[slot] = c (-1 when dead); returns the count
    int compact_live(int *liveId, int *compactOf) {
        const int nt = hcnt->ntets;
        ExportArgs<D> xa{mesh, liveId, compactOf};
        VOR_LAUNCH(ExportArgs<D>, export_flag_body<D>, nt, xa, stream);
        const long long n = scan_exclusive(compactOf, nt);
        VOR_LAUNCH(ExportArgs<D>, export_mark_body<D>, nt, xa, stream);
        return (int)n;
    }

    // locate (delaunay_tree.rs:33-58) for nq host query points: conflict regions as export indices (the index space
    // of export_simplices), `cap` entries per query.  counts[i] = size, -1 = does not fit, -2 = outside.
    void locate(const double *h_q, int nq, int cap, int *h_out, int *h_counts) {
        const int nt = hcnt->ntets;
        DevTmp<double> dq((size_t)nq * D);
        DevTmp<int> dseed((size_t)nq), dout((size_t)nq * cap), dcount((size_t)nq), liveId((size_t)nt), compactOf((size_t)nt);
        be::h2d(dq.p, h_q, sizeof(double) * (size_t)nq * D, stream);
        compact_live(liveId.p, compactOf.p);
        // seeds: root simplex of set 0 when nothing is inserted, else the simplex of the last inserted vertex (forwarded)
        fill_i(dseed.p, nv > nsuper ? -1 : 0, (size_t)nq);
        if (nv > nsuper) {
            QuerySeedArgs<D> qs{keysAll, mesh.pts, mesh.ptTet, dq.p, d_boxLo, d_boxHi, dseed.p, nsuper, refLo, refHi, axisBits};
            VOR_LAUNCH(QuerySeedArgs<D>, query_seed_body<D>, nq, qs, stream);
        }
        LocateQueryArgs<D> la{mesh, dq.p, dseed.p, compactOf.p, dout.p, dcount.p, cap};
        VOR_LAUNCH(LocateQueryArgs<D>, locate_query_body<D>, nq, la, stream);
        be::d2h(h_out, dout.p, sizeof(int) * (size_t)nq * cap, stream);
        be::d2h(h_counts, dcount.p, sizeof(int) * (size_t)nq, stream);
        be::sync(stream);
    }
    // vertices in reference id order (super k -> k, ghost copies, input i -> 2*M*nsets + i) with their incident live
    // simplices as CSR of export indices.  Returns the number of vertex ids; *n_inc = number of incidences.
    // Any output pointer may be null.
    long long export_vertices(double *h_coords, long long *h_off, int *h_simps, long long cap, long long *n_inc) {
        const int nt = hcnt->ntets;
        const int idOffset = 2 * M * nsets;
        const long long nIds = (long long)idOffset + ninput;
        DevTmp<int> liveId((size_t)nt), compactOf((size_t)nt), deg((size_t)nIds + 1), cursor((size_t)nIds + 1);
        const int nlive = compact_live(liveId.p, compactOf.p);
        be::dmemset(deg.p, 0, sizeof(int) * (size_t)(nIds + 1), stream);
        be::dmemset(cursor.p, 0, sizeof(int) * (size_t)(nIds + 1), stream);
        IncidArgs<D> ia{mesh, liveId.p, inputIdx, deg.p, cursor.p, nullptr, idOffset, 0};
        VOR_LAUNCH(IncidArgs<D>, incid_body<D>, nlive, ia, stream);
        const long long total = scan_exclusive(deg.p, (int)nIds + 1);
        if (n_inc) *n_inc = total;
        if (h_off) {
            std::vector<int> off((size_t)nIds + 1);
            be::d2h(off.data(), deg.p, sizeof(int) * (size_t)(nIds + 1), stream);
            be::sync(stream);
            for (long long i = 0; i <= nIds; i++) h_off[i] = off[(size_t)i];
        }
        if (h_simps) {
            if (cap < total) fail(ERR_ARG, "incidence buffer too small");
            DevTmp<int> simps((size_t)std::max(total, 1LL));
            ia.simps = simps.p;
            ia.pass = 1;
            VOR_LAUNCH(IncidArgs<D>, incid_body<D>, nlive, ia, stream);
            RowSortPlainArgs ra{deg.p, simps.p};
            VOR_LAUNCH(RowSortPlainArgs, row_sort_plain_body, (int)nIds, ra, stream);
            be::d2h(h_simps, simps.p, sizeof(int) * (size_t)total, stream);
            be::sync(stream);
        }
        if (h_coords) {
            DevTmp<double> dc((size_t)nIds * D);
            be::dmemset(dc.p, 0, sizeof(double) * (size_t)nIds * D, stream);
            CoordArgs<D> ca{mesh.pts, inputIdx, dc.p, nsuper, idOffset};
            VOR_LAUNCH(CoordArgs<D>, coord_body<D>, nv - nsuper, ca, stream);
            be::d2h(h_coords, dc.p, sizeof(double) * (size_t)nIds * D, stream);
            be::sync(stream);
            // super vertices, then their ghost copies (delaunay_tree.rs:407-412 / :559-566)
            static const int ghostOf3[4] = {0, 0, 0, 1}, ghostOf2[3] = {0, 1, 2};
            for (int s = 0; s < nsets; s++)
                for (int k = 0; k < M; k++) {
                    const double *sv = &superXYZ[((size_t)s * M + k) * D];
                    const int g = D == 3 ? ghostOf3[k] : ghostOf2[k];
                    const double *gv = &superXYZ[((size_t)s * M + g) * D];
                    for (int d = 0; d < D; d++) {
                        h_coords[((size_t)s * M + k) * D + d] = sv[d];
                        h_coords[((size_t)(nsets + s) * M + k) * D + d] = gv[d];
                    }
                }
        }
        return nIds;
    }
    // make_queue (scheduler.rs:6-28): footprints of nq host query points as sorted unique export indices, padded to
    // `fcap` per query.  counts[i] = size, -1 = conflict region or footprint does not fit, -2 = outside.
    void make_queue(const double *h_q, int nq, int kcap, int fcap, int *h_fp, int *h_counts) {
        const int nt = hcnt->ntets;
        DevTmp<double> dq((size_t)nq * D);
        DevTmp<int> dseed((size_t)nq), dk((size_t)nq * kcap), dkc((size_t)nq), dfp((size_t)nq * fcap), dfc((size_t)nq), liveId((size_t)nt),
            compactOf((size_t)nt);
        be::h2d(dq.p, h_q, sizeof(double) * (size_t)nq * D, stream);
        compact_live(liveId.p, compactOf.p);
        fill_i(dseed.p, nv > nsuper ? -1 : 0, (size_t)nq);
        if (nv > nsuper) {
            QuerySeedArgs<D> qs{keysAll, mesh.pts, mesh.ptTet, dq.p, d_boxLo, d_boxHi, dseed.p, nsuper, refLo, refHi, axisBits};
            VOR_LAUNCH(QuerySeedArgs<D>, query_seed_body<D>, nq, qs, stream);
        }
        LocateQueryArgs<D> la{mesh, dq.p, dseed.p, nullptr, dk.p, dkc.p, kcap};
        VOR_LAUNCH(LocateQueryArgs<D>, locate_query_body<D>, nq, la, stream);
        FootprintArgs<D> fa{mesh, dk.p, dkc.p, compactOf.p, dfp.p, dfc.p, kcap, fcap};
        VOR_LAUNCH(FootprintArgs<D>, footprint_body<D>, nq, fa, stream);
        std::vector<int> kc((size_t)nq);
        be::d2h(h_fp, dfp.p, sizeof(int) * (size_t)nq * fcap, stream);
        be::d2h(h_counts, dfc.p, sizeof(int) * (size_t)nq, stream);
        be::d2h(kc.data(), dkc.p, sizeof(int) * (size_t)nq, stream);
        be::sync(stream);
        for (int i = 0; i < nq; i++)
            if (kc[i] == -2) h_counts[i] = -2;
    }
    template <class T> struct DevTmp {
        T *p;
        explicit DevTmp(size_t n) : p((T *)be::dmalloc(sizeof(T) * (n ? n : 1))) {}
        ~DevTmp() { be::dfree(p); }
        DevTmp(const DevTmp &) = delete;
    };

    // compact export of live simplices; vertex ids: super k -> k, input i -> idOffset + i.  Returns the count.
    // Any output pointer may be null.  Two-phase use: call with all null to get the count.
    int export_simplices(int *h_verts, int *h_nbrs, double *h_center, double *h_radius, int idOffset) {
        const int nt = hcnt->ntets;
        int *liveId = (int *)be::dmalloc(sizeof(int) * (size_t)nt);
        int *compactOf = (int *)be::dmalloc(sizeof(int) * (size_t)nt);
        const int n = compact_live(liveId, compactOf);
        if (h_verts || h_nbrs || h_center || h_radius) {
            int *d_verts = (int *)be::dmalloc(sizeof(int) * (size_t)n * M);
            int *d_nbrs = (int *)be::dmalloc(sizeof(int) * (size_t)n * M);
            const bool spheres = h_center || h_radius;   // either may be asked for alone
            double *d_c = spheres ? (double *)be::dmalloc(sizeof(double) * (size_t)n * D) : nullptr;
            double *d_r = spheres ? (double *)be::dmalloc(sizeof(double) * (size_t)n) : nullptr;
            ExportFillArgs<D> fa{mesh, liveId, compactOf, inputIdx, d_verts, d_nbrs, d_c, d_r, idOffset};
            VOR_LAUNCH(ExportFillArgs<D>, export_fill_body<D>, n, fa, stream);
            if (h_verts) be::d2h(h_verts, d_verts, sizeof(int) * (size_t)n * M, stream);
            if (h_nbrs) be::d2h(h_nbrs, d_nbrs, sizeof(int) * (size_t)n * M, stream);
            if (h_center) be::d2h(h_center, d_c, sizeof(double) * (size_t)n * D, stream);
            if (h_radius && d_r) be::d2h(h_radius, d_r, sizeof(double) * (size_t)n, stream);
            be::sync(stream);
            be::dfree(d_verts); be::dfree(d_nbrs); be::dfree(d_c); be::dfree(d_r);
        }
        be::dfree(liveId); be::dfree(compactOf);
        return n;
    }
};

} // namespace vor
