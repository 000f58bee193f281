// setup_kernels.cuh -- bootstrap and ordering kernels.
//
//   bbox_chunk / count_outside  -> bounding_sphere, /root/reference/src/geometry.rs:99-142.  min/max folds are
//                                   exact and order independent; the 1.5x growth loop (:132-140) fires for the first
//                                   point that fails the strict in_sphere test and can never fire twice (1.5*r0
//                                   covers the whole box), so "any point outside r0" reproduces it bit for bit.
//   keys / gather / stage_bounds -> insertion order.  The reference inserts in input order (lib.rs:110-120) and
//                                   lets make_queue/find_placement pick rounds (scheduler.rs:6-55); here the order is a
//                                   BRIO: random doubling stages, Morton order inside a stage (the final DT is unique,
//                                   SURVEY.md §0 D2, so the order is free).
//   init_seeds                   -> replaces the kd-tree nearest-vertex seed (delaunay_tree.rs:37): binary search of
//                                   the point's Morton key among the previous stage's points.
#pragma once
#include "kernels.cuh"

namespace vor {

constexpr int STAGE_SHIFT = 58;       // key = stage(6) | set | morton
constexpr int BBOX_CHUNK = 1024;

struct ChunkDesc { int set; int lo; int hi; }; // points [lo,hi) of the input belong to `set`

template <int D> struct BboxArgs {
    const double *in;          // n x D row-major input points
    const ChunkDesc *chunks;
    double *partial;           // per chunk: lo[D], hi[D]
};
template <int D> VOR_HD void bbox_chunk_body(const BboxArgs<D> &A, int c) {
    const ChunkDesc ch = A.chunks[c];
    double lo[D], hi[D];
    for (int k = 0; k < D; k++) { lo[k] = INFINITY; hi[k] = -INFINITY; }
    for (int i = ch.lo; i < ch.hi; i++)
        for (int k = 0; k < D; k++) {
            const double x = A.in[(size_t)i * D + k];
            lo[k] = fmin(lo[k], x);
            hi[k] = fmax(hi[k], x);
        }
    for (int k = 0; k < D; k++) { A.partial[(size_t)c * 2 * D + k] = lo[k]; A.partial[(size_t)c * 2 * D + D + k] = hi[k]; }
}

template <int D> struct OutsideArgs {
    const double *in;
    const ChunkDesc *chunks;
    const double *center;      // per set: D doubles
    const double *radius;      // per set: r0
    int *outside;              // per set: number of points failing the strict in_sphere test
};
template <int D> VOR_HD void count_outside_body(const OutsideArgs<D> &A, int c) {
    const ChunkDesc ch = A.chunks[c];
    const double r = A.radius[ch.set];
    int cnt = 0;
    for (int i = ch.lo; i < ch.hi; i++) {
        double dist = 0.0;
        for (int k = 0; k < D; k++) {
            const double d = A.center[(size_t)ch.set * D + k] - A.in[(size_t)i * D + k];
            dist += d * d;   // geometry.rs:93-95, left-to-right, no FMA (-fmad=false)
        }
        if (!(dist < r * r)) cnt++;
    }
    if (cnt) atomic_add_i(&A.outside[ch.set], cnt);
}

VOR_HD uint64_t spread_bits3(uint64_t v) { // 21 bits -> every third bit
    v &= 0x1fffff;
    v = (v | v << 32) & 0x1f00000000ffffULL;
    v = (v | v << 16) & 0x1f0000ff0000ffULL;
    v = (v | v << 8) & 0x100f00f00f00f00fULL;
    v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
    v = (v | v << 2) & 0x1249249249249249ULL;
    return v;
}
VOR_HD uint64_t spread_bits2(uint64_t v) { // 29 bits -> every second bit
    v &= 0x1fffffff;
    v = (v | v << 16) & 0x0000ffff0000ffffULL;
    v = (v | v << 8) & 0x00ff00ff00ff00ffULL;
    v = (v | v << 4) & 0x0f0f0f0f0f0f0f0fULL;
    v = (v | v << 2) & 0x3333333333333333ULL;
    v = (v | v << 1) & 0x5555555555555555ULL;
    return v;
}

template <int D> struct KeyArgs {
    const double *in;          // this call's points, n x D
    const int *setOff;         // nsets+1 offsets into `in`
    const double *boxLo;       // per set D doubles (from create)
    const double *boxHi;
    const int *setS0;          // per set: size of stage 0 (>= 256, >= points already inserted in the set)
    uint64_t *keys;
    uint32_t *vals;
    int nsets;
    int setBits;               // bits of the set field
    int axisBits;              // Morton bits per axis
    uint64_t salt;
    int stageLog;              // stage sizes grow by 2^stageLog (1: doubling)
};
template <int D> VOR_HD void keys_body(const KeyArgs<D> &A, int i) {
    // set of point i
    int s = 0;
    if (A.nsets > 1) {
        int lo = 0, hi = A.nsets;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (A.setOff[mid] <= i) lo = mid; else hi = mid; }
        s = lo;
    }
    const int cnt = A.setOff[s + 1] - A.setOff[s];
    const uint64_t r = mix64(A.salt ^ (uint64_t)(i - A.setOff[s])) % (uint64_t)cnt;
    const uint64_t s0 = (uint64_t)A.setS0[s];
    int stage = 0;
    if (r >= s0) { uint64_t x = r / s0; int lg = 0; while (x > 1) { x >>= 1; lg++; } stage = 1 + lg / A.stageLog; }
    uint64_t code = 0;
    const double scale = (double)((1u << A.axisBits) - 1u);
    for (int k = 0; k < D; k++) {
        const double lo = A.boxLo[(size_t)s * D + k], hi = A.boxHi[(size_t)s * D + k];
        const double ext = hi - lo;
        double u = ext > 0.0 ? (A.in[(size_t)i * D + k] - lo) / ext : 0.0;
        u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
        const uint64_t qv = (uint64_t)(u * scale);
        code |= (D == 3 ? spread_bits3(qv) : spread_bits2(qv)) << k;
    }
    A.keys[i] = ((uint64_t)stage << STAGE_SHIFT) | ((uint64_t)s << (D * A.axisBits)) | code;
    A.vals[i] = (uint32_t)i;
}

template <int D> struct GatherArgs {
    const double *in;
    const uint32_t *vals;      // sorted position -> index in `in`
    typename Dim<D>::Pt *pts;
    int *inputIdx;             // per vertex: global input index
    int *vidOfInput;           // per global input index: vertex id
    int *seed;
    int *ptTet;
    int vbase;                 // vertex id of sorted position 0
    int inputBase;             // global input index of in[0]
};
VOR_HD void store_pt(double4 *p, int v, const double *src) { double4 q; q.x = src[0]; q.y = src[1]; q.z = src[2]; q.w = 0.0; p[v] = q; }
VOR_HD void store_pt(double2 *p, int v, const double *src) { double2 q; q.x = src[0]; q.y = src[1]; p[v] = q; }
template <int D> VOR_HD void gather_body(const GatherArgs<D> &A, int pos) {
    const int idx = (int)A.vals[pos];
    const int v = A.vbase + pos;
    store_pt(A.pts, v, A.in + (size_t)idx * D);
    A.inputIdx[v] = A.inputBase + idx;
    A.vidOfInput[A.inputBase + idx] = v;
    A.seed[v] = 0;
    A.ptTet[v] = -1;   // set by the insertion; stays -1 if the point is dropped as a duplicate
}

struct StageBoundsArgs { const uint64_t *keys; int *stageLo; };
VOR_HD void stage_bounds_body(const StageBoundsArgs &A, int pos) {
    const int st = (int)(A.keys[pos] >> STAGE_SHIFT);
    if (pos == 0 || st != (int)(A.keys[pos - 1] >> STAGE_SHIFT)) A.stageLo[st] = pos;
}

template <int D> struct SeedArgs {
    const uint64_t *keysAll;   // Morton keys of every real vertex, indexed by v - nsuper
    const typename Dim<D>::Pt *pts;
    const int *ptTet;
    const int *owner;
    int *seed;
    int nsuper;
    int lo;                    // first vertex of the stage being started
    int plo, phi;              // vertex range of the reference stage (already inserted), may be empty
    int setShift;              // D * axisBits
};
VOR_HD double dist2(const double4 &a, const double4 &b) { return (a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z); }
VOR_HD double dist2(const double2 &a, const double2 &b) { return (a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y); }
// Seed of a point that enters the active list: among the 4 Morton neighbours of its key in the reference stage take
// the vertex closest in space, start from the simplex created by that vertex's insertion and resolve the forwarding
// chain HERE (bulk, fully parallel) instead of on the critical path of the first attempt.
template <int D> VOR_HD void init_seeds_body(const SeedArgs<D> &A, int j) {
    const int v = A.lo + j;
    const uint64_t mask = (1ULL << STAGE_SHIFT) - 1ULL;
    const uint64_t key = A.keysAll[v - A.nsuper] & mask;
    const int set = (int)(key >> A.setShift);
    int seed = set; // root simplex of the set (dead roots forward to live simplices)
    if (A.phi > A.plo) {
        int lo = A.plo, hi = A.phi; // lower_bound of key in the reference stage
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((A.keysAll[mid - A.nsuper] & mask) < key) lo = mid + 1; else hi = mid;
        }
        const typename Dim<D>::Pt p = A.pts[v];
        int best = -1;
        double bestd = INFINITY;
        for (int c = lo - 2; c <= lo + 1; c++) {
            if (c < A.plo || c >= A.phi) continue;
            const uint64_t kc = A.keysAll[c - A.nsuper] & mask;
            if ((int)(kc >> A.setShift) != set) continue;
            if (A.ptTet[c] < 0) continue;   // dropped duplicate
            const double d = dist2(A.pts[c], p);
            if (d < bestd) { bestd = d; best = c; }
        }
        // every candidate pending (a duplicate, or a straggler the reference stage handed over): look further along the curve
        for (int r = 3; best < 0 && r <= 18; r++)
            for (int side = 0; side < 2 && best < 0; side++) {
                const int c = side ? lo + r - 1 : lo - r;
                if (c < A.plo || c >= A.phi) continue;
                if ((int)((A.keysAll[c - A.nsuper] & mask) >> A.setShift) != set) continue;
                if (A.ptTet[c] >= 0) best = c;
            }
        if (best >= 0) seed = A.ptTet[best];
    }
    int o;
    while ((o = A.owner[OWS * (size_t)seed]) < 0) seed = ~o;
    A.seed[v] = seed;
}

// seed of a read-only query point (vor_tree_locate): Morton neighbour of the query among the vertices of the last
// completed stage of set 0 (queries address single-set trees)
template <int D> struct QuerySeedArgs {
    const uint64_t *keysAll;
    const typename Dim<D>::Pt *pts;
    const int *ptTet;
    const double *q;
    const double *boxLo, *boxHi;
    int *seed;
    int nsuper, plo, phi, axisBits;
};
template <int D> VOR_HD void query_seed_body(const QuerySeedArgs<D> &A, int qi) {
    const uint64_t mask = (1ULL << STAGE_SHIFT) - 1ULL;
    uint64_t code = 0;
    const double scale = (double)((1u << A.axisBits) - 1u);
    for (int k = 0; k < D; k++) {
        const double lo = A.boxLo[k], ext = A.boxHi[k] - lo;
        double u = ext > 0.0 ? (A.q[(size_t)qi * D + k] - lo) / ext : 0.0;
        u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
        const uint64_t qv = (uint64_t)(u * scale);
        code |= (D == 3 ? spread_bits3(qv) : spread_bits2(qv)) << k;
    }
    int lo = A.plo, hi = A.phi;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((A.keysAll[mid - A.nsuper] & mask) < code) lo = mid + 1; else hi = mid;
    }
    int best = -1;
    for (int c = lo - 1; c <= lo + 1 && best < 0; c++)
        if (c >= A.plo && c < A.phi && A.ptTet[c] >= 0) best = c;
    for (int c = A.plo; c < A.phi && best < 0; c++)
        if (A.ptTet[c] >= 0) best = c;
    A.seed[qi] = best >= 0 ? A.ptTet[best] : 0;
}

// Bulk point location at the start of a stage: every point of the stage walks from its seed to the simplex that
// contains it (thread per point, the mesh is static here).  The walk of a point's FIRST attempt is the longest one
// (about 9 steps from the Morton neighbour's simplex); doing it here, with the whole stage in flight, takes it off the
// latency-critical path of the round kernels, which then only re-walk 0-3 steps after a neighbour's insertion.
template <int D> struct LocateArgs {
    Mesh<D> m;
    int lo;
    int stats;
};
template <int D> VOR_HD void locate_body(const LocateArgs<D> &A, int j) {
    constexpr int M = Dim<D>::M;
    using G = Geo<D>;
    const Mesh<D> &m = A.m;
    const int v = A.lo + j;
    int s = m.seed[v];
    if (s < 0) return;
    PredCtx cx{m.cnt};
    const typename G::Pt p = m.pts[v];
    unsigned rot = (unsigned)v * 2654435761u;
    unsigned steps = 0;
    for (;;) {
        int4 stv, stn;
        load_rec(m, s, stv, stn);
        const typename G::Verts tvv = G::load(m, stv);
        const int mk = G::beyond_mask(cx, tvv, p);
        if (mk == 0) break;
        int go = 0;
        const int r0 = (int)((rot >> 16) % (unsigned)M);
        for (int k = 0; k < M; k++) {
            const int i = (r0 + k) % M;
            if ((mk >> i) & 1) { go = i; break; }
        }
        const int code = get4(stn, go);
        if (code < 0) { set_err(m.cnt, ERR_OUTSIDE); return; }
        s = code >> 2;
        rot = rot * 1664525u + 1013904223u;
        if (++steps > (1u << 22)) { set_err(m.cnt, ERR_WALK); return; }
    }
    m.seed[v] = s;
    if (A.stats) atomic_add_ull(&m.cnt->walk_steps, steps);
}

// generic exclusive scan over int arrays: chunk sums -> serial scan of sums -> apply
struct ScanArgs { int *a; int *sums; int n; int chunk; int nchunks; long long *total; };
VOR_HD void scan_sum_body(const ScanArgs &A, int c) {
    const int lo = c * A.chunk, hi = lo + A.chunk < A.n ? lo + A.chunk : A.n;
    int s = 0;
    for (int i = lo; i < hi; i++) s += A.a[i];
    A.sums[c] = s;
}
VOR_HD void scan_serial_body(const ScanArgs &A, int) {
    long long run = 0;
    for (int c = 0; c < A.nchunks; c++) { const int s = A.sums[c]; A.sums[c] = (int)run; run += s; }
    if (A.total) *A.total = run;
}
VOR_HD void scan_apply_body(const ScanArgs &A, int c) {
    const int lo = c * A.chunk, hi = lo + A.chunk < A.n ? lo + A.chunk : A.n;
    int run = A.sums[c];
    for (int i = lo; i < hi; i++) { const int x = A.a[i]; A.a[i] = run; run += x; }
}

} // namespace vor
