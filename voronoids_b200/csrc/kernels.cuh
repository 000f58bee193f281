// kernels.cuh -- kernel bodies of the parallel incremental Delaunay engine
// (dimension-generic: D=3 tetrahedra / in-sphere, D=2 triangles / in-circle).
//
// Hot path named by BASELINE.json north_star, reference side:
//   locate + find_all_neighbors   /root/reference/src/delaunay_tree.rs:33-75     -> attempt_body (walk + flood)
//   make_queue / find_placement   /root/reference/src/scheduler.rs:6-55          -> atomicMin reservation in attempt_body,
//                                                                                   ownership check in check_body
//   get_new_simplices             /root/reference/src/delaunay_tree.rs:77-123    -> boundary facets recorded by the flood
//   insert_points_parallel        /root/reference/src/delaunay_tree.rs:213-334   -> check_body (allocator) + retri_body
//   pair_simplices                /root/reference/src/delaunay_tree.rs:674-695   -> edge/vertex pivot inside the cavity (retri_body)
//
// Store (SoA, all in HBM):
//   pts[v]    coordinates (double4 in 3D: one 32 B sector per vertex; double2 in 2D)
//   tet[2t]   int4 vertex ids of simplex t (2D uses x,y,z)                      } one 32 B record (sector) per simplex:
//   tet[2t+1] int4 neighbour codes: (neighbour << 2 | slot in neighbour that    } testing a simplex also brings its
//             points back), -1 = outside the super simplex; slot i is opposite  } adjacency into L1/L2 for the next
//             vertex i.  The reference keeps an unordered Vec (delaunay_tree.rs:15).   BFS level / walk step
//   owner[8t..8t+7]  one 32 B block per simplex (sphere.cuh): ownership words + certified circumsphere filter, so that
//               ONE 256-bit gather decides a conflict test (the reference caches centre/radius too, delaunay_tree.rs:11-16)
//   owner[8t]   "kill word"  >= 0: smallest key of the points that want to KILL t this round (OWNER_FREE when untouched)
//               <  0: simplex is dead, ~owner = a simplex created by the insertion that killed it (forwarding)
//   owner[8t+1] "ring word"  smallest key of the points that have t in the OUTER RING of their cavity this round.
//               Two winners may share an outer-ring simplex (they patch different neighbour slots of it); only
//               kill/kill and kill/ring overlaps exclude each other, and exactly the point with the worse key loses.
//   owner[8t+2..6]  float centre (relative to Mesh::sref origin), rin2, rout2
//   seed[v]   pending point: a simplex to start its walk from; -1 once inserted
//   ptTet[v]  a simplex created by v's insertion (seed for later points near v)
//
// A round (host loop in engine.cuh):
//   attempt  one thread per pending point selected this round: follow forwarding, visibility walk to the
//            containing simplex, flood the conflict region with exact in-sphere tests, atomicMin the point's
//            priority key on every killed simplex (key) and every surviving neighbour across the cavity
//            boundary (key|1); give up as soon as a better key is seen.
//   check    a point wins iff it still owns its whole footprint; winners are compacted and get a block of new
//            simplex slots from the bump allocator (warp prefix sum + one atomic per warp).
//   retri    winners only: write new simplices, patch outer back-pointers, link siblings by pivoting around the
//            shared edge (3D) / vertex (2D) through the dead cavity, then mark the cavity dead with forwarding.
// Two winners of one round have disjoint footprints (killed + outer ring), so their writes never overlap and the
// conflict region of one is unchanged by the other (new circumspheres lie inside the union of the two old ones).
#pragma once
#include "predicates.cuh"
#include "sphere.cuh"

namespace vor {

template <int D> struct Dim;
template <> struct Dim<3> { using Pt = double4; static constexpr int M = 4; };
template <> struct Dim<2> { using Pt = double2; static constexpr int M = 3; };

template <int D> struct Mesh {
    typename Dim<D>::Pt *pts;
    int4 *tet;    // interleaved records: tet[2t] = vertex ids, tet[2t+1] = neighbour codes (one 32 B sector per simplex)
    int *owner;   // OWS = 8 words per simplex: kill word, ring word, sphere filter (see above, sphere.cuh)
    int *seed;
    int *ptTet;
    Counters *cnt;
    int cap;      // simplex slots allocated
    int nsuper;   // vertices [0, nsuper) are super vertices
    SphereRef sref;   // origin of the float centres + query rounding allowance
};

template <int D> VOR_HD int &OWK(const Mesh<D> &m, int t) { return m.owner[OWS * (size_t)t]; }       // kill word / dead + forwarding
template <int D> VOR_HD int &OWR(const Mesh<D> &m, int t) { return m.owner[OWS * (size_t)t + 1]; }   // ring word
template <int D> VOR_HD int4 &TV(const Mesh<D> &m, int t) { return m.tet[REC4 * (size_t)t + TVO4]; }
template <int D> VOR_HD int4 &TN(const Mesh<D> &m, int t) { return m.tet[REC4 * (size_t)t + TVO4 + 1]; }
template <int D> VOR_HD int &TNI(const Mesh<D> &m, int t, int i) { return reinterpret_cast<int *>(m.tet)[4 * (REC4 * (size_t)t + TVO4) + 4 + i]; }

// One simplex record (vertex ids + neighbour codes, one 32 B sector) or one 3D vertex (double4, one sector) moves with
// ONE 256-bit instruction (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256).  Measured on the B200 (tools/micro/gather_bench.cu):
// once the gather footprint exceeds the TLB reach (~256 MB) the SMs sustain ~40 G scattered load INSTRUCTIONS per
// second per thread-lane whatever their width, so a 32 B record fetched as 2 x 128 bit gathers at half the rate
// (20.9 vs 38.2 G records/s over 16 GB).
template <int D> VOR_HD void load_rec(const Mesh<D> &m, int t, int4 &tv, int4 &tn) {
#ifdef __CUDA_ARCH__
    asm volatile("ld.global.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(tv.x), "=r"(tv.y), "=r"(tv.z), "=r"(tv.w), "=r"(tn.x), "=r"(tn.y), "=r"(tn.z), "=r"(tn.w)
                 : "l"(m.tet + REC4 * (size_t)t + TVO4));
#else
    tv = TV(m, t); tn = TN(m, t);
#endif
}
template <int D> VOR_HD void load_rec_cg(const Mesh<D> &m, int t, int4 &tv, int4 &tn) {   // L2 only (like __ldcg)
#ifdef __CUDA_ARCH__
    asm volatile("ld.global.cg.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(tv.x), "=r"(tv.y), "=r"(tv.z), "=r"(tv.w), "=r"(tn.x), "=r"(tn.y), "=r"(tn.z), "=r"(tn.w)
                 : "l"(m.tet + REC4 * (size_t)t + TVO4));
#else
    tv = TV(m, t); tn = TN(m, t);
#endif
}
template <int D> VOR_HD void store_rec(const Mesh<D> &m, int t, const int4 &tv, const int4 &tn) {
#ifdef __CUDA_ARCH__
    asm volatile("st.global.v8.s32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
                 :: "r"(tv.x), "r"(tv.y), "r"(tv.z), "r"(tv.w), "r"(tn.x), "r"(tn.y), "r"(tn.z), "r"(tn.w), "l"(m.tet + REC4 * (size_t)t + TVO4)
                 : "memory");
#else
    TV(m, t) = tv; TN(m, t) = tn;
#endif
}
VOR_HD double4 load_pt(const double4 *p) {
#ifdef __CUDA_ARCH__
    double4 r;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
#else
    return *p;
#endif
}
VOR_HD double2 load_pt(const double2 *p) { return *p; }

// ---- the 32 B ownership + sphere block of a simplex (sphere.cuh)
struct OwnBlk { int kill, ring; float cx, cy, cz, rin2, rout2; int spare; };
// one 256-bit gather, L2 only (the ownership words are updated by other SMs during the attempt kernel)
template <int D> VOR_HD OwnBlk load_blk(const Mesh<D> &m, int t) {
    OwnBlk b;
#ifdef __CUDA_ARCH__
    int c0, c1, c2, c3, c4;
    asm volatile("ld.global.cg.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(b.kill), "=r"(b.ring), "=r"(c0), "=r"(c1), "=r"(c2), "=r"(c3), "=r"(c4), "=r"(b.spare)
                 : "l"(m.owner + OWS * (size_t)t));
    b.cx = __int_as_float(c0); b.cy = __int_as_float(c1); b.cz = __int_as_float(c2); b.rin2 = __int_as_float(c3); b.rout2 = __int_as_float(c4);
#else
    const int *w = m.owner + OWS * (size_t)t;
    b.kill = w[0]; b.ring = w[1]; b.cx = i2f(w[2]); b.cy = i2f(w[3]); b.cz = i2f(w[4]); b.rin2 = i2f(w[5]); b.rout2 = i2f(w[6]); b.spare = w[7];
#endif
    return b;
}
// a NEW simplex: ownership words free, sphere filter of its vertices (one 256-bit store)
template <int D> VOR_HD void store_blk(const Mesh<D> &m, int t, const SphereBlk &s) {
#ifdef __CUDA_ARCH__
    asm volatile("st.global.v8.s32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
                 :: "r"(OWNER_FREE), "r"(OWNER_FREE), "r"(__float_as_int(s.cx)), "r"(__float_as_int(s.cy)), "r"(__float_as_int(s.cz)),
                    "r"(__float_as_int(s.rin2)), "r"(__float_as_int(s.rout2)), "r"(0), "l"(m.owner + OWS * (size_t)t)
                 : "memory");
#else
    int *w = m.owner + OWS * (size_t)t;
    w[0] = OWNER_FREE; w[1] = OWNER_FREE; w[2] = f2i(s.cx); w[3] = f2i(s.cy); w[4] = f2i(s.cz); w[5] = f2i(s.rin2); w[6] = f2i(s.rout2); w[7] = 0;
#endif
}
// query point relative to the origin of the float centres
struct RelPt { double x, y, z; };
template <int D> VOR_HD RelPt rel_pt(const Mesh<D> &m, const double4 &p) { RelPt q; q.x = p.x - m.sref.ox; q.y = p.y - m.sref.oy; q.z = p.z - m.sref.oz; return q; }
template <int D> VOR_HD RelPt rel_pt(const Mesh<D> &m, const double2 &p) { RelPt q; q.x = p.x - m.sref.ox; q.y = p.y - m.sref.oy; q.z = 0.0; return q; }
VOR_HD int sphere_test(const OwnBlk &b, const RelPt &q) { return sphere_test(b.cx, b.cy, b.cz, b.rin2, b.rout2, q.x, q.y, q.z); }

struct Scratch {
    int *killed, *bfacet, *bouter;     // contiguous per slot: entry j of slot s at [s * cap + j] (coalesced for a lane group)
    int *slotAct, *slotNk, *slotNb, *slotStatus, *slotBig;
    int nslots, capk, capb;
    int *bigK, *bigF, *bigO;           // overflow slots, contiguous per slot
    int nbig, bigCapK, bigCapB;
    int *winners, *wbase;
    int *slowSlots;                    // slots queued for the exact twin of the attempt kernel this round (coop_kernels.cuh)
    int4 *slotInfo;                    // cooperative kernels: {status | flagged << 2 | (overflow slot + 1) << 3, nk, nb, vertex} of a slot in ONE
                                       // 16 B word: the commit kernel starts from one load instead of five over three dependent levels
};

enum : int { ST_LOST = 0, ST_OK = 1 };

struct ScrView { int *k, *f, *o; int stride, capk, capb; };
VOR_HD ScrView scr_view(const Scratch &s, int slot, int big) {
    ScrView v;
    if (big < 0) {
        v.k = s.killed + (size_t)slot * s.capk; v.f = s.bfacet + (size_t)slot * s.capb; v.o = s.bouter + (size_t)slot * s.capb;
        v.stride = 1; v.capk = s.capk; v.capb = s.capb;
    } else {
        v.k = s.bigK + (size_t)big * s.bigCapK; v.f = s.bigF + (size_t)big * s.bigCapB; v.o = s.bigO + (size_t)big * s.bigCapB;
        v.stride = 1; v.capk = s.bigCapK; v.capb = s.bigCapB;
    }
    return v;
}

// warp-aggregated increment of a global counter (one atomic per warp)
VOR_HD int agg_inc(int *ctr) {
#ifdef __CUDA_ARCH__
    const unsigned mask = __activemask();
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(ctr, __popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
#else
    return atomic_add_i(ctr, 1);
#endif
}

// ------------------------------------------------------------------------------------------
// geometry helpers on the store
// ------------------------------------------------------------------------------------------
template <int D> struct Geo;

template <> struct Geo<3> {
    using Pt = double4;
    struct Verts { Pt p0, p1, p2, p3; };
    static VOR_HD Verts load(const Mesh<3> &m, const int4 &v) {
        Verts r; r.p0 = load_pt(m.pts + v.x); r.p1 = load_pt(m.pts + v.y); r.p2 = load_pt(m.pts + v.z); r.p3 = load_pt(m.pts + v.w); return r;
    }
    // bit i set <=> p is strictly beyond face i (orientation with vertex i replaced by p is negative)
    template <class CX> static VOR_HD int beyond_mask(CX &cx, const Verts &t, const Pt &p) {
        int mk = 0;
        if (orient3d(cx, p, t.p1, t.p2, t.p3) < 0) mk |= 1;
        if (orient3d(cx, t.p0, p, t.p2, t.p3) < 0) mk |= 2;
        if (orient3d(cx, t.p0, t.p1, p, t.p3) < 0) mk |= 4;
        if (orient3d(cx, t.p0, t.p1, t.p2, p) < 0) mk |= 8;
        return mk;
    }
    template <class CX> static VOR_HD int conflict(CX &cx, const Verts &t, const Pt &p) { return insphere(cx, t.p0, t.p1, t.p2, t.p3, p) > 0; }
    template <class CX> static VOR_HD int orient(CX &cx, const Verts &t) { return orient3d(cx, t.p0, t.p1, t.p2, t.p3); }
};

template <> struct Geo<2> {
    using Pt = double2;
    struct Verts { Pt p0, p1, p2; };
    static VOR_HD Verts load(const Mesh<2> &m, const int4 &v) {
        Verts r; r.p0 = m.pts[v.x]; r.p1 = m.pts[v.y]; r.p2 = m.pts[v.z]; return r;
    }
    template <class CX> static VOR_HD int beyond_mask(CX &cx, const Verts &t, const Pt &p) {
        int mk = 0;
        if (orient2d(cx, p, t.p1, t.p2) < 0) mk |= 1;
        if (orient2d(cx, t.p0, p, t.p2) < 0) mk |= 2;
        if (orient2d(cx, t.p0, t.p1, p) < 0) mk |= 4;
        return mk;
    }
    template <class CX> static VOR_HD int conflict(CX &cx, const Verts &t, const Pt &p) { return incircle(cx, t.p0, t.p1, t.p2, p) > 0; }
    template <class CX> static VOR_HD int orient(CX &cx, const Verts &t) { return orient2d(cx, t.p0, t.p1, t.p2); }
};

// sphere filter of a simplex given its vertex ids (gathers the coordinates)
VOR_HD SphereBlk sphere_of(const Mesh<3> &m, const int4 &v) {
    return sphere_make(load_pt(m.pts + v.x), load_pt(m.pts + v.y), load_pt(m.pts + v.z), load_pt(m.pts + v.w), m.sref);
}
VOR_HD SphereBlk sphere_of(const Mesh<2> &m, const int4 &v) { return sphere_make(m.pts[v.x], m.pts[v.y], m.pts[v.z], m.sref); }

// conflict test of simplex n against p (q = p relative to the sphere origin): the stored sphere decides, the
// determinant predicate (FP64 filter -> exact) takes the undecided shell.  delaunay_tree.rs:40-46 / geometry.rs:91-97.
template <int D, class CX> VOR_HD int conflict_at(CX &cx, const Mesh<D> &m, int n, const OwnBlk &b, const typename Dim<D>::Pt &p, const RelPt &q) {
    const int sv = sphere_test(b, q);
    if (sv) return sv > 0;
    atomic_add_ull(&m.cnt->sph_undecided, 1ULL);
    return Geo<D>::conflict(cx, Geo<D>::load(m, TV(m, n)), p);
}

// ------------------------------------------------------------------------------------------
// attempt: locate + conflict region + reservation
// ------------------------------------------------------------------------------------------
template <int D> struct AttemptArgs {
    Mesh<D> m;
    Scratch scr;
    const int *act;     // pending vertex ids of this stage
    int bits;           // priority bits (2^bits >= number of active entries)
    uint32_t salt;      // per-round salt of the priority hash
    uint32_t thr;       // random selection: attempt iff priority < thr
    int stride;         // stratified selection (stride > 0): attempt iff (a + offset) % stride == 0 -- one point per
    int offset;         //   run of `stride` consecutive (Morton-ordered) active entries, rotating every round
    int keybase;        // epoch << (bits + 1)
    int stats;          // accumulate W/E counters
    int *slowFlag;      // cooperative kernels: per vertex, != 0 = handed to the exact ("slow") twin (see coop_kernels.cuh); may be null
};

template <int D> VOR_HD void attempt_body(const AttemptArgs<D> &A, int a) {
    constexpr int M = Dim<D>::M;
    using G = Geo<D>;
    const Mesh<D> &m = A.m;
    const int v = A.act[a];
    int s = m.seed[v];
    if (s < 0) return;                                   // already inserted
    const uint32_t q = bij_hash((uint32_t)a, A.bits, A.salt);
    if (A.stride > 0 ? ((a + A.offset) % A.stride != 0) : (q >= A.thr)) return;   // not selected this round
    const int slot = agg_inc(&m.cnt->nslots);
    if (slot >= A.scr.nslots) return;                    // scratch exhausted: wait for a later round
    A.scr.slotAct[slot] = a;
    A.scr.slotStatus[slot] = ST_LOST;
    A.scr.slotBig[slot] = -1;
    PredCtx cx{m.cnt};
    const typename G::Pt p = m.pts[v];
    const RelPt rq = rel_pt(m, p);

    // -- forwarding: a dead seed points at a simplex created by its killer
    int o;
    while ((o = OWK(m, s)) < 0) s = ~o;

    // -- visibility walk; it may stop at ANY simplex in conflict with p (the conflict region is connected and the
    // flood below finds all of it from any member): the stored sphere of the simplex the walk stands in is tested first
    unsigned rot = (unsigned)v * 2654435761u;
    unsigned steps = 0;
    typename G::Verts tvv = G::load(m, TV(m, s));
    bool hit = false;
    for (;;) {
        if (sphere_test(load_blk(m, s), rq) > 0) { hit = true; break; }
        const int mk = G::beyond_mask(cx, tvv, p);
        if (mk == 0) break;
        int go = 0;
        const int r0 = (int)((rot >> 16) % (unsigned)M);
        for (int k = 0; k < M; k++) {
            const int i = (r0 + k) % M;
            if ((mk >> i) & 1) { go = i; break; }
        }
        const int code = get4(TN(m, s), go);
        if (code < 0) { set_err(m.cnt, ERR_OUTSIDE); return; }
        s = code >> 2;
        rot = rot * 1664525u + 1013904223u;
        if (++steps > (1u << 22)) { set_err(m.cnt, ERR_WALK); return; }
        tvv = G::load(m, TV(m, s));
    }
    m.seed[v] = s;

    // -- conflict region with reservation
    const int key_k = A.keybase | (int)(q << 1);
    const int key_o = key_k | 1;
    unsigned tests = 1;
    if (!hit && !G::conflict(cx, tvv, p)) {
        // p coincides with a vertex of its containing simplex: duplicate input point.  Drop it.
        m.seed[v] = -1;
        atomic_add_i(&m.cnt->ndup, 1);
        return;
    }
    if (OWR(m, s) < key_k) goto lost;                       // a better point keeps s in its outer ring
    if (atomic_min_i(&OWK(m, s), key_k) < key_k) goto lost;
    {
        ScrView sv = scr_view(A.scr, slot, -1);
        int big = -1;
        int nk = 1, nb = 0;
        sv.k[0] = s;
        for (int head = 0; head < nk; head++) {
            const int t = sv.k[(size_t)head * sv.stride];
            const int4 nbr = TN(m, t);
            for (int i = 0; i < M; i++) {
                const int code = get4(nbr, i);
                int isout = 1;
                if (code >= 0) {
                    const int n = code >> 2;
                    const OwnBlk blk = load_blk(m, n);
                    const int ow = blk.kill, orr = blk.ring;
                    if (ow == key_k) continue;               // already in my cavity
                    if (ow < key_k) goto lost;               // a better point kills it (or it is dead)
                    if (orr != key_o) {                      // not yet classified by me (or a better point shares the ring)
                        tests++;
                        if (conflict_at(cx, m, n, blk, p, rq)) {
                            if (orr < key_k) goto lost;          // a better point keeps n in its outer ring
                            if (atomic_min_i(&OWK(m, n), key_k) < key_k) goto lost;
                            isout = 0;
                            if (nk == sv.capk) {
                                if (big >= 0) { set_err(m.cnt, ERR_CAPACITY); goto lost; }
                                big = atomic_add_i(&m.cnt->nbig, 1);
                                if (big >= A.scr.nbig) goto lost;
                                const ScrView bv = scr_view(A.scr, slot, big);
                                for (int j = 0; j < nk; j++) bv.k[j] = sv.k[(size_t)j * sv.stride];
                                for (int j = 0; j < nb; j++) { bv.f[j] = sv.f[(size_t)j * sv.stride]; bv.o[j] = sv.o[(size_t)j * sv.stride]; }
                                sv = bv;
                                A.scr.slotBig[slot] = big;
                            }
                            sv.k[(size_t)nk * sv.stride] = n;
                            nk++;
                        } else {
                            atomic_min_i(&OWR(m, n), key_o);     // sharing the ring with a better point is fine
                        }
                    }
                }
                if (isout) {
                    if (nb == sv.capb) {
                        if (big >= 0) { set_err(m.cnt, ERR_CAPACITY); goto lost; }
                        big = atomic_add_i(&m.cnt->nbig, 1);
                        if (big >= A.scr.nbig) goto lost;
                        const ScrView bv = scr_view(A.scr, slot, big);
                        for (int j = 0; j < nk; j++) bv.k[j] = sv.k[(size_t)j * sv.stride];
                        for (int j = 0; j < nb; j++) { bv.f[j] = sv.f[(size_t)j * sv.stride]; bv.o[j] = sv.o[(size_t)j * sv.stride]; }
                        sv = bv;
                        A.scr.slotBig[slot] = big;
                    }
                    sv.f[(size_t)nb * sv.stride] = t * 4 + i;
                    sv.o[(size_t)nb * sv.stride] = code;
                    nb++;
                }
            }
        }
        A.scr.slotNk[slot] = nk;
        A.scr.slotNb[slot] = nb;
        A.scr.slotStatus[slot] = ST_OK;
        if (A.stats) {
            atomic_add_ull(&m.cnt->walk_steps, steps);
            atomic_add_ull(&m.cnt->tests, tests);
            atomic_add_ull(&m.cnt->attempts, 1ULL);
        }
        return;
    }
lost:
    if (A.stats) {
        atomic_add_ull(&m.cnt->walk_steps, steps);
        atomic_add_ull(&m.cnt->tests, tests);
        atomic_add_ull(&m.cnt->attempts, 1ULL);
        atomic_add_ull(&m.cnt->aborted, 1ULL);
    }
}

// ------------------------------------------------------------------------------------------
// check: ownership of the whole footprint -> winners, simplex slot allocation
// ------------------------------------------------------------------------------------------
template <int D> struct CheckArgs {
    Mesh<D> m;
    Scratch scr;
    int bits;
    uint32_t salt;
    int keybase;
    const int *slowFlag;
};

// `valid` = tid addresses a claimed slot.  No early return before the allocation so that whole warps reach it.
template <int D> VOR_HD void check_body(const CheckArgs<D> &A, int slot, bool valid) {
    const Mesh<D> &m = A.m;
    int win = 0, nb = 0;
    // the launch covers an upper bound of slots; only those claimed THIS round hold current data
    if (valid && slot < m.cnt->nslots && slot < A.scr.nslots && A.scr.slotStatus[slot] == ST_OK) {
        const int a = A.scr.slotAct[slot];
        const uint32_t q = bij_hash((uint32_t)a, A.bits, A.salt);
        const int key_k = A.keybase | (int)(q << 1);
        const int key_o = key_k | 1;
        const ScrView sv = scr_view(A.scr, slot, A.scr.slotBig[slot]);
        const int nk = A.scr.slotNk[slot];
        nb = A.scr.slotNb[slot];
        win = 1;
        (void)key_o;
        for (int j = 0; j < nk && win; j++) {
            const int t = sv.k[(size_t)j * sv.stride];
            if (OWK(m, t) != key_k || OWR(m, t) < key_k) win = 0;   // best killer, and no better point has t in its ring
        }
        for (int j = 0; j < nb && win; j++) {
            const int code = sv.o[(size_t)j * sv.stride];
            if (code >= 0 && OWK(m, code >> 2) < key_k) win = 0;    // no better point kills my outer ring
        }
    }
#ifdef __CUDA_ARCH__
    // warp prefix sums of (win, nb) -> one atomic pair per warp
    const int lane = threadIdx.x & 31;
    const unsigned wmask = __ballot_sync(0xffffffffu, win);
    if (wmask == 0) return;
    int incl = win ? nb : 0;
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int wb = 0, tb = 0;
    if (lane == 0) {
        wb = atomicAdd(&m.cnt->nwinners, __popc(wmask));
        tb = atomicAdd(&m.cnt->ntets, total);
    }
    wb = __shfl_sync(0xffffffffu, wb, 0);
    tb = __shfl_sync(0xffffffffu, tb, 0);
    if (win) {
        const int w = wb + __popc(wmask & ((1u << lane) - 1u));
        A.scr.winners[w] = slot;
        A.scr.wbase[w] = tb + incl - nb;
    }
#else
    if (win) {
        const int w = atomic_add_i(&m.cnt->nwinners, 1);
        A.scr.winners[w] = slot;
        A.scr.wbase[w] = atomic_add_i(&m.cnt->ntets, nb);
    }
#endif
}

// ------------------------------------------------------------------------------------------
// retriangulate: winners write their new simplices and repair adjacency
// ------------------------------------------------------------------------------------------
template <int D> struct RetriArgs {
    Mesh<D> m;
    Scratch scr;
    const int *act;
    int stats;
};

template <int D> VOR_HD void retri_body(const RetriArgs<D> &A, int w) {
    constexpr int M = Dim<D>::M;
    const Mesh<D> &m = A.m;
    const int slot = A.scr.winners[w];
    const int base = A.scr.wbase[w];
    const int a = A.scr.slotAct[slot];
    const int v = A.act[a];
    const ScrView sv = scr_view(A.scr, slot, A.scr.slotBig[slot]);
    const int nk = A.scr.slotNk[slot];
    const int nb = A.scr.slotNb[slot];
    if (base + nb > m.cap) { set_err(m.cnt, ERR_OOM); return; }

    // phase A: new simplex j hangs on boundary facet j; link it to the outer simplex and leave a marker
    // -(code)-2 in the dead simplex so that the pivots of phase B can find it.
    for (int j = 0; j < nb; j++) {
        const int fc = sv.f[(size_t)j * sv.stride];
        const int t = fc >> 2, i = fc & 3;
        const int outer = sv.o[(size_t)j * sv.stride];
        const int T = base + j;
        int4 verts = TV(m, t);
        set4(verts, i, v);
        TV(m, T) = verts;
        store_blk(m, T, sphere_of(m, verts));
        TNI(m, T, i) = outer;
        if (M == 3) TNI(m, T, 3) = -1;
        if (outer >= 0) TNI(m, outer >> 2, outer & 3) = T * 4 + i;
        TNI(m, t, i) = -(T * 4 + i) - 2;
    }
    // phase B: sibling links.  The face of T opposite slot k (k != i) contains v and the ridge R = T's vertices
    // other than slots i,k.  Pivot around R through the dead cavity until a marker is met.
    for (int j = 0; j < nb; j++) {
        const int fc = sv.f[(size_t)j * sv.stride];
        const int t = fc >> 2, i = fc & 3;
        const int T = base + j;
        const int4 tverts = TV(m, t);
        for (int k = 0; k < M; k++) {
            if (k == i) continue;
            int cur = t, enter = i, exitf = k;
            int4 cv = tverts;
            // ridge vertices (D-1 of them): slots of t other than i,k
            int r0 = -1, r1 = -1;
            for (int sidx = 0; sidx < M; sidx++) {
                if (sidx == i || sidx == k) continue;
                if (r0 < 0) r0 = get4(cv, sidx); else r1 = get4(cv, sidx);
            }
            for (;;) {
                const int e = TNI(m, cur, exitf);
                if (e <= -2) {
                    const int sc = -(e + 2);
                    TNI(m, T, k) = (sc >> 2) * 4 + enter;
                    break;
                }
                const int nxt = e >> 2, jb = e & 3;
                cv = TV(m, nxt);
                int y = -1;
                for (int sidx = 0; sidx < M; sidx++) {
                    if (sidx == jb) continue;
                    const int vv = get4(cv, sidx);
                    if (vv != r0 && vv != r1) y = sidx;
                }
                cur = nxt; enter = jb; exitf = y;
            }
        }
    }
    // phase C: the cavity dies; dead simplices forward to a new one
    for (int j = 0; j < nk; j++) OWK(m, sv.k[(size_t)j * sv.stride]) = ~base;
    m.ptTet[v] = base;
    m.seed[v] = -1;
    atomic_add_ull(&m.cnt->win_total, 1ULL);
    atomic_add_ull(&m.cnt->created_all, (unsigned long long)nb);
    if (A.stats) {
        atomic_add_ull(&m.cnt->killed, (unsigned long long)nk);
        atomic_add_ull(&m.cnt->created, (unsigned long long)nb);
    }
}

// ------------------------------------------------------------------------------------------
// small maintenance kernels
// ------------------------------------------------------------------------------------------
struct ResetOwnerArgs { int *owner; const Counters *cnt; };
VOR_HD void reset_owner_body(const ResetOwnerArgs &A, int t) {
    if (t >= A.cnt->ntets) return;   // the launch covers the whole store: the host's simplex count may be a batch behind
    if (A.owner[OWS * (size_t)t] >= 0) A.owner[OWS * (size_t)t] = OWNER_FREE;
    A.owner[OWS * (size_t)t + 1] = OWNER_FREE;
}

struct MarkDeadArgs { int *owner; int first; };   // slots handed out beyond the capacity of the store: dead, never referenced
VOR_HD void mark_dead_body(const MarkDeadArgs &A, int i) { A.owner[OWS * (size_t)(A.first + i)] = -1; }

struct FillArgs { int *p; int val; };
VOR_HD void fill_body(const FillArgs &A, int i) { A.p[i] = A.val; }

struct IotaArgs { int *p; int first; };
VOR_HD void iota_body(const IotaArgs &A, int i) { A.p[i] = A.first + i; }

// order-preserving compaction of the active list (3 passes: count, scan of block counts, scatter)
struct CompactArgs { const int *act; const int *seed; int *out; int *blockCnt; int n; int chunk; };
VOR_HD void compact_count_body(const CompactArgs &A, int b) {
    const int lo = b * A.chunk, hi = lo + A.chunk < A.n ? lo + A.chunk : A.n;
    int c = 0;
    for (int i = lo; i < hi; i++) c += A.seed[A.act[i]] >= 0;
    A.blockCnt[b] = c;
}
VOR_HD void compact_scatter_body(const CompactArgs &A, int b) {
    const int lo = b * A.chunk, hi = lo + A.chunk < A.n ? lo + A.chunk : A.n;
    int w = A.blockCnt[b]; // exclusive prefix after the scan
    for (int i = lo; i < hi; i++) {
        const int v = A.act[i];
        if (A.seed[v] >= 0) A.out[w++] = v;
    }
}

} // namespace vor
