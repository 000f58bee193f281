// coop_kernels.cuh -- lane-group-cooperative round kernels (CUDA only): the product path on the B200.
//
// G lanes (3D: a whole warp, north_star's "one warp per point"; 2D: 16, two points per warp) work on ONE pending point.  The
// serial dependency chain of the thread-per-point bodies in kernels.cuh (one in-sphere test after the other) becomes one chain
// per BFS LEVEL of the conflict region.  A round is three launches, chained by programmatic dependent launch on small rounds:
//   k_attempt_hot   forwarding -> walk on the power distance of the cached spheres (lane i < M gathers the 64 B line of neighbour
//            i: block + neighbour codes) -> flood: each lane takes one (frontier simplex, facet) item from the cavity staged in
//            shared memory, gathers the neighbour's 32 B block -- ownership words + certified sphere filter (sphere.cuh) -- with
//            ONE 256-bit load and decides from it; reservation by fire-and-forget red.min; new killed simplices / boundary
//            facets are appended with ballot + popc prefix sums.  NO determinant code: 56 registers, spill-free.  Its MID twin
//            (template parameter) puts the FP64 determinant filter behind the sphere filter for input that keeps leaving it;
//            k_attempt_slow / k_attempt_coop are the exact twin (visibility walk with exact orient, filter -> FP64 determinant ->
//            double-double -> exact integers) for the points the hot kernels hand over.  The result of a slot leaves as one 16 B word.
//   k_commit_coop   (check + retriangulate fused) starts from that word, stages the cavity in shared memory with one level of
//            gathers, lanes vote on ownership; a winner takes its block of simplex slots with one atomicAdd and retriangulates
//            inside the SM (neighbour codes translated to local indices once, one lane per new simplex pivots around its ridges).
//   k_spheres       block (free ownership words + sphere filter) of every simplex the commit just created.
// What bounds them: on full rounds DRAM 26 / 41 / 53 % of peak with ~50 % of the issue slots busy; on the ~600 rounds below ~32 k
// slots the latency of ONE attempt + commit + sphere pass (profiles/r2_round_ncu.md).  The compile-time switches
// below record the alternatives that were measured and lost.
// Same scratch format and same semantics as kernels.cuh (reference: delaunay_tree.rs:33-123, :213-334, scheduler.rs:6-55),
// so the two implementations are interchangeable (option "coop"); tests/emu exercises the thread-per-point bodies on
// the CPU, tests/test_gpu_* exercise these against the oracle on the B200.
#pragma once
#include "kernels.cuh"

#if VOR_GPU
namespace vor {

// Programmatic dependent launch (sm_90+): the three kernels of a round depend on each other completely, but the NEXT kernel's
// blocks can be set up and made resident while the current one drains -- they wait here until it has completed and flushed.
// pdl_wait() must come before the first global access; pdl_trigger() lets the dependent launch begin early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <int G> __device__ __forceinline__ unsigned group_mask() {
    if (G == 32) return 0xffffffffu;
    const int lane = threadIdx.x & 31;
    return ((1u << G) - 1u) << ((lane / G) * G);
}

// Selection is implicit and stratified: slot g of a round attempts active entry a = g * stride + offset, i.e. one
// point per run of `stride` consecutive entries of the Morton-ordered active list (rotating offset), so no selection
// kernel and no compaction of the selected set are needed.  A round is two launches: attempt, commit.
struct RoundSel { int nact; int stride; int offset; int nsel; int first; int last; };   // this launch works on slots [first, last) of the round's nsel
__device__ __forceinline__ int slot_entry(const RoundSel &rs, int slot) { return slot * rs.stride + rs.offset; }
// per-slot result word of an attempt (Scratch::slotInfo); a slot without a complete cavity only needs its status
__device__ __forceinline__ void slot_lost(const Scratch &scr, int slot) { reinterpret_cast<int *>(scr.slotInfo + slot)[0] = ST_LOST; }
__device__ __forceinline__ int4 slot_pack(int status, bool flagged, int big, int nk, int nb, int v) {
    return make_int4(status | (flagged ? 4 : 0) | ((big + 1) << 3), nk, nb, v);
}

template <int D> struct GeoCoop;
template <> struct GeoCoop<3> {
    // orientation of the simplex with vertex `k` replaced by p (lane k of the group evaluates facet k)
    template <class CX> static __device__ __forceinline__ int orient_repl(CX &cx, const Geo<3>::Verts &t, const double4 &p, int k) {
        const double4 a = k == 0 ? p : t.p0, b = k == 1 ? p : t.p1, c = k == 2 ? p : t.p2, d = k == 3 ? p : t.p3;
        return orient3d(cx, a, b, c, d);
    }
};
template <> struct GeoCoop<2> {
    template <class CX> static __device__ __forceinline__ int orient_repl(CX &cx, const Geo<2>::Verts &t, const double2 &p, int k) {
        const double2 a = k == 0 ? p : t.p0, b = k == 1 ? p : t.p1, c = k == 2 ? p : t.p2;
        return orient2d(cx, a, b, c);
    }
};

// ------------------------------------------------------------------------------------------
// attempt
// ------------------------------------------------------------------------------------------
// Small blocks: the groups of a block are independent, and a block keeps its registers until its SLOWEST group is
// done; with 8 warps per block the achieved occupancy was 30 % of the 50 % the 64 registers allow (ncu, round 1);
// measured attempt-kernel time on the 10M-point run: 135 ms (256 threads), 126 (128), 124 (64), 123 (32).
#ifndef VOR_COOP_BLOCK
#define VOR_COOP_BLOCK 64
#endif
#ifndef VOR_SK
#define VOR_SK 64                 // killed simplices of a cavity staged in shared memory (ids + neighbour codes: 20 B each)
#endif
#ifndef VOR_ATTEMPT_BLOCK
#define VOR_ATTEMPT_BLOCK 32      // threads per block of the attempt kernel: one warp, so that a finished attempt frees its
                                  // registers at once (measured on the 10M-point run: 108 ms vs 116 ms with 64 threads)
#endif
#ifndef VOR_ATTEMPT_REGS
#define VOR_ATTEMPT_REGS 64       // registers per thread of the attempt kernel with the exact path (measured best of 80/64/48)
#endif
// The determinant test of one simplex (record + 4 vertex gathers + FP64 filter [+ exact integers]) for the conflict
// tests the stored sphere leaves undecided (~1e-5 of them on uniform input).  Out of line, operands by value: the hot loop
// must not pay registers for it.  Returns 1 in conflict, 0 not, 2 = EX == false and the FP64 filter failed too.
template <int D, bool EX>
__device__ __noinline__ int conflict_slow(const typename Dim<D>::Pt *pts, const int4 *tet, Counters *cnt, int n, double px, double py, double pz) {
    Mesh<D> mm{};
    mm.pts = const_cast<typename Dim<D>::Pt *>(pts);
    mm.tet = const_cast<int4 *>(tet);
    mm.cnt = cnt;
    PredCtxT<EX> cx{cnt};
    typename Dim<D>::Pt p;
    if constexpr (D == 3) { p.x = px; p.y = py; p.z = pz; p.w = 0.0; } else { p.x = px; p.y = py; }
    const int c = Geo<D>::conflict(cx, Geo<D>::load(mm, TV(mm, n)), p);
    if (!EX && cx.failed) return 2;
    return c;
}
template <int D> __device__ __forceinline__ double pt_z(const typename Dim<D>::Pt &p) { if constexpr (D == 3) return p.z; else return 0.0; }

// RED = 1: the kill reservation is a fire-and-forget reduction too (no round trip on the critical path of a flood
// level): lanes of one batch that reach the same simplex are deduplicated with match.any, and a better killer that
// slips in between the owner read and the reduction is caught by the ownership check of commit.
// EXACT = 0: the hot twin without the exact predicates in its call tree (PredCtxT<false>): an attempt whose filter fails
// is abandoned and its point flagged (A.slowFlag[v] = key base of the round); it skips flagged points.  EXACT = 1 with
// A.thr == 2: the slow twin, launched behind it while flagged points are pending, attempts ONLY points flagged in an
// EARLIER round (a point flagged in this round has left marks under this round's key).  A.thr == 0: everything.
//
// A conflict test is ONE 256-bit gather: the 32 B block of the neighbour holds its ownership words and its certified
// circumsphere filter (sphere.cuh).  Only killed simplices have their record read (neighbour codes, by the next level).
#ifndef VOR_SPEC
#define VOR_SPEC 1                // 1: the neighbour codes of a tested simplex are gathered together with its block (one round
                                  // trip per flood level); 0: only once the sphere says "in conflict" (a dependent second trip)
#endif
VOR_HD int4 load_nbr(const int4 *tet, int t) {
#ifdef __CUDA_ARCH__
    return __ldg(tet + REC4 * (size_t)t + TVO4 + 1);
#else
    return tet[REC4 * (size_t)t + TVO4 + 1];
#endif
}
template <int D, int G, int RED, int EXACT>
__device__ __forceinline__ void attempt_one(const AttemptArgs<D> &A, const RoundSel &rsel, const int gid, int *const sk, int4 *const sn) {
    constexpr int M = Dim<D>::M;
    constexpr int SK = VOR_SK;
    using Gm = Geo<D>;
    const Mesh<D> &m = A.m;
    const int gl = threadIdx.x & (G - 1);                          // lane inside the group
    const unsigned gmask = group_mask<G>();
    const int gshift = (threadIdx.x & 31) & ~(G - 1);              // first lane of the group inside the warp
    if (gid >= rsel.nsel) return;
    const int slot = gid;
    const int a = slot_entry(rsel, slot);
    if (a >= rsel.nact) { if (gl == 0) slot_lost(A.scr, slot); return; }
    const int v = A.act[a];
    int s = m.seed[v];
    if (s < 0) { if (gl == 0) slot_lost(A.scr, slot); return; }   // already inserted
    bool flagged = false;                                          // a point the hot twin handed over (commit counts it as done)
    if (A.slowFlag) {
        const int fl = A.slowFlag[v];
        flagged = fl != 0;
        if (EXACT ? (A.thr == 2u && (fl == 0 || fl == A.keybase)) : fl != 0) {
            // not this twin's point.  The hot twin runs first and owns the slot's status; the slow twin leaves the
            // status of the points it skips alone
            if (!EXACT && gl == 0) slot_lost(A.scr, slot);
            return;
        }
    }
    PredCtxT<(EXACT != 0)> cx{m.cnt};
    const typename Gm::Pt p = m.pts[v];
    const RelPt rq = rel_pt(m, p);
    const uint32_t q = bij_hash((uint32_t)slot, A.bits, A.salt);   // unique among the slots of this round
    const int key_k = A.keybase | (int)(q << 1);
    const int key_o = key_k | 1;

    int status = ST_LOST, nk = 0, nb = 0, big = -1;
    unsigned steps = 0, tests = 0;

    // -- forwarding (all lanes follow the same chain: broadcast loads); the block that says "dead, go there" or "alive"
    // also holds the sphere of the simplex: a seed whose sphere certainly contains p starts the flood at once
    OwnBlk sb = load_blk(m, s);
    int4 stn = load_nbr(m.tet, s);         // neighbour codes of the simplex the flood starts from (speculative for a dead seed)
    while (sb.kill < 0) { s = ~sb.kill; sb = load_blk(m, s); stn = load_nbr(m.tet, s); }
    bool hit = sphere_test(sb, rq) > 0;
    bool fail = false;
    typename Gm::Verts tvv;

    // -- visibility walk: lane k < M tests facet k; it stops at the containing simplex or at any simplex whose stored
    // sphere certainly contains p (the conflict region is connected: the flood finds all of it from any member)
    if (!hit) {
        unsigned rot = (unsigned)v * 2654435761u;
        int4 stv;                          // record of the simplex the walk stands in: one 256-bit gather per step
        load_rec(m, s, stv, stn);
        tvv = Gm::load(m, stv);
        for (;;) {
            const int ok = gl < M ? GeoCoop<D>::orient_repl(cx, tvv, p, gl) : 1;
            if (!EXACT && __any_sync(gmask, cx.failed)) break;
            const unsigned bal = (__ballot_sync(gmask, ok < 0) >> gshift) & ((1u << M) - 1u);
            if (bal == 0) break;
            int go = 0;
            const int r0 = (int)((rot >> 16) % (unsigned)M);
            for (int k = 0; k < M; k++) {
                const int i = (r0 + k) % M;
                if ((bal >> i) & 1) { go = i; break; }
            }
            const int code = get4(stn, go);
            if (code < 0) { if (gl == 0) set_err(m.cnt, ERR_OUTSIDE); fail = true; break; }
            s = code >> 2;
            rot = rot * 1664525u + 1013904223u;
            if (++steps > (1u << 22)) { if (gl == 0) set_err(m.cnt, ERR_WALK); fail = true; break; }
            sb = load_blk(m, s);           // independent of the record gather below
            load_rec(m, s, stv, stn);
            if (sphere_test(sb, rq) > 0) { hit = true; break; }
            tvv = Gm::load(m, stv);
        }
    }

    bool needSlow = !EXACT && __any_sync(gmask, cx.failed);
    if (needSlow) fail = true;
    if (!fail) {
        if (gl == 0) m.seed[v] = s;
        tests = gl == 0 ? 1u : 0u;
        if (!hit) {
            // -- containing simplex must be in conflict, otherwise p duplicates one of its vertices
            int c0 = 0;
            if (gl == 0) c0 = Gm::conflict(cx, tvv, p);
            c0 = __shfl_sync(gmask, c0, gshift);
            if (!EXACT && __any_sync(gmask, cx.failed)) { needSlow = true; fail = true; }
            else if (!c0) {
                if (gl == 0) {
                    m.seed[v] = -1;
                    atomicAdd(&m.cnt->ndup, 1);
                    if (EXACT && A.slowFlag && A.slowFlag[v] != 0) atomicAdd(&m.cnt->nflag_done, 1);   // a flagged point leaves as a duplicate
                }
                fail = true;
            }
        }
    }
    if (!fail) {
        // reservation of the first simplex: early out on the block already in registers (a better point kills it or keeps
        // it in its outer ring); anything that slips in later is caught by the ownership check of commit
        if (sb.kill < key_k || sb.ring < key_k) fail = true;
        else if (gl == 0) atomicMin(&OWK(m, s), key_k);
    }
    if (!fail) {
        ScrView sv = scr_view(A.scr, slot, -1);
        if (gl == 0) { sv.k[0] = s; sk[0] = s; sn[0] = stn; }
        __syncwarp(gmask);
        nk = 1;
        int head = 0;
        bool lost = false;
        while (head < nk && !lost) {
            const int tail = nk;
            const int items = (tail - head) * M;
            for (int base = 0; base < items && !lost; base += G) {
                const int j = base + gl;
                bool pushK = false, pushB = false, lostLane = false;
                int newT = 0, fcode = 0, ocode = 0;
                int claim = -1 - gl;   // RED: simplex this lane wants to kill (unique dummy otherwise)
                int4 nnb = make_int4(-1, -1, -1, -1);
                if (j < items) {
                    const int e = head + j / M;
                    const int i = j % M;
                    int t, code;
                    if (e < SK) { t = sk[e]; code = reinterpret_cast<const int *>(sn)[e * 4 + i]; }   // the cavity so far is staged in shared memory
                    else { t = sv.k[e]; code = TNI(m, t, i); }
                    if (code < 0) {
                        pushB = true; fcode = t * 4 + i; ocode = code;
                    } else {
                        const int n = code >> 2;
                        const OwnBlk blk = load_blk(m, n);      // ownership words + sphere: ONE 256-bit gather
                        if (VOR_SPEC) nnb = load_nbr(m.tet, n);
                        if (blk.kill == key_k) {
                            // already in my cavity
                        } else if (blk.kill < key_k) {
                            lostLane = true;           // a better point kills n (or n is dead)
                        } else if (blk.ring == key_o) {
                            pushB = true; fcode = t * 4 + i; ocode = code;   // already tested by me: not in conflict
                        } else {
                            tests++;
                            int conf = sphere_test(blk, rq);
                            if (conf == 0) {
                                if (A.stats) atomicAdd(&m.cnt->sph_undecided, 1ULL);   // one address: only when counters are asked for
                                conf = conflict_slow<D, (EXACT != 0)>(m.pts, m.tet, m.cnt, n, p.x, p.y, pt_z<D>(p));
                                if (conf == 2) { cx.failed = true; conf = 0; }
                            }
                            if (conf > 0) {
                                if (!VOR_SPEC) nnb = load_nbr(m.tet, n);
                                if (blk.ring < key_k) lostLane = true;   // a better point keeps n in its outer ring
                                else if (RED) {
                                    atomicMin(&OWK(m, n), key_k);
                                    claim = n;
                                } else {
                                    const int old = atomicMin(&OWK(m, n), key_k);
                                    if (old < key_k) lostLane = true;
                                    else if (old != key_k) { pushK = true; newT = n; }   // first lane to claim it appends it
                                }
                            } else if (!cx.failed) {
                                // outer-ring mark: fire and forget (RED, no round trip).  Rings may be shared; a better
                                // point that KILLS n was either seen above or is caught by the ownership check of commit.
                                atomicMin(&OWR(m, n), key_o);
                                pushB = true; fcode = t * 4 + i; ocode = code;
                            }
                        }
                    }
                }
                if (__any_sync(gmask, lostLane || (!EXACT && cx.failed))) { lost = true; break; }
                if (RED) {
                    const unsigned same = __match_any_sync(gmask, claim);
                    if (claim >= 0 && (__ffs(same) - 1) == (threadIdx.x & 31)) { pushK = true; newT = claim; }
                }
                const unsigned mk = (__ballot_sync(gmask, pushK) >> gshift) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
                const unsigned mb = (__ballot_sync(gmask, pushB) >> gshift) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
                const int ck = __popc(mk), cb = __popc(mb);
                if (nk + ck > sv.capk || nb + cb > sv.capb) {
                    // spill to an overflow slot (contiguous, much larger)
                    if (big >= 0) { if (gl == 0) set_err(m.cnt, ERR_CAPACITY); lost = true; break; }
                    if (gl == 0) big = atomicAdd(&m.cnt->nbig, 1);
                    big = __shfl_sync(gmask, big, gshift);
                    if (big >= A.scr.nbig) { lost = true; break; }
                    const ScrView bv = scr_view(A.scr, slot, big);
                    for (int x = gl; x < nk; x += G) bv.k[x] = sv.k[x];
                    for (int x = gl; x < nb; x += G) { bv.f[x] = sv.f[x]; bv.o[x] = sv.o[x]; }
                    sv = bv;
                    __syncwarp(gmask);
                    if (nk + ck > sv.capk || nb + cb > sv.capb) { if (gl == 0) set_err(m.cnt, ERR_CAPACITY); lost = true; break; }
                }
                const unsigned lt = (1u << gl) - 1u;
                if (pushK) {
                    const int pos = nk + __popc(mk & lt);
                    sv.k[pos] = newT;
                    if (pos < SK) { sk[pos] = newT; sn[pos] = nnb; }
                }
                if (pushB) { const int pos = nb + __popc(mb & lt); sv.f[pos] = fcode; sv.o[pos] = ocode; }
                nk += ck;
                nb += cb;
                __syncwarp(gmask);
            }
            head = tail;
        }
        if (!lost) status = ST_OK;
    }
    if (!EXACT) {
        needSlow = needSlow || __any_sync(gmask, cx.failed);
        if (needSlow && gl == 0 && A.slowFlag) {
            // hand the point to the slow twin (from the next round on); marks left under this round's key are stale then
            A.slowFlag[v] = A.keybase;
            atomicAdd(&m.cnt->nflag_set, 1);
            status = ST_LOST;
        }
    }
    if (gl == 0) A.scr.slotInfo[slot] = slot_pack(status, flagged, big, nk, nb, v);
    if (A.stats) {
        // every lane counted its own tests
        for (int d = G / 2; d > 0; d >>= 1) tests += __shfl_xor_sync(gmask, tests, d);
        if (gl == 0) {
            atomicAdd(&m.cnt->walk_steps, (unsigned long long)steps);
            atomicAdd(&m.cnt->tests, (unsigned long long)tests);
            atomicAdd(&m.cnt->attempts, 1ULL);
            if (status != ST_OK) atomicAdd(&m.cnt->aborted, 1ULL);
            else atomicAdd(&m.cnt->tests_ok, (unsigned long long)tests);
        }
    }
}

// ------------------------------------------------------------------------------------------
// hot attempt kernel: one warp per point, NO determinant code at all.
// Everything it decides, it decides from the 32 B blocks (ownership words + certified sphere filter):
//   locate   walk on the power distance pow(p, t) = |p - c_t|^2 - r_t^2.  The radical hyperplane of the circumspheres of two
//            adjacent simplices is the hyperplane of their common facet, so "the neighbour across facet i has a smaller
//            power" <=> "p lies beyond facet i": descending the power IS the visibility walk of the orientation
//            predicate, evaluated from the cached spheres (lane i < M gathers the block and the neighbour codes of neighbour
//            i: one round trip per step, steepest descent).  It stops at the first simplex whose sphere CERTAINLY
//            contains p: any member of the conflict region will do, the flood finds the rest.  The float values only
//            steer the walk; a walk that gets stuck (p within the filter's shell of a facet or a sphere, a duplicate
//            point, a simplex without a filter) hands the point to the exact twin.
//   flood    as attempt_one, but a test the sphere filter leaves undecided (~1e-5 of them) abandons the attempt and
//            hands the point to the exact twin (slowFlag[v] = key base of this round; from the next round on this kernel
//            appends the point's slot to the round's list of slots for the exact twin, k_attempt_slow).
// With neither orient3d nor insphere in its call tree the kernel runs spill-free at 56 registers (116 needed before: the
// version with the walk predicates inside moved 6x more local-memory than global-memory sectors, ncu round 2).
// ------------------------------------------------------------------------------------------
#ifndef VOR_HOT_BLOCK
#define VOR_HOT_BLOCK 64
#endif
#ifndef VOR_HOT_REGS
#define VOR_HOT_REGS 56
#endif
#ifndef VOR_HOT_WALK
#define VOR_HOT_WALK 48           // power-descent steps before the point goes to the exact twin
#endif
// Lane groups of the hot kernel: G lanes per attempted point.  A 3D flood level has (frontier x 4) items -- 4, ~12, ~24, ~30 -- and
// a 2D one 3, 6, 6: a whole warp per point leaves most lanes idle, and what bounds the kernel is the number of attempts in
// flight, so sub-warp groups put 2 (3D) / 4 (2D) attempts into one warp's registers.
#ifndef VOR_HOT_G3
#define VOR_HOT_G3 32
#endif
#ifndef VOR_HOT_G2
#define VOR_HOT_G2 16
#endif
#ifndef VOR_SK2
#define VOR_SK2 16                // 2D: killed simplices staged in shared memory (mean cavity: 4)
#endif
template <int D> struct HotCfg {
    static constexpr int G = D == 3 ? VOR_HOT_G3 : VOR_HOT_G2;
    static constexpr int SK = D == 3 ? VOR_SK : (VOR_HOT_G2 == 32 ? VOR_SK : VOR_SK2);
};
__device__ __forceinline__ double pow_mid(const OwnBlk &b, const RelPt &q) {
    const double dx = q.x - (double)b.cx, dy = q.y - (double)b.cy, dz = q.z - (double)b.cz;
    // a simplex without a filter (rout2 = inf) never attracts the walk
    return (dx * dx + dy * dy + dz * dz) - 0.5 * ((double)b.rin2 + (double)b.rout2);
}
// Middle stage of the hot kernel's conflict test, for the tests the cached sphere leaves undecided (~1e-5 of them on uniform
// input, a third on the jittered lattice): the determinant test with the FP64 static filter, WITHOUT the exact arithmetic
// (conflict_slow<D, false>, out of line: the hot loop pays no registers for it; the point is re-read, it is not kept in registers).
// +1 in conflict, -1 not, 0 = the FP64 filter failed too (the exact twin's business).
// MID is a template parameter of the hot kernel: the call site costs the determinant-free kernel its spill-free register
// allocation (144 B of stack, attempt 51.4 -> 58.2 ms per 10M uniform points), so the engine starts with MID = 0 and switches to the
// MID = 1 twin only for input that keeps leaving the sphere filter (jittered lattice: attempt 82 -> 69 ms per 5M points).
template <int D, int MID> __device__ __forceinline__ int hot_mid_test(const Mesh<D> &m, int n, int v, int stats) {
    if constexpr (MID != 0) {
    const typename Dim<D>::Pt pp = m.pts[v];
    if (stats) atomicAdd(&m.cnt->sph_undecided, 1ULL);   // one address: only when counters are asked for
    const int c = conflict_slow<D, false>(m.pts, m.tet, m.cnt, n, pp.x, pp.y, pt_z<D>(pp));
    return c == 2 ? 0 : (c ? 1 : -1);
    } else return 0;
}
template <int D, int G, int MID>
__device__ __forceinline__ void attempt_hot_one(const AttemptArgs<D> &A, const RoundSel &rsel, const int slot, int *const sk, int4 *const sn) {
    constexpr int M = Dim<D>::M;
    constexpr int SK = HotCfg<D>::SK;
    constexpr unsigned GFULL = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
    const Mesh<D> &m = A.m;
    const int gl = threadIdx.x & (G - 1);                          // lane inside the group
    const unsigned gmask = group_mask<G>();
    const int gshift = (threadIdx.x & 31) & ~(G - 1);              // first lane of the group inside the warp
    const int a = slot_entry(rsel, slot);
    if (a >= rsel.nact) { if (gl == 0) slot_lost(A.scr, slot); return; }
    const int v = A.act[a];
    int s = m.seed[v];
    if (s < 0) { if (gl == 0) slot_lost(A.scr, slot); return; }   // already inserted
    const int fl = A.slowFlag[v];
    if (fl != 0) {
        // the exact twin's point: queue its slot for this round unless the flag is of this very round
        if (gl == 0) {
            slot_lost(A.scr, slot);
            if (fl != A.keybase) A.scr.slowSlots[atomicAdd(&m.cnt->nslow, 1)] = slot;
        }
        return;
    }
    const RelPt rq = rel_pt(m, m.pts[v]);
    const uint32_t q = bij_hash((uint32_t)slot, A.bits, A.salt);   // unique among the slots of this round
    const int key_k = A.keybase | (int)(q << 1);
    const int key_o = key_k | 1;
    int nk = 0, nb = 0, big = -1;
    unsigned steps = 0, tests = 0;
    bool give = false;                     // hand the point to the exact twin
    bool lost = false;

    // -- forwarding, then power descent to a simplex whose sphere certainly contains p
    OwnBlk sb = load_blk(m, s);
    int4 stn = load_nbr(m.tet, s);
    while (sb.kill < 0) { s = ~sb.kill; sb = load_blk(m, s); stn = load_nbr(m.tet, s); }
    for (;;) {
        const int st = sphere_test(sb, rq);
        if (st > 0) break;
        if (MID != 0 && st == 0) {
            // p inside the shell of this simplex's filter (near-cospherical input: a third of all tests on the jittered lattice):
            // the FP64 determinant filter decides whether the flood may start here; the walk goes on if it says "outside"
            const int c = hot_mid_test<D, MID>(m, s, v, 0);
            if (c > 0) break;
            if (c == 0) { give = true; break; }
        }
        if (++steps > VOR_HOT_WALK) { give = true; break; }
        const double pw0 = pow_mid(sb, rq);
        double pw = INFINITY;
        OwnBlk nbk = sb;
        int4 nnn = stn;
        const int code = gl < M ? get4(stn, gl) : -1;
        if (code >= 0) {
            nbk = load_blk(m, code >> 2);
            nnn = load_nbr(m.tet, code >> 2);
            pw = pow_mid(nbk, rq);
        }
        // steepest descent: lane with the smallest power among the M neighbours
        double best = pw;
        int who = gl;
#pragma unroll
        for (int d = 1; d < 4; d <<= 1) {
            const double ob = __shfl_xor_sync(gmask, best, d, G);
            const int ow = __shfl_xor_sync(gmask, who, d, G);
            if (ob < best || (ob == best && ow < who)) { best = ob; who = ow; }
        }
        best = __shfl_sync(gmask, best, 0, G);
        who = __shfl_sync(gmask, who, 0, G);
        if (!(best < pw0)) { give = true; break; }          // stuck: undecidable from the filters
        s = __shfl_sync(gmask, code, who, G) >> 2;
        sb.kill = __shfl_sync(gmask, nbk.kill, who, G); sb.ring = __shfl_sync(gmask, nbk.ring, who, G);
        sb.cx = __shfl_sync(gmask, nbk.cx, who, G); sb.cy = __shfl_sync(gmask, nbk.cy, who, G); sb.cz = __shfl_sync(gmask, nbk.cz, who, G);
        sb.rin2 = __shfl_sync(gmask, nbk.rin2, who, G); sb.rout2 = __shfl_sync(gmask, nbk.rout2, who, G);
        stn.x = __shfl_sync(gmask, nnn.x, who, G); stn.y = __shfl_sync(gmask, nnn.y, who, G);
        stn.z = __shfl_sync(gmask, nnn.z, who, G); stn.w = __shfl_sync(gmask, nnn.w, who, G);
    }
    if (!give) {
        if (gl == 0 && steps) m.seed[v] = s;
        // reservation of the first simplex: early out on the block in registers; what slips in later is caught by commit
        if (sb.kill < key_k || sb.ring < key_k) lost = true;
        else if (gl == 0) atomicMin(&OWK(m, s), key_k);
    }
    if (!give && !lost) {
        tests = gl == 0 ? 1u : 0u;
        ScrView sv = scr_view(A.scr, slot, -1);
        if (gl == 0) { sv.k[0] = s; sk[0] = s; sn[0] = stn; }
        __syncwarp(gmask);
        nk = 1;
        int head = 0;
        while (head < nk && !lost && !give) {
            const int tail = nk;
            const int items = (tail - head) * M;
            for (int base = 0; base < items; base += G) {
                const int j = base + gl;
                bool pushB = false, lostLane = false, giveLane = false;
                int fcode = 0, ocode = 0;
                int claim = -1 - gl;       // simplex this lane wants to kill (unique dummy otherwise)
                int4 nnb = make_int4(-1, -1, -1, -1);
                if (j < items) {
                    const int e = head + j / M;
                    const int i = j % M;
                    int t, code;
                    if (e < SK) { t = sk[e]; code = reinterpret_cast<const int *>(sn)[e * 4 + i]; }   // the cavity so far is staged in shared memory
                    else { t = sv.k[e]; code = TNI(m, t, i); }
                    if (code < 0) {
                        pushB = true; fcode = t * 4 + i; ocode = code;
                    } else {
                        const int n = code >> 2;
                        const OwnBlk blk = load_blk(m, n);      // ownership words + sphere: ONE 256-bit gather
                        if (VOR_SPEC) nnb = load_nbr(m.tet, n);
                        if (blk.kill == key_k) {
                            // already in my cavity
                        } else if (blk.kill < key_k) {
                            lostLane = true;           // a better point kills n (or n is dead)
                        } else if (blk.ring == key_o) {
                            pushB = true; fcode = t * 4 + i; ocode = code;   // already tested by me: not in conflict
                        } else {
                            tests++;
                            int conf = sphere_test(blk, rq);
                            if (conf == 0) conf = hot_mid_test<D, MID>(m, n, v, A.stats);   // inside the filter's shell: FP64 determinant filter (out of line)
                            if (conf > 0) {
                                if (!VOR_SPEC) nnb = load_nbr(m.tet, n);
                                if (blk.ring < key_k) lostLane = true;   // a better point keeps n in its outer ring
                                else { atomicMin(&OWK(m, n), key_k); claim = n; }
                            } else if (conf < 0) {
                                atomicMin(&OWR(m, n), key_o);            // outer-ring mark: fire and forget
                                pushB = true; fcode = t * 4 + i; ocode = code;
                            } else giveLane = true;                      // the FP64 filter fails too: the exact twin decides
                        }
                    }
                }
                if (__any_sync(gmask, giveLane)) { give = true; break; }
                if (__any_sync(gmask, lostLane)) { lost = true; break; }
                const unsigned same = __match_any_sync(gmask, claim);
                const bool pushK = claim >= 0 && (__ffs(same) - 1) == (int)(threadIdx.x & 31);
                const unsigned mk = (__ballot_sync(gmask, pushK) >> gshift) & GFULL;
                const unsigned mb = (__ballot_sync(gmask, pushB) >> gshift) & GFULL;
                const int ck = __popc(mk), cb = __popc(mb);
                if (nk + ck > sv.capk || nb + cb > sv.capb) {
                    // spill to an overflow slot (contiguous, much larger)
                    if (big >= 0) { if (gl == 0) set_err(m.cnt, ERR_CAPACITY); lost = true; break; }
                    if (gl == 0) big = atomicAdd(&m.cnt->nbig, 1);
                    big = __shfl_sync(gmask, big, 0, G);
                    if (big >= A.scr.nbig) { lost = true; break; }
                    const ScrView bv = scr_view(A.scr, slot, big);
                    for (int x = gl; x < nk; x += G) bv.k[x] = sv.k[x];
                    for (int x = gl; x < nb; x += G) { bv.f[x] = sv.f[x]; bv.o[x] = sv.o[x]; }
                    sv = bv;
                    __syncwarp(gmask);
                    if (nk + ck > sv.capk || nb + cb > sv.capb) { if (gl == 0) set_err(m.cnt, ERR_CAPACITY); lost = true; break; }
                }
                const unsigned lt = (1u << gl) - 1u;
                if (pushK) {
                    const int pos = nk + __popc(mk & lt);
                    sv.k[pos] = claim;
                    if (pos < SK) { sk[pos] = claim; sn[pos] = nnb; }
                }
                if (pushB) { const int pos = nb + __popc(mb & lt); sv.f[pos] = fcode; sv.o[pos] = ocode; }
                nk += ck;
                nb += cb;
                __syncwarp(gmask);
            }
            head = tail;
        }
    }
    const int status = (give || lost) ? ST_LOST : ST_OK;
    if (gl == 0) {
        if (give) {
            // hand the point to the exact twin (from the next round on); marks left under this round's key are stale then
            A.slowFlag[v] = A.keybase;
            atomicAdd(&m.cnt->nflag_set, 1);
        }
        A.scr.slotInfo[slot] = slot_pack(status, false, big, nk, nb, v);
    }
    if (A.stats) {
        for (int d = G / 2; d > 0; d >>= 1) tests += __shfl_xor_sync(gmask, tests, d, G);
        if (gl == 0) {
            atomicAdd(&m.cnt->walk_steps, (unsigned long long)steps);
            atomicAdd(&m.cnt->tests, (unsigned long long)tests);
            atomicAdd(&m.cnt->attempts, 1ULL);
            if (status != ST_OK) atomicAdd(&m.cnt->aborted, 1ULL);
            else atomicAdd(&m.cnt->tests_ok, (unsigned long long)tests);
        }
    }
}

// Resident warps with a static stride over the slots (grid = what fits on the machine, engine.cuh): a third to a half of
// the slots of a round hold points that are already inserted or are the exact twin's; with one block per pair of slots
// the SMs spent their time launching blocks that exit at once (ncu, round 2: 24 of 36 warps resident on average).
#ifndef VOR_MID_REGS
#define VOR_MID_REGS 56           // register budget of the MID = 1 twin
#endif
template <int MID> struct HotLaunch {
    static constexpr int regs = MID ? VOR_MID_REGS : VOR_HOT_REGS;
    static constexpr int minBlocks = (65536 / (regs * VOR_HOT_BLOCK)) > 32 ? 32 : (65536 / (regs * VOR_HOT_BLOCK));
};
template <int D, int G, int MID>
__global__ void __launch_bounds__(VOR_HOT_BLOCK, HotLaunch<MID>::minBlocks)
k_attempt_hot(AttemptArgs<D> A, RoundSel rsel) {
    __shared__ int s_kid[VOR_HOT_BLOCK / G][HotCfg<D>::SK];
    __shared__ int4 s_knb[VOR_HOT_BLOCK / G][HotCfg<D>::SK];
    pdl_trigger();
    pdl_wait();
    if (blockIdx.x == 0 && threadIdx.x == 0) A.m.cnt->sph_lo = A.m.cnt->ntets;   // k_spheres of the last round is done
    const unsigned gmask = group_mask<G>();
    const int ngroups = (gridDim.x * blockDim.x) / G;
    for (int slot = rsel.first + (blockIdx.x * blockDim.x + threadIdx.x) / G; slot < rsel.last; slot += ngroups) {
        attempt_hot_one<D, G, MID>(A, rsel, slot, s_kid[threadIdx.x / G], s_knb[threadIdx.x / G]);
        __syncwarp(gmask);
    }
}

// Tiled form of the same kernel (option "tiled"; a recorded NEGATIVE result, off by default: 10M points 138.6 ms against 115.3 ms,
// attempt 60.6 vs 52.9 ms, commit 56.7 vs 40.5 ms).  Resident blocks of 4 warps pull TILES of `tile` consecutive
// slots from a device-side queue (one atomicAdd per tile); the threads of the block first classify the tile's slots in
// parallel (entry out of range / point already inserted / exact twin's point -> status written at once), the live slots
// are compacted into shared memory, and the warps take live slots from that list until it is empty.  A third to a half of
// the slots of a round are dead: with one block per pair of slots the SMs kept launching blocks that exit at once (24 of
// 36 resident warps busy on average, ncu round 2); here every resident warp works on a live attempt, the load is balanced
// inside the block by the shared list and across blocks by the tile queue.  `tile` shrinks with the round (engine.cuh)
// so that a small round still spreads one slot per warp over the whole machine.
// Why it loses (as the per-SM queues of round 1 and the strided resident warps did): with one block per pair of slots the
// hardware scheduler hands the slots out IN ORDER, so the ~5,000 attempts in flight at any moment are one contiguous run
// of the Morton-ordered list -- neighbours in space, whose conflict regions share simplices, 64 B lines and DRAM pages
// at (nearly) the same time.  Any scheme that gives a block or a warp a private range spreads the attempts in flight over
// the whole round and loses that sharing (L2 hit rate of the attempt kernel: 41 % with the in-order window).
#define VOR_TILE_BLOCK 128
template <int D>
__global__ void __launch_bounds__(VOR_TILE_BLOCK, (65536 / (VOR_HOT_REGS * VOR_TILE_BLOCK)) > 16 ? 16 : (65536 / (VOR_HOT_REGS * VOR_TILE_BLOCK)))
k_attempt_hot_tiled(AttemptArgs<D> A, RoundSel rsel, int tile) {
    constexpr int W = VOR_TILE_BLOCK / 32;
    __shared__ int s_kid[W][VOR_SK];
    __shared__ int4 s_knb[W][VOR_SK];
    __shared__ int s_list[VOR_TILE_BLOCK];
    __shared__ int s_n, s_next, s_tile;
    const Mesh<D> &m = A.m;
    const int w = threadIdx.x >> 5, gl = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) { m.cnt->sph_lo = m.cnt->ntets; m.cnt->q_commit = 0; }
    const int ntiles = (rsel.nsel + tile - 1) / tile;
    for (;;) {
        if (threadIdx.x == 0) { s_tile = atomicAdd(&m.cnt->q_attempt, 1); s_n = 0; s_next = 0; }
        __syncthreads();
        const int t = s_tile;
        if (t >= ntiles) break;
        // classify the slots of the tile, one thread each
        if ((int)threadIdx.x < tile) {
            const int slot = t * tile + (int)threadIdx.x;
            if (slot < rsel.nsel) {
                const int a = slot_entry(rsel, slot);
                bool live = false;
                if (a < rsel.nact) {
                    const int v = A.act[a];
                    if (m.seed[v] >= 0) {
                        const int fl = A.slowFlag[v];
                        if (fl == 0) live = true;
                        else if (fl != A.keybase) A.scr.slowSlots[atomicAdd(&m.cnt->nslow, 1)] = slot;   // the exact twin's point
                    }
                }
                if (live) s_list[atomicAdd(&s_n, 1)] = slot;
                else slot_lost(A.scr, slot);
            }
        }
        __syncthreads();
        const int n = s_n;
        for (;;) {
            int i = 0;
            if (gl == 0) i = atomicAdd(&s_next, 1);
            i = __shfl_sync(0xffffffffu, i, 0);
            if (i >= n) break;
            attempt_hot_one<D, 32, 0>(A, rsel, s_list[i], s_kid[w], s_knb[w]);
            __syncwarp();
        }
        __syncthreads();
    }
}

// the exact twin behind the hot kernel: the slots the hot kernel queued this round (points flagged in earlier rounds)
template <int D, int RED>
__global__ void __launch_bounds__(VOR_ATTEMPT_BLOCK, 65536 / (VOR_ATTEMPT_REGS * VOR_ATTEMPT_BLOCK)) k_attempt_slow(AttemptArgs<D> A, RoundSel rsel) {
    __shared__ int s_kid[VOR_ATTEMPT_BLOCK / 32][VOR_SK];
    __shared__ int4 s_knb[VOR_ATTEMPT_BLOCK / 32][VOR_SK];
    pdl_trigger();
    pdl_wait();
    const int n = min(A.m.cnt->nslow, A.scr.nslots);
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; idx < n; idx += nwarps) {
        attempt_one<D, 32, RED, 1>(A, rsel, A.scr.slowSlots[idx], s_kid[threadIdx.x >> 5], s_knb[threadIdx.x >> 5]);
        __syncwarp();
    }
}

// One group per attempt slot (grid = slots).  The hot twin needs fewer registers than the kernel with the exact path.
template <int EXACT> struct AttemptLaunch {
    static constexpr int block = VOR_ATTEMPT_BLOCK;
    static constexpr int regs = VOR_ATTEMPT_REGS;
    static constexpr int minBlocks = (65536 / (regs * block)) > 32 ? 32 : (65536 / (regs * block));
};
template <int D, int G, int RED, int EXACT>
__global__ void __launch_bounds__(AttemptLaunch<EXACT>::block, AttemptLaunch<EXACT>::minBlocks) k_attempt_coop(AttemptArgs<D> A, RoundSel rsel) {
    // the cavity found so far, staged in shared memory: ids and neighbour codes of the first SK killed simplices, so that a
    // flood level starts from shared memory instead of two dependent round trips (scratch, then record)
    __shared__ int s_kid[AttemptLaunch<EXACT>::block / G][VOR_SK];
    __shared__ int4 s_knb[AttemptLaunch<EXACT>::block / G][VOR_SK];
    const int gid = rsel.first + (blockIdx.x * blockDim.x + threadIdx.x) / G;   // group = attempt slot
    if (blockIdx.x == 0 && threadIdx.x == 0) A.m.cnt->sph_lo = A.m.cnt->ntets;   // k_spheres of the last round is done
    if (gid >= rsel.last) return;
    attempt_one<D, G, RED, EXACT>(A, rsel, gid, s_kid[threadIdx.x / G], s_knb[threadIdx.x / G]);
}

// ------------------------------------------------------------------------------------------
// commit = ownership check + allocation + retriangulation, fused
// ------------------------------------------------------------------------------------------
// A group whose point still owns its whole footprint is a winner: it takes a block of simplex slots from the bump
// allocator (one atomicAdd) and retriangulates at once.  This is safe while other groups are still checking: a check
// only reads owner[] of its own footprint, and a winner only changes owner[] on its own killed simplices, which any
// group that shares them has lost anyway (it reads the winner's key or the dead mark, never its own key).
// Retriangulation of one cavity through the global store (any cavity size; used for the rare cavity that does not
// fit the shared-memory staging of k_commit_coop): markers and pivots go through the dead simplices' records in HBM.
template <int D, int G>
__device__ __noinline__ void commit_global(const Mesh<D> &m, const ScrView sv, int nk, int nb, int v, int base, int gl, unsigned gmask) {
    constexpr int M = Dim<D>::M;
    // phase A: one lane per boundary facet: new simplex, outer back-pointer, marker in the dead simplex
    for (int j = gl; j < nb; j += G) {
        const int fc = sv.f[j];
        const int t = fc >> 2, i = fc & 3;
        const int outer = sv.o[j];
        const int T = base + j;
        int4 verts = TV(m, t);
        set4(verts, i, v);
        TV(m, T) = verts;
        TNI(m, T, i) = outer;
        if (M == 3) TNI(m, T, 3) = -1;
        if (outer >= 0) TNI(m, outer >> 2, outer & 3) = T * 4 + i;
        TNI(m, t, i) = -(T * 4 + i) - 2;
        OWK(m, t) = ~T;     // dead; forwards to a new simplex that shares a facet with it (any of them: benign race)
    }
    __syncwarp(gmask);
    // phase B: one lane per (new simplex, facet containing v): pivot around the ridge through the dead cavity
    const int nitems = nb * (M - 1);
    for (int it = gl; it < nitems; it += G) {
        const int j = it / (M - 1);
        const int fc = sv.f[j];
        const int t = fc >> 2, i = fc & 3;
        int k = it % (M - 1);
        if (k >= i) k++;
        const int T = base + j;
        int4 cv = TV(m, t);
        int r0 = -1, r1 = -1;
        for (int sidx = 0; sidx < M; sidx++) {
            if (sidx == i || sidx == k) continue;
            if (r0 < 0) r0 = get4(cv, sidx); else r1 = get4(cv, sidx);
        }
        int cur = t, enter = i, exitf = k;
        for (;;) {
            const int e = TNI(m, cur, exitf);   // plain load: markers were written by this group (phase A)
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
    // phase C: the cavity dies (forwarding to the first new simplex)
    for (int j = gl; j < nk; j += G) {
        const int t = sv.k[j];
        if (__ldcg(&OWK(m, t)) >= 0) OWK(m, t) = ~base;   // interior of the cavity (no boundary facet)
    }
}

#ifndef VOR_CK
#define VOR_CK 48                 // commit: killed simplices of a cavity staged in shared memory (mean 20 in 3D, 4 in 2D)
#endif
#ifndef VOR_CB
#define VOR_CB 100                // commit: boundary facets staged in shared memory (mean 27 / 6)
#endif
// The whole cavity (records of the killed simplices, boundary facets, a small id -> local index hash) is pulled into
// shared memory with ONE level of independent gathers, checked, and retriangulated there: markers, the ridge pivots
// of pair_simplices (delaunay_tree.rs:674-695) and the forwarding choice never leave the SM; HBM sees one full 32 B
// record per new simplex, one back-pointer per outer facet and one dead mark per killed simplex.
constexpr int COMMIT_HS = 128;
#ifndef VOR_RIDGE_HASH
#define VOR_RIDGE_HASH 0          // 1: sibling links through a hash of the cavity's boundary ridges in shared memory instead of pivoting
                                  // around every ridge through the staged cavity.  Parity-green, fewer instructions, but 64 instead of
                                  // 40 registers and 64-bit shared-memory CAS: commit 43.2 vs 40.3 ms per 10M points (11.5 vs 12.0 per 1M)
#endif
constexpr int RIDGE_HS = 256;     // >= 1.5 x VOR_CB ridges, power of two
// Lane groups of the commit kernel (as for the hot kernel): a 2D cavity has ~4 killed simplices and ~6 boundary edges
#ifndef VOR_COMMIT_G3
#define VOR_COMMIT_G3 32
#endif
#ifndef VOR_COMMIT_G2
#define VOR_COMMIT_G2 16
#endif
#ifndef VOR_CK2
#define VOR_CK2 16
#endif
#ifndef VOR_CB2
#define VOR_CB2 36
#endif
template <int D> struct CommitCfg {
    static constexpr int G = D == 3 ? VOR_COMMIT_G3 : VOR_COMMIT_G2;
    static constexpr bool SMALL = D == 2 && VOR_COMMIT_G2 != 32;     // sub-warp groups in 2D: staging sized for 2D cavities
    static constexpr int CK = SMALL ? VOR_CK2 : VOR_CK;
    static constexpr int CB = SMALL ? VOR_CB2 : VOR_CB;
    static constexpr int HS = SMALL ? 32 : COMMIT_HS;                // id -> local index hash, at most half full
    static constexpr int HSHIFT = SMALL ? 27 : 25;                   // 32 - log2(HS)
};
template <int D> struct CommitSmem {
    int4 tv[CommitCfg<D>::CK];
#if VOR_RIDGE_HASH
    unsigned long long rkey[RIDGE_HS];
    int rval[RIDGE_HS][2];
#else
    int4 tn[CommitCfg<D>::CK];
#endif
    int id[CommitCfg<D>::CK], fw[CommitCfg<D>::CK], hash[CommitCfg<D>::HS], f[CommitCfg<D>::CB], o[CommitCfg<D>::CB];
};
template <int D, int G>
__device__ __forceinline__ void commit_one(const CheckArgs<D> &A, const int *act, const RoundSel &rsel, int stats, const int slot, CommitSmem<D> &sm) {
    constexpr int M = Dim<D>::M;
    constexpr int CK = CommitCfg<D>::CK, CB = CommitCfg<D>::CB, HS = CommitCfg<D>::HS, HSHIFT = CommitCfg<D>::HSHIFT;
    static_assert(HS >= 2 * CK && (1 << (32 - HSHIFT)) == HS, "hash must stay at most half full");
    const Mesh<D> &m = A.m;
    const int gid = slot;
    const int gl = threadIdx.x & (G - 1);
    const unsigned gmask = group_mask<G>();
    const int gshift = (threadIdx.x & 31) & ~(G - 1);
    const int4 info = A.scr.slotInfo[slot];                        // status, cavity sizes, overflow slot and vertex in one load
    if ((info.x & 3) != ST_OK) return;
    const int v = info.w;
    const uint32_t q = bij_hash((uint32_t)slot, A.bits, A.salt);   // unique among the slots of this round
    const int key_k = A.keybase | (int)(q << 1);
    const ScrView sv = scr_view(A.scr, slot, (info.x >> 3) - 1);
    const int nk = info.y, nb = info.z;
    const bool fast = !(stats & 2) && nk <= CK && nb <= CB;   // stats bit 1: force the global-store path (A/B switch)
    int4 *const tvs = sm.tv;
    int *const ids = sm.id, *const fw = sm.fw, *const hash = sm.hash, *const sf = sm.f, *const so = sm.o;
#if !VOR_RIDGE_HASH
    int4 *const tns = sm.tn;
    int *const tni = reinterpret_cast<int *>(tns);
#endif

    // -- ownership check; the fast path loads the cavity in the same level of gathers
    bool bad = false;
    if (fast) {
        for (int e = gl; e < nk; e += G) {
            const int t = sv.k[e];
            const int2 ow = __ldcg(reinterpret_cast<const int2 *>(&OWK(m, t)));
            int4 tv, tn;
            load_rec_cg(m, t, tv, tn);
            if (ow.x != key_k || ow.y < key_k) bad = true;     // best killer, and no better point has it in its ring
            ids[e] = t; tvs[e] = tv; fw[e] = 0;
#if !VOR_RIDGE_HASH
            tns[e] = tn;
#endif
        }
        for (int j = gl; j < nb; j += G) {
            const int f = sv.f[j], code = sv.o[j];
            if (code >= 0 && __ldcg(&OWK(m, code >> 2)) < key_k) bad = true;   // no better point kills my outer ring
            sf[j] = f; so[j] = code;
        }
        for (int h = gl; h < HS; h += G) hash[h] = -1;
#if VOR_RIDGE_HASH
        for (int h = gl; h < RIDGE_HS; h += G) { sm.rkey[h] = ~0ULL; sm.rval[h][0] = -1; }
#endif
    } else {
        for (int j = gl; j < nk; j += G) {
            const int2 ow = __ldcg(reinterpret_cast<const int2 *>(&OWK(m, sv.k[j])));
            if (ow.x != key_k || ow.y < key_k) bad = true;
        }
        for (int j = gl; j < nb; j += G) {
            const int code = sv.o[j];
            if (code >= 0 && __ldcg(&OWK(m, code >> 2)) < key_k) bad = true;
        }
    }
    if (__any_sync(gmask, bad)) return;
    const int nfresh = nb;
    int base = 0;
    if (gl == 0) base = atomicAdd(&m.cnt->ntets, nfresh);
    base = __shfl_sync(gmask, base, gshift);
    if (base + nfresh > m.cap) {
        // no room: leave the mesh untouched (the point stays pending), retire the part of the block that exists and
        // tell the host to grow the store
        for (int j = gl; j < nfresh; j += G)
            if (base + j < m.cap) store_rec(m, base + j, make_int4(-1, -1, -1, -1), make_int4(-1, -1, -1, -1));   // k_spheres marks it dead
        if (gl == 0) m.cnt->oom_soft = 1;
        return;
    }
    int first = base;     // a simplex created by this insertion (seed for later points)
    if (!fast) {
        commit_global<D, G>(m, sv, nk, nb, v, base, gl, gmask);
    } else {
        auto slot_of = [&](int j) -> int { return base + j; };
        // id -> local index (open addressing, at most 3/8 full)
        for (int e = gl; e < nk; e += G) {
            unsigned h = ((unsigned)ids[e] * 2654435761u) >> HSHIFT;
            while (atomicCAS(&hash[h], -1, e) != -1) h = (h + 1) & (HS - 1);
        }
        __syncwarp(gmask);
        auto local_of = [&](int t) -> int {
            unsigned h = ((unsigned)t * 2654435761u) >> HSHIFT;
            for (int probe = 0; probe < HS; probe++) {
                const int e = hash[h];
                if (e < 0) break;
                if (ids[e] == t) return e;
                h = (h + 1) & (HS - 1);
            }
            set_err(m.cnt, ERR_CUDA);   // a pivot left the cavity: cannot happen on a consistent mesh
            return -1;
        };
#if VOR_RIDGE_HASH
        // Sibling links (pair_simplices, delaunay_tree.rs:674-695, O(k^2 M^2) there) through a hash of the boundary RIDGES: the
        // boundary of a cavity is a closed surface (3D: every edge of it lies in exactly two boundary triangles; 2D: every vertex
        // in two boundary edges), and the two new simplices on those two facets are each other's neighbours across the
        // facet that contains the new point and the ridge.  One lane per new simplex: insert its M-1 ridges (key = the
        // ridge's vertex ids, value = new simplex and the slot opposite the shared facet), then read the partner of each.
        int myh[M], e0s = -1, is = 0;
        int4 verts = make_int4(-1, -1, -1, -1);
        for (int j0 = 0; j0 < nb; j0 += G) {          // nb <= VOR_CB: at most ceil(CB / G) passes, state of ONE facet per lane and pass
            const int j = j0 + gl;
            if (j < nb) {
                const int f = sf[j];
                const int e0 = local_of(f >> 2), i = f & 3;
                e0s = e0; is = i;
                if (e0 >= 0) {
                    fw[e0] = j;                // any of its boundary facets (benign race)
                    const int4 cv0 = tvs[e0];
                    verts = cv0;
                    set4(verts, i, v);
#pragma unroll
                    for (int k = 0; k < M; k++) {
                        myh[k] = -1;
                        if (k == i) continue;
                        unsigned r0 = 0xffffffffu, r1 = 0xffffffffu;
#pragma unroll
                        for (int sidx = 0; sidx < M; sidx++) {
                            if (sidx == i || sidx == k) continue;
                            if (r0 == 0xffffffffu) r0 = (unsigned)get4(cv0, sidx); else r1 = (unsigned)get4(cv0, sidx);
                        }
                        const unsigned lo = r0 < r1 ? r0 : r1, hi = r0 < r1 ? r1 : r0;     // 2D: hi stays 0xffffffff
                        const unsigned long long key = ((unsigned long long)lo << 32) | hi;
                        unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ULL) >> 56) & (RIDGE_HS - 1);
                        for (int probe = 0; probe < RIDGE_HS; probe++) {
                            const unsigned long long old = atomicCAS(&sm.rkey[h], ~0ULL, key);
                            if (old == ~0ULL || old == key) break;
                            h = (h + 1) & (RIDGE_HS - 1);
                        }
                        myh[k] = (int)h;
                        if (atomicCAS(&sm.rval[h][0], -1, j * 4 + k) != -1) sm.rval[h][1] = j * 4 + k;
                    }
                }
            }
            __syncwarp(gmask);
            if (nb <= G) break;                        // the common case: one pass, links resolved below from registers
            // more facets than lanes: resolve this pass's links after ALL passes have inserted (second loop below)
        }
        if (nb <= G) {
            const int j = gl;
            if (j < nb && e0s >= 0) {
                const int i = is, outer = so[j], T = slot_of(j);
                int4 nbr = make_int4(-1, -1, -1, -1);
                set4(nbr, i, outer);
#pragma unroll
                for (int k = 0; k < M; k++) {
                    if (k == i) continue;
                    const int a = sm.rval[myh[k]][0], b = sm.rval[myh[k]][1];
                    const int partner = a == j * 4 + k ? b : a;
                    set4(nbr, k, slot_of(partner >> 2) * 4 + (partner & 3));
                }
                store_rec(m, T, verts, nbr);   // its ownership + sphere block is written by k_spheres, next in the stream
                if (outer >= 0) TNI(m, outer >> 2, outer & 3) = T * 4 + i;
            }
        } else {
            for (int j = gl; j < nb; j += G) {
                const int f = sf[j];
                const int e0 = local_of(f >> 2), i = f & 3;
                if (e0 < 0) continue;
                const int outer = so[j], T = slot_of(j);
                const int4 cv0 = tvs[e0];
                int4 vv = cv0;
                set4(vv, i, v);
                int4 nbr = make_int4(-1, -1, -1, -1);
                set4(nbr, i, outer);
#pragma unroll
                for (int k = 0; k < M; k++) {
                    if (k == i) continue;
                    unsigned r0 = 0xffffffffu, r1 = 0xffffffffu;
#pragma unroll
                    for (int sidx = 0; sidx < M; sidx++) {
                        if (sidx == i || sidx == k) continue;
                        if (r0 == 0xffffffffu) r0 = (unsigned)get4(cv0, sidx); else r1 = (unsigned)get4(cv0, sidx);
                    }
                    const unsigned lo = r0 < r1 ? r0 : r1, hi = r0 < r1 ? r1 : r0;
                    const unsigned long long key = ((unsigned long long)lo << 32) | hi;
                    unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ULL) >> 56) & (RIDGE_HS - 1);
                    for (int probe = 0; probe < RIDGE_HS && sm.rkey[h] != key; probe++) h = (h + 1) & (RIDGE_HS - 1);
                    const int a = sm.rval[h][0], b = sm.rval[h][1];
                    const int partner = a == j * 4 + k ? b : a;
                    set4(nbr, k, slot_of(partner >> 2) * 4 + (partner & 3));
                }
                store_rec(m, T, vv, nbr);
                if (outer >= 0) TNI(m, outer >> 2, outer & 3) = T * 4 + i;
            }
        }
        __syncwarp(gmask);
#else
        // phase A: markers on the boundary facets of the staged cavity, forwarding choice per killed simplex
        for (int j = gl; j < nb; j += G) {
            const int f = sf[j];
            const int e = local_of(f >> 2), i = f & 3;
            if (e < 0) continue;
            tni[e * 4 + i] = -j - 2;
            fw[e] = j;                 // any of its boundary facets (benign race)
            sf[j] = e * 4 + i;
        }
        __syncwarp(gmask);
        // every facet of the staged cavity is now a marker (boundary) or leads to another killed simplex: translate those
        // neighbour codes to LOCAL indices once, all lanes in parallel, so that the pivots below are plain shared-memory
        // index chasing without a hash probe per step
        for (int idx = gl; idx < nk * 4; idx += G) {
            const int code = tni[idx];
            if (code >= 0) {
                const int ln = local_of(code >> 2);
                tni[idx] = ln < 0 ? -1 : ln * 4 + (code & 3);       // -1: inconsistent mesh (error already set)
            }
        }
        __syncwarp(gmask);
        // phase B: one lane per new simplex: vertices, outer neighbour, and the M-1 siblings found by pivoting around
        // each ridge through the staged cavity
        for (int j = gl; j < nb; j += G) {
            const int f = sf[j];
            const int e0 = f >> 2, i = f & 3;
            const int outer = so[j];
            const int T = slot_of(j);
            const int4 cv0 = tvs[e0];
            int4 verts = cv0;
            set4(verts, i, v);
            int4 nbr = make_int4(-1, -1, -1, -1);
            set4(nbr, i, outer);
#pragma unroll
            for (int k = 0; k < M; k++) {
                if (k == i) continue;
                int r0 = -1, r1 = -1;
#pragma unroll
                for (int sidx = 0; sidx < M; sidx++) {
                    if (sidx == i || sidx == k) continue;
                    if (r0 < 0) r0 = get4(cv0, sidx); else r1 = get4(cv0, sidx);
                }
                int cur = e0, enter = i, exitf = k;
                for (int guard = 0;; guard++) {
                    const int code = tni[cur * 4 + exitf];
                    if (code <= -2) { set4(nbr, k, slot_of(-code - 2) * 4 + enter); break; }
                    if (guard >= 4 * CK) { set_err(m.cnt, ERR_CUDA); break; }
                    if (code < 0) break;
                    const int ln = code >> 2, jb = code & 3;
                    const int4 cv = tvs[ln];
                    int y = -1;
#pragma unroll
                    for (int sidx = 0; sidx < M; sidx++) {
                        if (sidx == jb) continue;
                        const int vv = get4(cv, sidx);
                        if (vv != r0 && vv != r1) y = sidx;
                    }
                    cur = ln; enter = jb; exitf = y;
                }
            }
            store_rec(m, T, verts, nbr);   // its ownership + sphere block is written by k_spheres, next in the stream
            if (outer >= 0) TNI(m, outer >> 2, outer & 3) = T * 4 + i;
        }
#endif
        // phase C: the killed simplices die; a dead simplex forwards to a new simplex on one of its own boundary
        // facets (interior simplices: to the first new simplex)
        for (int e = gl; e < nk; e += G) OWK(m, ids[e]) = ~slot_of(fw[e]);
    }
    if (gl == 0) {
        m.ptTet[v] = first;
        m.seed[v] = -1;
        atomicAdd(&m.cnt->part[gid & (NPART - 1)][0], (1ULL << 40) | (unsigned long long)nb);   // win_total, created_all
        if (info.x & 4) atomicAdd(&m.cnt->nflag_done, 1);     // a point of the exact twin is done
        if (stats & 1) {
            atomicAdd(&m.cnt->killed, (unsigned long long)nk);
            atomicAdd(&m.cnt->created, (unsigned long long)nb);
        }
    }
}
// resident groups with a static stride over the slots (see k_attempt_hot): most slots of a round hold no winner
template <int D, int G>
__global__ void __launch_bounds__(VOR_COOP_BLOCK) k_commit_coop(CheckArgs<D> A, const int *act, RoundSel rsel, int stats) {
    __shared__ CommitSmem<D> s_cav[VOR_COOP_BLOCK / G];
    const unsigned gmask = group_mask<G>();
    pdl_trigger();
    pdl_wait();
    if (blockIdx.x == 0 && threadIdx.x == 0) { A.m.cnt->nbig = 0; A.m.cnt->nslow = 0; }   // overflow slots and the exact twin's queue are per round (attempt is over)
    const int ngroups = (gridDim.x * blockDim.x) / G;
    for (int slot = rsel.first + (blockIdx.x * blockDim.x + threadIdx.x) / G; slot < rsel.last; slot += ngroups) {
        commit_one<D, G>(A, act, rsel, stats, slot, s_cav[threadIdx.x / G]);
        __syncwarp(gmask);
    }
}

// tiled form (see k_attempt_hot_tiled): most slots of a round hold no winner candidate
template <int D>
__global__ void __launch_bounds__(VOR_TILE_BLOCK) k_commit_tiled(CheckArgs<D> A, const int *act, RoundSel rsel, int stats, int tile) {
    constexpr int W = VOR_TILE_BLOCK / 32;
    __shared__ CommitSmem<D> s_cav[W];
    __shared__ int s_list[VOR_TILE_BLOCK];
    __shared__ int s_n, s_next, s_tile;
    const Mesh<D> &m = A.m;
    const int w = threadIdx.x >> 5, gl = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) { m.cnt->nbig = 0; m.cnt->nslow = 0; m.cnt->q_attempt = 0; }
    const int ntiles = (rsel.nsel + tile - 1) / tile;
    for (;;) {
        if (threadIdx.x == 0) { s_tile = atomicAdd(&m.cnt->q_commit, 1); s_n = 0; s_next = 0; }
        __syncthreads();
        const int t = s_tile;
        if (t >= ntiles) break;
        if ((int)threadIdx.x < tile) {
            const int slot = t * tile + (int)threadIdx.x;
            if (slot < rsel.nsel && (reinterpret_cast<const int *>(A.scr.slotInfo + slot)[0] & 3) == ST_OK) s_list[atomicAdd(&s_n, 1)] = slot;
        }
        __syncthreads();
        const int n = s_n;
        for (;;) {
            int i = 0;
            if (gl == 0) i = atomicAdd(&s_next, 1);
            i = __shfl_sync(0xffffffffu, i, 0);
            if (i >= n) break;
            commit_one<D, 32>(A, act, rsel, stats, s_list[i], s_cav[w]);
            __syncwarp();
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// spheres: ownership words (free) + certified circumsphere filter (sphere.cuh) of the simplices created by the commit
// kernel just before it in the stream -- the slots [cnt->sph_lo, cnt->ntets) of the bump allocator.  One thread per new
// simplex: its record and block are contiguous (coalesced), the new simplices of one cavity are neighbours in the store
// and share their vertices (L1 hits).  Kept out of the commit kernel: the ~150 FP64 operations of a sphere would cost the
// ridge pivots their registers (measured: commit 43 -> 132 ms per 10M points when fused).  cnt->sph_lo is advanced by the
// first thread of the next attempt kernel (the allocator does not move during an attempt kernel).
// ------------------------------------------------------------------------------------------
#ifndef VOR_SPH_MINB
#define VOR_SPH_MINB 3
#endif
template <int D>
__global__ void __launch_bounds__(256, VOR_SPH_MINB) k_spheres(Mesh<D> m) {
    pdl_trigger();
    pdl_wait();
    const int lo = m.cnt->sph_lo, hi = min(m.cnt->ntets, m.cap);
    for (int t = lo + blockIdx.x * blockDim.x + threadIdx.x; t < hi; t += gridDim.x * blockDim.x) {
        const int4 tv = TV(m, t);
        if (tv.x < 0) { OWK(m, t) = -1; continue; }              // slot retired by a winner that found no room: dead
        store_blk(m, t, sphere_of(m, tv));
    }
}

// ------------------------------------------------------------------------------------------
// order-preserving compaction of the active list (pending entries keep their Morton order): block counts,
// single-block exclusive scan of the counts, scatter with ballot ranks
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_compact_count(const int *act, const int *seed, int *blockCnt, int n) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int keep = (i < n) && (seed[act[i]] >= 0);
    const int c = __syncthreads_count(keep);
    if (threadIdx.x == 0) blockCnt[blockIdx.x] = c;
}
__global__ void __launch_bounds__(1024) k_compact_scan(int *blockCnt, int nb, long long *total) {
    __shared__ int part[1024];
    const int per = (nb + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(lo + per, nb);
    int s = 0;
    for (int i = lo; i < hi; i++) s += blockCnt[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {   // Hillis-Steele inclusive scan
        const int v = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;        // exclusive prefix of this thread's segment
    for (int i = lo; i < hi; i++) { const int x = blockCnt[i]; blockCnt[i] = run; run += x; }
    if (threadIdx.x == 1023) *total = part[1023];
}
__global__ void __launch_bounds__(256) k_compact_scatter(const int *act, const int *seed, const int *blockCnt, int *out, int n) {
    __shared__ int warpCnt[8];
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int v = i < n ? act[i] : 0;
    const int keep = (i < n) && (seed[v] >= 0);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warpCnt[w] = __popc(bal);
    __syncthreads();
    int base = blockCnt[blockIdx.x];
    for (int k = 0; k < w; k++) base += warpCnt[k];
    if (keep) out[base + __popc(bal & ((1u << lane) - 1u))] = v;
}

} // namespace vor
#endif
