// coop_kernels.cuh -- lane-group-cooperative round kernels (CUDA only): the product path on the B200.
//
// G lanes (G = 32: a whole warp, north_star's "one warp per point"; G = 8 optional) work on ONE pending point.  The
// serial dependency chain of the thread-per-point bodies in kernels.cuh (one in-sphere test after the other: ~60-100
// tests x 3-4 dependent gathers each) becomes one chain per BFS LEVEL of the conflict region:
//   attempt  walk: lane k evaluates the orientation of facet k (ballot -> facet to cross), one 256-bit record gather per
//            step; flood: each lane takes one (frontier simplex, facet) item: gathers the neighbour code, its owner pair,
//            its record and its 4 vertex sectors (256-bit loads) and evaluates the exact in-sphere test; reservation by
//            red.min / atomicMin; new killed simplices / boundary facets are appended with ballot + popc prefix sums
//            into the point's contiguous scratch.  One warp per block, 64 registers.
//   commit   (check + retriangulate fused) the cavity is staged in shared memory, lanes vote on ownership; a winner
//            allocates its block of simplex slots and retriangulates inside the SM (one lane per new simplex).
// What bounds them on the B200 is the scattered-load-instruction rate beyond the TLB reach, hence
// the wide loads; the compile-time switches below record the alternatives that were measured and lost.
// Same scratch format and same semantics as kernels.cuh (reference: delaunay_tree.rs:33-123, :213-334, scheduler.rs:6-55),
// so the two implementations are interchangeable (option "coop"); tests/emu exercises the thread-per-point bodies on
// the CPU, tests/test_gpu_* exercise these against the oracle on the B200.
#pragma once
#include "kernels.cuh"

#if VOR_GPU
namespace vor {

template <int G> __device__ __forceinline__ unsigned group_mask() {
    if (G == 32) return 0xffffffffu;
    const int lane = threadIdx.x & 31;
    return ((1u << G) - 1u) << ((lane / G) * G);
}

// Selection is implicit and stratified: slot g of a round attempts active entry a = g * stride + offset, i.e. one
// point per run of `stride` consecutive entries of the Morton-ordered active list (rotating offset), so no selection
// kernel and no compaction of the selected set are needed.  A round is two launches: attempt, commit.
struct RoundSel { int nact; int stride; int offset; int nsel; };
__device__ __forceinline__ int slot_entry(const RoundSel &rs, int slot) { return slot * rs.stride + rs.offset; }

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
// walk (option "prewalk"): forwarding + visibility walk of the points selected this round, G = 4 lanes per point
// (8 independent walks per warp) or G = 1 (32 per warp).  In the warp-per-point attempt kernel a walk keeps 4 of 32
// lanes busy and is ONE dependent chain of gathers per warp; run here it leaves the attempt kernel with the flood only
// (its own walk loop ends at the first test, on sectors this kernel has just pulled into L2).
// ------------------------------------------------------------------------------------------
template <int D, int G>
__global__ void __launch_bounds__(128) k_walk_coop(AttemptArgs<D> A, RoundSel rsel) {
    constexpr int M = Dim<D>::M;
    using Gm = Geo<D>;
    const Mesh<D> &m = A.m;
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int gl = threadIdx.x & (G - 1);
    const unsigned gmask = G == 1 ? __activemask() : group_mask<G>();
    const int gshift = (threadIdx.x & 31) & ~(G - 1);
    if (gid >= rsel.nsel) return;
    const int a = slot_entry(rsel, gid);
    if (a >= rsel.nact) return;
    const int v = A.act[a];
    int s = m.seed[v];
    if (s < 0) return;
    PredCtx cx{m.cnt};
    const typename Gm::Pt p = m.pts[v];
    int o;
    while ((o = __ldcg(&OWK(m, s))) < 0) s = ~o;
    unsigned rot = (unsigned)v * 2654435761u;
    unsigned steps = 0;
    for (;;) {
        int4 stv, stn;
        load_rec(m, s, stv, stn);
        const typename Gm::Verts tvv = Gm::load(m, stv);
        unsigned bal;
        if (G == 1) bal = (unsigned)Gm::beyond_mask(cx, tvv, p);
        else {
            const int ok = gl < M ? GeoCoop<D>::orient_repl(cx, tvv, p, gl) : 1;
            bal = (__ballot_sync(gmask, ok < 0) >> gshift) & ((1u << M) - 1u);
        }
        if (bal == 0) break;
        int go = 0;
        const int r0 = (int)((rot >> 16) % (unsigned)M);
        for (int k = 0; k < M; k++) {
            const int i = (r0 + k) % M;
            if ((bal >> i) & 1) { go = i; break; }
        }
        const int code = get4(stn, go);
        if (code < 0) { if (gl == 0) set_err(m.cnt, ERR_OUTSIDE); return; }
        s = code >> 2;
        rot = rot * 1664525u + 1013904223u;
        if (++steps > (1u << 22)) { if (gl == 0) set_err(m.cnt, ERR_WALK); return; }
    }
    if (gl == 0) {
        m.seed[v] = s;
        if (A.stats) atomicAdd(&m.cnt->walk_steps, (unsigned long long)steps);
    }
}

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
#ifndef VOR_DEDUP
#define VOR_DEDUP 0               // 1: lanes of a batch that reach the same neighbour elect one tester (match.any + shuffle);
                                  // saves the ~6 % duplicate tests but measured 12 % SLOWER (107 vs 95 ms): the election
                                  // sits in front of every gather
#endif
#ifndef VOR_ATT_STAGE
#define VOR_ATT_STAGE 0           // 1: the flood reads the killed list and the neighbour codes from a shared-memory copy of
                                  // the cavity instead of the global store (shorter dependent chain; measured 9 % SLOWER:
                                  // the kernel is bound by the gather-instruction rate, not by the chain, and the
                                  // copy costs registers)
#endif
#ifndef VOR_ATTEMPT_REGS
#define VOR_ATTEMPT_REGS 64       // registers per thread of the attempt kernel (measured best of 80/64/48)
#endif
// RED = 1: the kill reservation is a fire-and-forget reduction too (no round trip on the critical path of a flood
// level): lanes of one batch that reach the same simplex are deduplicated with match.any, and a better killer that
// slips in between the owner read and the reduction is caught by the ownership check of commit.
// EXACT = 0: the hot twin without the exact predicates in its call tree (PredCtxT<false>): an attempt whose filter fails
// is abandoned and its point flagged (A.slowFlag[v] = key base of the round); it skips flagged points.  EXACT = 1 with
// A.thr == 2: the slow twin, launched behind it while flagged points are pending, attempts ONLY points flagged in an
// EARLIER round (a point flagged in this round has left marks under this round's key).  A.thr == 0: everything.
template <int D, int G, int RED, int STAGE, int EXACT>
__device__ __forceinline__ void attempt_one(const AttemptArgs<D> &A, const RoundSel &rsel, const int gid, int *const sk, int4 *const sn) {
    constexpr int M = Dim<D>::M;
    using Gm = Geo<D>;
    constexpr int SK = VOR_SK;
    const Mesh<D> &m = A.m;
    const int gl = threadIdx.x & (G - 1);                          // lane inside the group
    const unsigned gmask = group_mask<G>();
    const int gshift = (threadIdx.x & 31) & ~(G - 1);              // first lane of the group inside the warp
    if (gid >= rsel.nsel) return;
    const int slot = gid;
    const int a = slot_entry(rsel, slot);
    if (a >= rsel.nact) { if (gl == 0) A.scr.slotStatus[slot] = ST_LOST; return; }
    const int v = A.act[a];
    if (m.seed[v] < 0) { if (gl == 0) A.scr.slotStatus[slot] = ST_LOST; return; }   // already inserted
    if (A.slowFlag) {
        const int fl = A.slowFlag[v];
        if (EXACT ? (A.thr == 2u && (fl == 0 || fl == A.keybase)) : fl != 0) {
            // not this twin's point.  The hot twin runs first and owns the slot's status; the slow twin leaves the
            // status of the points it skips alone
            if (!EXACT && gl == 0) A.scr.slotStatus[slot] = ST_LOST;
            return;
        }
    }
    PredCtxT<(EXACT != 0)> cx{m.cnt};
    const typename Gm::Pt p = m.pts[v];
    const uint32_t q = bij_hash((uint32_t)slot, A.bits, A.salt);   // unique among the slots of this round
    const int key_k = A.keybase | (int)(q << 1);
    const int key_o = key_k | 1;

    int status = ST_LOST, nk = 0, nb = 0, big = -1;
    unsigned steps = 0, tests = 0;

    // -- forwarding (all lanes follow the same chain: broadcast loads)
    int s = m.seed[v];
    int o;
    while ((o = __ldcg(&OWK(m, s))) < 0) s = ~o;

    // -- visibility walk: lane k < M tests facet k
    unsigned rot = (unsigned)v * 2654435761u;
    int4 stv, stn;                     // record of the simplex the walk stands in: one 256-bit gather per step
    load_rec(m, s, stv, stn);
    typename Gm::Verts tvv = Gm::load(m, stv);
    bool fail = false;
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
        load_rec(m, s, stv, stn);
        tvv = Gm::load(m, stv);
    }

    bool needSlow = !EXACT && __any_sync(gmask, cx.failed);
    if (needSlow) fail = true;
    if (!fail) {
        if (gl == 0) m.seed[v] = s;
        // -- containing simplex must be in conflict, otherwise p duplicates one of its vertices
        int c0 = 0;
        if (gl == 0) c0 = Gm::conflict(cx, tvv, p);
        c0 = __shfl_sync(gmask, c0, gshift);
        tests = gl == 0 ? 1u : 0u;
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
    if (!fail) {
        int old0 = 0;
        if (gl == 0) {
            old0 = atomicMin(&OWK(m, s), key_k);
            if (__ldcg(&OWR(m, s)) < key_k) old0 = -1;     // a better point keeps s in its outer ring
        }
        old0 = __shfl_sync(gmask, old0, gshift);
        if (old0 < key_k) fail = true;
    }
    if (!fail) {
        ScrView sv = scr_view(A.scr, slot, -1);
        if (gl == 0) { sv.k[0] = s; if (STAGE) { sk[0] = s; sn[0] = stn; } }
        __syncwarp(gmask);
        nk = 1;
        int head = 0;
        bool lost = false;
        while (head < nk && !lost) {
            const int tail = nk;
            const int items = (tail - head) * M;
            for (int base = 0; base < items && !lost; base += G) {
#if VOR_DEDUP
                // Lanes of one batch that reach the same neighbour n (a ring simplex seen from two killed simplices of
                // the same level) elect ONE of them with match.any: it gathers and tests n, the others take its verdict
                // by shuffle.  Saves the ~6 % duplicate tests (7 gather instructions each) and makes the claim of a
                // newly killed simplex unique without looking at the value an atomic returns.
                const int j = base + gl;
                bool pushK = false, pushB = false, lostLane = false;
                int newT = 0, fcode = 0, ocode = 0;
                int4 nnb = make_int4(-1, -1, -1, -1);
                const bool active = j < items;
                int t = 0, i = 0, code = -1;
                int n = -1 - (int)(threadIdx.x & 31);      // unique dummy: inactive lanes and hull facets match nobody
                if (active) {
                    const int e = head + j / M;
                    i = j % M;
                    if (STAGE && e < SK) { t = sk[e]; code = reinterpret_cast<const int *>(sn)[e * 4 + i]; }
                    else { t = sv.k[e]; code = TNI(m, t, i); }
                    if (code >= 0) n = code >> 2;
                    else { pushB = true; fcode = t * 4 + i; ocode = code; }
                }
                const unsigned same = __match_any_sync(gmask, n);
                const int leader = __ffs(same) - 1;
                const bool lead = n >= 0 && leader == (int)(threadIdx.x & 31);
                enum { V_MINE = 0, V_LOST = 1, V_RING = 2, V_NEW = 3 };
                int verdict = V_MINE;
                if (lead) {
                    // the owner pair and the record of n are independent gathers: issue both before looking at either
                    const int2 ow = __ldcg(reinterpret_cast<const int2 *>(&OWK(m, n)));   // x = kill word, y = ring word
                    int4 nverts;
                    if (STAGE) load_rec_cg(m, n, nverts, nnb);   // vertex ids + neighbour codes: one 256-bit gather
                    else nverts = __ldcg(&TV(m, n));
                    if (ow.x == key_k) verdict = V_MINE;            // already in my cavity
                    else if (ow.x < key_k) verdict = V_LOST;        // a better point kills n (or n is dead)
                    else if (ow.y == key_o) verdict = V_RING;       // already tested by me: not in conflict
                    else {
                        tests++;
                        const typename Gm::Verts nv = Gm::load(m, nverts);
                        if (Gm::conflict(cx, nv, p)) {
                            if (ow.y < key_k) verdict = V_LOST;     // a better point keeps n in its outer ring
                            else if (RED) { atomicMin(&OWK(m, n), key_k); verdict = V_NEW; }
                            else {
                                const int old = atomicMin(&OWK(m, n), key_k);
                                verdict = old < key_k ? V_LOST : (old != key_k ? V_NEW : V_MINE);
                            }
                        } else {
                            // outer-ring mark: fire and forget (RED, no round trip).  Rings may be shared; a better
                            // point that KILLS n was either seen above or is caught by the ownership check of commit.
                            atomicMin(&OWR(m, n), key_o);
                            verdict = V_RING;
                        }
                    }
                }
                verdict = __shfl_sync(gmask, verdict, leader);
                if (n >= 0) {
                    if (verdict == V_LOST) lostLane = true;
                    else if (verdict == V_RING) { pushB = true; fcode = t * 4 + i; ocode = code; }
                    else if (verdict == V_NEW && lead) { pushK = true; newT = n; }
                }
                if (__any_sync(gmask, lostLane || (!EXACT && cx.failed))) { lost = true; break; }
#else
                const int j = base + gl;
                bool pushK = false, pushB = false, lostLane = false;
                int newT = 0, fcode = 0, ocode = 0;
                int claim = -1 - gl;   // RED: simplex this lane wants to kill (unique dummy otherwise)
                int4 nnb = make_int4(-1, -1, -1, -1);
                if (j < items) {
                    const int e = head + j / M;
                    const int i = j % M;
                    int t, code;
                    if (STAGE && e < SK) { t = sk[e]; code = reinterpret_cast<const int *>(sn)[e * 4 + i]; }
                    else { t = sv.k[e]; code = TNI(m, t, i); }
                    if (code < 0) {
                        pushB = true; fcode = t * 4 + i; ocode = code;
                    } else {
                        const int n = code >> 2;
                        // the owner pair and the record of n are independent gathers: issue both before looking at either
                        const int2 ow = __ldcg(reinterpret_cast<const int2 *>(&OWK(m, n)));   // x = kill word, y = ring word
                        int4 nverts;
                        if (STAGE) load_rec_cg(m, n, nverts, nnb);   // vertex ids + neighbour codes: one 256-bit gather
                        else nverts = __ldcg(&TV(m, n));
                        if (ow.x == key_k) {
                            // already in my cavity
                        } else if (ow.x < key_k) {
                            lostLane = true;           // a better point kills n (or n is dead)
                        } else if (ow.y == key_o) {
                            pushB = true; fcode = t * 4 + i; ocode = code;   // already tested by me: not in conflict
                        } else {
                            tests++;
                            const typename Gm::Verts nv = Gm::load(m, nverts);
                            if (Gm::conflict(cx, nv, p)) {
                                if (ow.y < key_k) lostLane = true;   // a better point keeps n in its outer ring
                                else if (RED) {
                                    atomicMin(&OWK(m, n), key_k);
                                    claim = n;
                                } else {
                                    const int old = atomicMin(&OWK(m, n), key_k);
                                    if (old < key_k) lostLane = true;
                                    else if (old != key_k) { pushK = true; newT = n; }   // first lane to claim it appends it
                                }
                            } else {
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
#endif
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
                    if (gl == 0) A.scr.slotBig[slot] = big;
                    __syncwarp(gmask);
                    if (nk + ck > sv.capk || nb + cb > sv.capb) { if (gl == 0) set_err(m.cnt, ERR_CAPACITY); lost = true; break; }
                }
                const unsigned lt = (1u << gl) - 1u;
                if (pushK) {
                    const int pos = nk + __popc(mk & lt);
                    sv.k[pos] = newT;
                    if (STAGE && pos < SK) { sk[pos] = newT; sn[pos] = nnb; }
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
    if (gl == 0) {
        A.scr.slotStatus[slot] = status;
        A.scr.slotNk[slot] = nk;
        A.scr.slotNb[slot] = nb;
        if (big < 0) A.scr.slotBig[slot] = -1;
    }
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

// One group per attempt slot (grid = slots), or -- option "persist" -- resident warps that pull slots from per-SM
// queues: SM i works through the i-th contiguous range of the Morton-ordered slots, so the simplices and vertices it
// gathers in one round come from one compact region of the mesh (and of the store) instead of every 148th block of
// it; the ranges of SMs that finish early are drained by the others.
// The hot twin needs fewer registers than the kernel with the exact path: two-warp blocks at 56 registers (36 warps per
// SM instead of 32) measured 88.7 vs 92.1 ms per 10M points; 48 registers spill too much (110 ms).
#ifndef VOR_HOT_BLOCK
#define VOR_HOT_BLOCK 64
#endif
#ifndef VOR_HOT_REGS
#define VOR_HOT_REGS 56
#endif
template <int EXACT> struct AttemptLaunch {
    static constexpr int block = EXACT ? VOR_ATTEMPT_BLOCK : VOR_HOT_BLOCK;
    static constexpr int regs = EXACT ? VOR_ATTEMPT_REGS : VOR_HOT_REGS;
    static constexpr int minBlocks = (65536 / (regs * block)) > 32 ? 32 : (65536 / (regs * block));
};
template <int D, int G, int RED, int STAGE, int EXACT>
__global__ void __launch_bounds__(AttemptLaunch<EXACT>::block, AttemptLaunch<EXACT>::minBlocks) k_attempt_coop(AttemptArgs<D> A, RoundSel rsel) {
    // the cavity found so far, staged in shared memory: ids and neighbour codes of the first SK killed simplices, so a
    // flood level starts from two shared-memory reads instead of two dependent L2 round trips (scratch, then record)
    __shared__ int s_kid[AttemptLaunch<EXACT>::block / G][STAGE ? VOR_SK : 1];
    __shared__ int4 s_knb[AttemptLaunch<EXACT>::block / G][STAGE ? VOR_SK : 1];
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / G;   // group = attempt slot
    if (gid >= rsel.nsel) return;
    attempt_one<D, G, RED, STAGE, EXACT>(A, rsel, gid, s_kid[threadIdx.x / G], s_knb[threadIdx.x / G]);
}

constexpr int NQUEUE = 148;       // per-SM slot queues of the persistent kernels (one per SM of the B200)
template <int D, int RED>
__global__ void __launch_bounds__(VOR_ATTEMPT_BLOCK, 65536 / (VOR_ATTEMPT_REGS * VOR_ATTEMPT_BLOCK)) k_attempt_persist(AttemptArgs<D> A, RoundSel rsel, int *qctr) {
    __shared__ int s_kid[VOR_ATTEMPT_BLOCK / 32][VOR_SK];
    __shared__ int4 s_knb[VOR_ATTEMPT_BLOCK / 32][VOR_SK];
    unsigned smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    const int lane = threadIdx.x & 31;
    const int per = (rsel.nsel + NQUEUE - 1) / NQUEUE;
    for (int r = 0; r < NQUEUE; r++) {
        const int qi = (int)((smid + (unsigned)r) % (unsigned)NQUEUE);
        const int lo = qi * per, size = min(per, rsel.nsel - lo);
        if (size <= 0) continue;
        if (r > 0 && *(volatile int *)&qctr[qi] >= size) continue;     // drained
        for (;;) {
            int g = 0;
            if (lane == 0) g = atomicAdd(&qctr[qi], 1);
            g = __shfl_sync(0xffffffffu, g, 0);
            if (g >= size) break;
            attempt_one<D, 32, RED, VOR_ATT_STAGE, 1>(A, rsel, lo + g, s_kid[threadIdx.x / 32], s_knb[threadIdx.x / 32]);
            __syncwarp();
        }
    }
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
template <int D, int G>
__global__ void __launch_bounds__(VOR_COOP_BLOCK) k_commit_coop(CheckArgs<D> A, const int *act, RoundSel rsel, int stats) {
    constexpr int M = Dim<D>::M;
    constexpr int CK = VOR_CK, CB = VOR_CB, HS = 128, GPB = VOR_COOP_BLOCK / G;
    static_assert(HS >= 2 * CK, "hash must stay at most half full");
    __shared__ int4 s_tv[GPB][CK], s_tn[GPB][CK];
    __shared__ int s_id[GPB][CK], s_fw[GPB][CK], s_hash[GPB][HS], s_f[GPB][CB], s_o[GPB][CB];
    const Mesh<D> &m = A.m;
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int gl = threadIdx.x & (G - 1);
    const unsigned gmask = group_mask<G>();
    const int gshift = (threadIdx.x & 31) & ~(G - 1);
    if (gid == 0 && gl == 0) m.cnt->nbig = 0;           // overflow slots are per round (attempt is over)
    if (gid >= rsel.nsel) return;
    const int slot = gid;
    if (A.scr.slotStatus[slot] != ST_OK) return;
    const int a = slot_entry(rsel, slot);
    const int v = act[a];
    const uint32_t q = bij_hash((uint32_t)slot, A.bits, A.salt);   // unique among the slots of this round
    const int key_k = A.keybase | (int)(q << 1);
    const ScrView sv = scr_view(A.scr, slot, A.scr.slotBig[slot]);
    const int nk = A.scr.slotNk[slot], nb = A.scr.slotNb[slot];
    const bool fast = !(stats & 2) && nk <= CK && nb <= CB;   // stats bit 1: force the global-store path (A/B switch)
    const int grp = threadIdx.x / G;
    int4 *const tvs = s_tv[grp], *const tns = s_tn[grp];
    int *const ids = s_id[grp], *const fw = s_fw[grp], *const hash = s_hash[grp], *const sf = s_f[grp], *const so = s_o[grp];
    int *const tni = reinterpret_cast<int *>(tns);

    // -- ownership check; the fast path loads the cavity in the same level of gathers
    bool bad = false;
    if (fast) {
        for (int e = gl; e < nk; e += G) {
            const int t = sv.k[e];
            const int2 ow = __ldcg(reinterpret_cast<const int2 *>(&OWK(m, t)));
            int4 tv, tn;
            load_rec_cg(m, t, tv, tn);
            if (ow.x != key_k || ow.y < key_k) bad = true;     // best killer, and no better point has it in its ring
            ids[e] = t; tvs[e] = tv; tns[e] = tn; fw[e] = 0;
        }
        for (int j = gl; j < nb; j += G) {
            const int f = sv.f[j], code = sv.o[j];
            if (code >= 0 && __ldcg(&OWK(m, code >> 2)) < key_k) bad = true;   // no better point kills my outer ring
            sf[j] = f; so[j] = code;
        }
        for (int h = gl; h < HS; h += G) hash[h] = -1;
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
    // Slot recycling (fast path): new simplex j < nk takes the slot of killed simplex j, only the surplus comes from
    // the bump allocator.  The store stays compact (no dead slots: 4x smaller footprint at 10M points), a new simplex
    // lies where the simplices it replaces lay (spatial locality of the store is inherited, not diluted by time), and
    // a pending point whose seed was killed finds a live simplex of the right neighbourhood in the same slot.
    // The kill word of a recycled slot keeps this winner's key: every other contender of this round has a worse key
    // on it and loses; from the next round on it is a stale mark (epochs count down).
    const bool reuse = fast && !(stats & 4);
    const int nfresh = reuse ? max(nb - nk, 0) : nb;
    int base = 0;
    if (gl == 0) base = atomicAdd(&m.cnt->ntets, nfresh);
    base = __shfl_sync(gmask, base, gshift);
    if (base + nfresh > m.cap) {
        // no room: leave the mesh untouched (the point stays pending), retire the part of the block that exists and
        // tell the host to grow the store
        for (int j = gl; j < nfresh; j += G)
            if (base + j < m.cap) OWK(m, base + j) = -1;
        if (gl == 0) m.cnt->oom_soft = 1;
        return;
    }
    int first = base;     // a simplex created by this insertion (seed for later points)
    if (!fast) {
        commit_global<D, G>(m, sv, nk, nb, v, base, gl, gmask);
    } else {
        const int nreuse = reuse ? min(nk, nb) : 0;
        auto slot_of = [&](int j) -> int { return j < nreuse ? ids[j] : base + (j - nreuse); };
        first = slot_of(0);
        // id -> local index (open addressing, at most 3/8 full)
        for (int e = gl; e < nk; e += G) {
            unsigned h = ((unsigned)ids[e] * 2654435761u) >> 25;
            while (atomicCAS(&hash[h], -1, e) != -1) h = (h + 1) & (HS - 1);
        }
        __syncwarp(gmask);
        auto local_of = [&](int t) -> int {
            unsigned h = ((unsigned)t * 2654435761u) >> 25;
            for (int probe = 0; probe < HS; probe++) {
                const int e = hash[h];
                if (e < 0) break;
                if (ids[e] == t) return e;
                h = (h + 1) & (HS - 1);
            }
            set_err(m.cnt, ERR_CUDA);   // a pivot left the cavity: cannot happen on a consistent mesh
            return -1;
        };
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
                    const int ln = local_of(code >> 2), jb = code & 3;
                    if (ln < 0) break;
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
            store_rec(m, T, verts, nbr);
            if (outer >= 0) TNI(m, outer >> 2, outer & 3) = T * 4 + i;
        }
        // phase C: killed simplices whose slot is not recycled die; a dead simplex forwards to a new simplex on one
        // of its own boundary facets (interior simplices: to the first new simplex)
        for (int e = gl + nreuse; e < nk; e += G) OWK(m, ids[e]) = ~slot_of(fw[e]);
    }
    if (gl == 0) {
        m.ptTet[v] = first;
        m.seed[v] = -1;
        atomicAdd(&m.cnt->part[gid & (NPART - 1)][0], (1ULL << 40) | (unsigned long long)nb);   // win_total, created_all
        if (A.slowFlag && A.slowFlag[v] != 0) atomicAdd(&m.cnt->nflag_done, 1);
        if (stats & 1) {
            atomicAdd(&m.cnt->killed, (unsigned long long)nk);
            atomicAdd(&m.cnt->created, (unsigned long long)nb);
        }
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
