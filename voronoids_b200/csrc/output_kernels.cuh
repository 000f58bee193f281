// output_kernels.cuh -- Delaunay-graph extraction, validation and export.
//
//   edges      SURVEY.md §8a row G.  The reference has no extraction function: the graph is implicit in
//              DelaunayTree.simplices[*].vertices and rebuilt as Python dicts on every getter call
//              (/root/reference/src/lib.rs:73-101).  Definition adopted: {lo,hi} input indices of two real vertices that
//              share a live simplex, lo<hi, lexicographically sorted, unique, little-endian u32 pairs.
//              Each edge is emitted once, by the lowest-numbered simplex of the ring around it (3D: pivot around
//              the edge; 2D: the lower of the two triangles), into CSR rows keyed by lo, then rows are sorted.
//   validate   structural + local-Delaunay check of every live simplex (local Delaunay on every interior facet is
//              equivalent to the reference's brute-force check_delaunay, delaunay_tree.rs:512-541, but O(S)).
//   circumsphere / export: the reference caches center/radius per simplex (delaunay_tree.rs:11-16); here they are
//              computed on demand for the Python getters with the reference's own formulas (geometry.rs:2-56).
#pragma once
#include "kernels.cuh"

namespace vor {

template <int D> VOR_HD bool simplex_live(const Mesh<D> &m, int t) { return OWK(m, t) >= 0; }

template <int D> struct EdgeArgs {
    Mesh<D> m;
    const int *inputIdx;   // vertex -> global input index
    int *deg;              // [nInput+1] row sizes, then exclusive offsets
    int *cursor;           // [nInput] fill cursors
    uint32_t *hi;          // CSR column array
    int pass;              // 0 = count, 1 = fill
    unsigned char *mask;   // 3D: per simplex slot, bit e = this simplex owns its e-th edge (written by the count pass so
                           // that the fill pass does not repeat the pivots around every edge); may be null
};

// true if simplex t is the lowest-numbered simplex around edge (slots sa, sb) -- 3D pivot around the edge
VOR_HD bool edge_owner3(const Mesh<3> &m, int t, const int4 &tv, int sa, int sb) {
    const int a = get4(tv, sa), b = get4(tv, sb);
    // leave t through the face opposite the first of the two other slots
    int exitf = -1, enterf = -1;
    for (int s = 0; s < 4; s++) {
        if (s == sa || s == sb) continue;
        if (exitf < 0) exitf = s; else enterf = s;
    }
    (void)enterf;
    int cur = t;
    for (int guard = 0; guard < (1 << 20); guard++) {
        const int code = get4(TN(m, cur), exitf);
        if (code < 0) return false; // cannot happen for an edge of two real vertices
        const int nxt = code >> 2, jb = code & 3;
        if (nxt == t) return true;
        if (nxt < t) return false;
        const int4 nv = TV(m, nxt);
        int y = -1;
        for (int s = 0; s < 4; s++) {
            if (s == jb) continue;
            const int vv = get4(nv, s);
            if (vv != a && vv != b) y = s;
        }
        cur = nxt;
        exitf = y;
    }
    return false;
}

template <int D> VOR_HD void edge_emit(const EdgeArgs<D> &A, int va, int vb) {
    const int ia = A.inputIdx[va], ib = A.inputIdx[vb];
    const int lo = ia < ib ? ia : ib, hi = ia < ib ? ib : ia;
    if (A.pass == 0) atomic_add_i(&A.deg[lo], 1);
    else {
        const int w = atomic_add_i(&A.cursor[lo], 1);
        A.hi[(size_t)A.deg[lo] + w] = (uint32_t)hi;
    }
}

template <int D> VOR_HD void edges_body(const EdgeArgs<D> &A, int t) {
    constexpr int M = Dim<D>::M;
    const Mesh<D> &m = A.m;
    if (!simplex_live(m, t)) return;
    const int4 tv = TV(m, t);
    if constexpr (D == 3) {
        const bool replay = A.pass == 1 && A.mask != nullptr;
        const unsigned have = replay ? A.mask[t] : 0u;
        unsigned own = 0;
        int e = 0;
        for (int sa = 0; sa < M; sa++)
            for (int sb = sa + 1; sb < M; sb++, e++) {
                const int va = get4(tv, sa), vb = get4(tv, sb);
                if (va < m.nsuper || vb < m.nsuper) continue;
                if (replay ? ((have >> e) & 1u) != 0u : edge_owner3(m, t, tv, sa, sb)) { own |= 1u << e; edge_emit(A, va, vb); }
            }
        if (A.pass == 0 && A.mask != nullptr) A.mask[t] = (unsigned char)own;
    } else {
        const int4 tn = TN(m, t);
        for (int i = 0; i < 3; i++) { // edge opposite slot i
            const int va = get4(tv, (i + 1) % 3), vb = get4(tv, (i + 2) % 3);
            if (va < m.nsuper || vb < m.nsuper) continue;
            const int code = get4(tn, i);
            if (code < 0 || t < (code >> 2)) edge_emit(A, va, vb);
        }
    }
}

// ---- 3D count pass WITHOUT walking around the edges (option edge_wedge).  The simplices around an interior edge (a, b) tile the
// full angle with their dihedral wedges, so exactly one of them contains the point q = p_lo + dir (p_lo = the endpoint with the
// lower vertex id, dir a fixed generic direction: q is the SAME double point for every simplex around the edge): that simplex owns
// the edge.  "q inside the wedge of t" = the orientation of t with the opposite vertex replaced by q is positive for both faces of
// t through the edge (exact predicate: FP64 filter -> double-double -> integers).  Four gathered vertices and 12 orientations per
// simplex instead of ~2.5 dependent record gathers per edge.  Ties (measure zero, but the rule must be consistent; provoked in
// tests/enginecases.py::case_edge_wedge_ties): q exactly on ONE of the two faces -- that face is the boundary between t and its
// neighbour across it, the lower simplex id takes the edge; q on BOTH planes, i.e. on the line through a and b -- every simplex
// around the edge sees that, all fall back to the pivot rule.  A kernel body of its own with the slots resolved at compile time:
// inside edges_body the point array and the predicates cost the replay pass and the pivot path their registers (count pass
// 15.8 -> 25.3 ms, replay 7 -> 21 ms when it was tried there).
struct EdgeWedgeArgs {
    Mesh<3> m;
    const int *inputIdx;
    int *deg;
    unsigned char *mask;
    double dx, dy, dz;
};
template <int K> VOR_HD const double4 &pick4(const double4 &p0, const double4 &p1, const double4 &p2, const double4 &p3) {
    if constexpr (K == 0) return p0; else if constexpr (K == 1) return p1; else if constexpr (K == 2) return p2; else return p3;
}
template <int K> VOR_HD int slot4(const int4 &v) {
    if constexpr (K == 0) return v.x; else if constexpr (K == 1) return v.y; else if constexpr (K == 2) return v.z; else return v.w;
}
template <int SA, int SB> VOR_HD bool wedge_own(PredCtx &cx, const EdgeWedgeArgs &A, int t, const int4 &tv, const double4 &p0, const double4 &p1,
                                                const double4 &p2, const double4 &p3) {
    constexpr int SC = (SA != 0 && SB != 0) ? 0 : ((SA != 1 && SB != 1) ? 1 : 2);
    constexpr int SD = 6 - SA - SB - SC;
    const bool aLo = slot4<SA>(tv) < slot4<SB>(tv);
    const double4 &pa = pick4<SA>(p0, p1, p2, p3), &pb = pick4<SB>(p0, p1, p2, p3);
    double4 q;
    q.x = (aLo ? pa.x : pb.x) + A.dx; q.y = (aLo ? pa.y : pb.y) + A.dy; q.z = (aLo ? pa.z : pb.z) + A.dz; q.w = 0.0;
    const int o1 = orient3d(cx, SC == 0 ? q : p0, SC == 1 ? q : p1, SC == 2 ? q : p2, SC == 3 ? q : p3);
    if (o1 < 0) return false;
    const int o2 = orient3d(cx, SD == 0 ? q : p0, SD == 1 ? q : p1, SD == 2 ? q : p2, SD == 3 ? q : p3);
    if (o2 < 0) return false;
    if (o1 > 0 && o2 > 0) return true;
    if (o1 == 0 && o2 == 0) return edge_owner3(A.m, t, tv, SA, SB);
    const int code = get4(TN(A.m, t), o1 == 0 ? SC : SD);     // the neighbour across the face q lies on
    return code >= 0 && t < (code >> 2);
}
VOR_HD void edges_wedge_body(const EdgeWedgeArgs &A, int t) {
    const Mesh<3> &m = A.m;
    if (!simplex_live(m, t)) return;
    const int4 tv = TV(m, t);
    const int ns = m.nsuper;
    const bool r0 = tv.x >= ns, r1 = tv.y >= ns, r2 = tv.z >= ns, r3 = tv.w >= ns;
    if ((int)r0 + (int)r1 + (int)r2 + (int)r3 < 2) { A.mask[t] = 0; return; }   // no edge of two real vertices
    const Geo<3>::Verts vv = Geo<3>::load(m, tv);
    PredCtx cx{m.cnt};
    unsigned own = 0;
    auto emit = [&](int va, int vb) {
        const int ia = A.inputIdx[va], ib = A.inputIdx[vb];
        atomic_add_i(&A.deg[ia < ib ? ia : ib], 1);
    };
    // edge numbering of edges_body: (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)
    if (r0 && r1 && wedge_own<0, 1>(cx, A, t, tv, vv.p0, vv.p1, vv.p2, vv.p3)) { own |= 1u; emit(tv.x, tv.y); }
    if (r0 && r2 && wedge_own<0, 2>(cx, A, t, tv, vv.p0, vv.p1, vv.p2, vv.p3)) { own |= 2u; emit(tv.x, tv.z); }
    if (r0 && r3 && wedge_own<0, 3>(cx, A, t, tv, vv.p0, vv.p1, vv.p2, vv.p3)) { own |= 4u; emit(tv.x, tv.w); }
    if (r1 && r2 && wedge_own<1, 2>(cx, A, t, tv, vv.p0, vv.p1, vv.p2, vv.p3)) { own |= 8u; emit(tv.y, tv.z); }
    if (r1 && r3 && wedge_own<1, 3>(cx, A, t, tv, vv.p0, vv.p1, vv.p2, vv.p3)) { own |= 16u; emit(tv.y, tv.w); }
    if (r2 && r3 && wedge_own<2, 3>(cx, A, t, tv, vv.p0, vv.p1, vv.p2, vv.p3)) { own |= 32u; emit(tv.z, tv.w); }
    A.mask[t] = (unsigned char)own;
}

struct RowSortArgs { const int *off; uint32_t *hi; uint32_t *out; int n; };
// sort each CSR row (insertion sort: rows hold ~8 entries in 3D, ~3 in 2D) and write (lo,hi) pairs
VOR_HD void row_sort_body(const RowSortArgs &A, int r) {
    const int lo = A.off[r], hi = A.off[r + 1];
    for (int i = lo + 1; i < hi; i++) {
        const uint32_t x = A.hi[i];
        int j = i - 1;
        while (j >= lo && A.hi[j] > x) { A.hi[j + 1] = A.hi[j]; j--; }
        A.hi[j + 1] = x;
    }
    for (int i = lo; i < hi; i++) { A.out[2 * (size_t)i] = (uint32_t)r; A.out[2 * (size_t)i + 1] = A.hi[i]; }
}

// order-independent checksum of the canonical edge list (bench: result read-back without copying every edge)
struct EdgeSumArgs { const uint32_t *edges; unsigned long long *sum; };
VOR_HD void edge_sum_body(const EdgeSumArgs &A, int i) {
    const uint64_t k = ((uint64_t)A.edges[2 * (size_t)i] << 32) | A.edges[2 * (size_t)i + 1];
    atomic_add_ull(A.sum, mix64(k));
}

// per-set edge count and order-independent checksum of a batch tree's canonical edge list (sets are independent units,
// BASELINE.json configs[4]): the list is sorted by lo, so every set owns one contiguous range; a thread walks a run of
// edges, folds them locally (indices made local to the set) and flushes once per set it meets.
struct SetEdgeArgs { const uint32_t *edges; long long m; const int *setOff; int nsets; unsigned long long *cnt; unsigned long long *sum; int run; };
VOR_HD void set_edge_stats_body(const SetEdgeArgs &A, int b) {
    const long long lo = (long long)b * A.run, hi = lo + A.run < A.m ? lo + A.run : A.m;
    if (lo >= hi) return;
    int s = 0;
    {   // set of the first edge: last set whose offset is <= its lo endpoint
        const int first = (int)A.edges[2 * (size_t)lo];
        int a = 0, z = A.nsets;
        while (z - a > 1) { const int mid = (a + z) >> 1; if (A.setOff[mid] <= first) a = mid; else z = mid; }
        s = a;
    }
    unsigned long long c = 0, acc = 0;
    for (long long i = lo; i < hi; i++) {
        const int el = (int)A.edges[2 * (size_t)i], eh = (int)A.edges[2 * (size_t)i + 1];
        while (el >= A.setOff[s + 1]) {
            if (c) { atomic_add_ull(&A.cnt[s], c); atomic_add_ull(&A.sum[s], acc); }
            c = 0; acc = 0; s++;
        }
        const uint64_t k = ((uint64_t)(uint32_t)(el - A.setOff[s]) << 32) | (uint32_t)(eh - A.setOff[s]);
        acc += mix64(k);
        c++;
    }
    if (c) { atomic_add_ull(&A.cnt[s], c); atomic_add_ull(&A.sum[s], acc); }
}

// ---- slab certification.  A simplex of the triangulation of (own points + halo + coarse sample) is a
// simplex of the GLOBAL triangulation iff no point of the global set lies strictly inside its circumsphere.  This tree
// holds every global point whose `axis` coordinate is in [lo, hi]; all points lie in the data box.  So it suffices that
// (open ball) n (data box) stays inside [lo, hi] along `axis`.  The ball is taken from the certified sphere block
// (sphere.cuh): the true open ball is contained in { q : |q - c|^2 <= rout2 } (c relative to the origin), hence the
// test errs only on the side of "not certified".  Only simplices with a vertex this slab OWNS need the certificate.
VOR_HD double atomic_min_d(double *p, double v) {
#ifdef __CUDA_ARCH__
    unsigned long long *a = reinterpret_cast<unsigned long long *>(p), old = *a, assumed;
    do { assumed = old; if (__longlong_as_double((long long)assumed) <= v) break; old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v)); } while (old != assumed);
    return __longlong_as_double((long long)old);
#else
    const double o = *p; if (v < o) *p = v; return o;
#endif
}
VOR_HD double atomic_max_d(double *p, double v) {
#ifdef __CUDA_ARCH__
    unsigned long long *a = reinterpret_cast<unsigned long long *>(p), old = *a, assumed;
    do { assumed = old; if (__longlong_as_double((long long)assumed) >= v) break; old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v)); } while (old != assumed);
    return __longlong_as_double((long long)old);
#else
    const double o = *p; if (v > o) *p = v; return o;
#endif
}
template <int D> struct CertifyArgs {
    Mesh<D> m;
    const int *inputIdx;          // vertex -> input index (-1: super vertex)
    const unsigned char *owned;   // per input index
    unsigned long long *count;
    double *need;                 // [2]: extent along `axis` the uncertified simplices reach
    int axis;
    double lo, hi;
    double shell;                 // the tree also holds every global point within `shell` of a lateral face of the data box
    double boxLo[3], boxHi[3];
    // optional list of the uncertified simplices (vor_tree_uncertified_slab): M x D vertex coordinates in the order of the mesh
    // record (positively oriented) and the reach [elo, ehi] of each, up to listCap of them; *count keeps counting beyond it
    double *listVerts = nullptr;
    double *listReach = nullptr;
    int listCap = 0;
};
template <int D> VOR_HD void certify_body(const CertifyArgs<D> &A, int t) {
    constexpr int M = Dim<D>::M;
    const Mesh<D> &m = A.m;
    if (!simplex_live(m, t)) return;
    const int4 tv = TV(m, t);
    bool mine = false;
    for (int k = 0; k < M; k++) {
        const int v = get4(tv, k);
        if (v >= m.nsuper && A.owned[A.inputIdx[v]]) mine = true;
    }
    if (!mine) return;
    // extent along the axis of (ball of centre c, squared radius r2) n (data box minus the lateral shell); false: it misses the box
    auto extent = [&](const double c[3], double r2, double &elo, double &ehi) -> bool {
        elo = A.boxLo[A.axis]; ehi = A.boxHi[A.axis];           // no bound (r2 = inf): the whole box
        if (!(r2 < 1e300)) return true;
        // squared distance of the centre to the part of the box this tree may be missing points from: the box shrunk by the
        // shell in the OTHER axes (an empty rest: every point of the box is held, nothing to certify)
        double d2 = 0.0;
        const double sh = A.shell * (1.0 - 1e-9);
        for (int k = 0; k < D; k++) {
            if (k == A.axis) continue;
            const double l = A.boxLo[k] + sh, h = A.boxHi[k] - sh;
            if (l >= h) return false;
            const double d = c[k] < l ? l - c[k] : (c[k] > h ? c[k] - h : 0.0);
            d2 += d * d;
        }
        d2 *= (1.0 - 1e-12);
        if (r2 <= d2) return false;                             // the ball misses the data box: nothing can be inside it
        const double half = sqrt(r2 - d2) * (1.0 + 1e-12) + 1e-300;
        const double slack = 4.0 * SPH_EPS * (fabs(c[A.axis]) + half);
        elo = fmax(elo, c[A.axis] - half - slack);
        ehi = fmin(ehi, c[A.axis] + half + slack);
        return elo <= ehi;                                      // else: misses the box along the axis
    };
    const OwnBlk b = load_blk(m, t);
    const double cb[3] = {(double)b.cx + m.sref.ox, (double)b.cy + m.sref.oy, (double)b.cz + m.sref.oz};
    double elo, ehi;
    if (!extent(cb, (double)b.rout2 * (1.0 + 1e-12), elo, ehi)) return;
    if (elo >= A.lo && ehi <= A.hi) return;                     // certified by the stored block
    // second chance in f64: the float block of a hull simplex (radius of 1e3..1e5 box widths) is good enough for conflict tests
    // but moves the reach of its cap into the box by tenths of the box (sphere.cuh, sphere_ball_d)
    {
        const typename Geo<D>::Verts vv = Geo<D>::load(m, tv);
        double cd[3], Rout;
        bool ok;
        if constexpr (D == 3) ok = sphere_ball_d(vv.p0, vv.p1, vv.p2, vv.p3, cd, Rout);
        else ok = sphere_ball_d(vv.p0, vv.p1, vv.p2, cd, Rout);
        if (ok) {
            double e0, e1;
            if (!extent(cd, Rout * Rout * (1.0 + 16.0 * SPH_EPS), e0, e1)) return;
            elo = fmax(elo, e0); ehi = fmin(ehi, e1);           // both are valid bounds of the same set
            if (elo >= A.lo && ehi <= A.hi) return;
        }
    }
    const unsigned long long idx = atomic_add_ull(A.count, 1ULL);
    if (elo < A.lo) atomic_min_d(&A.need[0], elo);
    if (ehi > A.hi) atomic_max_d(&A.need[1], ehi);
    if (A.listVerts && idx < (unsigned long long)A.listCap) {
        for (int k = 0; k < M; k++) {
            const typename Dim<D>::Pt q = m.pts[get4(tv, k)];
            A.listVerts[(idx * M + k) * D + 0] = q.x;
            A.listVerts[(idx * M + k) * D + 1] = q.y;
            if constexpr (D == 3) A.listVerts[(idx * M + k) * D + 2] = q.z;
        }
        A.listReach[2 * idx] = elo;
        A.listReach[2 * idx + 1] = ehi;
    }
}

// slab mode, the certificate that is not a ball: how many of n points lie STRICTLY inside the circumsphere of each of k simplices
// (exact predicate: FP64 filter -> exact).  A rank asks its peers this about the handful of simplices whose ball it cannot bound
// (slivers on the hull); zero everywhere = the simplex is a simplex of the global triangulation.
template <int D> struct InSpheresArgs {
    const double *pts;            // n x D
    const double *simp;           // k x M x D, positively oriented
    int k;
    unsigned long long *inside;   // [k]
    Counters *cnt;
};
template <int D> VOR_HD void in_spheres_body(const InSpheresArgs<D> &A, int i) {
    constexpr int M = Dim<D>::M;
    typename Dim<D>::Pt p;
    p.x = A.pts[(size_t)i * D]; p.y = A.pts[(size_t)i * D + 1];
    if constexpr (D == 3) { p.z = A.pts[(size_t)i * D + 2]; p.w = 0.0; }
    PredCtx cx{A.cnt};
    for (int j = 0; j < A.k; j++) {
        typename Geo<D>::Verts vv;
        typename Dim<D>::Pt q[4];
        for (int v = 0; v < M; v++) {
            const double *c = A.simp + ((size_t)j * M + v) * D;
            q[v].x = c[0]; q[v].y = c[1];
            if constexpr (D == 3) { q[v].z = c[2]; q[v].w = 0.0; }
        }
        vv.p0 = q[0]; vv.p1 = q[1]; vv.p2 = q[2];
        if constexpr (D == 3) vv.p3 = q[3];
        if (Geo<D>::conflict(cx, vv, p)) atomic_add_ull(&A.inside[j], 1ULL);
    }
}

// slab mode: this slab's part of the GLOBAL canonical edge list.  An edge of the local list (local input indices) is
// emitted iff the endpoint with the lower GLOBAL index is owned by this slab; key = lo << 32 | hi in global indices,
// ~0 for edges that belong to another slab (they sort to the end).
struct SlabEdgeArgs { const uint32_t *edges; const long long *gmap; const unsigned char *owned; unsigned long long *keys; };
VOR_HD void slab_edge_key_body(const SlabEdgeArgs &A, int i) {
    const uint32_t a = A.edges[2 * (size_t)i], b = A.edges[2 * (size_t)i + 1];
    const long long ga = A.gmap[a], gb = A.gmap[b];
    const bool mine = ga < gb ? A.owned[a] != 0 : A.owned[b] != 0;
    const unsigned long long lo = (unsigned long long)(ga < gb ? ga : gb), hi = (unsigned long long)(ga < gb ? gb : ga);
    A.keys[i] = mine ? ((lo << 32) | hi) : ~0ULL;
}
struct KeyCountArgs { const unsigned long long *keys; unsigned long long *count; };
VOR_HD void key_count_body(const KeyCountArgs &A, int i) {   // sorted keys: the first ~0 marks the end of the part
    if (A.keys[i] != ~0ULL && A.keys[i + 1] == ~0ULL) *A.count = (unsigned long long)i + 1;
}
struct KeyUnpackArgs { const unsigned long long *keys; uint32_t *out; };
VOR_HD void key_unpack_body(const KeyUnpackArgs &A, int i) {
    A.out[2 * (size_t)i] = (uint32_t)(A.keys[i] >> 32);
    A.out[2 * (size_t)i + 1] = (uint32_t)(A.keys[i] & 0xffffffffULL);
}

// ---- validation
template <int D> struct ValidateArgs {
    Mesh<D> m;
    int *fail;   // [8] failure counters: 0 orientation, 1 dead neighbour, 2 asymmetric adjacency, 3 facet mismatch, 4 not Delaunay,
                 // 5 stored sphere filter certifies a verdict the exact predicate contradicts
    unsigned long long *nlive;
};
template <int D> VOR_HD void validate_body(const ValidateArgs<D> &A, int t) {
    constexpr int M = Dim<D>::M;
    using G = Geo<D>;
    const Mesh<D> &m = A.m;
    if (!simplex_live(m, t)) return;
    atomic_add_ull(A.nlive, 1ULL);
    PredCtx cx{m.cnt};
    const int4 tv = TV(m, t);
    const int4 tn = TN(m, t);
    const typename G::Verts vt = G::load(m, tv);
    if (G::orient(cx, vt) <= 0) atomic_add_i(&A.fail[0], 1);
    const OwnBlk blk = load_blk(m, t);
    for (int k = 0; k < M; k++)   // own vertices lie ON the sphere: never certainly inside
        if (sphere_test(blk, rel_pt(m, m.pts[get4(tv, k)])) > 0) atomic_add_i(&A.fail[5], 1);
    for (int i = 0; i < M; i++) {
        const int code = get4(tn, i);
        if (code < 0) continue;
        const int nb = code >> 2, jb = code & 3;
        if (!simplex_live(m, nb)) { atomic_add_i(&A.fail[1], 1); continue; }
        const int back = get4(TN(m, nb), jb);
        if (back != t * 4 + i) { atomic_add_i(&A.fail[2], 1); continue; }
        const int4 nv = TV(m, nb);
        bool ok = true;
        for (int k = 0; k < M; k++) {
            if (k == jb) continue;
            const int x = get4(nv, k);
            bool found = false;
            for (int q = 0; q < M; q++) found |= (q != i && get4(tv, q) == x);
            ok &= found;
        }
        if (!ok) { atomic_add_i(&A.fail[3], 1); continue; }
        const typename G::Pt opp = m.pts[get4(nv, jb)];   // closest vertices to the sphere: the filter's hardest queries
        const int exact = G::conflict(cx, vt, opp);
        if (exact) atomic_add_i(&A.fail[4], 1);
        const int sv = sphere_test(blk, rel_pt(m, opp));
        if (sv != 0 && (sv > 0) != (exact != 0)) atomic_add_i(&A.fail[5], 1);
    }
}

// ---- export of live simplices: compact index = rank of the slot among the live slots (deterministic, so the index
// spaces of export_simplices and locate agree from call to call)
template <int D> struct ExportArgs {
    Mesh<D> m;
    int *liveId;     // compact index -> simplex slot
    int *compactOf;  // simplex slot -> compact index (or -1); pass 0 writes the live flag, the host scans it
};
template <int D> VOR_HD void export_flag_body(const ExportArgs<D> &A, int t) { A.compactOf[t] = simplex_live(A.m, t) ? 1 : 0; }
template <int D> VOR_HD void export_mark_body(const ExportArgs<D> &A, int t) {
    if (!simplex_live(A.m, t)) { A.compactOf[t] = -1; return; }
    A.liveId[A.compactOf[t]] = t;
}
template <int D> struct ExportFillArgs {
    Mesh<D> m;
    const int *liveId;
    const int *compactOf;
    const int *inputIdx;
    int *verts;      // [n x M] vertex ids: super vertex k -> k, input point i -> idOffset + i
    int *nbrs;       // [n x M] compact neighbour index, hull facets -> -1 - (slot)  (resolved by the host)
    double *center;  // [n x D] reference circumcentre (geometry.rs), may be null
    double *radius;  // [n]
    int idOffset;
};
VOR_HD void circum_ref(const Geo<3>::Verts &t, double *c, double *r) {
    // geometry.rs:24-56 restated (LU with partial pivoting as in nalgebra; see oracle/ref_geometry.h)
    const double v[12] = {t.p0.x, t.p0.y, t.p0.z, t.p1.x, t.p1.y, t.p1.z, t.p2.x, t.p2.y, t.p2.z, t.p3.x, t.p3.y, t.p3.z};
    double a[3][3], b[3];
    for (int i = 0; i < 3; i++) {
        double s = 0.0;
        for (int k = 0; k < 3; k++) {
            a[i][k] = v[3 * (i + 1) + k] - v[k];
            const double mid = (v[3 * (i + 1) + k] + v[k]) / 2.0;
            s += a[i][k] * mid;
        }
        b[i] = s;
    }
    int perm[3] = {0, 1, 2};
    for (int i = 0; i < 3; i++) {
        int piv = i;
        double best = fabs(a[i][i]);
        for (int q = i + 1; q < 3; q++) if (fabs(a[q][i]) > best) { best = fabs(a[q][i]); piv = q; }
        const double diag = a[piv][i];
        if (diag == 0.0) continue;
        if (piv != i) {
            for (int k = 0; k < 3; k++) { const double tmp = a[i][k]; a[i][k] = a[piv][k]; a[piv][k] = tmp; }
            const int tp = perm[i]; perm[i] = perm[piv]; perm[piv] = tp;
        }
        const double inv = 1.0 / diag;
        for (int q = i + 1; q < 3; q++) a[q][i] *= inv;
        for (int k = i + 1; k < 3; k++)
            for (int q = i + 1; q < 3; q++) a[q][k] = (-a[i][k]) * a[q][i] + a[q][k];
    }
    double x[3] = {b[perm[0]], b[perm[1]], b[perm[2]]};
    for (int i = 0; i < 3; i++) { const double cf = x[i]; for (int q = i + 1; q < 3; q++) x[q] = (-cf) * a[q][i] + x[q]; }
    for (int i = 2; i >= 0; i--) { const double cf = x[i] / a[i][i]; x[i] = cf; for (int q = 0; q < i; q++) x[q] = (-cf) * a[q][i] + x[q]; }
    c[0] = x[0]; c[1] = x[1]; c[2] = x[2];
    *r = sqrt((v[0] - x[0]) * (v[0] - x[0]) + (v[1] - x[1]) * (v[1] - x[1]) + (v[2] - x[2]) * (v[2] - x[2]));
}
VOR_HD void circum_ref(const Geo<2>::Verts &t, double *c, double *r) {
    // geometry.rs:2-22 restated
    const double x1 = t.p0.x, y1 = t.p0.y, x2 = t.p1.x, y2 = t.p1.y, x3 = t.p2.x, y3 = t.p2.y;
    const double d0 = (x1 + x2) / 2.0, d1 = (y1 + y2) / 2.0, e0 = (x2 + x3) / 2.0, e1 = (y2 + y3) / 2.0;
    const double m_ab = (y2 - y1) / (x2 - x1), m_bc = (y3 - y2) / (x3 - x2);
    const double m_d = -1. / m_ab, m_e = -1. / m_bc;
    const double x = (m_d * d0 - m_e * e0 + e1 - d1) / (m_d - m_e);
    const double y = m_d * (x - d0) + d1;
    c[0] = x; c[1] = y;
    *r = sqrt((x - x1) * (x - x1) + (y - y1) * (y - y1));
}
template <int D> VOR_HD void export_fill_body(const ExportFillArgs<D> &A, int c) {
    constexpr int M = Dim<D>::M;
    const Mesh<D> &m = A.m;
    const int t = A.liveId[c];
    const int4 tv = TV(m, t);
    const int4 tn = TN(m, t);
    for (int k = 0; k < M; k++) {
        const int v = get4(tv, k);
        A.verts[(size_t)c * M + k] = v < m.nsuper ? v : A.idOffset + A.inputIdx[v];
        const int code = get4(tn, k);
        A.nbrs[(size_t)c * M + k] = code < 0 ? -1 : A.compactOf[code >> 2];
    }
    if (A.center) {
        const typename Geo<D>::Verts vt = Geo<D>::load(m, tv);
        circum_ref(vt, A.center + (size_t)c * D, A.radius + c);
    }
}

// ---- locate (reference: DelaunayTree::locate, delaunay_tree.rs:33-58): conflict region of query points, read only.
// Thread per query: seed from the Morton neighbour (same rule as init_seeds), visibility walk, flood; the visited test
// is a scan of the query's own result list (the owner words are not touched, so queries can run between inserts).
template <int D> struct LocateQueryArgs {
    Mesh<D> m;
    const double *q;          // nq x D query points
    const int *seedSimplex;   // per query: a live simplex to start from
    const int *compactOf;     // simplex slot -> export index
    int *out;                 // [nq x cap] export indices of the conflict region (unsorted)
    int *count;               // per query: region size, or -1 if it does not fit `cap`, -2 if outside the super simplex
    int cap;
};
VOR_HD double4 make_pt(const double *s, double4 *) { return double4{s[0], s[1], s[2], 0.0}; }
VOR_HD double2 make_pt(const double *s, double2 *) { return double2{s[0], s[1]}; }
template <int D> VOR_HD void locate_query_body(const LocateQueryArgs<D> &A, int qi) {
    constexpr int M = Dim<D>::M;
    using G = Geo<D>;
    const Mesh<D> &m = A.m;
    PredCtx cx{m.cnt};
    const typename G::Pt p = make_pt(A.q + (size_t)qi * D, (typename G::Pt *)nullptr);
    int s = A.seedSimplex[qi];
    int o;
    while ((o = OWK(m, s)) < 0) s = ~o;
    unsigned rot = (unsigned)qi * 2654435761u;
    typename G::Verts tvv = G::load(m, TV(m, s));
    for (unsigned steps = 0;; steps++) {
        const int mk = G::beyond_mask(cx, tvv, p);
        if (mk == 0) break;
        int go = 0;
        const int r0 = (int)((rot >> 16) % (unsigned)M);
        for (int k = 0; k < M; k++) {
            const int i = (r0 + k) % M;
            if ((mk >> i) & 1) { go = i; break; }
        }
        const int code = TNI(m, s, go);
        if (code < 0 || steps > (1u << 22)) { A.count[qi] = -2; return; }
        s = code >> 2;
        rot = rot * 1664525u + 1013904223u;
        tvv = G::load(m, TV(m, s));
    }
    int *res = A.out + (size_t)qi * A.cap;
    int nk = 0;
    if (!G::conflict(cx, tvv, p)) { A.count[qi] = 0; return; }   // p coincides with a vertex: empty region (the reference panics)
    res[nk++] = s;
    for (int head = 0; head < nk; head++) {
        const int4 nbr = TN(m, res[head]);
        for (int i = 0; i < M; i++) {
            const int code = get4(nbr, i);
            if (code < 0) continue;
            const int n = code >> 2;
            bool seen = false;
            for (int j = 0; j < nk; j++) seen |= res[j] == n;
            if (seen) continue;
            if (G::conflict(cx, G::load(m, TV(m, n)), p)) {
                if (nk == A.cap) { A.count[qi] = -1; return; }
                res[nk++] = n;
            }
        }
    }
    if (A.compactOf)
        for (int j = 0; j < nk; j++) res[j] = A.compactOf[res[j]];
    A.count[qi] = nk;
}

// ---- scheduler (reference API: scheduler::{make_queue, find_placement}, /root/reference/src/scheduler.rs:6-55).
// The device path does not schedule with these (it reserves simplices with atomicMin, kernels.cuh); they exist so that
// callers of the reference's scheduler find the same functions, computed on the device from the same store.
// make_queue: footprint of a query = sorted unique neighbours-of-neighbours of its conflict region (scheduler.rs:14-24).
// The reference's ghost simplices (outside the super simplex) do not exist in this store: a hull facet contributes
// nothing, so footprints differ from the reference's only for queries whose 2-ring reaches the super simplex's hull.
template <int D> struct FootprintArgs {
    Mesh<D> m;
    const int *killed;      // [nq x kcap] conflict regions as simplex slots (locate_query_body with compactOf == nullptr)
    const int *kcount;      // per query: region size (<= 0: empty / error)
    const int *compactOf;   // simplex slot -> export index
    int *fp;                // [nq x fcap] out: footprint as sorted unique export indices
    int *fcount;            // per query: footprint size, -1 if it does not fit
    int kcap, fcap;
};
template <int D> VOR_HD void footprint_body(const FootprintArgs<D> &A, int qi) {
    constexpr int M = Dim<D>::M;
    const Mesh<D> &m = A.m;
    const int nk = A.kcount[qi];
    int *out = A.fp + (size_t)qi * A.fcap;
    int n = 0;
    if (nk <= 0) { A.fcount[qi] = nk < 0 ? -1 : 0; return; }
    // sorted insertion with binary search (footprints hold 100-300 ids)
    for (int j = 0; j < nk; j++) {
        const int4 na = TN(m, A.killed[(size_t)qi * A.kcap + j]);
        for (int ia = 0; ia < M; ia++) {
            const int ca = get4(na, ia);
            if (ca < 0) continue;
            const int4 nb = TN(m, ca >> 2);
            for (int ib = 0; ib < M; ib++) {
                const int cb = get4(nb, ib);
                if (cb < 0) continue;
                const int id = A.compactOf[cb >> 2];
                int lo = 0, hi = n;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (out[mid] < id) lo = mid + 1; else hi = mid; }
                if (lo < n && out[lo] == id) continue;
                if (n == A.fcap) { A.fcount[qi] = -1; return; }
                for (int x = n; x > lo; x--) out[x] = out[x - 1];
                out[lo] = id;
                n++;
            }
        }
    }
    A.fcount[qi] = n;
}
// find_placement (scheduler.rs:30-55): greedy round of every queue entry in queue order.  round[p] = 1 + the largest
// round among the previous occupants of p's footprint simplices (0 when p is the first occupant of all of them) --
// the same recurrence as the reference's occupancy lists, which only ever look at the LAST previous occupant.
// Inherently sequential in p (the reference's loop is serial too): one thread walks the queue.
struct PlacementArgs { const long long *off; const int *ids; int nq; int *last; unsigned long long *round; };
VOR_HD void placement_body(const PlacementArgs &A, int) {
    for (int p = 0; p < A.nq; p++) {
        int best = 0;
        for (long long x = A.off[p]; x < A.off[p + 1]; x++) { const int l = A.last[A.ids[x]]; best = l > best ? l : best; }
        best += 1;
        for (long long x = A.off[p]; x < A.off[p + 1]; x++) A.last[A.ids[x]] = best;
        A.round[p] = (unsigned long long)best;
    }
}

// ---- export of vertices: coordinates in reference id order + incident live simplices (Vertex.simplex,
// delaunay_tree.rs:20-24) as CSR of export indices.  Ids: super vertex k -> k, input point i -> idOffset + i (the
// same ids export_fill_body writes into `verts`).  Two passes like the edge list: count, scan (host), fill; rows sorted.
template <int D> struct IncidArgs {
    Mesh<D> m;
    const int *liveId;
    const int *inputIdx;
    int *deg;        // [nIds + 1] counts, then exclusive offsets
    int *cursor;     // [nIds]
    int *simps;      // CSR column array (export indices)
    int idOffset;
    int pass;
};
template <int D> VOR_HD void incid_body(const IncidArgs<D> &A, int c) {
    constexpr int M = Dim<D>::M;
    const int4 tv = TV(A.m, A.liveId[c]);
    for (int k = 0; k < M; k++) {
        const int v = get4(tv, k);
        const int id = v < A.m.nsuper ? v : A.idOffset + A.inputIdx[v];
        if (A.pass == 0) atomic_add_i(&A.deg[id], 1);
        else A.simps[(size_t)A.deg[id] + atomic_add_i(&A.cursor[id], 1)] = c;
    }
}
struct RowSortPlainArgs { const int *off; int *val; };
VOR_HD void row_sort_plain_body(const RowSortPlainArgs &A, int r) {
    const int lo = A.off[r], hi = A.off[r + 1];
    for (int i = lo + 1; i < hi; i++) {
        const int x = A.val[i];
        int j = i - 1;
        while (j >= lo && A.val[j] > x) { A.val[j + 1] = A.val[j]; j--; }
        A.val[j + 1] = x;
    }
}
template <int D> struct CoordArgs { const typename Dim<D>::Pt *pts; const int *inputIdx; double *out; int nsuper; int idOffset; };
VOR_HD void put_coords(double *o, const double4 &p) { o[0] = p.x; o[1] = p.y; o[2] = p.z; }
VOR_HD void put_coords(double *o, const double2 &p) { o[0] = p.x; o[1] = p.y; }
template <int D> VOR_HD void coord_body(const CoordArgs<D> &A, int j) {   // j-th real vertex (insertion order) -> row idOffset + input index
    const int v = A.nsuper + j;
    const int i = A.inputIdx[v];
    if (i >= 0) put_coords(A.out + (size_t)(A.idOffset + i) * D, A.pts[v]);
}

// ---- batched geometry entry points (reference API: geometry::{circumsphere, in_sphere})
template <int D> struct CircumBatchArgs { const double *verts; double *center; double *radius; };
VOR_HD void circum_batch_body(const CircumBatchArgs<3> &A, int i) {
    const double *v = A.verts + (size_t)i * 12;
    Geo<3>::Verts t;
    t.p0 = double4{v[0], v[1], v[2], 0}; t.p1 = double4{v[3], v[4], v[5], 0}; t.p2 = double4{v[6], v[7], v[8], 0}; t.p3 = double4{v[9], v[10], v[11], 0};
    circum_ref(t, A.center + (size_t)i * 3, A.radius + i);
}
VOR_HD void circum_batch_body(const CircumBatchArgs<2> &A, int i) {
    const double *v = A.verts + (size_t)i * 6;
    Geo<2>::Verts t;
    t.p0 = double2{v[0], v[1]}; t.p1 = double2{v[2], v[3]}; t.p2 = double2{v[4], v[5]};
    circum_ref(t, A.center + (size_t)i * 2, A.radius + i);
}
struct InSphereBatchArgs { const double *p; const double *c; const double *r; int *out; int dim; };
VOR_HD void in_sphere_batch_body(const InSphereBatchArgs &A, int i) {
    double dist = 0.0;
    for (int k = 0; k < A.dim; k++) {
        const double d = A.c[(size_t)i * A.dim + k] - A.p[(size_t)i * A.dim + k];
        dist += d * d;
    }
    A.out[i] = dist < A.r[i] * A.r[i];
}
// exact predicate batches (K1/K2 of SURVEY.md §2.3): rows of M+1 (insphere) or M (orient) points
struct PredBatchArgs { const double *rows; int *out; Counters *cnt; int kind; }; // kind: 0 orient2d 1 orient3d 2 incircle 3 insphere
VOR_HD void pred_batch_body(const PredBatchArgs &A, int i) {
    PredCtx cx{A.cnt};
    if (A.kind == 0) {
        const double *r = A.rows + (size_t)i * 6;
        A.out[i] = orient2d(cx, double2{r[0], r[1]}, double2{r[2], r[3]}, double2{r[4], r[5]});
    } else if (A.kind == 1) {
        const double *r = A.rows + (size_t)i * 12;
        A.out[i] = orient3d(cx, double4{r[0], r[1], r[2], 0}, double4{r[3], r[4], r[5], 0}, double4{r[6], r[7], r[8], 0}, double4{r[9], r[10], r[11], 0});
    } else if (A.kind == 2) {
        const double *r = A.rows + (size_t)i * 8;
        A.out[i] = incircle(cx, double2{r[0], r[1]}, double2{r[2], r[3]}, double2{r[4], r[5]}, double2{r[6], r[7]});
    } else {
        const double *r = A.rows + (size_t)i * 15;
        A.out[i] = insphere(cx, double4{r[0], r[1], r[2], 0}, double4{r[3], r[4], r[5], 0}, double4{r[6], r[7], r[8], 0}, double4{r[9], r[10], r[11], 0},
                            double4{r[12], r[13], r[14], 0});
    }
}

// cached-circumsphere filter (sphere.cuh) on packed rows of M simplex vertices + 1 query: out = +1 certainly strictly
// inside, -1 certainly not, 0 undecided; blocks (optional) = cx, cy, cz, rin2, rout2 as stored
struct SphereBatchArgs { const double *rows; int *out; float *blocks; SphereRef ref; int dim; };
VOR_HD void sphere_batch_body(const SphereBatchArgs &A, int i) {
    SphereBlk b;
    double qx, qy, qz = 0.0;
    if (A.dim == 3) {
        const double *r = A.rows + (size_t)i * 15;
        b = sphere_make(double4{r[0], r[1], r[2], 0}, double4{r[3], r[4], r[5], 0}, double4{r[6], r[7], r[8], 0}, double4{r[9], r[10], r[11], 0}, A.ref);
        qx = r[12] - A.ref.ox; qy = r[13] - A.ref.oy; qz = r[14] - A.ref.oz;
    } else {
        const double *r = A.rows + (size_t)i * 8;
        b = sphere_make(double2{r[0], r[1]}, double2{r[2], r[3]}, double2{r[4], r[5]}, A.ref);
        qx = r[6] - A.ref.ox; qy = r[7] - A.ref.oy;
    }
    A.out[i] = sphere_test(b.cx, b.cy, b.cz, b.rin2, b.rout2, qx, qy, qz);
    if (A.blocks) { float *o = A.blocks + (size_t)i * 5; o[0] = b.cx; o[1] = b.cy; o[2] = b.cz; o[3] = b.rin2; o[4] = b.rout2; }
}

} // namespace vor
