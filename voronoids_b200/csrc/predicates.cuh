// predicates.cuh -- filtered exact predicates for the device engine (FP64 SIMT).
//
// Replaces the reference's float test `dist^2 < radius*radius` on a cached
// circumsphere (/root/reference/src/geometry.rs:91-97 with :24-56) by the sign
// of the in-sphere determinant: a semi-static filter (Shewchuk's stage-A error
// bound, permanent-based) decides almost every call in plain FP64; the rest go
// to exact_int.cuh.  orient2d/orient3d have no counterpart in the reference
// (SURVEY.md §0 D3): they drive the visibility walk and keep simplices
// positively oriented.
//
// Build with -fmad=false: the error bounds assume every operation is rounded
// individually, and it keeps device and host-emulation decisions identical.
//
// Conventions (same as oracle/predicates.c):
//   orient3d(a,b,c,d) = sign det[a-d; b-d; c-d]
//   insphere(a,b,c,d,e) > 0  <=>  e strictly inside sphere(a,b,c,d) when orient3d(a,b,c,d) > 0
//   orient2d(a,b,c) > 0 <=> counter-clockwise; incircle(a,b,c,d) > 0 <=> d strictly inside when ccw
#pragma once
#include "exact_int.cuh"

namespace vor {

constexpr double EPSH = 1.1102230246251565e-16; // 2^-53

// EX = false: a predicate that leaves the FP64 filter does NOT call the exact path; it sets `failed` and returns 0, and
// the caller gives the work item to a kernel instantiated with EX = true.  Measured: the exact-integer code in the
// call tree of the attempt kernel costs its hot loop 10 % (95 vs 85 ms per 10M points) although it is never executed
// on such inputs -- so the hot kernel is built without it and a slow twin picks up the flagged points (engine.cuh).
template <bool EX> struct PredCtxT {
    Counters *cnt;
    bool failed = false;
    static constexpr bool exact = EX;
};
using PredCtx = PredCtxT<true>;

// exact_calls counts the predicates that left the FP64 filter, whether the double-double stage inside *_exact
// (dd_stage.cuh) or the integers settled them; exact_zero counts true zeros.
template <class CX> VOR_HD int finish_exact(CX &cx, int s, int range_err) {
    atomic_add_ull(&cx.cnt->exact_calls, 1ULL);
    if (range_err) set_err(cx.cnt, ERR_RANGE);
    else if (s == 0) atomic_add_ull(&cx.cnt->exact_zero, 1ULL);
    return s;
}

template <class CX> VOR_HD int orient2d(CX &cx, const double2 &a, const double2 &b, const double2 &c) {
    const double l = (a.x - c.x) * (b.y - c.y);
    const double r = (a.y - c.y) * (b.x - c.x);
    const double det = l - r;
    const double bound = (3.0 + 16.0 * EPSH) * EPSH * (fabs(l) + fabs(r));
    if (det > bound) return 1;
    if (-det > bound) return -1;
    if constexpr (!CX::exact) { cx.failed = true; return 0; }
    const double A[2] = {a.x, a.y}, B[2] = {b.x, b.y}, C[2] = {c.x, c.y};
    int re = 0;
    const int s = orient2d_exact(A, B, C, &re);
    return finish_exact(cx, s, re);
}

template <class CX> VOR_HD int orient3d(CX &cx, const double4 &a, const double4 &b, const double4 &c, const double4 &d) {
    const double adx = a.x - d.x, bdx = b.x - d.x, cdx = c.x - d.x;
    const double ady = a.y - d.y, bdy = b.y - d.y, cdy = c.y - d.y;
    const double adz = a.z - d.z, bdz = b.z - d.z, cdz = c.z - d.z;
    const double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy;
    const double cdxady = cdx * ady, adxcdy = adx * cdy;
    const double adxbdy = adx * bdy, bdxady = bdx * ady;
    const double det = adz * (bdxcdy - cdxbdy) + bdz * (cdxady - adxcdy) + cdz * (adxbdy - bdxady);
    const double perm = (fabs(bdxcdy) + fabs(cdxbdy)) * fabs(adz) + (fabs(cdxady) + fabs(adxcdy)) * fabs(bdz) +
                        (fabs(adxbdy) + fabs(bdxady)) * fabs(cdz);
    const double bound = (7.0 + 56.0 * EPSH) * EPSH * perm;
    if (det > bound) return 1;
    if (-det > bound) return -1;
    if constexpr (!CX::exact) { cx.failed = true; return 0; }
    const double A[3] = {a.x, a.y, a.z}, B[3] = {b.x, b.y, b.z}, C[3] = {c.x, c.y, c.z}, D[3] = {d.x, d.y, d.z};
    int re = 0;
    const int s = orient3d_exact(A, B, C, D, &re);
    return finish_exact(cx, s, re);
}

// (noinline slow paths take everything BY VALUE: a reference parameter would force the caller to keep the operands in
// local memory and store them there before every test, not only in the rare branch that calls this)
VOR_HD_NOINLINE int incircle_slow(Counters *cnt, double ax, double ay, double bx, double by, double cx_, double cy_, double dx, double dy) {
    PredCtx cx{cnt};
    const double2 a{ax, ay}, b{bx, by}, c{cx_, cy_}, d{dx, dy};
    const double adx = a.x - d.x, ady = a.y - d.y;
    const double bdx = b.x - d.x, bdy = b.y - d.y;
    const double cdx = c.x - d.x, cdy = c.y - d.y;
    const double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy;
    const double cdxady = cdx * ady, adxcdy = adx * cdy;
    const double adxbdy = adx * bdy, bdxady = bdx * ady;
    const double al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;
    const double det = al * (bdxcdy - cdxbdy) + bl * (cdxady - adxcdy) + cl * (adxbdy - bdxady);
    const double perm = (fabs(bdxcdy) + fabs(cdxbdy)) * al + (fabs(cdxady) + fabs(adxcdy)) * bl +
                        (fabs(adxbdy) + fabs(bdxady)) * cl;
    const double bound = (10.0 + 96.0 * EPSH) * EPSH * perm;
    if (det > bound) return 1;
    if (-det > bound) return -1;
    const double A[2] = {a.x, a.y}, B[2] = {b.x, b.y}, C[2] = {c.x, c.y}, D[2] = {d.x, d.y};
    int re = 0;
    const int s = incircle_exact(A, B, C, D, &re);
    return finish_exact(cx, s, re);
}

// Static filter in front of the semi-static one: with X, Y = largest |x|, |y| difference, every term of the permanent
// is bounded by the same expression in X, Y (rounding is monotone), so permanent <= 6 fl(XY) fl(X^2 + Y^2) (1+eps)^4 and
// |det| > 61 eps fl(XY) fl(X^2+Y^2) >= (10 + 96 eps) eps permanent certifies the sign.  It keeps 2 extra values
// live instead of the 6 products of the permanent.
template <class CX> VOR_HD int incircle(CX &cx, const double2 &a, const double2 &b, const double2 &c, const double2 &d) {
    const double adx = a.x - d.x, ady = a.y - d.y;
    const double bdx = b.x - d.x, bdy = b.y - d.y;
    const double cdx = c.x - d.x, cdy = c.y - d.y;
    const double mx = fmax(fmax(fabs(adx), fabs(bdx)), fabs(cdx));
    const double my = fmax(fmax(fabs(ady), fabs(bdy)), fabs(cdy));
    const double al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;
    const double det = al * (bdx * cdy - cdx * bdy) + bl * (cdx * ady - adx * cdy) + cl * (adx * bdy - bdx * ady);
    const double bound = ((61.0 * EPSH) * (mx * my)) * (mx * mx + my * my);
    if (det > bound) return 1;
    if (-det > bound) return -1;
    if constexpr (!CX::exact) { cx.failed = true; return 0; }
    return incircle_slow(cx.cnt, a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y);
}

VOR_HD_NOINLINE int insphere_slow(Counters *cnt, double ax, double ay, double az, double bx, double by, double bz, double cx_, double cy_, double cz_,
                                  double dx, double dy, double dz, double ex, double ey, double ez) {
    PredCtx cx{cnt};
    double4 a, b, c, d, e;
    a.x = ax; a.y = ay; a.z = az; a.w = 0.0; b.x = bx; b.y = by; b.z = bz; b.w = 0.0; c.x = cx_; c.y = cy_; c.z = cz_; c.w = 0.0;
    d.x = dx; d.y = dy; d.z = dz; d.w = 0.0; e.x = ex; e.y = ey; e.z = ez; e.w = 0.0;
    const double aex = a.x - e.x, bex = b.x - e.x, cex = c.x - e.x, dex = d.x - e.x;
    const double aey = a.y - e.y, bey = b.y - e.y, cey = c.y - e.y, dey = d.y - e.y;
    const double aez = a.z - e.z, bez = b.z - e.z, cez = c.z - e.z, dez = d.z - e.z;
    const double aexbey = aex * bey, bexaey = bex * aey, ab = aexbey - bexaey;
    const double bexcey = bex * cey, cexbey = cex * bey, bc = bexcey - cexbey;
    const double cexdey = cex * dey, dexcey = dex * cey, cd = cexdey - dexcey;
    const double dexaey = dex * aey, aexdey = aex * dey, da = dexaey - aexdey;
    const double aexcey = aex * cey, cexaey = cex * aey, ac = aexcey - cexaey;
    const double bexdey = bex * dey, dexbey = dex * bey, bd = bexdey - dexbey;
    const double abc = aez * bc - bez * ac + cez * ab;
    const double bcd = bez * cd - cez * bd + dez * bc;
    const double cda = cez * da + dez * ac + aez * cd;
    const double dab = dez * ab + aez * bd + bez * da;
    const double al = aex * aex + aey * aey + aez * aez;
    const double bl = bex * bex + bey * bey + bez * bez;
    const double cl = cex * cex + cey * cey + cez * cez;
    const double dl = dex * dex + dey * dey + dez * dez;
    const double det = (dl * abc - cl * dab) + (bl * cda - al * bcd);
    const double aezp = fabs(aez), bezp = fabs(bez), cezp = fabs(cez), dezp = fabs(dez);
    const double perm =
        ((fabs(cexdey) + fabs(dexcey)) * bezp + (fabs(dexbey) + fabs(bexdey)) * cezp + (fabs(bexcey) + fabs(cexbey)) * dezp) * al +
        ((fabs(dexaey) + fabs(aexdey)) * cezp + (fabs(aexcey) + fabs(cexaey)) * dezp + (fabs(cexdey) + fabs(dexcey)) * aezp) * bl +
        ((fabs(aexbey) + fabs(bexaey)) * dezp + (fabs(bexdey) + fabs(dexbey)) * aezp + (fabs(dexaey) + fabs(aexdey)) * bezp) * cl +
        ((fabs(bexcey) + fabs(cexbey)) * aezp + (fabs(cexaey) + fabs(aexcey)) * bezp + (fabs(aexbey) + fabs(bexaey)) * cezp) * dl;
    const double bound = (16.0 + 224.0 * EPSH) * EPSH * perm;
    if (det > bound) return 1;
    if (-det > bound) return -1;
    const double A[3] = {a.x, a.y, a.z}, B[3] = {b.x, b.y, b.z}, C[3] = {c.x, c.y, c.z}, D[3] = {d.x, d.y, d.z}, E[3] = {e.x, e.y, e.z};
    int re = 0;
    const int s = insphere_exact(A, B, C, D, E, &re);
    return finish_exact(cx, s, re);
}

template <class CX> VOR_HD int insphere_semi(CX &cx, const double4 &a, const double4 &b, const double4 &c, const double4 &d, const double4 &e) {
    const double aex = a.x - e.x, bex = b.x - e.x, cex = c.x - e.x, dex = d.x - e.x;
    const double aey = a.y - e.y, bey = b.y - e.y, cey = c.y - e.y, dey = d.y - e.y;
    const double aez = a.z - e.z, bez = b.z - e.z, cez = c.z - e.z, dez = d.z - e.z;
    const double aexbey = aex * bey, bexaey = bex * aey, ab = aexbey - bexaey;
    const double bexcey = bex * cey, cexbey = cex * bey, bc = bexcey - cexbey;
    const double cexdey = cex * dey, dexcey = dex * cey, cd = cexdey - dexcey;
    const double dexaey = dex * aey, aexdey = aex * dey, da = dexaey - aexdey;
    const double aexcey = aex * cey, cexaey = cex * aey, ac = aexcey - cexaey;
    const double bexdey = bex * dey, dexbey = dex * bey, bd = bexdey - dexbey;
    const double abc = aez * bc - bez * ac + cez * ab;
    const double bcd = bez * cd - cez * bd + dez * bc;
    const double cda = cez * da + dez * ac + aez * cd;
    const double dab = dez * ab + aez * bd + bez * da;
    const double al = aex * aex + aey * aey + aez * aez;
    const double bl = bex * bex + bey * bey + bez * bez;
    const double cl = cex * cex + cey * cey + cez * cez;
    const double dl = dex * dex + dey * dey + dez * dez;
    const double det = (dl * abc - cl * dab) + (bl * cda - al * bcd);
    const double aezp = fabs(aez), bezp = fabs(bez), cezp = fabs(cez), dezp = fabs(dez);
    const double perm =
        ((fabs(cexdey) + fabs(dexcey)) * bezp + (fabs(dexbey) + fabs(bexdey)) * cezp + (fabs(bexcey) + fabs(cexbey)) * dezp) * al +
        ((fabs(dexaey) + fabs(aexdey)) * cezp + (fabs(aexcey) + fabs(cexaey)) * dezp + (fabs(cexdey) + fabs(dexcey)) * aezp) * bl +
        ((fabs(aexbey) + fabs(bexaey)) * dezp + (fabs(bexdey) + fabs(dexbey)) * aezp + (fabs(dexaey) + fabs(aexdey)) * bezp) * cl +
        ((fabs(bexcey) + fabs(cexbey)) * aezp + (fabs(cexaey) + fabs(aexcey)) * bezp + (fabs(aexbey) + fabs(bexaey)) * cezp) * dl;
    const double bound = (16.0 + 224.0 * EPSH) * EPSH * perm;
    if (det > bound) return 1;
    if (-det > bound) return -1;
    if constexpr (!CX::exact) { cx.failed = true; return 0; }
    const double A[3] = {a.x, a.y, a.z}, B[3] = {b.x, b.y, b.z}, C[3] = {c.x, c.y, c.z}, D[3] = {d.x, d.y, d.z}, E[3] = {e.x, e.y, e.z};
    int re = 0;
    const int s = insphere_exact(A, B, C, D, E, &re);
    return finish_exact(cx, s, re);
}

// Static filter in front of the semi-static one (same determinant, same operation order).  With X, Y, Z = largest
// |x|, |y|, |z| difference and L = fl(X^2 + Y^2 + Z^2): every |product| of the permanent is <= fl(XY), every lift <= L
// (rounding is monotone), so permanent <= 24 fl(fl(XY) Z) L (1+eps)^5, and
// |det| > 385 eps fl(fl(XY) Z) L >= (16 + 224 eps) eps permanent certifies the sign.  The permanent's 12 products do
// not have to stay live next to the determinant (the semi-static form spilled ~45 registers per test at the 64
// registers the attempt kernel runs with; ncu: 7x more local than global requests).
#ifndef VOR_STATIC_FILTER
#define VOR_STATIC_FILTER 0   // measured on the 10M-point run: attempt kernel 98.7 ms with it, 96.2 ms without
#endif
template <class CX> VOR_HD int insphere(CX &cx, const double4 &a, const double4 &b, const double4 &c, const double4 &d, const double4 &e) {
    if (!VOR_STATIC_FILTER) return insphere_semi(cx, a, b, c, d, e);
    const double aex = a.x - e.x, bex = b.x - e.x, cex = c.x - e.x, dex = d.x - e.x;
    const double aey = a.y - e.y, bey = b.y - e.y, cey = c.y - e.y, dey = d.y - e.y;
    const double aez = a.z - e.z, bez = b.z - e.z, cez = c.z - e.z, dez = d.z - e.z;
    const double mx = fmax(fmax(fabs(aex), fabs(bex)), fmax(fabs(cex), fabs(dex)));
    const double my = fmax(fmax(fabs(aey), fabs(bey)), fmax(fabs(cey), fabs(dey)));
    const double mz = fmax(fmax(fabs(aez), fabs(bez)), fmax(fabs(cez), fabs(dez)));
    const double al = aex * aex + aey * aey + aez * aez;
    const double bl = bex * bex + bey * bey + bez * bez;
    const double cl = cex * cex + cey * cey + cez * cez;
    const double dl = dex * dex + dey * dey + dez * dez;
    const double ab = aex * bey - bex * aey;
    const double bc = bex * cey - cex * bey;
    const double cd = cex * dey - dex * cey;
    const double da = dex * aey - aex * dey;
    const double ac = aex * cey - cex * aey;
    const double bd = bex * dey - dex * bey;
    const double abc = aez * bc - bez * ac + cez * ab;
    const double bcd = bez * cd - cez * bd + dez * bc;
    const double cda = cez * da + dez * ac + aez * cd;
    const double dab = dez * ab + aez * bd + bez * da;
    const double det = (dl * abc - cl * dab) + (bl * cda - al * bcd);
    const double bound = ((385.0 * EPSH) * ((mx * my) * mz)) * (mx * mx + my * my + mz * mz);
    if (det > bound) return 1;
    if (-det > bound) return -1;
    if constexpr (!CX::exact) { cx.failed = true; return 0; }
    return insphere_slow(cx.cnt, a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z, d.x, d.y, d.z, e.x, e.y, e.z);
}

} // namespace vor
