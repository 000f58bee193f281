// dd_stage.cuh -- double-double re-evaluation of the predicates, tried at the head of the exact functions
// (exact_int.cuh).  ON in the product build since the hot attempt kernel has no predicate code in its call tree (round 2):
// on the jittered-lattice workload 0.02 % of the in-sphere tests leave the FP64 filter, each exact call is ~30 us on one lane
// and a kernel cannot end before its slowest warp; the stage takes the attempt kernels of that workload from 68.8 to 52.7 ms
// per 5M points (117 -> 88 ms in all) and leaves the uniform run alone (50.9 vs 51.3 ms, tools/r2_exp31.sh).  In round 1, with
// ONE attempt kernel that had the predicates inside, the mere presence of this code changed ptxas' register allocation of the
// hot loop and cost the uniform 10M-point run 7-8 %, wherever it was called from -- it was off then (-DVOR_DD=0 still builds
// without it).  tests/emu compiles it in as well: predicates against fractions.Fraction, engine parity.
#pragma once
#include "vor_common.cuh"

#ifndef VOR_DD
#define VOR_DD 1
#endif

namespace vor {

// ---- double-double stage between the semi-static filter and the exact integers.
// The exact path costs ~30 us on one lane and a round's attempt kernel cannot end before its slowest warp: on the
// jittered lattice (0.02 % of the in-sphere tests fail the FP64 filter, ~90 per round) that tail cost 23 ms per 5M
// points.  Here the determinant is re-evaluated in double-double arithmetic (differences are exact as a pair; sloppy
// addition and FMA-based products, each with an error below 2^-104 of the MAGNITUDES of its operands), so the absolute
// error stays below 13 * 2^-104 * permanent; a result larger than 2^-90 * permanent is certified (factor 2^10 to
// spare), anything else -- and any input outside the range where no intermediate can overflow or lose its low word --
// goes on to the exact path.  fma() is an explicit FMA and is not affected by -fmad=false.
struct DD { double hi, lo; };
VOR_HD DD dd_two_sum(double a, double b) { const double s = a + b, bb = s - a; return DD{s, (a - (s - bb)) + (b - bb)}; }
VOR_HD DD dd_quick(double a, double b) { const double s = a + b; return DD{s, b - (s - a)}; }
VOR_HD DD dd_prod(double a, double b) { const double p = a * b; return DD{p, fma(a, b, -p)}; }
VOR_HD DD dd_add(const DD &a, const DD &b) { const DD s = dd_two_sum(a.hi, b.hi); return dd_quick(s.hi, s.lo + (a.lo + b.lo)); }
VOR_HD DD dd_neg(const DD &a) { return DD{-a.hi, -a.lo}; }
VOR_HD DD dd_sub(const DD &a, const DD &b) { return dd_add(a, dd_neg(b)); }
VOR_HD DD dd_mul(const DD &a, const DD &b) { DD p = dd_prod(a.hi, b.hi); p.lo += a.hi * b.lo + a.lo * b.hi; return dd_quick(p.hi, p.lo); }
VOR_HD DD dd_diff(double a, double b) { return dd_two_sum(a, -b); }   // exact
VOR_HD double dd_abs_hi(const DD &a) { return fabs(a.hi); }
// sign (+1 / -1) when certified, 0 = undecided
VOR_HD_NOINLINE int insphere_dd(double ax, double ay, double az, double bx, double by, double bz, double cx_, double cy_, double cz_, double dx,
                                double dy, double dz, double ex, double ey, double ez) {
    const DD aex = dd_diff(ax, ex), bex = dd_diff(bx, ex), cex = dd_diff(cx_, ex), dex = dd_diff(dx, ex);
    const DD aey = dd_diff(ay, ey), bey = dd_diff(by, ey), cey = dd_diff(cy_, ey), dey = dd_diff(dy, ey);
    const DD aez = dd_diff(az, ez), bez = dd_diff(bz, ez), cez = dd_diff(cz_, ez), dez = dd_diff(dz, ez);
    double mx = fmax(fmax(dd_abs_hi(aex), dd_abs_hi(bex)), fmax(dd_abs_hi(cex), dd_abs_hi(dex)));
    mx = fmax(mx, fmax(fmax(dd_abs_hi(aey), dd_abs_hi(bey)), fmax(dd_abs_hi(cey), dd_abs_hi(dey))));
    mx = fmax(mx, fmax(fmax(dd_abs_hi(aez), dd_abs_hi(bez)), fmax(dd_abs_hi(cez), dd_abs_hi(dez))));
    if (!(mx < 1e55)) return 0;                        // a degree-5 product could overflow (or NaN)
    const DD aexbey = dd_mul(aex, bey), bexaey = dd_mul(bex, aey), ab = dd_sub(aexbey, bexaey);
    const DD bexcey = dd_mul(bex, cey), cexbey = dd_mul(cex, bey), bc = dd_sub(bexcey, cexbey);
    const DD cexdey = dd_mul(cex, dey), dexcey = dd_mul(dex, cey), cd = dd_sub(cexdey, dexcey);
    const DD dexaey = dd_mul(dex, aey), aexdey = dd_mul(aex, dey), da = dd_sub(dexaey, aexdey);
    const DD aexcey = dd_mul(aex, cey), cexaey = dd_mul(cex, aey), ac = dd_sub(aexcey, cexaey);
    const DD bexdey = dd_mul(bex, dey), dexbey = dd_mul(dex, bey), bd = dd_sub(bexdey, dexbey);
    const DD abc = dd_add(dd_sub(dd_mul(aez, bc), dd_mul(bez, ac)), dd_mul(cez, ab));
    const DD bcd = dd_add(dd_sub(dd_mul(bez, cd), dd_mul(cez, bd)), dd_mul(dez, bc));
    const DD cda = dd_add(dd_add(dd_mul(cez, da), dd_mul(dez, ac)), dd_mul(aez, cd));
    const DD dab = dd_add(dd_add(dd_mul(dez, ab), dd_mul(aez, bd)), dd_mul(bez, da));
    const DD al = dd_add(dd_add(dd_mul(aex, aex), dd_mul(aey, aey)), dd_mul(aez, aez));
    const DD bl = dd_add(dd_add(dd_mul(bex, bex), dd_mul(bey, bey)), dd_mul(bez, bez));
    const DD cl = dd_add(dd_add(dd_mul(cex, cex), dd_mul(cey, cey)), dd_mul(cez, cez));
    const DD dl = dd_add(dd_add(dd_mul(dex, dex), dd_mul(dey, dey)), dd_mul(dez, dez));
    const DD det = dd_add(dd_sub(dd_mul(dl, abc), dd_mul(cl, dab)), dd_sub(dd_mul(bl, cda), dd_mul(al, bcd)));
    const double aezp = fabs(aez.hi), bezp = fabs(bez.hi), cezp = fabs(cez.hi), dezp = fabs(dez.hi);
    const double perm =
        ((fabs(cexdey.hi) + fabs(dexcey.hi)) * bezp + (fabs(dexbey.hi) + fabs(bexdey.hi)) * cezp + (fabs(bexcey.hi) + fabs(cexbey.hi)) * dezp) * al.hi +
        ((fabs(dexaey.hi) + fabs(aexdey.hi)) * cezp + (fabs(aexcey.hi) + fabs(cexaey.hi)) * dezp + (fabs(cexdey.hi) + fabs(dexcey.hi)) * aezp) * bl.hi +
        ((fabs(aexbey.hi) + fabs(bexaey.hi)) * dezp + (fabs(bexdey.hi) + fabs(dexbey.hi)) * aezp + (fabs(dexaey.hi) + fabs(aexdey.hi)) * bezp) * cl.hi +
        ((fabs(bexcey.hi) + fabs(cexbey.hi)) * aezp + (fabs(cexaey.hi) + fabs(aexcey.hi)) * bezp + (fabs(aexbey.hi) + fabs(bexaey.hi)) * cezp) * dl.hi;
    if (!(perm > 1e-200)) return 0;                    // low words could fall below the subnormal threshold
    const double bound = 8.077935669463161e-28 * perm; // 2^-90 * permanent
    if (det.hi > bound) return 1;
    if (-det.hi > bound) return -1;
    return 0;
}
VOR_HD_NOINLINE int orient3d_dd(double ax, double ay, double az, double bx, double by, double bz, double cx_, double cy_, double cz_, double dx,
                                double dy, double dz) {
    const DD adx = dd_diff(ax, dx), bdx = dd_diff(bx, dx), cdx = dd_diff(cx_, dx);
    const DD ady = dd_diff(ay, dy), bdy = dd_diff(by, dy), cdy = dd_diff(cy_, dy);
    const DD adz = dd_diff(az, dz), bdz = dd_diff(bz, dz), cdz = dd_diff(cz_, dz);
    double mx = fmax(fmax(dd_abs_hi(adx), dd_abs_hi(bdx)), dd_abs_hi(cdx));
    mx = fmax(mx, fmax(fmax(dd_abs_hi(ady), dd_abs_hi(bdy)), dd_abs_hi(cdy)));
    mx = fmax(mx, fmax(fmax(dd_abs_hi(adz), dd_abs_hi(bdz)), dd_abs_hi(cdz)));
    if (!(mx < 1e90)) return 0;                        // degree 3
    const DD bdxcdy = dd_mul(bdx, cdy), cdxbdy = dd_mul(cdx, bdy), cdxady = dd_mul(cdx, ady), adxcdy = dd_mul(adx, cdy);
    const DD adxbdy = dd_mul(adx, bdy), bdxady = dd_mul(bdx, ady);
    const DD det = dd_add(dd_add(dd_mul(adz, dd_sub(bdxcdy, cdxbdy)), dd_mul(bdz, dd_sub(cdxady, adxcdy))), dd_mul(cdz, dd_sub(adxbdy, bdxady)));
    const double perm = (fabs(bdxcdy.hi) + fabs(cdxbdy.hi)) * fabs(adz.hi) + (fabs(cdxady.hi) + fabs(adxcdy.hi)) * fabs(bdz.hi) +
                        (fabs(adxbdy.hi) + fabs(bdxady.hi)) * fabs(cdz.hi);
    if (!(perm > 1e-200)) return 0;
    const double bound = 8.077935669463161e-28 * perm; // 2^-90 * permanent (error of the evaluation < 8 * 2^-104 * permanent)
    if (det.hi > bound) return 1;
    if (-det.hi > bound) return -1;
    return 0;
}
VOR_HD_NOINLINE int incircle_dd(double ax, double ay, double bx, double by, double cx_, double cy_, double dx, double dy) {
    const DD adx = dd_diff(ax, dx), ady = dd_diff(ay, dy), bdx = dd_diff(bx, dx), bdy = dd_diff(by, dy), cdx = dd_diff(cx_, dx), cdy = dd_diff(cy_, dy);
    const double mx = fmax(fmax(fmax(dd_abs_hi(adx), dd_abs_hi(ady)), fmax(dd_abs_hi(bdx), dd_abs_hi(bdy))), fmax(dd_abs_hi(cdx), dd_abs_hi(cdy)));
    if (!(mx < 1e70)) return 0;                        // degree 4
    const DD bdxcdy = dd_mul(bdx, cdy), cdxbdy = dd_mul(cdx, bdy), cdxady = dd_mul(cdx, ady), adxcdy = dd_mul(adx, cdy);
    const DD adxbdy = dd_mul(adx, bdy), bdxady = dd_mul(bdx, ady);
    const DD al = dd_add(dd_mul(adx, adx), dd_mul(ady, ady)), bl = dd_add(dd_mul(bdx, bdx), dd_mul(bdy, bdy)), cl = dd_add(dd_mul(cdx, cdx), dd_mul(cdy, cdy));
    const DD det = dd_add(dd_add(dd_mul(al, dd_sub(bdxcdy, cdxbdy)), dd_mul(bl, dd_sub(cdxady, adxcdy))), dd_mul(cl, dd_sub(adxbdy, bdxady)));
    const double perm = (fabs(bdxcdy.hi) + fabs(cdxbdy.hi)) * al.hi + (fabs(cdxady.hi) + fabs(adxcdy.hi)) * bl.hi + (fabs(adxbdy.hi) + fabs(bdxady.hi)) * cl.hi;
    if (!(perm > 1e-200)) return 0;
    const double bound = 8.077935669463161e-28 * perm; // 2^-90 * permanent (error of the evaluation < 10 * 2^-104 * permanent)
    if (det.hi > bound) return 1;
    if (-det.hi > bound) return -1;
    return 0;
}


} // namespace vor
