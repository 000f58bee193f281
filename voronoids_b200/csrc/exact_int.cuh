// exact_int.cuh -- exact sign of the orientation / in-sphere determinants by
// scaled multi-word integer arithmetic (device + host-emulation).
//
// Replaces nothing in the reference: the reference has no exact predicate, it
// compares against a cached float circumsphere (/root/reference/src/geometry.rs:91-97).
// north_star asks for "a static error filter and an exact fallback"; this file
// is the fallback, predicates.cuh holds the filter.
//
// Method: every input coordinate is a double m*2^e.  With emin the smallest
// exponent among the (at most 15) coordinates of one predicate call, X = x*2^-emin
// is an integer of at most B = emax-emin bits.  The determinant is then an
// integer polynomial evaluated without rounding in sign-magnitude limbs
// (32-bit limbs, schoolbook multiply).  Limb budgets below are sized for
// B <= 318 bits of coordinate dynamic range (e.g. every |x| in [2^-265, 2^53));
// wider ranges raise ERR_RANGE instead of returning a wrong sign.
//
// The CPU oracle uses a different exact method (floating-point expansions,
// oracle/predicates.c); tests pin both against Python fractions.Fraction.
#pragma once
#include "vor_common.cuh"
#include "dd_stage.cuh"

namespace vor {

constexpr int BIG_CW = 10;                 // limbs of a translated coordinate (320 bits)
constexpr int BIG_RANGE_BITS = 32 * BIG_CW - 2;

template <int W> struct Big {
    int sign;        // -1, 0, +1
    int len;         // limbs in use; d[len-1] != 0 when sign != 0
    uint32_t d[W];
};

VOR_HD int mag_cmp(const uint32_t *a, int la, const uint32_t *b, int lb) {
    if (la != lb) return la > lb ? 1 : -1;
    for (int i = la - 1; i >= 0; i--)
        if (a[i] != b[i]) return a[i] > b[i] ? 1 : -1;
    return 0;
}
// r = a + b (magnitudes), returns length
VOR_HD int mag_add(uint32_t *r, const uint32_t *a, int la, const uint32_t *b, int lb) {
    if (la < lb) { const uint32_t *t = a; a = b; b = t; int tl = la; la = lb; lb = tl; }
    uint64_t c = 0;
    int i = 0;
    for (; i < lb; i++) { c += (uint64_t)a[i] + b[i]; r[i] = (uint32_t)c; c >>= 32; }
    for (; i < la; i++) { c += a[i]; r[i] = (uint32_t)c; c >>= 32; }
    if (c) r[i++] = (uint32_t)c;
    return i;
}
// r = a - b (magnitudes, a >= b), returns normalised length
VOR_HD int mag_sub(uint32_t *r, const uint32_t *a, int la, const uint32_t *b, int lb) {
    int64_t br = 0;
    int i = 0;
    for (; i < lb; i++) { int64_t v = (int64_t)a[i] - b[i] + br; r[i] = (uint32_t)v; br = v >> 32; }
    for (; i < la; i++) { int64_t v = (int64_t)a[i] + br; r[i] = (uint32_t)v; br = v >> 32; }
    while (i > 0 && r[i - 1] == 0) i--;
    return i;
}
// r = a * b (magnitudes), r must not alias; returns normalised length
VOR_HD int mag_mul(uint32_t *r, const uint32_t *a, int la, const uint32_t *b, int lb) {
    if (la == 0 || lb == 0) return 0;
    for (int i = 0; i < la + lb; i++) r[i] = 0;
    for (int i = 0; i < la; i++) {
        uint64_t c = 0;
        const uint64_t ai = a[i];
        for (int j = 0; j < lb; j++) {
            c += ai * b[j] + r[i + j];
            r[i + j] = (uint32_t)c;
            c >>= 32;
        }
        r[i + lb] = (uint32_t)c;
    }
    int l = la + lb;
    while (l > 0 && r[l - 1] == 0) l--;
    return l;
}

// r = a + s*b  (s = +1 / -1)
template <int WO, int WA, int WB>
VOR_HD void big_addsub(Big<WO> &r, const Big<WA> &a, const Big<WB> &b, int s) {
    const int bs = b.sign * s;
    if (bs == 0) { r.sign = a.sign; r.len = a.len; for (int i = 0; i < a.len; i++) r.d[i] = a.d[i]; return; }
    if (a.sign == 0) { r.sign = bs; r.len = b.len; for (int i = 0; i < b.len; i++) r.d[i] = b.d[i]; return; }
    if (a.sign == bs) {
        r.len = mag_add(r.d, a.d, a.len, b.d, b.len);
        r.sign = a.sign;
        return;
    }
    const int c = mag_cmp(a.d, a.len, b.d, b.len);
    if (c == 0) { r.sign = 0; r.len = 0; return; }
    if (c > 0) { r.len = mag_sub(r.d, a.d, a.len, b.d, b.len); r.sign = a.sign; }
    else { r.len = mag_sub(r.d, b.d, b.len, a.d, a.len); r.sign = bs; }
}
template <int WO, int WA, int WB>
VOR_HD void big_mul(Big<WO> &r, const Big<WA> &a, const Big<WB> &b) {
    r.sign = a.sign * b.sign;
    r.len = r.sign ? mag_mul(r.d, a.d, a.len, b.d, b.len) : 0;
}

// ---- double -> (sign, 53-bit mantissa with trailing zeros stripped, exponent)
struct Dec { int sign; int e; uint64_t m; };
VOR_HD Dec decompose(double x) {
    union { double f; uint64_t u; } cv;
    cv.f = x;
    Dec r;
    const uint64_t frac = cv.u & 0xFFFFFFFFFFFFFULL;
    const int ex = (int)((cv.u >> 52) & 0x7ff);
    r.sign = (cv.u >> 63) ? -1 : 1;
    if (ex == 0) { r.m = frac; r.e = -1074; }
    else { r.m = frac | (1ULL << 52); r.e = ex - 1075; }
    if (r.m == 0) { r.sign = 0; r.e = 0; return r; }
#ifdef __CUDA_ARCH__
    const int tz = __ffsll((long long)r.m) - 1;
#else
    const int tz = __builtin_ctzll(r.m);
#endif
    r.m >>= tz;
    r.e += tz;
    return r;
}
VOR_HD int bitlen64(uint64_t m) {
#ifdef __CUDA_ARCH__
    return 64 - __clzll((long long)m);
#else
    return m ? 64 - __builtin_clzll(m) : 0;
#endif
}

struct ExpRange { int emin, emax; };
VOR_HD void range_init(ExpRange &r) { r.emin = 1 << 30; r.emax = -(1 << 30); }
VOR_HD void range_add(ExpRange &r, double x) {
    const Dec d = decompose(x);
    if (d.sign == 0) return;
    if (d.e < r.emin) r.emin = d.e;
    const int top = d.e + bitlen64(d.m);
    if (top > r.emax) r.emax = top;
}

// X = x * 2^-emin as a Big (requires the range check to have passed)
template <int W> VOR_HD void big_from(Big<W> &r, double x, int emin) {
    const Dec d = decompose(x);
    r.sign = d.sign;
    r.len = 0;
    if (d.sign == 0) return;
    const int s = d.e - emin;
    const int li = s >> 5, sh = s & 31;
    for (int i = 0; i < li; i++) r.d[i] = 0;
    // m << sh spans up to 3 limbs (53 + 31 bits)
    const uint64_t lo = d.m << sh;
    const uint32_t hi = sh ? (uint32_t)(d.m >> (64 - sh)) : 0u;
    // limbs at or beyond W are zero once the range check has passed
    r.d[li] = (uint32_t)lo;
    if (li + 1 < W) r.d[li + 1] = (uint32_t)(lo >> 32);
    if (li + 2 < W) r.d[li + 2] = hi;
    int l = li + 3 < W ? li + 3 : W;
    while (l > 0 && r.d[l - 1] == 0) l--;
    r.len = l;
}
// r = X(a) - X(b)
template <int W> VOR_HD void big_diff(Big<W> &r, double a, double b, int emin) {
    Big<W> x, y;
    big_from(x, a, emin);
    big_from(y, b, emin);
    big_addsub(r, x, y, -1);
}
// r = a*b - c*d
template <int WO, int WI> VOR_HD void big_minor2(Big<WO> &r, const Big<WI> &a, const Big<WI> &b, const Big<WI> &c, const Big<WI> &d) {
    Big<WO> t1, t2;
    big_mul(t1, a, b);
    big_mul(t2, c, d);
    big_addsub(r, t1, t2, -1);
}

constexpr int W2 = 2 * BIG_CW + 1;   // 21: 2x2 minors, lifts
constexpr int W3 = 3 * BIG_CW + 1;   // 31: 3x3 minors
constexpr int W4 = 4 * BIG_CW + 3;   // 43: incircle determinant (mag_mul zero-fills la+lb limbs)
constexpr int W5 = 5 * BIG_CW + 3;   // 53: insphere determinant

// Each exact routine returns the sign; *range_err is set when the inputs exceed the limb budget.

VOR_HD_NOINLINE int orient2d_exact(const double *a, const double *b, const double *c, int *range_err) {
    ExpRange rg; range_init(rg);
    for (int k = 0; k < 2; k++) { range_add(rg, a[k]); range_add(rg, b[k]); range_add(rg, c[k]); }
    if (rg.emax < rg.emin) return 0;
    if (rg.emax - rg.emin > BIG_RANGE_BITS) { *range_err = 1; return 0; }
    Big<BIG_CW> acx, acy, bcx, bcy;
    big_diff(acx, a[0], c[0], rg.emin); big_diff(acy, a[1], c[1], rg.emin);
    big_diff(bcx, b[0], c[0], rg.emin); big_diff(bcy, b[1], c[1], rg.emin);
    Big<W2> det;
    big_minor2(det, acx, bcy, acy, bcx);
    return det.sign;
}

VOR_HD_NOINLINE int orient3d_exact(const double *a, const double *b, const double *c, const double *d, int *range_err) {
    if (VOR_DD) { const int sdd = orient3d_dd(a[0], a[1], a[2], b[0], b[1], b[2], c[0], c[1], c[2], d[0], d[1], d[2]); if (sdd != 0) return sdd; }
    ExpRange rg; range_init(rg);
    for (int k = 0; k < 3; k++) { range_add(rg, a[k]); range_add(rg, b[k]); range_add(rg, c[k]); range_add(rg, d[k]); }
    if (rg.emax < rg.emin) return 0;
    if (rg.emax - rg.emin > BIG_RANGE_BITS) { *range_err = 1; return 0; }
    Big<BIG_CW> ad[3], bd[3], cd[3];
    for (int k = 0; k < 3; k++) {
        big_diff(ad[k], a[k], d[k], rg.emin);
        big_diff(bd[k], b[k], d[k], rg.emin);
        big_diff(cd[k], c[k], d[k], rg.emin);
    }
    // det = adz*(bdx*cdy - cdx*bdy) + bdz*(cdx*ady - adx*cdy) + cdz*(adx*bdy - bdx*ady)
    Big<W2> m;
    Big<W3> t, acc, acc2;
    big_minor2(m, bd[0], cd[1], cd[0], bd[1]);
    big_mul(acc, ad[2], m);
    big_minor2(m, cd[0], ad[1], ad[0], cd[1]);
    big_mul(t, bd[2], m);
    big_addsub(acc2, acc, t, 1);
    big_minor2(m, ad[0], bd[1], bd[0], ad[1]);
    big_mul(t, cd[2], m);
    big_addsub(acc, acc2, t, 1);
    return acc.sign;
}

VOR_HD_NOINLINE int incircle_exact(const double *a, const double *b, const double *c, const double *d, int *range_err) {
    if (VOR_DD) { const int sdd = incircle_dd(a[0], a[1], b[0], b[1], c[0], c[1], d[0], d[1]); if (sdd != 0) return sdd; }
    ExpRange rg; range_init(rg);
    for (int k = 0; k < 2; k++) { range_add(rg, a[k]); range_add(rg, b[k]); range_add(rg, c[k]); range_add(rg, d[k]); }
    if (rg.emax < rg.emin) return 0;
    if (rg.emax - rg.emin > BIG_RANGE_BITS) { *range_err = 1; return 0; }
    Big<BIG_CW> ad[2], bd[2], cd[2];
    for (int k = 0; k < 2; k++) {
        big_diff(ad[k], a[k], d[k], rg.emin);
        big_diff(bd[k], b[k], d[k], rg.emin);
        big_diff(cd[k], c[k], d[k], rg.emin);
    }
    Big<W2> m, l, s1, s2;
    Big<W4> t, acc, acc2;
    // det = alift*(bdx*cdy - cdx*bdy) + blift*(cdx*ady - adx*cdy) + clift*(adx*bdy - bdx*ady)
    big_mul(s1, ad[0], ad[0]); big_mul(s2, ad[1], ad[1]); big_addsub(l, s1, s2, 1);
    big_minor2(m, bd[0], cd[1], cd[0], bd[1]);
    big_mul(acc, l, m);
    big_mul(s1, bd[0], bd[0]); big_mul(s2, bd[1], bd[1]); big_addsub(l, s1, s2, 1);
    big_minor2(m, cd[0], ad[1], ad[0], cd[1]);
    big_mul(t, l, m);
    big_addsub(acc2, acc, t, 1);
    big_mul(s1, cd[0], cd[0]); big_mul(s2, cd[1], cd[1]); big_addsub(l, s1, s2, 1);
    big_minor2(m, ad[0], bd[1], bd[0], ad[1]);
    big_mul(t, l, m);
    big_addsub(acc, acc2, t, 1);
    return acc.sign;
}

template <int W> VOR_HD void big_lift3(Big<W> &l, const Big<BIG_CW> *v) {
    Big<W> s1, s2, s3;
    big_mul(s1, v[0], v[0]);
    big_mul(s2, v[1], v[1]);
    big_addsub(s3, s1, s2, 1);
    big_mul(s1, v[2], v[2]);
    big_addsub(l, s3, s1, 1);
}
// r = x*m1 + s2*y*m2 + s3*z*m3   (3x3 cofactor combination)
VOR_HD void big_comb3(Big<W3> &r, const Big<BIG_CW> &x, const Big<W2> &m1, int s2, const Big<BIG_CW> &y, const Big<W2> &m2,
                      int s3, const Big<BIG_CW> &z, const Big<W2> &m3) {
    Big<W3> t, u, v;
    big_mul(t, x, m1);
    big_mul(u, y, m2);
    big_addsub(v, t, u, s2);
    big_mul(t, z, m3);
    big_addsub(r, v, t, s3);
}

VOR_HD_NOINLINE int insphere_exact(const double *a, const double *b, const double *c, const double *d, const double *e, int *range_err) {
    if (VOR_DD) {
        const int sdd = insphere_dd(a[0], a[1], a[2], b[0], b[1], b[2], c[0], c[1], c[2], d[0], d[1], d[2], e[0], e[1], e[2]);
        if (sdd != 0) return sdd;
    }
    ExpRange rg; range_init(rg);
    for (int k = 0; k < 3; k++) { range_add(rg, a[k]); range_add(rg, b[k]); range_add(rg, c[k]); range_add(rg, d[k]); range_add(rg, e[k]); }
    if (rg.emax < rg.emin) return 0;
    if (rg.emax - rg.emin > BIG_RANGE_BITS) { *range_err = 1; return 0; }
    Big<BIG_CW> ae[3], be[3], ce[3], de[3];
    for (int k = 0; k < 3; k++) {
        big_diff(ae[k], a[k], e[k], rg.emin);
        big_diff(be[k], b[k], e[k], rg.emin);
        big_diff(ce[k], c[k], e[k], rg.emin);
        big_diff(de[k], d[k], e[k], rg.emin);
    }
    Big<W2> ab, bc, cd, da, ac, bd;
    big_minor2(ab, ae[0], be[1], be[0], ae[1]);
    big_minor2(bc, be[0], ce[1], ce[0], be[1]);
    big_minor2(cd, ce[0], de[1], de[0], ce[1]);
    big_minor2(da, de[0], ae[1], ae[0], de[1]);
    big_minor2(ac, ae[0], ce[1], ce[0], ae[1]);
    big_minor2(bd, be[0], de[1], de[0], be[1]);
    Big<W3> m3;
    Big<W2> lift;
    Big<W5> t, acc, acc2;
    // det = dlift*abc - clift*dab + blift*cda - alift*bcd
    big_comb3(m3, ae[2], bc, -1, be[2], ac, 1, ce[2], ab);      // abc = aez*bc - bez*ac + cez*ab
    big_lift3(lift, de);
    big_mul(acc, lift, m3);
    big_comb3(m3, de[2], ab, 1, ae[2], bd, 1, be[2], da);       // dab = dez*ab + aez*bd + bez*da
    big_lift3(lift, ce);
    big_mul(t, lift, m3);
    big_addsub(acc2, acc, t, -1);
    big_comb3(m3, ce[2], da, 1, de[2], ac, 1, ae[2], cd);       // cda = cez*da + dez*ac + aez*cd
    big_lift3(lift, be);
    big_mul(t, lift, m3);
    big_addsub(acc, acc2, t, 1);
    big_comb3(m3, be[2], cd, -1, ce[2], bd, 1, de[2], bc);      // bcd = bez*cd - cez*bd + dez*bc
    big_lift3(lift, ae);
    big_mul(t, lift, m3);
    big_addsub(acc2, acc, t, -1);
    return acc2.sign;
}

} // namespace vor
