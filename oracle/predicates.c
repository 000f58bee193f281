/*
 * oracle/predicates.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Exact geometric predicates for the CPU oracle (orient2d, orient3d, incircle,
 * insphere).  A semi-static floating-point filter decides the easy cases; the
 * rest are evaluated exactly with floating-point expansion arithmetic
 * (Shewchuk 1997, "Adaptive precision floating-point arithmetic and fast
 * robust geometric predicates": TWO-SUM, TWO-PRODUCT, SCALE-EXPANSION,
 * FAST-EXPANSION-SUM with zero elimination -- restated from the paper).
 *
 * Why the oracle is exact although the reference is not: the reference tests
 * `dist^2 < radius*radius` against a cached float circumsphere
 * (/root/reference/src/geometry.rs:91-97, :24-56).  In general position the
 * Delaunay triangulation is unique, so an exact engine and the reference agree
 * whenever the reference's float test takes no wrong decision.
 * The float restatement of the reference lives in oracle/refcpu.cpp.
 *
 * The GPU library uses a DIFFERENT exact method (scaled multi-word integers,
 * voronoids_b200/csrc/exact_int.cuh) so that the two exact paths check each
 * other; both are pinned against Python `fractions.Fraction` in tests/.
 *
 * Strictness: "in conflict" means strictly inside (sign > 0), on-sphere is
 * "not in conflict" -- mirrors the strict `<` at geometry.rs:96.
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction; explicit fma() only
 * inside two_prod).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EPS_HALF 1.1102230246251565e-16 /* 2^-53 */

/* counters (per process, not thread safe: the exact oracle is sequential) */
static uint64_t g_cnt_filter = 0, g_cnt_exact = 0, g_cnt_zero = 0;

void vo_pred_counters(uint64_t *filter, uint64_t *exact, uint64_t *zero, int reset) {
    if (filter) *filter = g_cnt_filter;
    if (exact) *exact = g_cnt_exact;
    if (zero) *zero = g_cnt_zero;
    if (reset) g_cnt_filter = g_cnt_exact = g_cnt_zero = 0;
}

/* ------------------------------------------------------------------ */
/* expansion arithmetic                                               */
/* ------------------------------------------------------------------ */

static inline void two_sum(double a, double b, double *x, double *y) {
    double s = a + b;
    double bv = s - a;
    double av = s - bv;
    *x = s;
    *y = (a - av) + (b - bv);
}
static inline void fast_two_sum(double a, double b, double *x, double *y) {
    double s = a + b; /* requires |a| >= |b| */
    *x = s;
    *y = b - (s - a);
}
static inline void two_diff(double a, double b, double *x, double *y) {
    double s = a - b;
    double bv = a - s;
    double av = s + bv;
    *x = s;
    *y = (a - av) + (bv - b);
}
static inline void two_prod(double a, double b, double *x, double *y) {
    double p = a * b;
    *x = p;
    *y = fma(a, b, -p);
}

/* bump arena for expansions: reset at every predicate call */
#define ARENA_DOUBLES (1u << 22)
static double *g_arena = NULL;
static size_t g_top = 0;

typedef struct {
    int n;
    double *c; /* increasing magnitude, nonoverlapping, no zeros */
} expn;

static double *arena_alloc(size_t n) {
    if (!g_arena) g_arena = (double *)malloc(sizeof(double) * ARENA_DOUBLES);
    if (g_top + n > ARENA_DOUBLES) abort(); /* oracle: fail loudly */
    double *p = g_arena + g_top;
    g_top += n;
    return p;
}

static expn ex_from_diff(double a, double b) {
    expn r;
    r.c = arena_alloc(2);
    double x, y;
    two_diff(a, b, &x, &y);
    r.n = 0;
    if (y != 0.0) r.c[r.n++] = y;
    if (x != 0.0) r.c[r.n++] = x;
    return r;
}

/* h = e * b (SCALE-EXPANSION with zero elimination) */
static expn ex_scale(expn e, double b) {
    expn h;
    h.n = 0;
    if (e.n == 0 || b == 0.0) {
        h.c = NULL;
        return h;
    }
    h.c = arena_alloc(2 * (size_t)e.n);
    double Q, hh, T, t, Qn;
    two_prod(e.c[0], b, &Q, &hh);
    if (hh != 0.0) h.c[h.n++] = hh;
    for (int i = 1; i < e.n; i++) {
        two_prod(e.c[i], b, &T, &t);
        two_sum(Q, t, &Qn, &hh);
        if (hh != 0.0) h.c[h.n++] = hh;
        fast_two_sum(T, Qn, &Q, &hh);
        if (hh != 0.0) h.c[h.n++] = hh;
    }
    if (Q != 0.0) h.c[h.n++] = Q;
    return h;
}

/* h = e + f (FAST-EXPANSION-SUM with zero elimination) */
static expn ex_add(expn e, expn f) {
    if (e.n == 0) return f;
    if (f.n == 0) return e;
    expn h;
    h.c = arena_alloc((size_t)e.n + f.n);
    h.n = 0;
    int ei = 0, fi = 0;
    double Q, Qn, hh, nxt;
    /* merge by increasing magnitude */
    if (fabs(f.c[0]) > fabs(e.c[0])) Q = e.c[ei++];
    else Q = f.c[fi++];
    int first = 1;
    while (ei < e.n || fi < f.n) {
        if (fi >= f.n || (ei < e.n && fabs(f.c[fi]) > fabs(e.c[ei]))) nxt = e.c[ei++];
        else nxt = f.c[fi++];
        if (first) {
            fast_two_sum(nxt, Q, &Qn, &hh);
            first = 0;
        } else {
            two_sum(Q, nxt, &Qn, &hh);
        }
        Q = Qn;
        if (hh != 0.0) h.c[h.n++] = hh;
    }
    if (Q != 0.0) h.c[h.n++] = Q;
    return h;
}

static expn ex_neg(expn e) {
    expn h;
    h.n = e.n;
    h.c = e.n ? arena_alloc((size_t)e.n) : NULL;
    for (int i = 0; i < e.n; i++) h.c[i] = -e.c[i];
    return h;
}
static expn ex_sub(expn e, expn f) { return ex_add(e, ex_neg(f)); }

/* h = e * f : distribute over the components of the shorter operand,
 * summing partial products pairwise (balanced) to keep the sums short. */
static expn ex_mul(expn e, expn f) {
    if (e.n < f.n) {
        expn t = e;
        e = f;
        f = t;
    }
    expn zero = {0, NULL};
    if (f.n == 0) return zero;
    expn parts[64];
    int np = 0;
    /* binary-counter style balanced accumulation */
    int level[64];
    for (int i = 0; i < f.n; i++) {
        expn cur = ex_scale(e, f.c[i]);
        int lv = 0;
        while (np > 0 && level[np - 1] == lv) {
            cur = ex_add(parts[np - 1], cur);
            np--;
            lv++;
        }
        parts[np] = cur;
        level[np] = lv;
        np++;
    }
    expn acc = parts[np - 1];
    for (int i = np - 2; i >= 0; i--) acc = ex_add(parts[i], acc);
    return acc;
}

static int ex_sign(expn e) {
    if (e.n == 0) return 0;
    return e.c[e.n - 1] > 0.0 ? 1 : -1;
}

/* ------------------------------------------------------------------ */
/* exact evaluations (no filter)                                      */
/* ------------------------------------------------------------------ */

int vo_orient2d_exact(const double *a, const double *b, const double *c) {
    g_top = 0;
    expn acx = ex_from_diff(a[0], c[0]), acy = ex_from_diff(a[1], c[1]);
    expn bcx = ex_from_diff(b[0], c[0]), bcy = ex_from_diff(b[1], c[1]);
    return ex_sign(ex_sub(ex_mul(acx, bcy), ex_mul(acy, bcx)));
}

int vo_orient3d_exact(const double *a, const double *b, const double *c, const double *d) {
    g_top = 0;
    expn adx = ex_from_diff(a[0], d[0]), ady = ex_from_diff(a[1], d[1]), adz = ex_from_diff(a[2], d[2]);
    expn bdx = ex_from_diff(b[0], d[0]), bdy = ex_from_diff(b[1], d[1]), bdz = ex_from_diff(b[2], d[2]);
    expn cdx = ex_from_diff(c[0], d[0]), cdy = ex_from_diff(c[1], d[1]), cdz = ex_from_diff(c[2], d[2]);
    expn m1 = ex_sub(ex_mul(bdy, cdz), ex_mul(bdz, cdy));
    expn m2 = ex_sub(ex_mul(cdy, adz), ex_mul(cdz, ady));
    expn m3 = ex_sub(ex_mul(ady, bdz), ex_mul(adz, bdy));
    expn det = ex_add(ex_add(ex_mul(adx, m1), ex_mul(bdx, m2)), ex_mul(cdx, m3));
    return ex_sign(det);
}

int vo_incircle_exact(const double *a, const double *b, const double *c, const double *d) {
    g_top = 0;
    expn adx = ex_from_diff(a[0], d[0]), ady = ex_from_diff(a[1], d[1]);
    expn bdx = ex_from_diff(b[0], d[0]), bdy = ex_from_diff(b[1], d[1]);
    expn cdx = ex_from_diff(c[0], d[0]), cdy = ex_from_diff(c[1], d[1]);
    expn al = ex_add(ex_mul(adx, adx), ex_mul(ady, ady));
    expn bl = ex_add(ex_mul(bdx, bdx), ex_mul(bdy, bdy));
    expn cl = ex_add(ex_mul(cdx, cdx), ex_mul(cdy, cdy));
    expn bc = ex_sub(ex_mul(bdx, cdy), ex_mul(cdx, bdy));
    expn ca = ex_sub(ex_mul(cdx, ady), ex_mul(adx, cdy));
    expn ab = ex_sub(ex_mul(adx, bdy), ex_mul(bdx, ady));
    expn det = ex_add(ex_add(ex_mul(al, bc), ex_mul(bl, ca)), ex_mul(cl, ab));
    return ex_sign(det);
}

int vo_insphere_exact(const double *a, const double *b, const double *c, const double *d, const double *e) {
    g_top = 0;
    expn aex = ex_from_diff(a[0], e[0]), aey = ex_from_diff(a[1], e[1]), aez = ex_from_diff(a[2], e[2]);
    expn bex = ex_from_diff(b[0], e[0]), bey = ex_from_diff(b[1], e[1]), bez = ex_from_diff(b[2], e[2]);
    expn cex = ex_from_diff(c[0], e[0]), cey = ex_from_diff(c[1], e[1]), cez = ex_from_diff(c[2], e[2]);
    expn dex = ex_from_diff(d[0], e[0]), dey = ex_from_diff(d[1], e[1]), dez = ex_from_diff(d[2], e[2]);
    expn ab = ex_sub(ex_mul(aex, bey), ex_mul(bex, aey));
    expn bc = ex_sub(ex_mul(bex, cey), ex_mul(cex, bey));
    expn cd = ex_sub(ex_mul(cex, dey), ex_mul(dex, cey));
    expn da = ex_sub(ex_mul(dex, aey), ex_mul(aex, dey));
    expn ac = ex_sub(ex_mul(aex, cey), ex_mul(cex, aey));
    expn bd = ex_sub(ex_mul(bex, dey), ex_mul(dex, bey));
    expn abc = ex_add(ex_sub(ex_mul(aez, bc), ex_mul(bez, ac)), ex_mul(cez, ab));
    expn bcd = ex_add(ex_sub(ex_mul(bez, cd), ex_mul(cez, bd)), ex_mul(dez, bc));
    expn cda = ex_add(ex_add(ex_mul(cez, da), ex_mul(dez, ac)), ex_mul(aez, cd));
    expn dab = ex_add(ex_add(ex_mul(dez, ab), ex_mul(aez, bd)), ex_mul(bez, da));
    expn al = ex_add(ex_add(ex_mul(aex, aex), ex_mul(aey, aey)), ex_mul(aez, aez));
    expn bl = ex_add(ex_add(ex_mul(bex, bex), ex_mul(bey, bey)), ex_mul(bez, bez));
    expn cl = ex_add(ex_add(ex_mul(cex, cex), ex_mul(cey, cey)), ex_mul(cez, cez));
    expn dl = ex_add(ex_add(ex_mul(dex, dex), ex_mul(dey, dey)), ex_mul(dez, dez));
    expn det = ex_add(ex_sub(ex_mul(dl, abc), ex_mul(cl, dab)), ex_sub(ex_mul(bl, cda), ex_mul(al, bcd)));
    return ex_sign(det);
}

/* ------------------------------------------------------------------ */
/* filtered predicates: sign in {-1,0,+1}                             */
/* ------------------------------------------------------------------ */

/* orient2d > 0  <=>  a,b,c counter-clockwise */
int vo_orient2d(const double *a, const double *b, const double *c) {
    double l = (a[0] - c[0]) * (b[1] - c[1]);
    double r = (a[1] - c[1]) * (b[0] - c[0]);
    double det = l - r;
    double bound = (3.0 + 16.0 * EPS_HALF) * EPS_HALF * (fabs(l) + fabs(r));
    if (det > bound) { g_cnt_filter++; return 1; }
    if (-det > bound) { g_cnt_filter++; return -1; }
    g_cnt_exact++;
    int s = vo_orient2d_exact(a, b, c);
    if (s == 0) g_cnt_zero++;
    return s;
}

/* orient3d(a,b,c,d) = sign det[a-d; b-d; c-d] */
int vo_orient3d(const double *a, const double *b, const double *c, const double *d) {
    double adx = a[0] - d[0], bdx = b[0] - d[0], cdx = c[0] - d[0];
    double ady = a[1] - d[1], bdy = b[1] - d[1], cdy = c[1] - d[1];
    double adz = a[2] - d[2], bdz = b[2] - d[2], cdz = c[2] - d[2];
    double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy;
    double cdxady = cdx * ady, adxcdy = adx * cdy;
    double adxbdy = adx * bdy, bdxady = bdx * ady;
    double det = adz * (bdxcdy - cdxbdy) + bdz * (cdxady - adxcdy) + cdz * (adxbdy - bdxady);
    double perm = (fabs(bdxcdy) + fabs(cdxbdy)) * fabs(adz) + (fabs(cdxady) + fabs(adxcdy)) * fabs(bdz) +
                  (fabs(adxbdy) + fabs(bdxady)) * fabs(cdz);
    double bound = (7.0 + 56.0 * EPS_HALF) * EPS_HALF * perm;
    if (det > bound) { g_cnt_filter++; return 1; }
    if (-det > bound) { g_cnt_filter++; return -1; }
    g_cnt_exact++;
    int s = vo_orient3d_exact(a, b, c, d);
    if (s == 0) g_cnt_zero++;
    return s;
}

/* incircle(a,b,c,d) > 0 <=> d strictly inside circle(a,b,c) when orient2d(a,b,c) > 0 */
int vo_incircle(const double *a, const double *b, const double *c, const double *d) {
    double adx = a[0] - d[0], ady = a[1] - d[1];
    double bdx = b[0] - d[0], bdy = b[1] - d[1];
    double cdx = c[0] - d[0], cdy = c[1] - d[1];
    double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy;
    double cdxady = cdx * ady, adxcdy = adx * cdy;
    double adxbdy = adx * bdy, bdxady = bdx * ady;
    double al = adx * adx + ady * ady, bl = bdx * bdx + bdy * bdy, cl = cdx * cdx + cdy * cdy;
    double det = al * (bdxcdy - cdxbdy) + bl * (cdxady - adxcdy) + cl * (adxbdy - bdxady);
    double perm = (fabs(bdxcdy) + fabs(cdxbdy)) * al + (fabs(cdxady) + fabs(adxcdy)) * bl +
                  (fabs(adxbdy) + fabs(bdxady)) * cl;
    double bound = (10.0 + 96.0 * EPS_HALF) * EPS_HALF * perm;
    if (det > bound) { g_cnt_filter++; return 1; }
    if (-det > bound) { g_cnt_filter++; return -1; }
    g_cnt_exact++;
    int s = vo_incircle_exact(a, b, c, d);
    if (s == 0) g_cnt_zero++;
    return s;
}

/* insphere(a,b,c,d,e) > 0 <=> e strictly inside sphere(a,b,c,d) when orient3d(a,b,c,d) > 0 */
int vo_insphere(const double *a, const double *b, const double *c, const double *d, const double *e) {
    double aex = a[0] - e[0], bex = b[0] - e[0], cex = c[0] - e[0], dex = d[0] - e[0];
    double aey = a[1] - e[1], bey = b[1] - e[1], cey = c[1] - e[1], dey = d[1] - e[1];
    double aez = a[2] - e[2], bez = b[2] - e[2], cez = c[2] - e[2], dez = d[2] - e[2];
    double aexbey = aex * bey, bexaey = bex * aey, ab = aexbey - bexaey;
    double bexcey = bex * cey, cexbey = cex * bey, bc = bexcey - cexbey;
    double cexdey = cex * dey, dexcey = dex * cey, cd = cexdey - dexcey;
    double dexaey = dex * aey, aexdey = aex * dey, da = dexaey - aexdey;
    double aexcey = aex * cey, cexaey = cex * aey, ac = aexcey - cexaey;
    double bexdey = bex * dey, dexbey = dex * bey, bd = bexdey - dexbey;
    double abc = aez * bc - bez * ac + cez * ab;
    double bcd = bez * cd - cez * bd + dez * bc;
    double cda = cez * da + dez * ac + aez * cd;
    double dab = dez * ab + aez * bd + bez * da;
    double al = aex * aex + aey * aey + aez * aez;
    double bl = bex * bex + bey * bey + bez * bez;
    double cl = cex * cex + cey * cey + cez * cez;
    double dl = dex * dex + dey * dey + dez * dez;
    double det = (dl * abc - cl * dab) + (bl * cda - al * bcd);
    double aezp = fabs(aez), bezp = fabs(bez), cezp = fabs(cez), dezp = fabs(dez);
    double aexbeyp = fabs(aexbey), bexaeyp = fabs(bexaey), bexceyp = fabs(bexcey), cexbeyp = fabs(cexbey);
    double cexdeyp = fabs(cexdey), dexceyp = fabs(dexcey), dexaeyp = fabs(dexaey), aexdeyp = fabs(aexdey);
    double aexceyp = fabs(aexcey), cexaeyp = fabs(cexaey), bexdeyp = fabs(bexdey), dexbeyp = fabs(dexbey);
    double perm = ((cexdeyp + dexceyp) * bezp + (dexbeyp + bexdeyp) * cezp + (bexceyp + cexbeyp) * dezp) * al +
                  ((dexaeyp + aexdeyp) * cezp + (aexceyp + cexaeyp) * dezp + (cexdeyp + dexceyp) * aezp) * bl +
                  ((aexbeyp + bexaeyp) * dezp + (bexdeyp + dexbeyp) * aezp + (dexaeyp + aexdeyp) * bezp) * cl +
                  ((bexceyp + cexbeyp) * aezp + (cexaeyp + aexceyp) * bezp + (aexbeyp + bexaeyp) * cezp) * dl;
    double bound = (16.0 + 224.0 * EPS_HALF) * EPS_HALF * perm;
    if (det > bound) { g_cnt_filter++; return 1; }
    if (-det > bound) { g_cnt_filter++; return -1; }
    g_cnt_exact++;
    int s = vo_insphere_exact(a, b, c, d, e);
    if (s == 0) g_cnt_zero++;
    return s;
}

/* batch entry points for tests: flat arrays, one predicate per row */
void vo_orient3d_batch(const double *abcd, int n, int *out, int exact_only) {
    for (int i = 0; i < n; i++) {
        const double *p = abcd + 12 * (size_t)i;
        out[i] = exact_only ? vo_orient3d_exact(p, p + 3, p + 6, p + 9) : vo_orient3d(p, p + 3, p + 6, p + 9);
    }
}
void vo_insphere_batch(const double *abcde, int n, int *out, int exact_only) {
    for (int i = 0; i < n; i++) {
        const double *p = abcde + 15 * (size_t)i;
        out[i] = exact_only ? vo_insphere_exact(p, p + 3, p + 6, p + 9, p + 12) : vo_insphere(p, p + 3, p + 6, p + 9, p + 12);
    }
}
void vo_orient2d_batch(const double *abc, int n, int *out, int exact_only) {
    for (int i = 0; i < n; i++) {
        const double *p = abc + 6 * (size_t)i;
        out[i] = exact_only ? vo_orient2d_exact(p, p + 2, p + 4) : vo_orient2d(p, p + 2, p + 4);
    }
}
void vo_incircle_batch(const double *abcd, int n, int *out, int exact_only) {
    for (int i = 0; i < n; i++) {
        const double *p = abcd + 8 * (size_t)i;
        out[i] = exact_only ? vo_incircle_exact(p, p + 2, p + 4, p + 6) : vo_incircle(p, p + 2, p + 4, p + 6);
    }
}
