/*
 * oracle/ref_geometry.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Float restatement of /root/reference/src/geometry.rs (operation order as
 * written there; build with -ffp-contract=off because rustc never contracts
 * a*b+c into an FMA):
 *   circumsphere_2d   geometry.rs:2-22
 *   circumsphere_3d   geometry.rs:24-56   (nalgebra 0.32.4 Matrix3::lu().solve())
 *   in_sphere         geometry.rs:91-97
 *   bounding_sphere   geometry.rs:99-142
 * and of the super-simplex coordinates of DelaunayTree::new
 *   3D  /root/reference/src/delaunay_tree.rs:392-406
 *   2D  /root/reference/src/delaunay_tree.rs:547-558
 *
 * nalgebra is NOT under /root/reference (Cargo.lock pins 0.32.4).  Its LU is
 * restated from its published algorithm: partial (row) pivoting by max |.|,
 * the pivot column scaled by the reciprocal of the pivot, axpy updates without
 * FMA, unit-lower forward substitution, upper back substitution dividing by
 * the diagonal.  Pinned only by the reference's single known-answer test
 * (tests/test_geometry.rs:5-15); beyond it bit-level parity of center/radius
 * is UNPINNED.
 */
#ifndef VO_REF_GEOMETRY_H
#define VO_REF_GEOMETRY_H
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

static inline int ref_in_sphere(int N, const double *vertex, const double *center, double radius) {
    double distance = 0.0;
    for (int i = 0; i < N; i++) distance += (center[i] - vertex[i]) * (center[i] - vertex[i]);
    return distance < radius * radius;
}

static inline void ref_circumsphere_2d(const double *v /*3x2*/, double *center, double *radius) {
    double x1 = v[0], y1 = v[1], x2 = v[2], y2 = v[3], x3 = v[4], y3 = v[5];
    double d0 = (x1 + x2) / 2.0, d1 = (y1 + y2) / 2.0;
    double e0 = (x2 + x3) / 2.0, e1 = (y2 + y3) / 2.0;
    double m_ab = (y2 - y1) / (x2 - x1);
    double m_bc = (y3 - y2) / (x3 - x2);
    double m_d = -1. / m_ab;
    double m_e = -1. / m_bc;
    double x = (m_d * d0 - m_e * e0 + e1 - d1) / (m_d - m_e);
    double y = m_d * (x - d0) + d1;
    double r = sqrt((x - x1) * (x - x1) + (y - y1) * (y - y1));
    center[0] = x;
    center[1] = y;
    *radius = r;
}

/* returns 0 on success, 1 if the LU solve fails (reference: .unwrap() panic, geometry.rs:49) */
static inline int ref_circumsphere_3d(const double *v /*4x3*/, double *center, double *radius) {
    double a[3][3], mid[3][3], b[3];
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++) {
            a[i][k] = v[3 * (i + 1) + k] - v[k];
            mid[i][k] = (v[3 * (i + 1) + k] + v[k]) / 2.0;
        }
    for (int i = 0; i < 3; i++) {
        double s = 0.0;
        for (int k = 0; k < 3; k++) s += a[i][k] * mid[i][k];
        b[i] = s;
    }
    /* LU with partial pivoting (nalgebra::linalg::LU::new) */
    int perm[3] = {0, 1, 2};
    for (int i = 0; i < 3; i++) {
        int piv = i;
        double best = fabs(a[i][i]);
        for (int r = i + 1; r < 3; r++)
            if (fabs(a[r][i]) > best) { best = fabs(a[r][i]); piv = r; }
        double diag = a[piv][i];
        if (diag == 0.0) continue;
        if (piv != i) {
            for (int k = 0; k < 3; k++) { double t = a[i][k]; a[i][k] = a[piv][k]; a[piv][k] = t; }
            int t = perm[i]; perm[i] = perm[piv]; perm[piv] = t;
        }
        double inv = 1.0 / diag;
        for (int r = i + 1; r < 3; r++) a[r][i] *= inv;
        for (int k = i + 1; k < 3; k++)
            for (int r = i + 1; r < 3; r++) a[r][k] = (-a[i][k]) * a[r][i] + a[r][k];
    }
    double x[3] = {b[perm[0]], b[perm[1]], b[perm[2]]};
    for (int i = 0; i < 3; i++) {
        double coeff = x[i];
        for (int r = i + 1; r < 3; r++) x[r] = (-coeff) * a[r][i] + x[r];
    }
    for (int i = 2; i >= 0; i--) {
        if (a[i][i] == 0.0) return 1;
        double coeff = x[i] / a[i][i];
        x[i] = coeff;
        for (int r = 0; r < i; r++) x[r] = (-coeff) * a[r][i] + x[r];
    }
    center[0] = x[0];
    center[1] = x[1];
    center[2] = x[2];
    *radius = sqrt((v[0] - x[0]) * (v[0] - x[0]) + (v[1] - x[1]) * (v[1] - x[1]) + (v[2] - x[2]) * (v[2] - x[2]));
    return 0;
}

/* geometry.rs:99-142 */
static inline void ref_bounding_sphere(int N, const double *pts, long n, double *center, double *radius) {
    double lo[3], hi[3];
    for (int i = 0; i < N; i++) {
        double l = INFINITY, h = -INFINITY;
        for (long p = 0; p < n; p++) {
            double c = pts[p * N + i];
            l = fmin(l, c); /* f64::min */
            h = fmax(h, c);
        }
        lo[i] = l;
        hi[i] = h;
        center[i] = (h + l) / 2.0;
    }
    double ud = 0.0, ld = 0.0;
    for (int i = 0; i < N; i++) ud += (hi[i] - center[i]) * (hi[i] - center[i]);
    for (int i = 0; i < N; i++) ld += (lo[i] - center[i]) * (lo[i] - center[i]);
    ud = sqrt(ud);
    ld = sqrt(ld);
    double r = ud > ld ? ud : ld;
    for (long p = 0; p < n; p++) {
        if (!ref_in_sphere(N, pts + p * N, center, r)) r = r * 1.5;
    }
    *radius = r;
}

/* super-simplex vertices, M x N row-major; also returns center and 10x radius */
static inline void ref_super_simplex(int N, const double *pts, long n, double *super, double *center, double *radius) {
    double r;
    ref_bounding_sphere(N, pts, n, center, &r);
    r *= 10.0;
    *radius = r;
    const double PI = 3.14159265358979323846264338327950288;
    double a1 = 2. * PI / 3., a2 = 4. * PI / 3.;
    if (N == 3) {
        super[0] = center[0]; super[1] = center[1]; super[2] = center[2] + r;
        super[3] = center[0] + r; super[4] = center[1]; super[5] = center[2] - r;
        super[6] = center[0] + r * cos(a1); super[7] = center[1] + r * sin(a1); super[8] = center[2] - r;
        super[9] = center[0] + r * cos(a2); super[10] = center[1] + r * sin(a2); super[11] = center[2] - r;
    } else {
        super[0] = center[0] + r; super[1] = center[1];
        super[2] = center[0] + r * cos(a1); super[3] = center[1] + r * sin(a1);
        super[4] = center[0] + r * cos(a2); super[5] = center[1] + r * sin(a2);
    }
}

#ifdef __cplusplus
}
#endif
#endif
