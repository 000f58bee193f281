/*
 * oracle/ref_geometry.c -- TEST INFRASTRUCTURE, not product code.
 * C-callable exports of the float restatement in ref_geometry.h
 * (see that header for the reference file:line each function follows).
 */
#include "ref_geometry.h"

int vo_ref_in_sphere(int N, const double *vertex, const double *center, double radius) {
    return ref_in_sphere(N, vertex, center, radius);
}
int vo_ref_circumsphere(int N, const double *verts, double *center, double *radius) {
    if (N == 2) { ref_circumsphere_2d(verts, center, radius); return 0; }
    return ref_circumsphere_3d(verts, center, radius);
}
void vo_ref_bounding_sphere(int N, const double *pts, long n, double *center, double *radius) {
    ref_bounding_sphere(N, pts, n, center, radius);
}
void vo_ref_super_simplex(int N, const double *pts, long n, double *super, double *center, double *radius) {
    ref_super_simplex(N, pts, n, super, center, radius);
}
