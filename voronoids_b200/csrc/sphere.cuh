// sphere.cuh -- certified circumsphere filter stored next to the ownership words of every simplex.
//
// The reference caches centre + radius per simplex and decides "in conflict" by dist^2 < r*r in plain f64
// (/root/reference/src/delaunay_tree.rs:11-16, src/geometry.rs:91-97).  Here the cached sphere is a FILTER with a
// proof obligation: the block of a simplex holds a float centre (relative to the tree's origin) and two float radii
//     rin2  : every query with |q - c|^2 <  rin2 is certainly STRICTLY INSIDE the true circumsphere
//     rout2 : every query with |q - c|^2 >  rout2 is certainly NOT strictly inside it
// and everything in between (a shell of relative thickness ~1e-5 at 10M uniform points) falls through to the
// determinant predicates of predicates.cuh (FP64 filter -> exact integers).  The decision taken is therefore always the
// exact one; the filter only removes the five scattered gathers (record + 4 vertices) of a determinant test: one 32 B
// block -- ownership words and sphere together -- decides a conflict test with ONE 256-bit load.
//
// Block layout (8 ints = one 32 B sector per simplex, Mesh::owner):
//   [0] kill word   [1] ring word   [2..4] float centre - origin   [5] rin2   [6] rout2   [7] unused
//
// Error analysis (EPS = 2^-53; all constants are rounded up generously):
//   a, b, c  = computed edge vectors p_i - p_0; the true centre c* relative to p_0 solves 2 M* x = s* (M* = true edge
//              vectors as rows, s*_i = |edge_i|^2).  For the computed x^ (Cramer's rule in f64) the residual of the TRUE
//              system is bounded by |r^_i| + 16 EPS (2 |edge_i|.|x^| + |edge_i|^2), with r^ evaluated in f64, and
//              |x^ - c*'| = |M*^-1 r*| / 2 <= ||adj M||_F |r*| / (2 (|det| - 8 EPS permanent(|M|))), with the cheap upper bounds
//              ||adj M||_F^2 <= |b|^2|c|^2 + |c|^2|a|^2 + |a|^2|b|^2 and permanent(|M|) <= product of the row 1-norms.
//              A simplex whose determinant bound is not positive gets no filter (rin2 = 0, rout2 = +inf).
//   storing   the centre as floats relative to the origin adds the measured deviation |float(g) - g| and the rounding of
//              the two additions; the query q = fl(p - origin) adds EPS |p - origin| (Mesh::sref.qerr bounds it).
//   radius    r^ = |p_0 - stored centre| differs from the true radius by at most the centre error rho (+ 4 EPS r^).
//   =>        Rin = r^ (1 - 8 EPS) - 2 rho, Rout = r^ (1 + 8 EPS) + 2 rho; squared with 16 EPS slack and rounded
//              down / up to float.  The test evaluates |q - c|^2 in f64 (relative error <= 6 EPS with or without FMA).
// tests/test_sphere_filter.py checks every certified verdict against the exact predicates on random, sliver and
// near-cospherical inputs (emulation build) and Engine::validate re-checks every stored block of a finished mesh against
// the opposite vertices of its neighbours (fail counter 5).
#pragma once
#include "vor_common.cuh"

namespace vor {

// Store layout of one simplex.  VOR_INTERLEAVE = 1 (default): ONE 64 B line per simplex -- [ownership + sphere block | vertex
// ids | neighbour codes] -- because the L2 fills from HBM in 64 B units: the block gather of a conflict test brings the
// neighbour codes the flood needs next (and the commit's ownership re-read brings the record) in the same DRAM access.
// Measured (ncu, round 2): with separate arrays the attempt kernel read 1.5x more DRAM sectors than it asked for.
// VOR_INTERLEAVE = 0: block array (8 ints per simplex) and record array (2 x int4 per simplex) apart.
#ifndef VOR_INTERLEAVE
#define VOR_INTERLEAVE 1
#endif
#if VOR_INTERLEAVE
constexpr int OWS = 16;                      // ints from one simplex's block to the next in Mesh::owner
constexpr int REC4 = 4;                      // int4 from one simplex's record to the next in Mesh::tet
constexpr int TVO4 = 2;                      // int4 offset of the record inside the simplex's line
#else
constexpr int OWS = 8;
constexpr int REC4 = 2;
constexpr int TVO4 = 0;
#endif
constexpr double SPH_EPS = 1.1102230246251565e-16;

struct SphereRef {
    double ox, oy, oz;   // origin of the float centres (centre of the bounding box of all sets)
    double qerr;         // >= |fl(p - origin) - (p - origin)| for every legal query point p
};

struct SphereBlk { float cx, cy, cz, rin2, rout2; };

VOR_HD float f32_down(double x) {
    float f = (float)x;
    if ((double)f > x) f = nextafterf(f, -INFINITY);
    return f;
}
VOR_HD float f32_up(double x) {
    float f = (float)x;
    if ((double)f < x) f = nextafterf(f, INFINITY);
    return f;
}

VOR_HD SphereBlk sphere_none() {
    SphereBlk o;
    o.cx = 0.0f; o.cy = 0.0f; o.cz = 0.0f; o.rin2 = 0.0f; o.rout2 = INFINITY;
    return o;
}

// finish: centre relative to p0 (ccx..) with error bound rho_c -> float block
VOR_HD SphereBlk sphere_finish(double p0x, double p0y, double p0z, double ccx, double ccy, double ccz, double rho_c, const SphereRef &R) {
    SphereBlk o = sphere_none();
    const double hx = p0x - R.ox, hy = p0y - R.oy, hz = p0z - R.oz;
    const double gx = hx + ccx, gy = hy + ccy, gz = hz + ccz;
    if (!(fabs(gx) < 1e30 && fabs(gy) < 1e30 && fabs(gz) < 1e30) || !(rho_c >= 0.0)) return o;   // also rejects NaN
    const float fx = (float)gx, fy = (float)gy, fz = (float)gz;
    const double ex = (double)fx - gx, ey = (double)fy - gy, ez = (double)fz - gz;
    const double rho_f = sqrt(ex * ex + ey * ey + ez * ez) * (1.0 + 1e-9);
    const double mag = fabs(hx) + fabs(gx) + fabs(hy) + fabs(gy) + fabs(hz) + fabs(gz) + fabs(ccx) + fabs(ccy) + fabs(ccz);
    const double rho = rho_c + rho_f + 4.0 * SPH_EPS * mag + R.qerr;
    const double ux = hx - (double)fx, uy = hy - (double)fy, uz = hz - (double)fz;
    const double rhat = sqrt(ux * ux + uy * uy + uz * uz);
    const double Rin = rhat * (1.0 - 8.0 * SPH_EPS) - 2.0 * rho;
    const double Rout = rhat * (1.0 + 8.0 * SPH_EPS) + 2.0 * rho;
    if (!(Rout < 1e150)) return o;
    o.cx = fx; o.cy = fy; o.cz = fz;
    o.rin2 = Rin > 0.0 ? f32_down(Rin * Rin * (1.0 - 16.0 * SPH_EPS)) : 0.0f;
    if (!(o.rin2 >= 0.0f)) o.rin2 = 0.0f;
    o.rout2 = f32_up(Rout * Rout * (1.0 + 16.0 * SPH_EPS));
    if (!(o.rout2 > 0.0f)) o.rout2 = INFINITY;   // NaN guard
    return o;
}

VOR_HD SphereBlk sphere_make(const double4 &p0, const double4 &p1, const double4 &p2, const double4 &p3, const SphereRef &R) {
    const double ax = p1.x - p0.x, ay = p1.y - p0.y, az = p1.z - p0.z;
    const double bx = p2.x - p0.x, by = p2.y - p0.y, bz = p2.z - p0.z;
    const double cx = p3.x - p0.x, cy = p3.y - p0.y, cz = p3.z - p0.z;
    // adjugate columns (cross products)
    const double bcx = by * cz - bz * cy, bcy = bz * cx - bx * cz, bcz = bx * cy - by * cx;
    const double cax = cy * az - cz * ay, cay = cz * ax - cx * az, caz = cx * ay - cy * ax;
    const double abx = ay * bz - az * by, aby = az * bx - ax * bz, abz = ax * by - ay * bx;
    const double det = ax * bcx + ay * bcy + az * bcz;
    const double sa = ax * ax + ay * ay + az * az, sb = bx * bx + by * by + bz * bz, sc = cx * cx + cy * cy + cz * cz;
    // cheap upper bounds in place of the permanents (the kernel that runs this is bound by the FP64 pipe): the permanent of
    // |M| is at most the product of the row 1-norms, and |u x v| <= |u| |v| bounds the Frobenius norm of the adjugate
    const double a1 = fabs(ax) + fabs(ay) + fabs(az), b1 = fabs(bx) + fabs(by) + fabs(bz), c1 = fabs(cx) + fabs(cy) + fabs(cz);
    const double detLow = fabs(det) - 8.0 * SPH_EPS * (a1 * b1 * c1) * (1.0 + 1e-9);
    if (!(detLow > 0.0)) return sphere_none();
    const double inv = 0.5 / det;
    const double ccx = (sa * bcx + sb * cax + sc * abx) * inv;
    const double ccy = (sa * bcy + sb * cay + sc * aby) * inv;
    const double ccz = (sa * bcz + sb * caz + sc * abz) * inv;
    const double acx = fabs(ccx), acy = fabs(ccy), acz = fabs(ccz);
    const double ra = fabs(2.0 * (ax * ccx + ay * ccy + az * ccz) - sa) + 16.0 * SPH_EPS * (2.0 * (fabs(ax) * acx + fabs(ay) * acy + fabs(az) * acz) + sa);
    const double rb = fabs(2.0 * (bx * ccx + by * ccy + bz * ccz) - sb) + 16.0 * SPH_EPS * (2.0 * (fabs(bx) * acx + fabs(by) * acy + fabs(bz) * acz) + sb);
    const double rc = fabs(2.0 * (cx * ccx + cy * ccy + cz * ccz) - sc) + 16.0 * SPH_EPS * (2.0 * (fabs(cx) * acx + fabs(cy) * acy + fabs(cz) * acz) + sc);
    const double adj2 = (sb * sc + sc * sa + sa * sb) * (1.0 + 1e-9);
    const double rho_c = 0.5 * sqrt(adj2 * (ra * ra + rb * rb + rc * rc)) / detLow * (1.0 + 1e-9);
    return sphere_finish(p0.x, p0.y, p0.z, ccx, ccy, ccz, rho_c, R);
}

VOR_HD SphereBlk sphere_make(const double2 &p0, const double2 &p1, const double2 &p2, const SphereRef &R) {
    const double ax = p1.x - p0.x, ay = p1.y - p0.y;
    const double bx = p2.x - p0.x, by = p2.y - p0.y;
    const double det = ax * by - ay * bx;
    const double permdet = fabs(ax * by) + fabs(ay * bx);
    const double detLow = fabs(det) - 4.0 * SPH_EPS * permdet;
    if (!(detLow > 0.0)) return sphere_none();
    const double sa = ax * ax + ay * ay, sb = bx * bx + by * by;
    const double inv = 0.5 / det;
    const double ccx = (sa * by - sb * ay) * inv;
    const double ccy = (sb * ax - sa * bx) * inv;
    const double acx = fabs(ccx), acy = fabs(ccy);
    const double ra = fabs(2.0 * (ax * ccx + ay * ccy) - sa) + 16.0 * SPH_EPS * (2.0 * (fabs(ax) * acx + fabs(ay) * acy) + sa);
    const double rb = fabs(2.0 * (bx * ccx + by * ccy) - sb) + 16.0 * SPH_EPS * (2.0 * (fabs(bx) * acx + fabs(by) * acy) + sb);
    const double adj2 = ax * ax + ay * ay + bx * bx + by * by;
    const double rho_c = 0.5 * sqrt(adj2) * sqrt(ra * ra + rb * rb) / detLow * (1.0 + 1e-9);
    return sphere_finish(p0.x, p0.y, 0.0, ccx, ccy, 0.0, rho_c, R);
}

// The same ball WITHOUT the float rounding of the stored block: absolute centre and an outer radius Rout in f64 with
//     true open ball  subset of  { q : |q - c| <= Rout }.
// For the slab certification (output_kernels.cuh): the circumsphere of a simplex on the hull is nearly a plane (radius 1e3..1e5
// box widths), and how far its cap reaches into the data box is sqrt(Rout^2 - d^2) with d within a hair of Rout -- the 6e-8
// relative rounding of a float rout2 moves that reach by tenths of the box.  Same formulas and bounds as sphere_make /
// sphere_finish minus the float storage terms.  false: no bound (degenerate simplex).
VOR_HD bool sphere_ball_d(const double4 &p0, const double4 &p1, const double4 &p2, const double4 &p3, double c[3], double &Rout) {
    const double ax = p1.x - p0.x, ay = p1.y - p0.y, az = p1.z - p0.z;
    const double bx = p2.x - p0.x, by = p2.y - p0.y, bz = p2.z - p0.z;
    const double cx = p3.x - p0.x, cy = p3.y - p0.y, cz = p3.z - p0.z;
    const double bcx = by * cz - bz * cy, bcy = bz * cx - bx * cz, bcz = bx * cy - by * cx;
    const double cax = cy * az - cz * ay, cay = cz * ax - cx * az, caz = cx * ay - cy * ax;
    const double abx = ay * bz - az * by, aby = az * bx - ax * bz, abz = ax * by - ay * bx;
    const double det = ax * bcx + ay * bcy + az * bcz;
    const double sa = ax * ax + ay * ay + az * az, sb = bx * bx + by * by + bz * bz, sc = cx * cx + cy * cy + cz * cz;
    // the permanent of |M| itself (this is not a hot path): the product of the row 1-norms that sphere_make uses overestimates it by
    // orders of magnitude for a hull simplex (two long edges to super vertices, one short edge) and denies it a bound
    const double perm = fabs(ax) * (fabs(by * cz) + fabs(bz * cy)) + fabs(ay) * (fabs(bz * cx) + fabs(bx * cz)) + fabs(az) * (fabs(bx * cy) + fabs(by * cx));
    const double detLow = fabs(det) - 8.0 * SPH_EPS * perm * (1.0 + 1e-9);
    if (!(detLow > 0.0)) return false;
    const double inv = 0.5 / det;
    const double ccx = (sa * bcx + sb * cax + sc * abx) * inv;
    const double ccy = (sa * bcy + sb * cay + sc * aby) * inv;
    const double ccz = (sa * bcz + sb * caz + sc * abz) * inv;
    const double acx = fabs(ccx), acy = fabs(ccy), acz = fabs(ccz);
    const double ra = fabs(2.0 * (ax * ccx + ay * ccy + az * ccz) - sa) + 16.0 * SPH_EPS * (2.0 * (fabs(ax) * acx + fabs(ay) * acy + fabs(az) * acz) + sa);
    const double rb = fabs(2.0 * (bx * ccx + by * ccy + bz * ccz) - sb) + 16.0 * SPH_EPS * (2.0 * (fabs(bx) * acx + fabs(by) * acy + fabs(bz) * acz) + sb);
    const double rc = fabs(2.0 * (cx * ccx + cy * ccy + cz * ccz) - sc) + 16.0 * SPH_EPS * (2.0 * (fabs(cx) * acx + fabs(cy) * acy + fabs(cz) * acz) + sc);
    const double adj2 = (sb * sc + sc * sa + sa * sb) * (1.0 + 1e-9);
    const double rho_c = 0.5 * sqrt(adj2 * (ra * ra + rb * rb + rc * rc)) / detLow * (1.0 + 1e-9);
    c[0] = p0.x + ccx; c[1] = p0.y + ccy; c[2] = p0.z + ccz;
    // |c - c*| <= rho_c + rounding of the three additions; true radius <= |p0 - c| + that
    const double rho = rho_c + 4.0 * SPH_EPS * (fabs(p0.x) + acx + fabs(p0.y) + acy + fabs(p0.z) + acz);
    const double rhat = sqrt(ccx * ccx + ccy * ccy + ccz * ccz);
    Rout = rhat * (1.0 + 8.0 * SPH_EPS) + 2.0 * rho;
    return Rout < 1e150 && rho >= 0.0;
}
VOR_HD bool sphere_ball_d(const double2 &p0, const double2 &p1, const double2 &p2, double c[3], double &Rout) {
    const double ax = p1.x - p0.x, ay = p1.y - p0.y;
    const double bx = p2.x - p0.x, by = p2.y - p0.y;
    const double det = ax * by - ay * bx;
    const double detLow = fabs(det) - 4.0 * SPH_EPS * (fabs(ax * by) + fabs(ay * bx));
    if (!(detLow > 0.0)) return false;
    const double sa = ax * ax + ay * ay, sb = bx * bx + by * by;
    const double inv = 0.5 / det;
    const double ccx = (sa * by - sb * ay) * inv;
    const double ccy = (sb * ax - sa * bx) * inv;
    const double acx = fabs(ccx), acy = fabs(ccy);
    const double ra = fabs(2.0 * (ax * ccx + ay * ccy) - sa) + 16.0 * SPH_EPS * (2.0 * (fabs(ax) * acx + fabs(ay) * acy) + sa);
    const double rb = fabs(2.0 * (bx * ccx + by * ccy) - sb) + 16.0 * SPH_EPS * (2.0 * (fabs(bx) * acx + fabs(by) * acy) + sb);
    const double adj2 = ax * ax + ay * ay + bx * bx + by * by;
    const double rho_c = 0.5 * sqrt(adj2) * sqrt(ra * ra + rb * rb) / detLow * (1.0 + 1e-9);
    c[0] = p0.x + ccx; c[1] = p0.y + ccy; c[2] = 0.0;
    const double rho = rho_c + 4.0 * SPH_EPS * (fabs(p0.x) + acx + fabs(p0.y) + acy);
    const double rhat = sqrt(ccx * ccx + ccy * ccy);
    Rout = rhat * (1.0 + 8.0 * SPH_EPS) + 2.0 * rho;
    return Rout < 1e150 && rho >= 0.0;
}

// +1: certainly strictly inside, -1: certainly not strictly inside, 0: undecided (ask the determinant)
VOR_HD int sphere_test(float cx, float cy, float cz, float rin2, float rout2, double qx, double qy, double qz) {
    const double dx = qx - (double)cx, dy = qy - (double)cy, dz = qz - (double)cz;
    const double d2 = dx * dx + dy * dy + dz * dz;
    if (d2 < (double)rin2) return 1;
    if (d2 > (double)rout2) return -1;
    return 0;
}

VOR_HD int f2i(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_int(f);
#else
    int i; memcpy(&i, &f, 4); return i;
#endif
}
VOR_HD float i2f(int i) {
#ifdef __CUDA_ARCH__
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}

} // namespace vor
