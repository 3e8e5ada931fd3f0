// Exact geometric predicates on the 1e-6 Angstrom lattice, for sm_100a device code.
//
// The reference prints every coordinate with "%.6f" before qhull sees it (reference
// src/voronoi.c:130), so Delaunay sites are integers k (|k| < 2^35) in units of 1e-6 A. They are
// kept as doubles holding those integers exactly: differences are exact, which is what the
// Shewchuk-style forward error bounds of the filtered stage assume. When the filter cannot decide,
// the determinant is evaluated exactly in integer arithmetic (orient3d: 128 bit; insphere: 256 bit).
// This replaces qhull's epsilon-tolerant hyperplane tests (reference src/qhull/src/libqhull/geom.c
// qh_distplane, poly2.c) with exact ones; on inputs where qhull merges nothing the tetrahedra are
// identical (tests/test_oracle_vs_reference.py::test_oracle_delaunay_equals_qhull).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FPK_HD __host__ __device__ __forceinline__
#define FPK_HDN __host__ __device__ __noinline__
#else
#define FPK_HD inline
#define FPK_HDN
#endif

namespace fpk {

typedef __int128 i128;
typedef unsigned __int128 u128;

FPK_HD double dabs(double x) { return x < 0 ? -x : x; }

struct I256 { uint64_t w[4]; };

FPK_HD void i256_add_prod(I256 &acc, i128 a, i128 b, bool negate)
{
    bool neg = negate;
    u128 ua, ub;
    if (a < 0) { ua = (u128)(-a); neg = !neg; } else ua = (u128)a;
    if (b < 0) { ub = (u128)(-b); neg = !neg; } else ub = (u128)b;
    uint64_t a0 = (uint64_t)ua, a1 = (uint64_t)(ua >> 64), b0 = (uint64_t)ub, b1 = (uint64_t)(ub >> 64);
    u128 p00 = (u128)a0 * b0, p01 = (u128)a0 * b1, p10 = (u128)a1 * b0, p11 = (u128)a1 * b1;
    uint64_t r0, r1, r2, r3;
    r0 = (uint64_t)p00;
    u128 mid = (p00 >> 64) + (uint64_t)p01 + (uint64_t)p10;
    r1 = (uint64_t)mid;
    u128 hi = (mid >> 64) + (p01 >> 64) + (p10 >> 64) + (uint64_t)p11;
    r2 = (uint64_t)hi;
    r3 = (uint64_t)((hi >> 64) + (p11 >> 64));
    if (neg) {
        r0 = ~r0; r1 = ~r1; r2 = ~r2; r3 = ~r3;
        r0 += 1;
        if (r0 == 0) { r1 += 1; if (r1 == 0) { r2 += 1; if (r2 == 0) r3 += 1; } }
    }
    u128 s = (u128)acc.w[0] + r0; acc.w[0] = (uint64_t)s;
    s = (u128)acc.w[1] + r1 + (uint64_t)(s >> 64); acc.w[1] = (uint64_t)s;
    s = (u128)acc.w[2] + r2 + (uint64_t)(s >> 64); acc.w[2] = (uint64_t)s;
    s = (u128)acc.w[3] + r3 + (uint64_t)(s >> 64); acc.w[3] = (uint64_t)s;
}

FPK_HD int i256_sign(const I256 &a)
{
    if (a.w[3] >> 63) return -1;
    return (a.w[0] | a.w[1] | a.w[2] | a.w[3]) ? 1 : 0;
}

// exact determinant det[a-d; b-d; c-d] (Shewchuk's orient3d sign convention)
FPK_HDN i128 orient3d_exact_det(const double *a, const double *b, const double *c, const double *d)
{
    i128 adx = (long long)(a[0] - d[0]), ady = (long long)(a[1] - d[1]), adz = (long long)(a[2] - d[2]);
    i128 bdx = (long long)(b[0] - d[0]), bdy = (long long)(b[1] - d[1]), bdz = (long long)(b[2] - d[2]);
    i128 cdx = (long long)(c[0] - d[0]), cdy = (long long)(c[1] - d[1]), cdz = (long long)(c[2] - d[2]);
    return adx * (bdy * cdz - bdz * cdy) + bdx * (cdy * adz - cdz * ady) + cdx * (ady * bdz - adz * bdy);
}

// sign of orient3d: filtered double evaluation, exact 128-bit fallback
FPK_HD int orient3d(const double *a, const double *b, const double *c, const double *d)
{
    double adx = a[0] - d[0], ady = a[1] - d[1], adz = a[2] - d[2];
    double bdx = b[0] - d[0], bdy = b[1] - d[1], bdz = b[2] - d[2];
    double cdx = c[0] - d[0], cdy = c[1] - d[1], cdz = c[2] - d[2];
    double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy, cdxady = cdx * ady, adxcdy = adx * cdy, adxbdy = adx * bdy, bdxady = bdx * ady;
    double det = adz * (bdxcdy - cdxbdy) + bdz * (cdxady - adxcdy) + cdz * (adxbdy - bdxady);
    double permanent = (dabs(bdxcdy) + dabs(cdxbdy)) * dabs(adz) + (dabs(cdxady) + dabs(adxcdy)) * dabs(bdz) +
                       (dabs(adxbdy) + dabs(bdxady)) * dabs(cdz);
    // 2 x Shewchuk's o3derrboundA = (7 + 56 eps) eps, eps = 2^-53
    double errbound = 1.6e-15 * permanent;
    if (det > errbound) return 1;
    if (-det > errbound) return -1;
    i128 e = orient3d_exact_det(a, b, c, d);
    return e > 0 ? 1 : (e < 0 ? -1 : 0);
}

FPK_HDN int insphere_exact(const double *a, const double *b, const double *c, const double *d, const double *e)
{
    i128 aex = (long long)(a[0] - e[0]), aey = (long long)(a[1] - e[1]), aez = (long long)(a[2] - e[2]);
    i128 bex = (long long)(b[0] - e[0]), bey = (long long)(b[1] - e[1]), bez = (long long)(b[2] - e[2]);
    i128 cex = (long long)(c[0] - e[0]), cey = (long long)(c[1] - e[1]), cez = (long long)(c[2] - e[2]);
    i128 dex = (long long)(d[0] - e[0]), dey = (long long)(d[1] - e[1]), dez = (long long)(d[2] - e[2]);
    i128 ab = aex * bey - bex * aey, bc = bex * cey - cex * bey, cd = cex * dey - dex * cey;
    i128 da = dex * aey - aex * dey, ac = aex * cey - cex * aey, bd = bex * dey - dex * bey;
    i128 abc = aez * bc - bez * ac + cez * ab;
    i128 bcd = bez * cd - cez * bd + dez * bc;
    i128 cda = cez * da + dez * ac + aez * cd;
    i128 dab = dez * ab + aez * bd + bez * da;
    i128 alift = aex * aex + aey * aey + aez * aez, blift = bex * bex + bey * bey + bez * bez;
    i128 clift = cex * cex + cey * cey + cez * cez, dlift = dex * dex + dey * dey + dez * dez;
    I256 acc;
    acc.w[0] = acc.w[1] = acc.w[2] = acc.w[3] = 0;
    i256_add_prod(acc, dlift, abc, false);
    i256_add_prod(acc, clift, dab, true);
    i256_add_prod(acc, blift, cda, false);
    i256_add_prod(acc, alift, bcd, true);
    return i256_sign(acc);
}

// sign of insphere (> 0: e strictly inside the sphere of the positively oriented tet a,b,c,d); 0 = cospherical
FPK_HD int insphere(const double *a, const double *b, const double *c, const double *d, const double *e)
{
    double aex = a[0] - e[0], aey = a[1] - e[1], aez = a[2] - e[2];
    double bex = b[0] - e[0], bey = b[1] - e[1], bez = b[2] - e[2];
    double cex = c[0] - e[0], cey = c[1] - e[1], cez = c[2] - e[2];
    double dex = d[0] - e[0], dey = d[1] - e[1], dez = d[2] - e[2];
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
    double alift = aex * aex + aey * aey + aez * aez, blift = bex * bex + bey * bey + bez * bez;
    double clift = cex * cex + cey * cey + cez * cez, dlift = dex * dex + dey * dey + dez * dez;
    double det = (dlift * abc - clift * dab) + (blift * cda - alift * bcd);
    double az = dabs(aez), bz = dabs(bez), cz = dabs(cez), dz = dabs(dez);
    double pab = dabs(aexbey) + dabs(bexaey), pbc = dabs(bexcey) + dabs(cexbey), pcd = dabs(cexdey) + dabs(dexcey);
    double pda = dabs(dexaey) + dabs(aexdey), pac = dabs(aexcey) + dabs(cexaey), pbd = dabs(bexdey) + dabs(dexbey);
    double permanent = (pcd * bz + pbd * cz + pbc * dz) * alift + (pda * cz + pac * dz + pcd * az) * blift +
                       (pab * dz + pbd * az + pda * bz) * clift + (pbc * az + pac * bz + pab * cz) * dlift;
    // 2 x Shewchuk's isperrboundA = (16 + 224 eps) eps
    double errbound = 3.6e-15 * permanent;
    if (det > errbound) return 1;
    if (-det > errbound) return -1;
    return insphere_exact(a, b, c, d, e);
}

}  // namespace fpk
