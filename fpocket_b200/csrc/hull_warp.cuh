// K4b fast path: exact convex-hull volume of one pocket's sphere centres by ONE WARP (incremental hull in shared memory).
//
// Replaces reference src/voronoi.c:1235-1291 get_convex_hull_volume -> src/qhull/src/qconvex/qconvex.c:37 run_qconvex("FS").
// The block-parallel exact Delaunay of k_hull_volume (k_hull_md.cuh) computes the same number as the summed volume of the
// pocket's Delaunay tetrahedra; the hull alone is much less work: no insphere tests, no mesh in global memory, no block
// barriers. Points are the centres on the 1e-6 lattice (voronoi.c:1256-1262 prints them "%.6f"), kept as int32 offsets from
// the pocket's first centre; a face is three 10-bit point ids in one word. Per inserted point every lane tests its share of
// the faces (filtered f64 determinant, exact 128-bit integers when the filter cannot decide); STRICTLY visible faces
// (det < 0) are replaced by the cone over their horizon edges. With exact signs no zero-area face can appear (a point
// collinear with a horizon edge lies in the planes of both faces at that edge, so neither is strictly visible) and the
// strictly visible set of a convex polytope is a disc: V faces around I buried vertices have V + 2 - 2 I horizon edges. The
// volume is the EXACT integer sum over the final faces of det[a-q; b-q; c-q] (q = a hull vertex) = 6 x volume in lattice
// units, the same integer the Delaunay route sums, converted by the same expression: bit-identical results.
// Anything outside the fast path's limits (more than HW_NCAP centres, more than HW_FCAP faces, extent above 2^30 lattice
// units, a flat point set, a horizon that is not one loop) is handed to the Delaunay route of the same library, on the GPU.
//
// The routine is written once for both targets: on the device a warp runs it in lockstep, on the host (tests/hostemu) the
// same code runs as a "warp" of ONE lane, which checks the algorithm against the oracle without a GPU.
#pragma once
#include "predicates.cuh"

namespace fpk {

#ifndef FPK_HW_NCAP
#define FPK_HW_NCAP 1024
#endif
#ifndef FPK_HW_FCAP
#define FPK_HW_FCAP 1024
#endif
enum { HW_NCAP = FPK_HW_NCAP, HW_FCAP = FPK_HW_FCAP, HW_IDMASK = 1023, HW_MAXREL = 1 << 30 };
static_assert(HW_NCAP <= 1024, "face words hold 10-bit point ids");

struct HullWarp {                  // one warp's shared memory
    int px[HW_NCAP], py[HW_NCAP], pz[HW_NCAP];
    unsigned face[HW_FCAP];        // a | b << 10 | c << 20, interior on the positive side of det[a-q; b-q; c-q]
    unsigned nf[HW_FCAP];          // horizon edges of the current insertion: u | w << 10
    unsigned short vis[HW_FCAP];   // faces strictly visible from the current point, ascending
};

#if defined(__CUDA_ARCH__)
enum { HW_LANES = 32 };
__device__ __forceinline__ unsigned hw_ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
__device__ __forceinline__ void hw_sync() { __syncwarp(); }
__device__ __forceinline__ int hw_popc(unsigned m) { return __popc(m); }
// lane holding the largest value (ties: lowest index); every lane gets the index
__device__ __forceinline__ int hw_argmax(double v, int idx)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, d);
        int oi = __shfl_xor_sync(0xffffffffu, idx, d);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    return idx;
}
__device__ __forceinline__ i128 hw_sum128(i128 s)
{
    unsigned long long lo = (unsigned long long)s;
    long long hi = (long long)(s >> 64);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        unsigned long long ol = __shfl_xor_sync(0xffffffffu, lo, d);
        long long oh = __shfl_xor_sync(0xffffffffu, hi, d);
        i128 t = (i128)(((u128)(unsigned long long)hi << 64) | (u128)lo) + (i128)(((u128)(unsigned long long)oh << 64) | (u128)ol);
        lo = (unsigned long long)t; hi = (long long)(t >> 64);
    }
    return (i128)(((u128)(unsigned long long)hi << 64) | (u128)lo);
}
#else
enum { HW_LANES = 1 };
inline unsigned hw_ballot(bool p) { return p ? 1u : 0u; }
inline void hw_sync() {}
inline int hw_popc(unsigned m) { return __builtin_popcount(m); }
inline int hw_argmax(double, int idx) { return idx; }
inline i128 hw_sum128(i128 s) { return s; }
#endif

// exact det[a-q; b-q; c-q] on the int32 offsets
FPK_HDN i128 hw_det_exact(const HullWarp &H, int a, int b, int c, int qx, int qy, int qz)
{
    i128 adx = (long long)H.px[a] - qx, ady = (long long)H.py[a] - qy, adz = (long long)H.pz[a] - qz;
    i128 bdx = (long long)H.px[b] - qx, bdy = (long long)H.py[b] - qy, bdz = (long long)H.pz[b] - qz;
    i128 cdx = (long long)H.px[c] - qx, cdy = (long long)H.py[c] - qy, cdz = (long long)H.pz[c] - qz;
    return adx * (bdy * cdz - bdz * cdy) + bdx * (cdy * adz - cdz * ady) + cdx * (ady * bdz - adz * bdy);
}

// sign of det[a-q; b-q; c-q]: the filter of predicates.cuh orient3d() (differences of lattice integers are exact in f64)
FPK_HD int hw_side(const HullWarp &H, unsigned f, int qx, int qy, int qz)
{
    const int a = f & HW_IDMASK, b = (f >> 10) & HW_IDMASK, c = (f >> 20) & HW_IDMASK;
    double adx = (double)(H.px[a] - qx), ady = (double)(H.py[a] - qy), adz = (double)(H.pz[a] - qz);
    double bdx = (double)(H.px[b] - qx), bdy = (double)(H.py[b] - qy), bdz = (double)(H.pz[b] - qz);
    double cdx = (double)(H.px[c] - qx), cdy = (double)(H.py[c] - qy), cdz = (double)(H.pz[c] - qz);
    double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy, cdxady = cdx * ady, adxcdy = adx * cdy, adxbdy = adx * bdy, bdxady = bdx * ady;
    double det = adz * (bdxcdy - cdxbdy) + bdz * (cdxady - adxcdy) + cdz * (adxbdy - bdxady);
    double permanent = (dabs(bdxcdy) + dabs(cdxbdy)) * dabs(adz) + (dabs(cdxady) + dabs(adxcdy)) * dabs(bdz) +
                       (dabs(adxbdy) + dabs(bdxady)) * dabs(cdz);
    double errbound = 1.6e-15 * permanent;
    if (det > errbound) return 1;
    if (-det > errbound) return -1;
    i128 e = hw_det_exact(H, a, b, c, qx, qy, qz);
    return e > 0 ? 1 : (e < 0 ? -1 : 0);
}

FPK_HD unsigned hw_face(int a, int b, int c) { return (unsigned)a | ((unsigned)b << 10) | ((unsigned)c << 20); }

// Hull of the n points in H.px/py/pz (|offset| < HW_MAXREL, 4 <= n <= HW_NCAP), all lanes of the warp in lockstep.
// Returns 0 and *total = 6 x volume (lattice units, exact) on every lane, or 1 = outside the fast path (nothing decided).
FPK_HD int hull_warp_run(HullWarp &H, const int n, const int lane, i128 *total)
{
    const unsigned lt = (1u << lane) - 1u;
    // ---- a large first tetrahedron: lowest x, farthest from it, farthest from that line, farthest from that plane (all
    // approximate choices; only the exact test of the result matters) ------------------------------------------------------
    double bv = -1e300; int bi = 0;
    for (int i = lane; i < n; i += HW_LANES) { double v = -(double)H.px[i]; if (v > bv) { bv = v; bi = i; } }
    const int i0 = hw_argmax(bv, bi);
    const int x0 = H.px[i0], y0 = H.py[i0], z0 = H.pz[i0];
    bv = -1.0; bi = 0;
    for (int i = lane; i < n; i += HW_LANES) {
        double dx = (double)H.px[i] - x0, dy = (double)H.py[i] - y0, dz = (double)H.pz[i] - z0;
        double v = dx * dx + dy * dy + dz * dz;
        if (v > bv) { bv = v; bi = i; }
    }
    int i1 = hw_argmax(bv, bi);
    const double ux = (double)H.px[i1] - x0, uy = (double)H.py[i1] - y0, uz = (double)H.pz[i1] - z0;
    bv = -1.0; bi = 0;
    for (int i = lane; i < n; i += HW_LANES) {
        double dx = (double)H.px[i] - x0, dy = (double)H.py[i] - y0, dz = (double)H.pz[i] - z0;
        double cx = uy * dz - uz * dy, cy = uz * dx - ux * dz, cz = ux * dy - uy * dx;
        double v = cx * cx + cy * cy + cz * cz;
        if (v > bv) { bv = v; bi = i; }
    }
    int i2 = hw_argmax(bv, bi);
    const double vx = (double)H.px[i2] - x0, vy = (double)H.py[i2] - y0, vz = (double)H.pz[i2] - z0;
    const double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
    bv = -1.0; bi = 0;
    for (int i = lane; i < n; i += HW_LANES) {
        double dx = (double)H.px[i] - x0, dy = (double)H.py[i] - y0, dz = (double)H.pz[i] - z0;
        double v = dabs(nx * dx + ny * dy + nz * dz);
        if (v > bv) { bv = v; bi = i; }
    }
    const int i3 = hw_argmax(bv, bi);
    const i128 d0 = hw_det_exact(H, i0, i1, i2, H.px[i3], H.py[i3], H.pz[i3]);
    if (d0 == 0) return 1;                                    // flat (or the approximate choices failed): Delaunay route decides
    if (d0 < 0) { int t = i1; i1 = i2; i2 = t; }               // now det[i0-q; i1-q; i2-q] > 0 at q = i3
    int F = 4;
    if (lane == 0) {
        H.face[0] = hw_face(i0, i1, i2); H.face[1] = hw_face(i1, i3, i2);
        H.face[2] = hw_face(i0, i2, i3); H.face[3] = hw_face(i0, i3, i1);
    }
    hw_sync();
    // ---- insertions ------------------------------------------------------------------------------------------------
    for (int p = 0; p < n; p++) {
        if (p == i0 || p == i1 || p == i2 || p == i3) continue;
        const int qx = H.px[p], qy = H.py[p], qz = H.pz[p];
        int V = 0;
        for (int base = 0; base < F; base += HW_LANES) {
            const int f = base + lane;
            bool v = false;
            if (f < F) v = hw_side(H, H.face[f], qx, qy, qz) < 0;
            const unsigned m = hw_ballot(v);
            if (v) H.vis[V + hw_popc(m & lt)] = (unsigned short)f;
            V += hw_popc(m);
        }
        if (V == 0) continue;                                  // inside or on the hull
        hw_sync();
        int Hn = 0;
        for (int base = 0; base < 3 * V; base += HW_LANES) {
            const int e = base + lane;
            bool hz = false;
            unsigned u = 0, w = 0;
            if (e < 3 * V) {
                const unsigned g = H.face[H.vis[e / 3]];
                const int k = e % 3;
                const unsigned a = g & HW_IDMASK, b = (g >> 10) & HW_IDMASK, c = (g >> 20) & HW_IDMASK;
                u = k == 0 ? a : (k == 1 ? b : c);
                w = k == 0 ? b : (k == 1 ? c : a);
                hz = true;                                     // horizon edge: no visible face holds the reversed edge
                for (int j = 0; j < V; j++) {
                    const unsigned t = H.face[H.vis[j]];
                    const unsigned ta = t & HW_IDMASK, tb = (t >> 10) & HW_IDMASK, tc = (t >> 20) & HW_IDMASK;
                    if ((ta == w && tb == u) || (tb == w && tc == u) || (tc == w && ta == u)) { hz = false; break; }
                }
            }
            const unsigned m = hw_ballot(hz);
            if (hz) { const int idx = Hn + hw_popc(m & lt); if (idx < HW_FCAP) H.nf[idx] = u | (w << 10); }
            Hn += hw_popc(m);
        }
        // Euler: a disc of V triangles with I interior vertices (hull vertices the new point buries) has V + 2 - 2 I boundary edges
        if (Hn < 3 || Hn > V + 2 || ((V + 2 - Hn) & 1) || F + Hn - V > HW_FCAP) return 1;
        hw_sync();
        for (int j = lane; j < Hn; j += HW_LANES) {
            const int slot = j < V ? (int)H.vis[j] : F + (j - V);
            const unsigned e = H.nf[j];
            H.face[slot] = hw_face((int)(e & HW_IDMASK), (int)((e >> 10) & HW_IDMASK), p);
        }
        hw_sync();
        if (Hn < V && lane == 0) {                              // fewer new faces than holes: the tail moves into the holes left
            int lo = Hn, hi = V - 1, t = F - 1;
            while (lo <= hi) {
                if (t == (int)H.vis[hi]) { hi--; t--; }
                else { H.face[H.vis[lo]] = H.face[t]; lo++; t--; }
            }
        }
        F += Hn - V;
        hw_sync();
    }
    // ---- 6 x volume = sum over the faces of det[a-q; b-q; c-q], q = the first vertex ------------------------------------
    i128 sum = 0;
    for (int f = lane; f < F; f += HW_LANES) {
        const unsigned g = H.face[f];
        sum += hw_det_exact(H, (int)(g & HW_IDMASK), (int)((g >> 10) & HW_IDMASK), (int)((g >> 20) & HW_IDMASK), x0, y0, z0);
    }
    *total = hw_sum128(sum);
    return 0;
}

// the expression both hull routes use to turn the exact integer into the reference's float (sscanf "%f" of qhull's "%6.16g")
FPK_HD float hull_total_to_volume(i128 tot)
{
    long long h = (long long)(tot >> 64);
    unsigned long long l = (unsigned long long)tot;
    double vol = (double)h * 18446744073709551616.0 + (double)l;
    return (float)(vol / 6.0 * 1e-18);
}

}  // namespace fpk
