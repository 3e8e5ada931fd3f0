// Shared device-side definitions of libfpocket_cuda (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/fpocket_cuda.h"

namespace fpk {

// ---- per-structure descriptor and the packed batch (all arrays device resident, SoA) -----------------
struct StructDev {
    int atom_off, natoms;      // pdb atoms (ligand-stripped structure): offset of the COORDINATES (ax, ay, az)
    int prop_off;              // offset of the per-atom properties (radius, electroneg, bfactor, ids, flags); mdpocket frames share one copy
    int w_off, natoms_w;       // pdb_w_lig atoms; natoms_w == 0 -> same arrays as pdb (same_pdb)
    int site_off, nsites;      // Delaunay sites (heavy atoms): site2atom[site_off + i] = atom index inside the structure
    int xlig_off, n_xlig;
    int same_pdb;
    float avg_b, min_b, max_b;
    int sph_off, sph_cap;      // this structure's region in the sphere arrays
    int tet_off, tet_cap;      // region in tets_out (only with want_tets)
};

enum { CNT_TETS = 0, CNT_SPHERES = 1, CNT_CLUSTERS = 2, CNT_POCKETS = 3, CNT_STATUS = 4, CNT_SUBROUNDS = 5, CNT_RETRIES = 6,
       CNT_NBIG = 7, CNT_WALK = 8, CNT_LEVELS = 9, CNT_ATTEMPTS = 10, CNT_HOPS = 11,
       CNT_XP = 12,      // 1: the last cluster is dpocket's explicit pocket (spheres near the known ligand), not a pocket of the partition
       CNT_NX = 13,      // its number of spheres
       CNT_STRIDE = 16 };

struct BatchDev {
    int nstruct;
    const StructDev *S;
    const float *ax, *ay, *az, *arad, *aen, *abf;
    const int *aid, *ares;
    const int8_t *aaa;
    const uint8_t *afl;
    const float *wx, *wy, *wz, *wrad, *wen;
    const uint8_t *wfl;
    const int *site2atom;
    const float *xlig;
    fpk_params prm;
    // spheres (region [sph_off, sph_off+sph_cap) per structure), canonical order
    float4 *s_xyzr;
    float *s_bary;             // 3 per sphere
    int4 *s_neigh;             // atom indices (inside the structure), order = ascending site index
    int4 *s_key;               // the 4 site indices ascending (identity of the sphere)
    uint8_t *s_type;
    int *s_cluster;
    // clusters: region [sph_off + k, ...) with sph_cap + 1 entries for offsets
    int *cl_off;               // [sph_off + k + c]
    int *cl_sph;               // [sph_off + i]
    int *counts;               // [nstruct * CNT_STRIDE]
    const int *work_order;     // structures, largest first
    int *work_counter;
    int *tets_out;             // optional
    int k1_max_sites;          // structures with more sites were triangulated by the grid-wide kernel (k_delaunay_grid.cuh): K1 skips them
};

// ---- arithmetic mirrored from the reference (gcc x86-64, no FMA): compile with -fmad=false ---------------
// calc.c:51-58 dist(): float differences/products/sums; sqrt of the float sum through double == sqrt.rn.f32
__device__ __forceinline__ float f_dist(float x1, float y1, float z1, float x2, float y2, float z2)
{
    float xdif = x1 - x2, ydif = y1 - y2, zdif = z1 - z2;
    float s = (xdif * xdif) + (ydif * ydif) + (zdif * zdif);
    return __fsqrt_rn(s);
}
// calc.c:76-83 ddist()
__device__ __forceinline__ float f_ddist(float x1, float y1, float z1, float x2, float y2, float z2)
{
    float xdif = x1 - x2, ydif = y1 - y2, zdif = z1 - z2;
    return (xdif * xdif) + (ydif * ydif) + (zdif * zdif);
}

// ---- block-wide helpers (blockDim.x <= 1024) ----------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// exclusive prefix sum of data[0..n) in place, all threads of the block; returns the total. `sm` = 33 ints of shared memory.
__device__ int block_exclusive_scan(int *data, int n, int *sm)
{
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = (nt + 31) >> 5;
    const int chunk = (n + nt - 1) / nt;
    const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    int s = 0;
    for (int i = lo; i < hi; i++) s += data[i];
    int inc = warp_incl_scan(s);
    if (lane == 31) sm[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < nw ? sm[lane] : 0;
        int wi = warp_incl_scan(w);
        sm[lane] = wi - w;
        if (lane == 31) sm[32] = wi;
    }
    __syncthreads();
    int run = sm[wid] + inc - s;
    int total = sm[32];
    for (int i = lo; i < hi; i++) { int v = data[i]; data[i] = run; run += v; }
    __syncthreads();
    return total;
}

// in-place bitonic sort of npad (power of two) 64-bit keys by the whole block
__device__ void block_bitonic_sort(unsigned long long *keys, int npad)
{
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npad; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = keys[i], b = keys[ixj];
                    bool up = ((i & k) == 0);
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
}

}  // namespace fpk
