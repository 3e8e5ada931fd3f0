// K4b: convex-hull volume of a pocket's sphere centres; K5: mdpocket grid scatter.
#pragma once
#include "k_delaunay.cuh"
#include "k_pockets.cuh"

namespace fpk {

// Replaces reference src/voronoi.c:1235-1291 get_convex_hull_volume -> src/qhull/src/qconvex/qconvex.c:37 run_qconvex("FS")
// (text through temp files, a second qhull run per pocket). The hull volume equals the summed volume of the Delaunay
// tetrahedra of the same points, so the block-parallel exact Delaunay of delaunay.cuh is reused on the pocket's centres
// (on the 1e-6 lattice the reference's "%.6f" text puts them on, voronoi.c:1256-1262); tet volumes are summed EXACTLY
// (128-bit integers), so the result does not depend on thread timing.
enum { HULL_THREADS = 64 };
__global__ void __launch_bounds__(HULL_THREADS, 8) k_hull_volume(BatchDev B, PocketDev Pk, const DelScratch *scratch, int *work_counter, int pi_cap)
{
    __shared__ int s_item;
    BlockDel &BD = blk();
    DelaunayCtx &C = BD.C;
    __shared__ double s_box[6];
    __shared__ double s_bmin[HULL_THREADS / 32][3], s_bmax[HULL_THREADS / 32][3];
    __shared__ long long s_hi[HULL_THREADS];
    __shared__ unsigned long long s_lo[HULL_THREADS];
    const int tid = threadIdx.x, nt = blockDim.x;
    const DelScratch &X = scratch[blockIdx.x];
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= Pk.total_clusters) return;
        int lo = 0, hi = B.nstruct;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (Pk.clbase[mid] <= item) lo = mid; else hi = mid; }
        const int k = lo, c = item - Pk.clbase[k];
        const StructDev S = B.S[k];
        const int *cl_off = B.cl_off + S.sph_off + k;
        const int off = cl_off[c], n = cl_off[c + 1] - off;
        if (n < 10 || n < B.prm.min_pock_nb_asph) continue;     // voronoi.c:1244: volume 0 (desc is zero-initialised); always-dropped clusters (refine.c:124) too
        const int *L = B.cl_sph + S.sph_off + off;
        const float4 *sx = B.s_xyzr + S.sph_off;
        for (int i = tid; i < n; i += nt) {
            float4 v = sx[L[i]];
            X.P[3 * (size_t)i] = (double)__double2ll_rn(__dmul_rn((double)v.x, 1e6));
            X.P[3 * (size_t)i + 1] = (double)__double2ll_rn(__dmul_rn((double)v.y, 1e6));
            X.P[3 * (size_t)i + 2] = (double)__double2ll_rn(__dmul_rn((double)v.z, 1e6));
        }
        __syncthreads();
        block_bbox(X.P, n, s_box, s_bmin, s_bmax);
        const int err = block_delaunay(pi_cap, X, n, s_box);
        if (err) { if (tid == 0) Pk.desc[item].convex_hull_volume = -1.0f; continue; }      // voronoi.c:1276-1279
        i128 sum = 0;
        const int ntet = C.cnt[C_NTET];
        for (int t = tid; t < ntet; t += nt) {
            int4 q = reinterpret_cast<const int4 *>(X.tet)[2 * (size_t)t];
            if (q.x < 0 || q.y < 0 || q.z < 0 || q.w < 0) continue;
            sum += orient3d_exact_det(X.P + 3 * (size_t)q.x, X.P + 3 * (size_t)q.y, X.P + 3 * (size_t)q.z, X.P + 3 * (size_t)q.w);
        }
        s_hi[tid] = (long long)(sum >> 64);
        s_lo[tid] = (unsigned long long)sum;
        __syncthreads();
        if (tid == 0) {
            i128 tot = 0;
            for (int t = 0; t < nt; t++) tot += (i128)(((u128)(unsigned long long)s_hi[t] << 64) | (u128)s_lo[t]);
            long long h = (long long)(tot >> 64);
            unsigned long long l = (unsigned long long)tot;
            double vol = (double)h * 18446744073709551616.0 + (double)l;
            Pk.desc[item].convex_hull_volume = (float)(vol / 6.0 * 1e-18);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K5. Replaces reference src/mdpbase.c:108-169 calculate_md_dens_grid, :188-262 update_md_grid and :275 reset_grid
// (a full-grid clear per frame) for a batch of frames: one block per frame, one thread per kept sphere walks its stencil.
// The per-frame "already counted" grid (refgrid) becomes a per-block bit mask.
struct MdGrid {
    float origin[3];
    int nx, ny, nz;
    float *dens, *freq;
    unsigned *mask;            // [gridDim.x * nwords]
    int nwords;
    int use_drug_score;
};

__device__ __forceinline__ int md_idx(int o, float v, float org, float res, int ctr)
{
    if (o < 0) return (int)ceil((double)(((float)o + v - org) / res));
    if (o > 0) return (int)floor((double)(((float)o + v - org) / res));
    return ctr;
}

__global__ void __launch_bounds__(256) k_md_scatter(BatchDev B, PocketDev Pk, MdGrid G, unsigned long long *nscattered)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    unsigned *mask = G.mask + (size_t)blockIdx.x * G.nwords;
    const float res = 1.0f;                                   // M_MDP_GRID_RESOLUTION (mdpbase.h:41)
    const int nx = G.nx, ny = G.ny, nz = G.nz;
    for (int k = blockIdx.x; k < B.nstruct; k += gridDim.x) {
        const int *counts = B.counts + (size_t)k * CNT_STRIDE;
        __syncthreads();
        if (counts[CNT_STATUS] != FPK_OK) continue;
        const StructDev S = B.S[k];
        const int npk = counts[CNT_POCKETS];
        const int *cl_off = B.cl_off + S.sph_off + k, *cl_sph = B.cl_sph + S.sph_off;
        const int *rank = Pk.rank_to_cluster + Pk.sbase[k] + k;
        const fpk_desc *D = Pk.desc + Pk.clbase[k];
        const float4 *sx = B.s_xyzr + S.sph_off;
        for (int w = tid; w < G.nwords; w += nt) mask[w] = 0u;          // reset_grid(refgrid), mdpbase.c:275
        __syncthreads();
        for (int r = 0; r < npk; r++) {                                 // pockets in rank order: "first toucher" is per pocket
            const int c = rank[r];
            const float incre = G.use_drug_score ? D[c].drug_score : 1.0f;
            const int off = cl_off[c], n = cl_off[c + 1] - off;
            for (int q = tid; q < n; q += nt) {
                const float4 v = sx[cl_sph[off + q]];
                const int xidx = (int)roundf((v.x - G.origin[0]) / res), yidx = (int)roundf((v.y - G.origin[1]) / res),
                          zidx = (int)roundf((v.z - G.origin[2]) / res);
                // density: 26-neighbourhood (mdpbase.c:119: s = (int)((2.0/2)/res) = 1)
                for (int sxi = -1; sxi <= 1; sxi++) for (int syi = -1; syi <= 1; syi++) for (int szi = -1; szi <= 1; szi++) {
                    if ((sxi | syi | szi) == 0) continue;
                    int gx = md_idx(sxi, v.x, G.origin[0], res, xidx), gy = md_idx(syi, v.y, G.origin[1], res, yidx),
                        gz = md_idx(szi, v.z, G.origin[2], res, zidx);
                    if (((gx != xidx) || (gy != yidx) || (gz != zidx)) && gx >= 0 && gy >= 0 && gz >= 0 && gx < nx && gy < ny && gz < nz)
                        atomicAdd(&G.dens[((size_t)gx * ny + gy) * nz + gz], incre);
                }
                // frequency: half-width (int)((ray/2)/res) (mdpbase.c:218), each cell at most once per frame
                const int s = (int)(((double)v.w / 2.0) / (double)res);
                for (int sxi = -s; sxi <= s; sxi++) for (int syi = -s; syi <= s; syi++) for (int szi = -s; szi <= s; szi++) {
                    if ((sxi | syi | szi) == 0) continue;
                    int gx = md_idx(sxi, v.x, G.origin[0], res, xidx), gy = md_idx(syi, v.y, G.origin[1], res, yidx),
                        gz = md_idx(szi, v.z, G.origin[2], res, zidx);
                    if (((gx != xidx) || (gy != yidx) || (gz != zidx)) && gx >= 0 && gy >= 0 && gz >= 0 && gx < nx && gy < ny && gz < nz) {
                        size_t g = ((size_t)gx * ny + gy) * nz + gz;
                        unsigned bit = 1u << (g & 31);
                        if (!(atomicOr(&mask[g >> 5], bit) & bit)) atomicAdd(&G.freq[g], incre);
                    }
                }
                // the centre cell is incremented once per POCKET, with the indices of its LAST sphere (mdpbase.c:164-166, :255-260)
                if (q == n - 1 && xidx < nx && yidx < ny && zidx < nz && xidx >= 0 && yidx >= 0 && zidx >= 0) {
                    size_t g = ((size_t)xidx * ny + yidx) * nz + zidx;
                    atomicAdd(&G.dens[g], incre);
                }
            }
            __syncthreads();
            // the frequency centre cell comes after every sphere of the pocket (update_md_grid: outside the sphere loop)
            if (tid == 0 && n > 0) {
                const float4 v = sx[cl_sph[off + n - 1]];
                const int xidx = (int)roundf((v.x - G.origin[0]) / res), yidx = (int)roundf((v.y - G.origin[1]) / res),
                          zidx = (int)roundf((v.z - G.origin[2]) / res);
                if (xidx < nx && yidx < ny && zidx < nz && xidx >= 0 && yidx >= 0 && zidx >= 0) {
                    size_t g = ((size_t)xidx * ny + yidx) * nz + zidx;
                    unsigned bit = 1u << (g & 31);
                    if (!(atomicOr(&mask[g >> 5], bit) & bit)) atomicAdd(&G.freq[g], incre);
                }
                atomicAdd(nscattered, (unsigned long long)n);
            }
            __syncthreads();
        }
    }
}

// init_md_grid (mdpbase.c:444-502) from float_get_min_max_from_pockets (:403-416) over ALL kept spheres of one frame
__global__ void k_md_geometry(BatchDev B, int k, float *out /* origin[3], dims[3] as floats */)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const StructDev S = B.S[k];
    const int ns = B.counts[(size_t)k * CNT_STRIDE + CNT_SPHERES];
    const float4 *sx = B.s_xyzr + S.sph_off;
    float mn[3] = {50000.f, 50000.f, 50000.f}, mx[3] = {-50000.f, -50000.f, -50000.f};
    for (int z = 0; z < ns; z++) {
        float4 v = sx[z];
        float c[3] = {v.x, v.y, v.z};
        for (int q = 0; q < 3; q++) {
            if (mn[q] > c[q]) mn[q] = c[q];
            else if (mx[q] < c[q]) mx[q] = c[q];
        }
    }
    const float span = (float)(2.0 / 2.0), res = 1.0f;
    for (int q = 0; q < 3; q++) {
        out[3 + q] = (float)(int)((float)1 + (float)((int)((double)mx[q] + 30. * (double)span - (double)mn[q])) / res);
        out[q] = (float)((double)mn[q] - 15. * (double)span);
    }
}

}  // namespace fpk
