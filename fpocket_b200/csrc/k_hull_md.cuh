// K4b: convex-hull volume of a pocket's sphere centres; K5: mdpocket grid scatter.
#pragma once
#include "k_delaunay.cuh"
#include "k_delaunay_grid.cuh"
#include "k_pockets.cuh"
#include "hull_warp.cuh"

namespace fpk {

// K4b, first kernel: one WARP per pocket, incremental exact hull in shared memory (hull_warp.cuh). Warps take chunks of 32
// clusters from a counter, every lane decides whether its cluster gets a hull volume at all (voronoi.c:1244: fewer than 10
// centres -> 0; clusters refine.c:124 always drops are never described), then the warp works through the chunk's pockets.
// A pocket outside the fast path's limits goes on the list the Delaunay kernel below consumes.
#ifndef FPK_HW_WARPS
#define FPK_HW_WARPS 2         // 22 KB of shared memory per warp: 2 x 5 blocks per SM measured best (profiles/r01_k4b_hull_capacity_ab.log)
#endif
#ifndef FPK_HW_MINBLOCKS
#define FPK_HW_MINBLOCKS 5
#endif
enum { HW_WARPS = FPK_HW_WARPS, HW_THREADS = 32 * FPK_HW_WARPS };
__global__ void __launch_bounds__(HW_THREADS, FPK_HW_MINBLOCKS) k_hull_warp(BatchDev B, PocketDev Pk, int *chunk_counter, int *fb_list, int *fb_count,
                                                                           int enable)
{
    extern __shared__ __align__(16) unsigned char hw_smem[];
    const int lane = threadIdx.x & 31;
    HullWarp &H = reinterpret_cast<HullWarp *>(hw_smem)[threadIdx.x >> 5];
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(chunk_counter, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= Pk.total_clusters) return;
        const int item = base + lane;
        int k = 0, off = 0, n = 0;
        bool want = false;
        if (item < Pk.total_clusters) {
            int lo = 0, hi = B.nstruct;
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (Pk.clbase[mid] <= item) lo = mid; else hi = mid; }
            k = lo;
            const int c = item - Pk.clbase[k];
            const int *cl_off = B.cl_off + B.S[k].sph_off + k;
            off = cl_off[c]; n = cl_off[c + 1] - off;
            const int *kc = B.counts + (size_t)k * CNT_STRIDE;
            const bool xp = kc[CNT_XP] == 1 && c == kc[CNT_CLUSTERS] - 1;      // dpocket's explicit pocket: always described
            want = !(n < 10 || (n < B.prm.min_pock_nb_asph && !xp));
        }
        unsigned todo = __ballot_sync(0xffffffffu, want);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int it = base + src, ks = __shfl_sync(0xffffffffu, k, src), os = __shfl_sync(0xffffffffu, off, src),
                      ns = __shfl_sync(0xffffffffu, n, src);
            bool fast = enable && ns <= HW_NCAP;
            if (fast) {
                const int so = B.S[ks].sph_off;
                const int *L = B.cl_sph + so + os;
                const float4 *sx = B.s_xyzr + so;
                const float4 v0 = sx[L[0]];
                const long long ox = __double2ll_rn(__dmul_rn((double)v0.x, 1e6)), oy = __double2ll_rn(__dmul_rn((double)v0.y, 1e6)),
                                oz = __double2ll_rn(__dmul_rn((double)v0.z, 1e6));
                bool wide = false;
                __syncwarp();
                for (int i = lane; i < ns; i += 32) {
                    const float4 v = sx[L[i]];
                    const long long rx = __double2ll_rn(__dmul_rn((double)v.x, 1e6)) - ox, ry = __double2ll_rn(__dmul_rn((double)v.y, 1e6)) - oy,
                                    rz = __double2ll_rn(__dmul_rn((double)v.z, 1e6)) - oz;
                    wide |= llabs(rx) >= HW_MAXREL || llabs(ry) >= HW_MAXREL || llabs(rz) >= HW_MAXREL;
                    H.px[i] = (int)rx; H.py[i] = (int)ry; H.pz[i] = (int)rz;
                }
                __syncwarp();
                if (__any_sync(0xffffffffu, wide)) fast = false;
                else {
                    i128 tot = 0;
                    if (hull_warp_run(H, ns, lane, &tot)) fast = false;
                    else if (lane == 0) Pk.desc[it].convex_hull_volume = hull_total_to_volume(tot);
                }
            }
            if (!fast && lane == 0) fb_list[atomicAdd(fb_count, 1)] = it;
        }
    }
}

// K4b, second kernel: the pockets the warp kernel handed over (more than HW_NCAP centres, degenerate, ...), one block each.
// Replaces reference src/voronoi.c:1235-1291 get_convex_hull_volume -> src/qhull/src/qconvex/qconvex.c:37 run_qconvex("FS")
// (text through temp files, a second qhull run per pocket). The hull volume equals the summed volume of the Delaunay
// tetrahedra of the same points, so the block-parallel exact Delaunay of delaunay.cuh is reused on the pocket's centres
// (on the 1e-6 lattice the reference's "%.6f" text puts them on, voronoi.c:1256-1262); tet volumes are summed EXACTLY
// (128-bit integers), so the result does not depend on thread timing.
#ifndef FPK_HULL_MINBLOCKS
#define FPK_HULL_MINBLOCKS 8
#endif
#ifndef FPK_HULL_THREADS
#define FPK_HULL_THREADS 64
#endif
enum { HULL_THREADS = FPK_HULL_THREADS };
__global__ void __launch_bounds__(HULL_THREADS, FPK_HULL_MINBLOCKS) k_hull_volume(BatchDev B, PocketDev Pk, const DelScratch *scratch, int *work_counter, int pi_cap, const int *fb_list,
                                                                                  const int *fb_count)
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
        if (tid == 0) { const int w = atomicAdd(work_counter, 1); s_item = w < *fb_count ? fb_list[w] : -1; }
        __syncthreads();
        const int item = s_item;
        if (item < 0) return;
        int lo = 0, hi = B.nstruct;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (Pk.clbase[mid] <= item) lo = mid; else hi = mid; }
        const int k = lo, c = item - Pk.clbase[k];
        const StructDev S = B.S[k];
        const int *cl_off = B.cl_off + S.sph_off + k;
        const int off = cl_off[c], n = cl_off[c + 1] - off;
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
            Pk.desc[item].convex_hull_volume = hull_total_to_volume(tot);
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
    // -S: the increments are floats (the pocket's drug score). Float atomics would make the sums depend on the order in which blocks
    // reach a cell, i.e. differ in the last bits from run to run; they are accumulated as 32.32 fixed-point integers instead (exact,
    // order-free) and converted to the float grids when those are read (k_md_fixed_to_float).
    long long *dens_fx, *freq_fx;
};
#define FPK_MD_FX_ONE 4294967296.0
__device__ __forceinline__ void md_add(float *grid, long long *fx, size_t g, float incre, long long incre_fx)
{
    if (fx) atomicAdd(reinterpret_cast<unsigned long long *>(fx) + g, (unsigned long long)incre_fx);
    else atomicAdd(&grid[g], incre);
}
__global__ void k_md_fixed_to_float(const long long *fx, float *grid, size_t ncell)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < ncell; i += (size_t)gridDim.x * blockDim.x)
        grid[i] = (float)((double)fx[i] / FPK_MD_FX_ONE);
}

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
            const long long incre_fx = G.dens_fx ? __double2ll_rn((double)incre * FPK_MD_FX_ONE) : 0;
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
                        md_add(G.dens, G.dens_fx, ((size_t)gx * ny + gy) * nz + gz, incre, incre_fx);
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
                        if (!(atomicOr(&mask[g >> 5], bit) & bit)) md_add(G.freq, G.freq_fx, g, incre, incre_fx);
                    }
                }
                // the centre cell is incremented once per POCKET, with the indices of its LAST sphere (mdpbase.c:164-166, :255-260)
                if (q == n - 1 && xidx < nx && yidx < ny && zidx < nz && xidx >= 0 && yidx >= 0 && zidx >= 0) {
                    size_t g = ((size_t)xidx * ny + yidx) * nz + zidx;
                    md_add(G.dens, G.dens_fx, g, incre, incre_fx);
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
                    if (!(atomicOr(&mask[g >> 5], bit) & bit)) md_add(G.freq, G.freq_fx, g, incre, incre_fx);
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

// ---------------------------------------------------------------------------------------------------------------------
// K7: mdpocket's characterize mode. Replaces, per snapshot, reference src/mdpocket.c:695-745 extract_wanted_vertices: the spheres of the
// SURVIVING pockets, in rank order and pocket order, listed once for EVERY point of the wanted zone within (float)M_MDP_CUBE_SIDE / 2.0 = 1.0
// of the centre (dist() <= 1.0: a sphere near three points is listed three times). The list is described as ONE pocket by
// get_pocket_contacted_atms + set_descriptors (mdpocket.c:446-453): here the entries become the spheres of a second batch ("zone batch":
// one structure per frame, one cluster each) that the hull and descriptor kernels process unchanged.
struct ZoneDev {
    const float *zone;         // [3 * nzone]
    int nzone;
    int *ecount;               // [nstruct] entries of every frame (pass 0)
    const int *ebase;          // [nstruct + 1] exclusive scan (pass 1)
    int *entry_sphere;         // [E] sphere index inside the frame
    int *entry_pocket;         // [E] 1-based rank of the pocket it belongs to (vertice->resid)
    float4 *z_xyzr; float *z_bary; int4 *z_neigh; uint8_t *z_type;      // [E] the zone batch's sphere arrays
    int *z_cl_sph, *z_cl_off;  // [E], [E + 2 nstruct + 2]
    int *z_counts;             // [nstruct * CNT_STRIDE]
    int *z_mcid;               // [nstruct]
};
__global__ void __launch_bounds__(256) k_zone_select(BatchDev B, PocketDev Pk, ZoneDev Z, int pass)
{
    __shared__ int s_w[8], s_base;
    const int k = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const StructDev S = B.S[k];
    const int *counts = B.counts + (size_t)k * CNT_STRIDE;
    const int npk = counts[CNT_STATUS] == FPK_OK ? counts[CNT_POCKETS] : 0;
    const int *cl_off = B.cl_off + S.sph_off + k, *cl_sph = B.cl_sph + S.sph_off;
    const int *rank = Pk.rank_to_cluster + Pk.sbase[k] + k;
    const float4 *sx = B.s_xyzr + S.sph_off;
    int nsurv = 0;
    for (int r = 0; r < npk; r++) nsurv += cl_off[rank[r] + 1] - cl_off[rank[r]];
    if (tid == 0) s_base = 0;
    __syncthreads();
    const int e0 = pass ? Z.ebase[k] : 0;
    for (int j0 = 0; j0 < nsurv; j0 += nt) {
        const int j = j0 + tid;
        int s = -1, cnt = 0, r = 0;
        if (j < nsurv) {
            int acc = 0;
            for (; r < npk; r++) { const int m = cl_off[rank[r] + 1] - cl_off[rank[r]]; if (j < acc + m) break; acc += m; }
            s = cl_sph[cl_off[rank[r]] + (j - acc)];
            const float4 v = sx[s];
            for (int z = 0; z < Z.nzone; z++)
                cnt += (double)f_dist(Z.zone[3 * z], Z.zone[3 * z + 1], Z.zone[3 * z + 2], v.x, v.y, v.z) <= (double)(float)2.0 / 2.0;
        }
        const int inc = warp_incl_scan(cnt);
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < wid; w++) before += s_w[w];
        if (pass && cnt > 0) {
            const float4 v = sx[s];
            for (int q = 0; q < cnt; q++) {
                const int loc = before + inc - cnt + q, e = e0 + loc;
                Z.entry_sphere[e] = s; Z.entry_pocket[e] = r + 1;
                Z.z_xyzr[e] = v; Z.z_neigh[e] = B.s_neigh[S.sph_off + s]; Z.z_type[e] = B.s_type[S.sph_off + s];
                Z.z_bary[3 * (size_t)e] = B.s_bary[3 * (size_t)(S.sph_off + s)]; Z.z_bary[3 * (size_t)e + 1] = B.s_bary[3 * (size_t)(S.sph_off + s) + 1];
                Z.z_bary[3 * (size_t)e + 2] = B.s_bary[3 * (size_t)(S.sph_off + s) + 2];
                Z.z_cl_sph[e] = loc;
            }
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < (nt >> 5); w++) t += s_w[w]; s_base += t; }
        __syncthreads();
    }
    if (tid == 0) {
        const int E = s_base;
        if (!pass) Z.ecount[k] = E;
        else {
            Z.z_cl_off[e0 + k] = 0; Z.z_cl_off[e0 + k + 1] = E;
            int *zc = Z.z_counts + (size_t)k * CNT_STRIDE;
            for (int q = 0; q < CNT_STRIDE; q++) zc[q] = 0;
            zc[CNT_SPHERES] = E; zc[CNT_CLUSTERS] = E > 0 ? 1 : 0; zc[CNT_STATUS] = E > 0 ? FPK_OK : FPK_ERR_NO_POCKETS;
            Z.z_mcid[k] = counts[CNT_CLUSTERS] + 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// K6: dpocket's ligand queries. Replaces reference src/neighbor.c:192 get_mol_vert_neigh, :313 get_mol_ctd_atm_neigh (interface
// search), :490 count_pocket_lig_vert_ovlp, :598 count_atm_prop_vert_neigh, src/atom.c:264 atm_corsp and the barycentre distance of
// src/dpocket.c:253-261. The reference finds the pairs by sweeping an x-sorted merged array (sort.c:86 get_sorted_list); the RESULTS
// are sets and counts of the strict "dist < d" relation (calc.c:51 dist in f32), which is what is computed here, one block per
// structure: lanes over spheres x ligand atoms, lists in ascending index order.
struct LigDev {
    const int *lig_off;        // [nstruct + 1] offsets into lig_idx / lig_mol
    const int *lig_idx;        // indices into the pdb_w_lig atom arrays (pdb arrays when there is no separate copy)
    const int *lig_mol;
    const int *name_key;       // [natoms_total] or nullptr
    int *xpos;                 // [sphcap_total + nstruct + 1] scratch of the selection, per structure at sph_off + k
    int *ipos;                 // [natoms_total + nstruct + 1] scratch, per structure at atom_off + k
    int *ilist;                // interface atoms (ascending) per structure at atom_off + k
    int *ni;                   // [nstruct]
    float *crit;               // [4 * total_clusters]: ovlp, dst, c4, c5 of the pocket ranked r of structure k at 4 * (clbase[k] + r)
    float *lig_vol;            // [nstruct] Monte-Carlo volume of the ligand (atom.c:156)
};
enum { LIG_MAXMOL = 64 };

// K6a (after K3): the explicit pocket = spheres within lig_interface_dist of a ligand atom, appended to the structure's cluster list as
// one more cluster (K4 / K4b then give it a descriptor record like any pocket, dpocket.c:350 set_descriptors), and the interface atoms.
__global__ void __launch_bounds__(256) k_ligand_select(BatchDev B, LigDev G)
{
    __shared__ int s_sm[40];
    const int k = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const StructDev S = B.S[k];
    int *counts = B.counts + (size_t)k * CNT_STRIDE;
    const int L0 = G.lig_off[k], nl = G.lig_off[k + 1] - L0;
    if (tid == 0) { G.ni[k] = 0; counts[CNT_XP] = 0; counts[CNT_NX] = 0; }
    if (nl <= 0 || counts[CNT_STATUS] != FPK_OK || B.prm.skip_clustering) return;
    const int ns = counts[CNT_SPHERES], na = S.natoms, nc = counts[CNT_CLUSTERS];
    const int W0 = S.natoms_w ? S.w_off : S.atom_off;
    const float *wx = S.natoms_w ? B.wx : B.ax, *wy = S.natoms_w ? B.wy : B.ay, *wz = S.natoms_w ? B.wz : B.az;
    const int *lig = G.lig_idx + L0;
    const float4 *sx = B.s_xyzr + S.sph_off;
    const int4 *sn = B.s_neigh + S.sph_off;
    int *xpos = G.xpos + S.sph_off + k, *ipos = G.ipos + S.atom_off + k, *ilist = G.ilist + S.atom_off + k;
    int *cl_off = B.cl_off + S.sph_off + k, *cl_sph = B.cl_sph + S.sph_off;
    const float dcrit = B.prm.lig_interface_dist;
    const bool method2 = B.prm.lig_interface_method == 2;
    for (int i = tid; i <= na; i += nt) ipos[i] = 0;
    __syncthreads();
    if (method2) {
        // dparams.h:44 M_INTERFACE_METHOD2 (dpocket.c:331-336, neighbor.c:67 get_mol_atm_neigh): the atoms of the complex with dist() < dcrit to
        // a ligand atom, the ligand's own atoms excluded. pdb is pdb_w_lig without the ligand, in the same order: index in pdb = index in
        // pdb_w_lig minus the ligand atoms before it.
        const int nwl = S.natoms_w ? S.natoms_w : S.natoms;
        for (int a = tid; a < nwl; a += nt) {
            int before = 0;
            bool isl = false;
            for (int l = 0; l < nl; l++) { before += lig[l] < a; isl |= lig[l] == a; }
            const int pi = a - before;
            if (isl || pi >= na) continue;
            const float x = wx[W0 + a], y = wy[W0 + a], z = wz[W0 + a];
            bool nearl = false;
            for (int l = 0; l < nl && !nearl; l++) nearl = f_dist(x, y, z, wx[W0 + lig[l]], wy[W0 + lig[l]], wz[W0 + lig[l]]) < dcrit;
            if (nearl) ipos[pi] = 1;
        }
    }
    for (int s = tid; s <= ns; s += nt) {
        int near = 0;
        if (s < ns) {
            const float4 v = sx[s];
            const int4 q = sn[s];
            for (int l = 0; l < nl; l++) {
                const float lx = wx[W0 + lig[l]], ly = wy[W0 + lig[l]], lz = wz[W0 + lig[l]];
                if (f_dist(v.x, v.y, v.z, lx, ly, lz) < dcrit) {
                    near = 1;
                    if (method2) continue;
                    const int a[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int j = 0; j < 4; j++)         // neighbor.h:30 M_INTERFACE_SEARCH_DIST 8.0
                        if ((double)f_dist(B.ax[S.atom_off + a[j]], B.ay[S.atom_off + a[j]], B.az[S.atom_off + a[j]], lx, ly, lz) < 8.0) ipos[a[j]] = 1;
                }
            }
        }
        xpos[s] = near;
    }
    __syncthreads();
    const int nx = block_exclusive_scan(xpos, ns + 1, s_sm);
    const int ni = block_exclusive_scan(ipos, na + 1, s_sm);
    const bool fits = nx > 0 && ns + nx <= S.sph_cap;
    if (fits) for (int s = tid; s < ns; s += nt) if (xpos[s + 1] != xpos[s]) cl_sph[ns + xpos[s]] = s;
    for (int a = tid; a < na; a += nt) if (ipos[a + 1] != ipos[a]) ilist[ipos[a]] = a;
    if (tid == 0) {
        G.ni[k] = ni;
        if (fits) { cl_off[nc + 1] = ns + nx; counts[CNT_CLUSTERS] = nc + 1; counts[CNT_XP] = 1; counts[CNT_NX] = nx; }
    }
}

// K6b (after K4c): per surviving pocket ovlp, dst, c4, c5 (one warp per pocket)
__global__ void __launch_bounds__(256) k_ligand(BatchDev B, PocketDev Pk, LigDev G)
{
    __shared__ int s_hit[8][LIG_MAXMOL], s_tot[8][LIG_MAXMOL];
    const int k = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nwarp = nt >> 5;
    const StructDev S = B.S[k];
    const int *counts = B.counts + (size_t)k * CNT_STRIDE;
    const int L0 = G.lig_off[k], nl = G.lig_off[k + 1] - L0;
    if (nl <= 0 || (counts[CNT_STATUS] != FPK_OK && counts[CNT_STATUS] != FPK_ERR_NO_POCKETS)) return;
    const int W0 = S.natoms_w ? S.w_off : S.atom_off;
    const float *wx = S.natoms_w ? B.wx : B.ax, *wy = S.natoms_w ? B.wy : B.ay, *wz = S.natoms_w ? B.wz : B.az;
    const int *lig = G.lig_idx + L0, *mol = G.lig_mol + L0;
    const float4 *sx = B.s_xyzr + S.sph_off;
    const int *ilist = G.ilist + S.atom_off + k;
    const int ni = G.ni[k];
    const float ccrit = B.prm.lig_crit_dist;
    const int npk = counts[CNT_STATUS] == FPK_OK ? counts[CNT_POCKETS] : 0;
    const int *cl_off = B.cl_off + S.sph_off + k, *cl_sph = B.cl_sph + S.sph_off;
    const int *rank = Pk.rank_to_cluster + Pk.sbase[k] + k;
    const fpk_desc *D = Pk.desc + Pk.clbase[k];
    int nmol = 0;
    for (int l = 0; l < nl; l++) nmol = max(nmol, mol[l] + 1);
    nmol = min(nmol, LIG_MAXMOL);
    if (wid == nwarp - 1) {
        // atom.c:156 get_mol_volume_ptr(ligand, nb_mcv_iter) (dpocket.c:245): box with the reference's else-if updates (one lane, in order),
        // samples over the lanes, first-hit loop over the ligand's atoms; Philox counter (sample, 0x11A, 0, 0), key mc_seed
        const float *wr = S.natoms_w ? B.wrad : B.arad;
        const int WQ0 = S.natoms_w ? S.w_off : S.prop_off;
        float xmin = 0, xmax = 0, ymin = 0, ymax = 0, zmin = 0, zmax = 0, t;
        for (int l = 0; l < nl; l++) {
            const float x = wx[W0 + lig[l]], y = wy[W0 + lig[l]], z = wz[W0 + lig[l]], rr = wr[WQ0 + lig[l]];
            if (l == 0) { xmin = x - rr; xmax = x + rr; ymin = y - rr; ymax = y + rr; zmin = z - rr; zmax = z + rr; }
            else {
                if (xmin > (t = x - rr)) xmin = t; else if (xmax < (t = x + rr)) xmax = t;
                if (ymin > (t = y - rr)) ymin = t; else if (ymax < (t = y + rr)) ymax = t;
                if (zmin > (t = z - rr)) zmin = t; else if (zmax < (t = z + rr)) zmax = t;
            }
        }
        const float vbox = (xmax - xmin) * (ymax - ymin) * (zmax - zmin);
        const int niter = B.prm.nb_mcv_iter;
        int nb_in = 0;
        for (int i = lane; i < niter; i += 32) {
            unsigned o[4];
            philox4x32((unsigned)i, 0x11Au, 0u, 0u, (unsigned)B.prm.mc_seed, (unsigned)(B.prm.mc_seed >> 32), o);
            const float xr = unif_from_bits(o[0] >> 1, xmin, xmax), yr = unif_from_bits(o[1] >> 1, ymin, ymax), zr = unif_from_bits(o[2] >> 1, zmin, zmax);
            bool in = false;
            for (int l = 0; l < nl && !in; l++) {
                const float xt = wx[W0 + lig[l]] - xr, yt = wy[W0 + lig[l]] - yr, zt = wz[W0 + lig[l]] - zr, rr = wr[WQ0 + lig[l]];
                in = (rr * rr) > (xt * xt + yt * yt + zt * zt);
            }
            nb_in += in;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) nb_in += __shfl_xor_sync(0xffffffffu, nb_in, d);
        if (lane == 0) G.lig_vol[k] = niter > 0 ? ((float)nb_in) / ((float)niter) * vbox : 0.0f;
    }
    for (int r = wid; r < npk; r += nwarp) {
        const int c = rank[r], off = cl_off[c], n = cl_off[c + 1] - off;
        const int *Lp = cl_sph + off;
        // c5 (neighbor.c:490): share of the pocket's spheres within ccrit of some ligand atom
        int c5n = 0;
        for (int i = lane; i < n; i += 32) {
            const float4 v = sx[Lp[i]];
            bool hit = false;
            for (int l = 0; l < nl && !hit; l++) hit = f_dist(v.x, v.y, v.z, wx[W0 + lig[l]], wy[W0 + lig[l]], wz[W0 + lig[l]]) < ccrit;
            c5n += hit;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c5n += __shfl_xor_sync(0xffffffffu, c5n, d);
        // c4 (neighbor.c:598): per ligand molecule, share of its atoms with a sphere of the pocket within ccrit; the best molecule counts
        for (int m = lane; m < nmol; m += 32) { s_hit[wid][m] = 0; s_tot[wid][m] = 0; }
        __syncwarp();
        float dmin = 3e38f;
        for (int l = lane; l < nl; l += 32) {
            const float lx = wx[W0 + lig[l]], ly = wy[W0 + lig[l]], lz = wz[W0 + lig[l]];
            bool hit = false;
            for (int i = 0; i < n && !hit; i++) { const float4 v = sx[Lp[i]]; hit = f_dist(v.x, v.y, v.z, lx, ly, lz) < ccrit; }
            const int m = min(mol[l], LIG_MAXMOL - 1);
            atomicAdd(&s_tot[wid][m], 1);
            if (hit) atomicAdd(&s_hit[wid][m], 1);
            dmin = fminf(dmin, f_dist(lx, ly, lz, D[c].bary[0], D[c].bary[1], D[c].bary[2]));      // dpocket.c:253-261
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, d));
        __syncwarp();
        // ovlp (atom.c:264 atm_corsp): % of the interface atoms found among the pocket's atoms (same res_id and same id or name)
        const int *uat = Pk.cl_atoms + 4 * (size_t)(Pk.sbase[k] + off);
        const int U = D[c].n_atoms;
        int found = 0;
        for (int i = lane; i < ni; i += 32) {
            const int a = S.prop_off + ilist[i];
            bool f = false;
            for (int j = 0; j < U && !f; j++) {
                const int b = S.prop_off + uat[j];
                f = B.ares[b] == B.ares[a] && (B.aid[b] == B.aid[a] || (G.name_key && G.name_key[b] == G.name_key[a]));
            }
            found += f;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) found += __shfl_xor_sync(0xffffffffu, found, d);
        if (lane == 0) {
            float c4 = 0.0f;
            for (int m = 0; m < nmol; m++) {
                const float t = (float)s_hit[wid][m] / (float)s_tot[wid][m];
                if (t > c4) c4 = t;
            }
            float *o = G.crit + 4 * (size_t)(Pk.clbase[k] + r);
            o[0] = (ni <= 0 || U <= 0) ? 0.0f : (float)((double)(((float)found) / ((float)ni)) * 100.0);
            o[1] = dmin;
            o[2] = c4;
            o[3] = (float)c5n / (float)n;
        }
        __syncwarp();
    }
}

}  // namespace fpk
