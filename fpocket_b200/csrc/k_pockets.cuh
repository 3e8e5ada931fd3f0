// K3 (clustering), K4 (per-pocket descriptors) and K4c (normalise / score / drop / rank).
//
// K3 replaces reference src/clusterlib.c:3886 treecluster('s','e') -> :3514 pslcluster (O(S^2) SLINK),
//    :3235 cuttree_distance, src/voronoi.c:260 transferClustersToVertices, src/pocket.c:77 assign_pockets,
//    src/refine.c:58 apply_clustering and :192 reIndexPockets by ONE pass: uniform cell grid ->
//    lock-free union-find over the  d <= clust_max_dist  graph (the single-linkage cut IS its connected
//    components) -> numbering by lowest sphere index -> ordered segment fill.
// K4 replaces src/pocket.c:575 set_pockets_descriptors: :1280 get_pocket_contacted_atms,
//    src/descriptors.c:158-513, src/asa.c:84-454, src/voronoi.c:1150-1233 (Monte-Carlo volume).
// K4c replaces src/pocket.c:664 set_normalized_descriptors, src/pscoring.c:241,325,
//    src/refine.c:114 dropSmallNpolarPockets and src/psorting.c:64 sort_pockets.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace fpk {

struct PocketDev {            // cluster-level arrays, sized after K3 (exact number of clusters known)
    const int *clbase;        // [nstruct + 1] exclusive scan of n_clusters
    const int *sbase;         // [nstruct + 1] exclusive scan of n_spheres (compact sphere numbering)
    fpk_desc *desc;           // [clbase[nstruct]]
    // sphere-level scratch / outputs, region of cluster c of structure k: R = sph_off + cl_off[c]
    int *cl_atoms;            // [4*R ...) unique contacted atoms (index inside the structure), first-seen order
    int *ent_u;               // [4*R ...) entry (sphere i, slot j) -> position of its atom in cl_atoms
    int *touch_off;           // [5*R ...) CSR offsets per unique atom
    int *touch_list;          // [4*R ...) spheres (local index) touching the atom
    int *acc;                 // [16*R ...) 4 int accumulators per unique atom: nn14, nn22, napol, npol
    float *fx_val;            // [16*R ...) per unique atom: prob, a0, dA, flag(abpa)
    int *fx_who;              // [2*natoms_total] per atom: highest cluster that wrote prob / abpa
    float *atom_fx;           // [4*natoms_total] result
    int *rank_to_cluster;     // [sph_off ...) per structure
    int total_clusters;
    const int *xp_ilist;      // dpocket: interface atoms of structure k at atom_off + k (ascending), count xp_ni[k]; nullptr without ligands
    const int *xp_ni;
    const int *sph_uid;       // nullptr, or per sphere of the batch its identity: mdpocket's zone list holds the same sphere several times, and
                              // descriptors.c:201 skips a sphere's own copies by POINTER (vc != vcur), not by position
    const int *mc_id;         // [nstruct] or nullptr: the cluster number the Monte-Carlo stream of structure k is keyed with instead of the
                              // cluster's own (mdpocket's zone record, k_hull_md.cuh K7: it is drawn as cluster n_clusters + 1 of its frame)
};

struct ClusterScratch {       // per resident block of K3
    int *cellcnt;             // [CELLMAX + 1]
    int *cellsph;             // [scap]
    int *tmp;                 // [scap + 1]
    int scap;
};
enum { CELLMAX = 1 << 18 };

__device__ __forceinline__ int uf_find(int *par, int i)
{
    int p = par[i];
    while (p != i) {
        int g = par[p];
        if (g != p) par[i] = g;
        i = p; p = g;
    }
    return i;
}
__device__ __forceinline__ void uf_union(int *par, int a, int b)
{
    for (;;) {
        a = uf_find(par, a); b = uf_find(par, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = atomicMin(&par[a], b);          // hook the larger root under the smaller one
        if (old == a) return;
        a = old;
    }
}

// the clustering distance of two sphere centres (f32 promoted to f64 as prepare_vertices_for_cluster_lib does, voronoi.c:283):
// -e e  clusterlib.c:933-1018 euclid      sqrt(sum of squares)
// -e b  clusterlib.c:1022-1083 cityblock  (|dx| + |dy| + |dz|) / 3 - the MEAN of the absolute differences (weights 1, tweight 3)
__device__ __forceinline__ double pair_distance(bool cityblock, double xi, double yi, double zi, const float4 &w)
{
    double result = 0.;
    if (cityblock) {
        double tweight = 0;
        double term = xi - (double)w.x; result = result + 1.0 * fabs(term); tweight += 1.0;
        term = yi - (double)w.y; result = result + 1.0 * fabs(term); tweight += 1.0;
        term = zi - (double)w.z; result = result + 1.0 * fabs(term); tweight += 1.0;
        return result / tweight;
    }
    double term = xi - (double)w.x; result += term * term;
    term = yi - (double)w.y; result += term * term;
    term = zi - (double)w.z; result += term * term;
    return sqrt(result);
}

__global__ void __launch_bounds__(256) k_cluster(BatchDev B, const ClusterScratch *scratch, int *work_counter)
{
    __shared__ int s_k, s_sm[40], s_dims[4];
    __shared__ float s_mn[8][3], s_mx[8][3], s_box[7];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const ClusterScratch &X = scratch[blockIdx.x];
    for (;;) {
        __syncthreads();
        if (tid == 0) s_k = atomicAdd(work_counter, 1);
        __syncthreads();
        if (s_k >= B.nstruct) return;
        const int k = B.work_order[s_k];
        const StructDev S = B.S[k];
        int *counts = B.counts + (size_t)k * CNT_STRIDE;
        if (S.nsites > B.k1_max_sites) continue;          // clustered by the whole GPU (k_cluster_grid)
        if (counts[CNT_STATUS] != FPK_OK) { if (tid == 0) counts[CNT_CLUSTERS] = 0; continue; }
        const int ns = counts[CNT_SPHERES];
        const float4 *sx = B.s_xyzr + S.sph_off;
        int *par = B.s_cluster + S.sph_off;
        int *cl_off = B.cl_off + S.sph_off + k, *cl_sph = B.cl_sph + S.sph_off;
        if (!B.prm.skip_clustering && ns < 2) {     // clusterlib.c:3969 -> fpocket.c:135-139
            if (tid == 0) { counts[CNT_CLUSTERS] = 0; counts[CNT_STATUS] = FPK_ERR_NO_SPHERES; }
            continue;
        }
        // no sphere at all with clustering skipped (-r / -P): assign_pockets returns NULL, search_pocket returns NULL (pocket.c:77, fpocket.c:160)
        if (ns == 0) { if (tid == 0) { counts[CNT_CLUSTERS] = 0; counts[CNT_STATUS] = FPK_ERR_NO_SPHERES; } continue; }
        if (B.prm.skip_clustering) {                 // explicit pocket: every sphere has the same resid -> one pocket
            for (int i = tid; i < ns; i += nt) { par[i] = 0; cl_sph[i] = i; }
            if (tid == 0) { cl_off[0] = 0; cl_off[1] = ns; counts[CNT_CLUSTERS] = 1; }
            continue;
        }
        // -C m | a | c: k_linkage (k_linkage.cuh) has already left par[i] = lowest sphere of i's cluster; only the numbering below remains
        const bool linked = B.prm.clustering_method == 'm' || B.prm.clustering_method == 'a' || B.prm.clustering_method == 'c';
        if (!linked) {
        // ---- bounding box of the centres --------------------------------------------------------------
        float mn[3] = {3e38f, 3e38f, 3e38f}, mx[3] = {-3e38f, -3e38f, -3e38f};
        for (int i = tid; i < ns; i += nt) {
            float4 v = sx[i];
            mn[0] = fminf(mn[0], v.x); mn[1] = fminf(mn[1], v.y); mn[2] = fminf(mn[2], v.z);
            mx[0] = fmaxf(mx[0], v.x); mx[1] = fmaxf(mx[1], v.y); mx[2] = fmaxf(mx[2], v.z);
            par[i] = i;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], d));
                mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], d));
            }
        if (lane == 0) for (int c = 0; c < 3; c++) { s_mn[wid][c] = mn[c]; s_mx[wid][c] = mx[c]; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < (nt >> 5); w++)
                for (int c = 0; c < 3; c++) { mn[c] = fminf(mn[c], s_mn[w][c]); mx[c] = fmaxf(mx[c], s_mx[w][c]); }
            // cell edge >= the largest per-axis difference a linked pair can have, so neighbours are within +-1 cell: the cut itself for the
            // euclidean distance, 3 x the cut for the city-block MEAN (|dx| <= |dx| + |dy| + |dz| = 3 x mean)
            float cell = (B.prm.distance_measure == 'b' ? 3.0f : 1.0f) * B.prm.clust_max_dist * 1.001f + 1e-4f;
            int d0, d1, d2;
            for (;;) {
                d0 = (int)((mx[0] - mn[0]) / cell) + 1; d1 = (int)((mx[1] - mn[1]) / cell) + 1; d2 = (int)((mx[2] - mn[2]) / cell) + 1;
                if ((long long)d0 * d1 * d2 <= CELLMAX) break;
                cell *= 1.26f;
            }
            s_dims[0] = d0; s_dims[1] = d1; s_dims[2] = d2;
            s_box[0] = mn[0]; s_box[1] = mn[1]; s_box[2] = mn[2]; s_box[3] = cell;
        }
        __syncthreads();
        const int d0 = s_dims[0], d1 = s_dims[1], d2 = s_dims[2], ncell = d0 * d1 * d2;
        const float bx = s_box[0], by = s_box[1], bz = s_box[2], cell = s_box[3];
        for (int i = tid; i <= ncell; i += nt) X.cellcnt[i] = 0;
        __syncthreads();
        for (int i = tid; i < ns; i += nt) {
            float4 v = sx[i];
            int cx = min(d0 - 1, max(0, (int)((v.x - bx) / cell))), cy = min(d1 - 1, max(0, (int)((v.y - by) / cell))),
                cz = min(d2 - 1, max(0, (int)((v.z - bz) / cell)));
            int cid = (cx * d1 + cy) * d2 + cz;
            X.tmp[i] = cid;
            atomicAdd(&X.cellcnt[cid], 1);
        }
        __syncthreads();
        block_exclusive_scan(X.cellcnt, ncell + 1, s_sm);
        // fill (slot order inside a cell is irrelevant); cursor = the upper end walking down
        for (int i = tid; i < ns; i += nt) {
            int cid = X.tmp[i];
            int slot = atomicAdd(&X.cellcnt[cid], 1);       // temporarily advances the start; restored below
            X.cellsph[slot] = i;
        }
        __syncthreads();
        // cellcnt[c] now holds the END of cell c (= old start of c+1); start of c = end of c-1
        // ---- union over the  d <= cut  graph (clusterlib.c:933 euclid in f64 on the f32 centres) --------------------------
        const double cut = (double)B.prm.clust_max_dist;
        const bool cityblock = B.prm.distance_measure == 'b';
        for (int i = tid; i < ns; i += nt) {
            float4 v = sx[i];
            int cid = X.tmp[i];
            int cz = cid % d2, cy = (cid / d2) % d1, cx = cid / (d2 * d1);
            double xi = (double)v.x, yi = (double)v.y, zi = (double)v.z;
            for (int ax = max(cx - 1, 0); ax <= min(cx + 1, d0 - 1); ax++)
                for (int ay = max(cy - 1, 0); ay <= min(cy + 1, d1 - 1); ay++)
                    for (int az = max(cz - 1, 0); az <= min(cz + 1, d2 - 1); az++) {
                        int c2 = (ax * d1 + ay) * d2 + az;
                        int lo = c2 ? X.cellcnt[c2 - 1] : 0, hi = X.cellcnt[c2];
                        for (int q = lo; q < hi; q++) {
                            int j = X.cellsph[q];
                            if (j >= i) continue;
                            float4 w = sx[j];
                            if (!(pair_distance(cityblock, xi, yi, zi, w) > cut)) uf_union(par, i, j);
                        }
                    }
        }
        }
        __syncthreads();
        // ---- flatten; number clusters by their lowest sphere (order apply_clustering leaves, refine.c:58-93) ------------------
        for (int i = tid; i < ns; i += nt) { int r = uf_find(par, i); X.tmp[i] = (r == i) ? 1 : 0; X.cellsph[i] = r; }
        __syncthreads();
        if (tid == 0) X.tmp[ns] = 0;
        __syncthreads();
        const int nc = block_exclusive_scan(X.tmp, ns + 1, s_sm);
        for (int i = tid; i < ns; i += nt) par[i] = X.tmp[X.cellsph[i]];     // cluster id of every sphere
        for (int i = tid; i <= nc; i += nt) cl_off[i] = 0;
        __syncthreads();
        for (int i = tid; i < ns; i += nt) atomicAdd(&cl_off[par[i]], 1);
        __syncthreads();
        block_exclusive_scan(cl_off, nc + 1, s_sm);
        for (int i = tid; i < nc; i += nt) X.tmp[i] = 0;                    // per-cluster fill cursor
        __syncthreads();
        if (wid == 0) {                                                     // ordered fill: members ascending
            for (int base = 0; base < ns; base += 32) {
                int i = base + lane;
                int c = i < ns ? par[i] : -1 - lane;
                unsigned m = __match_any_sync(0xffffffffu, c);
                int r = __popc(m & ((1u << lane) - 1u));
                if (i < ns) {
                    int cur = X.tmp[c];
                    cl_sph[cl_off[c] + cur + r] = i;
                }
                __syncwarp();
                if (i < ns && r == 0) X.tmp[c] += __popc(m);
                __syncwarp();
            }
        }
        if (tid == 0) counts[CNT_CLUSTERS] = nc;
    }
}

// K3g: the same clustering for ONE large structure by every block of a cooperative launch (BASELINE configs[3]: ~45,000 spheres, where a
// single block needs 45 ms). Phases are grid-strided with grid barriers between them; the union-find is the same lock-free one (its
// arbiter is the atomicMin on the root, so it never needed a block); member lists are filled through atomic cursors and then put in
// ascending order by a rank count inside each cluster (one warp per cluster), which is the order apply_clustering leaves (refine.c:58-93).
struct ClusterGridScratch {
    int *cellcnt;             // [CELLMAX + 2]
    int *cellsph;             // [scap + 1]
    int *tmp;                 // [scap + 2]
    float *bb;                // [6 * gridDim.x] per-block bounding boxes
    int *glob;                // [4]: number of clusters
    int scap;
};
__device__ __forceinline__ int uf_find_cg(int *par, int i)
{
    int p = __ldcg(&par[i]);
    while (p != i) {
        const int g = __ldcg(&par[p]);
        if (g != p) par[i] = g;
        i = p; p = g;
    }
    return i;
}
__device__ __forceinline__ void uf_union_cg(int *par, int a, int b)
{
    for (;;) {
        a = uf_find_cg(par, a); b = uf_find_cg(par, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        const int old = atomicMin(&par[a], b);
        if (old == a) return;
        a = old;
    }
}
__global__ void __launch_bounds__(256) k_cluster_grid(BatchDev B, ClusterGridScratch X, int k)
{
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ int s_sm[40];
    __shared__ float s_mn[8][3], s_mx[8][3];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int gtid = blockIdx.x * nt + tid, gnt = gridDim.x * nt, gw = gtid >> 5, ngw = gnt >> 5;
    const StructDev S = B.S[k];
    int *counts = B.counts + (size_t)k * CNT_STRIDE;
    if (counts[CNT_STATUS] != FPK_OK) { if (gtid == 0) counts[CNT_CLUSTERS] = 0; return; }
    const int ns = counts[CNT_SPHERES];
    const float4 *sx = B.s_xyzr + S.sph_off;
    int *par = B.s_cluster + S.sph_off;
    int *cl_off = B.cl_off + S.sph_off + k, *cl_sph = B.cl_sph + S.sph_off;
    if (!B.prm.skip_clustering && ns < 2) { if (gtid == 0) { counts[CNT_CLUSTERS] = 0; counts[CNT_STATUS] = FPK_ERR_NO_SPHERES; } return; }
    if (ns == 0) { if (gtid == 0) { counts[CNT_CLUSTERS] = 0; counts[CNT_STATUS] = FPK_ERR_NO_SPHERES; } return; }
    if (B.prm.skip_clustering) {
        for (int i = gtid; i < ns; i += gnt) { par[i] = 0; cl_sph[i] = i; }
        if (gtid == 0) { cl_off[0] = 0; cl_off[1] = ns; counts[CNT_CLUSTERS] = 1; }
        return;
    }
    // ---- bounding box of the centres: per block, then every block folds the per-block boxes -----------------------------------------
    float mn[3] = {3e38f, 3e38f, 3e38f}, mx[3] = {-3e38f, -3e38f, -3e38f};
    for (int i = gtid; i < ns; i += gnt) {
        const float4 v = sx[i];
        mn[0] = fminf(mn[0], v.x); mn[1] = fminf(mn[1], v.y); mn[2] = fminf(mn[2], v.z);
        mx[0] = fmaxf(mx[0], v.x); mx[1] = fmaxf(mx[1], v.y); mx[2] = fmaxf(mx[2], v.z);
        par[i] = i;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], d));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], d));
        }
    if (lane == 0) for (int c = 0; c < 3; c++) { s_mn[wid][c] = mn[c]; s_mx[wid][c] = mx[c]; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (nt >> 5); w++)
            for (int c = 0; c < 3; c++) { mn[c] = fminf(mn[c], s_mn[w][c]); mx[c] = fmaxf(mx[c], s_mx[w][c]); }
        for (int c = 0; c < 3; c++) { X.bb[6 * blockIdx.x + c] = mn[c]; X.bb[6 * blockIdx.x + 3 + c] = mx[c]; }
    }
    grid.sync();
    for (int c = 0; c < 3; c++) { mn[c] = 3e38f; mx[c] = -3e38f; }
    for (int b = 0; b < (int)gridDim.x; b++)
        for (int c = 0; c < 3; c++) { mn[c] = fminf(mn[c], __ldcg(&X.bb[6 * b + c])); mx[c] = fmaxf(mx[c], __ldcg(&X.bb[6 * b + 3 + c])); }
    float cell = (B.prm.distance_measure == 'b' ? 3.0f : 1.0f) * B.prm.clust_max_dist * 1.001f + 1e-4f;      // see k_cluster
    int d0, d1, d2;
    for (;;) {
        d0 = (int)((mx[0] - mn[0]) / cell) + 1; d1 = (int)((mx[1] - mn[1]) / cell) + 1; d2 = (int)((mx[2] - mn[2]) / cell) + 1;
        if ((long long)d0 * d1 * d2 <= CELLMAX) break;
        cell *= 1.26f;
    }
    const int ncell = d0 * d1 * d2;
    const float bx = mn[0], by = mn[1], bz = mn[2];
    for (int i = gtid; i <= ncell; i += gnt) X.cellcnt[i] = 0;
    grid.sync();
    for (int i = gtid; i < ns; i += gnt) {
        const float4 v = sx[i];
        const int cx = min(d0 - 1, max(0, (int)((v.x - bx) / cell))), cy = min(d1 - 1, max(0, (int)((v.y - by) / cell))),
                  cz = min(d2 - 1, max(0, (int)((v.z - bz) / cell)));
        const int cid = (cx * d1 + cy) * d2 + cz;
        X.tmp[i] = cid;
        atomicAdd(&X.cellcnt[cid], 1);
    }
    grid.sync();
    if (blockIdx.x == 0) block_exclusive_scan(X.cellcnt, ncell + 1, s_sm);
    grid.sync();
    for (int i = gtid; i < ns; i += gnt) {
        const int cid = X.tmp[i];
        const int slot = atomicAdd(&X.cellcnt[cid], 1);       // advances the start; afterwards cellcnt[c] = END of cell c
        X.cellsph[slot] = i;
    }
    grid.sync();
    // ---- union over the  d <= cut  graph (clusterlib.c:933 euclid in f64 on the f32 centres) ------------------------------------------
    const double cut = (double)B.prm.clust_max_dist;
    const bool cityblock = B.prm.distance_measure == 'b';
    for (int i = gtid; i < ns; i += gnt) {
        const float4 v = sx[i];
        const int cid = X.tmp[i];
        const int cz = cid % d2, cy = (cid / d2) % d1, cx = cid / (d2 * d1);
        const double xi = (double)v.x, yi = (double)v.y, zi = (double)v.z;
        for (int ax = max(cx - 1, 0); ax <= min(cx + 1, d0 - 1); ax++)
            for (int ay = max(cy - 1, 0); ay <= min(cy + 1, d1 - 1); ay++)
                for (int az = max(cz - 1, 0); az <= min(cz + 1, d2 - 1); az++) {
                    const int c2 = (ax * d1 + ay) * d2 + az;
                    const int lo = c2 ? X.cellcnt[c2 - 1] : 0, hi = X.cellcnt[c2];
                    for (int q = lo; q < hi; q++) {
                        const int j = X.cellsph[q];
                        if (j >= i) continue;
                        const float4 w = sx[j];
                        if (!(pair_distance(cityblock, xi, yi, zi, w) > cut)) uf_union_cg(par, i, j);
                    }
                }
    }
    grid.sync();
    // ---- flatten; number clusters by their lowest sphere -----------------------------------------------------------------------------------
    for (int i = gtid; i < ns; i += gnt) { const int r = uf_find_cg(par, i); X.tmp[i] = (r == i) ? 1 : 0; X.cellsph[i] = r; }
    if (gtid == 0) X.tmp[ns] = 0;
    grid.sync();
    if (blockIdx.x == 0) { const int nc0 = block_exclusive_scan(X.tmp, ns + 1, s_sm); if (tid == 0) X.glob[0] = nc0; }
    grid.sync();
    const int nc = __ldcg(&X.glob[0]);
    for (int i = gtid; i < ns; i += gnt) par[i] = X.tmp[X.cellsph[i]];     // cluster id of every sphere
    for (int i = gtid; i <= nc; i += gnt) cl_off[i] = 0;
    grid.sync();
    for (int i = gtid; i < ns; i += gnt) atomicAdd(&cl_off[par[i]], 1);
    for (int i = gtid; i < nc; i += gnt) X.tmp[i] = 0;                     // per-cluster fill cursor (tmp is free again)
    grid.sync();
    if (blockIdx.x == 0) block_exclusive_scan(cl_off, nc + 1, s_sm);
    grid.sync();
    for (int i = gtid; i < ns; i += gnt) {                                 // members in arrival order ...
        const int c = par[i];
        X.cellsph[cl_off[c] + atomicAdd(&X.tmp[c], 1)] = i;
    }
    grid.sync();
    for (int c = gw; c < nc; c += ngw) {                                   // ... then ascending: rank inside the cluster, one warp per cluster
        const int o = cl_off[c], n = cl_off[c + 1] - o;
        for (int a = lane; a < n; a += 32) {
            const int me = X.cellsph[o + a];
            int rank = 0;
            for (int b = 0; b < n; b++) rank += X.cellsph[o + b] < me;
            cl_sph[o + rank] = me;
        }
    }
    if (gtid == 0) counts[CNT_CLUSTERS] = nc;
}

// ---- uniform cell grid over the heavy atoms of pdb_w_lig, one per structure, built once per batch ---------------------------------------
// K4 needs, per pocket, the atoms within ray + 1.0 of one of its spheres (asa.c:84-112 get_surrounding_atoms_idx walks ALL atoms for
// every pocket: natoms x pockets distance loops). With the grid a pocket only meets the atoms of the cells its bounding box overlaps:
// what makes the 150,000-atom complex (3,540 clusters) affordable, and a few times fewer atoms per pocket at 2k-5k atoms.
struct AtomGrid {
    int *cellend;             // per structure at cbase[k]: END of every cell in cellatom (start = end of the cell before)
    int *cellatom;            // per structure at (natoms_w ? w_off : same_base + atom_off): heavy-atom indices grouped by cell (any order inside a cell)
    int same_base;            // = total pdb_w_lig atoms of the batch: structures without a separate pdb_w_lig list come after those with one
    int min_atoms;            // structures with at most this many atoms get no grid
    const int *cbase;         // [nstruct + 1]
    float4 *geom;             // [nstruct] (x0, y0, z0, cell edge)
    int4 *dims;               // [nstruct] (d0, d1, d2, 0)
    // every structure, whatever its size: bounding box of the heavy atoms of every run of 32 consecutive atoms of pdb_w_lig (two float4: min, max), at
    // chunk_base(k) + c. Atoms come in chain order, so such a run is a few residues, ~10 A across: k_descriptors_small scans only the runs whose box
    // meets the cluster's (in ascending order: the atom order asa.c:84-112 leaves is kept) instead of all atoms of the structure for every cluster.
    float4 *cbox;
};
__device__ __forceinline__ size_t chunk_base(const AtomGrid &AG, const StructDev &S, int k)
{
    return (size_t)((S.natoms_w ? S.w_off : AG.same_base + S.atom_off) >> 5) + (size_t)k;
}
__global__ void __launch_bounds__(256) k_atom_grid(BatchDev B, AtomGrid AG)
{
    __shared__ int s_sm[40], s_dims[4];
    __shared__ float s_mn[8][3], s_mx[8][3], s_box[4];
    const int k = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const StructDev S = B.S[k];
    const int W0 = S.natoms_w ? S.w_off : S.atom_off, WQ0 = S.natoms_w ? S.w_off : S.prop_off, nw = S.natoms_w ? S.natoms_w : S.natoms;
    const unsigned whmask = S.natoms_w ? FPK_ATOM_H : FPK_ATOM_H1;
    const float *wx = S.natoms_w ? B.wx : B.ax, *wy = S.natoms_w ? B.wy : B.ay, *wz = S.natoms_w ? B.wz : B.az;
    const uint8_t *wf = S.natoms_w ? B.wfl : B.afl;
    int *cellend = AG.cellend + AG.cbase[k];
    int *cellatom = AG.cellatom + (S.natoms_w ? S.w_off : AG.same_base + S.atom_off);
    const int ccap = AG.cbase[k + 1] - AG.cbase[k] - 2;
    {
        float4 *cbx = AG.cbox + 2 * chunk_base(AG, S, k);
        const int nch = (nw + 31) >> 5;
        for (int c = wid; c < nch; c += nt >> 5) {
            const int i = 32 * c + lane;
            float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
            if (i < nw && !(wf[WQ0 + i] & whmask)) { lo[0] = hi[0] = wx[W0 + i]; lo[1] = hi[1] = wy[W0 + i]; lo[2] = hi[2] = wz[W0 + i]; }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1)
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    lo[q] = fminf(lo[q], __shfl_xor_sync(0xffffffffu, lo[q], d));
                    hi[q] = fmaxf(hi[q], __shfl_xor_sync(0xffffffffu, hi[q], d));
                }
            if (lane == 0) { cbx[2 * c] = make_float4(lo[0], lo[1], lo[2], 0.0f); cbx[2 * c + 1] = make_float4(hi[0], hi[1], hi[2], 0.0f); }
        }
    }
    if (nw <= AG.min_atoms) return;          // nobody will ask: K4 scans the atoms of a small structure directly
    float mn[3] = {3e38f, 3e38f, 3e38f}, mx[3] = {-3e38f, -3e38f, -3e38f};
    for (int i = tid; i < nw; i += nt) {
        if (wf[WQ0 + i] & whmask) continue;
        const float x = wx[W0 + i], y = wy[W0 + i], z = wz[W0 + i];
        mn[0] = fminf(mn[0], x); mn[1] = fminf(mn[1], y); mn[2] = fminf(mn[2], z);
        mx[0] = fmaxf(mx[0], x); mx[1] = fmaxf(mx[1], y); mx[2] = fmaxf(mx[2], z);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], d));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], d));
        }
    if (lane == 0) for (int c = 0; c < 3; c++) { s_mn[wid][c] = mn[c]; s_mx[wid][c] = mx[c]; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (nt >> 5); w++)
            for (int c = 0; c < 3; c++) { mn[c] = fminf(mn[c], s_mn[w][c]); mx[c] = fmaxf(mx[c], s_mx[w][c]); }
        float cell = 6.0f;
        int d0 = 1, d1 = 1, d2 = 1;
        if (mn[0] <= mx[0])
            for (;;) {
                d0 = (int)((mx[0] - mn[0]) / cell) + 1; d1 = (int)((mx[1] - mn[1]) / cell) + 1; d2 = (int)((mx[2] - mn[2]) / cell) + 1;
                if ((long long)d0 * d1 * d2 <= ccap) break;
                cell *= 1.26f;
            }
        else { mn[0] = mn[1] = mn[2] = 0.0f; }
        s_dims[0] = d0; s_dims[1] = d1; s_dims[2] = d2;
        s_box[0] = mn[0]; s_box[1] = mn[1]; s_box[2] = mn[2]; s_box[3] = cell;
        AG.geom[k] = make_float4(mn[0], mn[1], mn[2], cell);
        AG.dims[k] = make_int4(d0, d1, d2, 0);
    }
    __syncthreads();
    const int d0 = s_dims[0], d1 = s_dims[1], d2 = s_dims[2], ncell = d0 * d1 * d2;
    const float bx = s_box[0], by = s_box[1], bz = s_box[2], cell = s_box[3];
    for (int i = tid; i <= ncell; i += nt) cellend[i] = 0;
    __syncthreads();
    for (int i = tid; i < nw; i += nt) {
        if (wf[WQ0 + i] & whmask) continue;
        const int cx = min(d0 - 1, max(0, (int)((wx[W0 + i] - bx) / cell))), cy = min(d1 - 1, max(0, (int)((wy[W0 + i] - by) / cell))),
                  cz = min(d2 - 1, max(0, (int)((wz[W0 + i] - bz) / cell)));
        atomicAdd(&cellend[(cx * d1 + cy) * d2 + cz], 1);
    }
    __syncthreads();
    block_exclusive_scan(cellend, ncell + 1, s_sm);
    for (int i = tid; i < nw; i += nt) {
        if (wf[WQ0 + i] & whmask) continue;
        const int cx = min(d0 - 1, max(0, (int)((wx[W0 + i] - bx) / cell))), cy = min(d1 - 1, max(0, (int)((wy[W0 + i] - by) / cell))),
                  cz = min(d2 - 1, max(0, (int)((wz[W0 + i] - bz) / cell)));
        cellatom[atomicAdd(&cellend[(cx * d1 + cy) * d2 + cz], 1)] = i;       // the start walks up to the end
    }
}

// ============================================= K4: descriptors ===============================================================
// aa.c:60-80 property table: volume score, hydrophobicity, charge, polarity
__constant__ float c_aa_vol[20] = {2, 7, 3, 3, 3, 4, 4, 1, 4, 5, 5, 6, 5, 6, 3, 2, 3, 8, 7, 4};
__constant__ float c_aa_hyd[20] = {41, -14, -28, -55, 49, -10, -31, 0, 8, 99, 97, -23, 74, 100, -46, -5, 13, 97, 63, 76};
__constant__ int c_aa_chg[20] = {0, 1, 0, -1, 0, 0, -1, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0};
__constant__ int c_aa_pol[20] = {0, 1, 1, 1, 0, 1, 1, 0, 1, 0, 0, 1, 0, 0, 0, 1, 1, 1, 1, 0};
__constant__ float c_spiral[300];     // asa.c:438-454 get_points_on_sphere(100), computed on the host with libm

// Philox4x32-10 (Salmon et al. SC'11): counter-based, so every sample of every pocket is independent of scheduling
__device__ __forceinline__ void philox4x32(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned out[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        unsigned h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        unsigned h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        unsigned n0 = h1 ^ c1 ^ k0, n1 = l1, n2 = h0 ^ c3 ^ k1, n3 = l0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// utils.c:113-128 rand_uniform on a 31-bit draw
__device__ __forceinline__ float unif_from_bits(unsigned r31, float mn, float mx)
{
    float rnd = ((float)(int)r31 / (float)2147483647) * (mx - mn);
    return mn + rnd;
}

#ifndef FPK_K4_STAGE_TOUCHING
#define FPK_K4_STAGE_TOUCHING 0
#endif
#ifndef FPK_K4_SEARCH_NEAR
#define FPK_K4_SEARCH_NEAR 0
#endif
#ifndef FPK_K4_MC_GRID_MIN
#define FPK_K4_MC_GRID_MIN 48      // pockets below it test every sample against every sphere: all lanes walk the same list, which beats the shorter but
                                  // divergent cell lists up to ~50 spheres (profiles/r02_k4_mc_grid_min.log: 12: 93.6 ms, 24: 90.4-91.1, 48: 88.1, 64: 88.2, 112: 89.7, 160: 91.2)
#endif
struct DescScratch {          // per resident block of K4
    int *sa;                  // [natoms_w max] surrounding atoms of the current pocket (asa.c:84-112), three ordered segments
    unsigned *bits;           // [natoms_w max / 32 + 1] large structures: one bit per atom of pdb_w_lig (marked by the pocket's spheres)
    // the pocket's spheres in pocket order, copied once per pocket of up to K4_STAGE_MAX spheres: warp 0's pair loops (n^2 / 2 reads in the
    // reference's order) then read consecutive 16-byte records from L1 instead of chasing  cl_sph -> s_xyzr  through L2 for every pair
    float4 *sph;              // [K4_STAGE_MAX] (x, y, z, ray)
    int *sphk;                // [K4_STAGE_MAX] bit 0: polar; bits 1..: the sphere's identity (zone batches list a sphere once per zone point)
    float4 *sphmc;            // [K4_STAGE_MAX] (x, y, z, (ray - 1.6)^2): the Monte-Carlo tile of a pocket too large for the shared one
    unsigned short *mcent;    // [K4_MC_GENT] Monte-Carlo cell entries of a pocket whose lists do not fit the shared array
};

enum { K4_THREADS = 128, K4_WORKERS = 96, MC_SMEM_SPH = 512, MC_CELLS = 512, MC_ENTRIES = 4096, ASA_CAND = 128,
       K4_GRID_MIN_ATOMS = 10000,        // from here on the surrounding atoms of a pocket come from the structure's cell grid (AtomGrid)
       K4_STAGE_MAX = 2048, K4_MC_GENT = 32768 };

// barrier among warps 1..3 only (warp 0 runs the reference-order sequential sums meanwhile)
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, 96;" ::: "memory"); }

// K4: one block per cluster. After the contacted-atom lists are built the block splits:
//   warp 0     - the reference's sequential float sums (pair loop of descriptors.c:186-236 etc.), in the reference's order;
//   warps 1..3 - surrounding atoms, the two SASA passes and the Monte-Carlo volume (integer counts: order-free).
#ifdef FPK_K4_PROFILE
// measurement build only (scripts/gpu_k4_split.sh): block cycles and counts of work items, [0,1] clusters below min_pock_nb_asph, [2,3] the others
__device__ unsigned long long g_k4prof[24];      // [12..21]: clusters of >= 256 spheres: count, spheres, then cycles by phase (see capi.cu)      // [4..8]: the first kind by phase (fetch + contacted atoms, CSR, sums / surroundings / SASA, per-atom epilogue, record)
#endif
__global__ void __launch_bounds__(K4_THREADS, 8) k_descriptors(BatchDev B, PocketDev Pk, const DescScratch *scratch, int *work_counter, AtomGrid AG,
                                                               const unsigned char *small_done)
{
    __shared__ int s_item, s_U, s_cnt[100], s_sm[40], s_segcnt[3], s_mcdim[4], s_mcmode, s_asanext, s_mcnext, s_nsa;
    __shared__ float s_mcbox[6], s_mch[3], s_bb[3][6];
    __shared__ float4 s_sph[MC_SMEM_SPH];                  // (x, y, z, (ray-1.6)^2) of the pocket's spheres
    __shared__ int s_cell[MC_CELLS + 1];
    __shared__ __align__(16) unsigned short s_ent[MC_ENTRIES];   // MC cell entries; in clusters without MC samples: warp 0's SASA candidates
    __shared__ float4 s_cand[3][ASA_CAND];                 // per worker warp: burial candidates of the current atom (x, y, z, radius)
    __shared__ unsigned char s_candapol[3][ASA_CAND];
    __shared__ fpk_desc s_desc;                            // warp 0's part of the record (everything but ASA / volume)
    __shared__ float s_spiral[300];                        // spiral points: the lanes read 32 different ones at once (constant memory would serialise that)
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < 300; i += nt) s_spiral[i] = c_spiral[i];       // visible after the first barrier of the work loop
    const DescScratch &X = scratch[blockIdx.x];
    const fpk_params &prm = B.prm;
#if FPK_K4_SEARCH_NEAR
    __shared__ int s_klast;           // in shared memory: a register that lives across the whole work item costs the kernel 300 bytes of spills
    if (tid == 0) s_klast = -1;
#endif
#ifdef FPK_K4_PROFILE
    long long prof_t = 0, prof_p[4] = {0, 0, 0, 0}; int prof_class = -1, prof_n = 0;
    __shared__ unsigned long long s_prof[4];      // end of warp 0's sums, of the workers' surroundings, of the last SASA atom
#endif
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(work_counter, 1);
#ifdef FPK_K4_PROFILE
        if (tid == 0) {
            const long long now = clock64();
            if (prof_class >= 0) { atomicAdd(&g_k4prof[2 * prof_class], (unsigned long long)(now - prof_t)); atomicAdd(&g_k4prof[2 * prof_class + 1], 1ull); }
            if (prof_class == 0) {
                atomicAdd(&g_k4prof[4], (unsigned long long)(prof_p[0] - prof_t)); atomicAdd(&g_k4prof[5], (unsigned long long)(prof_p[1] - prof_p[0]));
                atomicAdd(&g_k4prof[6], (unsigned long long)(prof_p[2] - prof_p[1])); atomicAdd(&g_k4prof[7], (unsigned long long)(prof_p[3] - prof_p[2]));
                atomicAdd(&g_k4prof[8], (unsigned long long)(now - prof_p[3]));
            }
            if (prof_class >= 0 && prof_n >= 256) {
                atomicAdd(&g_k4prof[12], 1ull); atomicAdd(&g_k4prof[13], (unsigned long long)prof_n);
                atomicAdd(&g_k4prof[14], (unsigned long long)(prof_p[0] - prof_t)); atomicAdd(&g_k4prof[15], (unsigned long long)(prof_p[1] - prof_p[0]));
                atomicAdd(&g_k4prof[16], s_prof[0] - (unsigned long long)prof_p[1]); atomicAdd(&g_k4prof[17], s_prof[1] - (unsigned long long)prof_p[1]);
                atomicAdd(&g_k4prof[18], s_prof[2] - (unsigned long long)prof_p[1]); atomicAdd(&g_k4prof[19], (unsigned long long)(prof_p[2] - prof_p[1]));
                atomicAdd(&g_k4prof[20], (unsigned long long)(prof_p[3] - prof_p[2])); atomicAdd(&g_k4prof[21], (unsigned long long)(now - prof_p[3]));
            }
            prof_t = now;
        }
#endif
        __syncthreads();
        const int item = s_item;
        if (item >= Pk.total_clusters) return;
        if (small_done && small_done[item]) continue;          // a small cluster that k_descriptors_small (one warp) has already described
        // (structure, cluster) of this work item
        int lo = 0, hi = B.nstruct;
#if FPK_K4_SEARCH_NEAR
        // EXPERIMENT, off by default, not yet measured on a GPU (DESIGN.md par. 9): items are handed out in order, so the structure of a block's next
        // item is almost always the one of its last item or the next one: two loads instead of a 12-step chain of dependent ones
        const int k_last = s_klast;       // written after the first barrier of the previous item, read after the two barriers of this one
        if (k_last >= 0) {
            if (Pk.clbase[k_last] <= item && item < Pk.clbase[k_last + 1]) { lo = k_last; hi = k_last + 1; }
            else if (k_last + 2 <= B.nstruct && Pk.clbase[k_last + 1] <= item && item < Pk.clbase[k_last + 2]) { lo = k_last + 1; hi = k_last + 2; }
        }
#endif
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (Pk.clbase[mid] <= item) lo = mid; else hi = mid; }
        const int k = lo, c = item - Pk.clbase[k];
        const StructDev S = B.S[k];
        const int *cl_off = B.cl_off + S.sph_off + k;
        const int off = cl_off[c], n = cl_off[c + 1] - off;
        const int R = Pk.sbase[k] + off;                   // pocket-level arrays are addressed through the compact sphere base
        const int *L = B.cl_sph + S.sph_off + off;         // sphere indices (inside the structure) of this pocket, ascending
        const float4 *sx = B.s_xyzr + S.sph_off;
        const int4 *sn = B.s_neigh + S.sph_off;
        const uint8_t *sty = B.s_type + S.sph_off;
        const float *sb = B.s_bary + 3 * (size_t)S.sph_off;
        int *uat = Pk.cl_atoms + 4 * (size_t)R, *entu = Pk.ent_u + 4 * (size_t)R;
        int *toff = Pk.touch_off + 5 * (size_t)R, *tlist = Pk.touch_list + 4 * (size_t)R;
        int *acc = Pk.acc + 16 * (size_t)R;
        float *fxv = Pk.fx_val + 16 * (size_t)R;
        fpk_desc *D = Pk.desc + item;
        const int A0 = S.atom_off, Q0 = S.prop_off;      // coordinates / properties of atom a: [A0 + a] / [Q0 + a]
        const bool same = S.same_pdb != 0;
        const int W0 = S.natoms_w ? S.w_off : S.atom_off, WQ0 = S.natoms_w ? S.w_off : S.prop_off, nw = S.natoms_w ? S.natoms_w : S.natoms;
        const unsigned whmask = S.natoms_w ? FPK_ATOM_H : FPK_ATOM_H1;
        const float *wx = S.natoms_w ? B.wx : B.ax, *wy = S.natoms_w ? B.wy : B.ay, *wz = S.natoms_w ? B.wz : B.az;
        const float *wr = S.natoms_w ? B.wrad : B.arad, *we = S.natoms_w ? B.wen : B.aen;
        const uint8_t *wf = S.natoms_w ? B.wfl : B.afl;
        const bool full = prm.do_asa_and_volume != 0;
        const int *kcnt = B.counts + (size_t)k * CNT_STRIDE;
        const bool explicit_pocket = kcnt[CNT_XP] == 1 && c == kcnt[CNT_CLUSTERS] - 1;     // dpocket.c:350: described from the interface atoms
        const bool observable = explicit_pocket || n >= prm.min_pock_nb_asph;      // below it the cluster is always dropped (refine.c:124): no MC samples
        const bool insm = n <= MC_SMEM_SPH;
        const bool stg = n <= K4_STAGE_MAX;           // the pocket's spheres are copied to X.sph / X.sphk (and X.sphmc beyond the shared tile)
        const float4 *Xs = X.sph;
        const int *Xk = X.sphk;
        const float correct = -1.6f;
#ifdef FPK_K4_PROFILE
        prof_class = observable ? 1 : 0; prof_n = n;
        if (tid == 0) { s_prof[0] = s_prof[1] = s_prof[2] = 0ull; }
#endif

        // ---- 1. warp 0: unique contacted atoms by atom id, first-seen order (pocket.c:1280-1320); warps 1..3: MC sphere tile ----------
        if (wid == 0) {
            int U = 0;
            int *uids = tlist;
            for (int base = 0; base < 4 * n; base += 32) {
                int e = base + lane;
                bool valid = e < 4 * n;
                int a = -1, id = -2 - lane;
                if (valid) {
                    int4 q = sn[L[e >> 2]];
                    int j = e & 3;
                    a = j == 0 ? q.x : (j == 1 ? q.y : (j == 2 ? q.z : q.w));
                    id = B.aid[Q0 + a];
                }
                int dupidx = -1;
#pragma unroll 4
                for (int u = 0; u < U; u++)
                    if (uids[u] == id) dupidx = u;          // the ids of the atoms found so far, kept beside the list (touch_list is idle until step 2)
                unsigned m = __match_any_sync(0xffffffffu, id);
                int leader = __ffs(m) - 1;
                bool isnew = valid && dupidx < 0 && lane == leader;
                unsigned bal = __ballot_sync(0xffffffffu, isnew);
                int pos = U + __popc(bal & ((1u << lane) - 1u));
                if (isnew) { uat[pos] = a; uids[pos] = id; }
                int lp = __shfl_sync(0xffffffffu, pos, leader);
                if (valid) entu[e] = dupidx >= 0 ? dupidx : lp;
                U += __popc(bal);
                __syncwarp();
            }
            if (lane == 0) s_U = U;
        } else if (stg) {
            for (int i = tid - 32; i < n; i += K4_WORKERS) {
                const int si = L[i];
                float4 v = sx[si];
                X.sph[i] = v;
                X.sphk[i] = ((Pk.sph_uid ? Pk.sph_uid[S.sph_off + si] : si) << 1) | (sty[si] != 0 ? 1 : 0);
                if (full) { v.w = (correct + v.w) * (v.w + correct); if (insm) s_sph[i] = v; else X.sphmc[i] = v; }
            }
        }
        if (wid != 0 && full) {
            if (tid == 32) {        // MC box (voronoi.c:1172-1195: +correct on both sides, else-if updates), one thread, reference order
                float xmin = 0, xmax = 0, ymin = 0, ymax = 0, zmin = 0, zmax = 0, xt, yt, zt;
                for (int i = 0; i < n; i++) {
                    float4 v = sx[L[i]];
                    if (i == 0) {
                        xmin = v.x - v.w + correct; xmax = v.x + v.w + correct;
                        ymin = v.y - v.w + correct; ymax = v.y + v.w + correct;
                        zmin = v.z - v.w + correct; zmax = v.z + v.w + correct;
                    } else {
                        if (xmin > (xt = v.x - v.w + correct)) xmin = xt; else if (xmax < (xt = v.x + v.w + correct)) xmax = xt;
                        if (ymin > (yt = v.y - v.w + correct)) ymin = yt; else if (ymax < (yt = v.y + v.w + correct)) ymax = yt;
                        if (zmin > (zt = v.z - v.w + correct)) zmin = zt; else if (zmax < (zt = v.z + v.w + correct)) zmax = zt;
                    }
                }
                s_mcbox[0] = xmin; s_mcbox[1] = xmax; s_mcbox[2] = ymin; s_mcbox[3] = ymax; s_mcbox[4] = zmin; s_mcbox[5] = zmax;
            }
            for (int i = tid - 32; i < 100; i += K4_WORKERS) s_cnt[i] = 0;
            if (tid == 33) { s_asanext = 0; s_mcnext = 0; s_nsa = 0; }
        }
        __syncthreads();
#ifdef FPK_K4_PROFILE
        if (tid == 0) { prof_p[0] = clock64(); }
#endif
#if FPK_K4_SEARCH_NEAR
        if (tid == 0) s_klast = k;
#endif
        const int U = s_U;
        // ---- 2. spheres touching each unique atom (CSR) ----------------------------------------------------------------------------
        for (int u = tid; u <= U; u += nt) toff[u] = 0;
        for (int u = tid; u < 4 * U; u += nt) acc[u] = 0;
        __syncthreads();
        for (int e = tid; e < 4 * n; e += nt) atomicAdd(&toff[entu[e]], 1);
        __syncthreads();
        block_exclusive_scan(toff, U + 1, s_sm);
        for (int e = tid; e < 4 * n; e += nt) {
            int u = entu[e];
            int slot = atomicAdd(&acc[u], 1);               // acc[0..U) doubles as the fill cursor here
            tlist[toff[u] + slot] = e >> 2;
        }
        __syncthreads();
        for (int u = tid; u < 4 * U; u += nt) acc[u] = 0;
        __syncthreads();
#ifdef FPK_K4_PROFILE
        if (tid == 0) { prof_p[1] = clock64(); }
#endif

        // ---- work shared by all four warps once warp 0 is through with its sequential sums (3a): contacted atoms of the SASA pass and
        // chunks of Monte-Carlo samples are handed out through two counters in shared memory; both only add integers, so who does what is free
        const int seg = (nw + 2) / 3;                       // the three ordered segments of X.sa (one per worker warp)
        auto asa_atom = [&](const int u, float4 *cand, unsigned char *candapol) {
            const int cura = uat[u];
            const float cx = B.ax[A0 + cura], cy = B.ay[A0 + cura], cz = B.az[A0 + cura], crad = B.arad[Q0 + cura];
            // burial candidates: surrounding atoms close enough to bury a point of this atom at either probe, in sa order
            const float reach = crad + 2.2f + 2.2f + 0.01f;
            int ncand = 0;
            bool overflow = false;
            for (int sg = 0; sg < 3 && !overflow; sg++) {
                const int s0 = min(nw, sg * seg), cnt = s_segcnt[sg];
                for (int base = 0; base < cnt && !overflow; base += 32) {
                    const int j = base + lane;
                    bool take = false;
                    int a = -1;
                    float ax_ = 0, ay_ = 0, az_ = 0, ar_ = 0;
                    if (j < cnt) {
                        a = X.sa[s0 + j];
                        if (!(same && a == cura)) {                 // asa.c:315 pointer inequality
                            ax_ = wx[W0 + a]; ay_ = wy[W0 + a]; az_ = wz[W0 + a]; ar_ = wr[WQ0 + a];
                            const float dx = ax_ - cx, dy = ay_ - cy, dz = az_ - cz, rr = reach + ar_;
                            take = dx * dx + dy * dy + dz * dz <= rr * rr;
                        }
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, take);
                    if (ncand + __popc(bal) > ASA_CAND) { overflow = true; break; }
                    if (take) {
                        const int pos = ncand + __popc(bal & ((1u << lane) - 1u));
                        cand[pos] = make_float4(ax_, ay_, az_, ar_);
                        candapol[pos] = (double)we[WQ0 + a] < 2.8 ? 1 : 0;
                    }
                    ncand += __popc(bal);
                }
            }
#if FPK_K4_STAGE_TOUCHING
            // EXPERIMENT, off by default, not yet measured or parity-checked on a GPU (DESIGN.md par. 9): the spheres touching this atom are read by every
            // one of the 200 tests through three dependent global loads (tlist -> L -> sx); copy them once behind the burial candidates instead
            const int t0 = toff[u], ntouch = toff[u + 1] - t0;
            const bool staged = !overflow && ncand + ntouch <= ASA_CAND;
            if (staged) for (int q = lane; q < ntouch; q += 32) cand[ASA_CAND - ntouch + q] = sx[L[tlist[t0 + q]]];
#endif
            __syncwarp();
            for (int it = lane; it < 200; it += 32) {
                const int pass = it >= 100 ? 1 : 0, kp = it - 100 * pass;
                const double probe = pass == 0 ? 1.4 : 2.2;
                const float tx = (float)((double)cx + (double)s_spiral[3 * kp] * ((double)crad + probe));
                const float ty = (float)((double)cy + (double)s_spiral[3 * kp + 1] * ((double)crad + probe));
                const float tz = (float)((double)cz + (double)s_spiral[3 * kp + 2] * ((double)crad + probe));
                bool vref = true;                                   // "vrefburried": not inside any touching alpha sphere
#if FPK_K4_STAGE_TOUCHING
                if (staged) {
                    for (int q = 0; q < ntouch && vref; q++) {
                        const float4 v = cand[ASA_CAND - ntouch + q];
                        if (f_ddist(tx, ty, tz, v.x, v.y, v.z) <= v.w * v.w) vref = false;
                    }
                } else
#endif
                for (int q = toff[u]; q < toff[u + 1] && vref; q++) {
                    float4 v = sx[L[tlist[q]]];
                    if (f_ddist(tx, ty, tz, v.x, v.y, v.z) <= v.w * v.w) vref = false;
                }
                if (vref) continue;
                int hit = -1;       // -1 none, 0 polar burying atom, 1 apolar burying atom
                if (!overflow) {
                    for (int j = 0; j < ncand; j++) {
                        const float4 q = cand[j];
                        const float dsq = f_ddist(tx, ty, tz, q.x, q.y, q.z);
                        const double rr = (double)q.w + probe;
                        if ((double)dsq < rr * rr) { hit = candapol[j]; break; }
                    }
                } else {
                    for (int sg = 0; sg < 3 && hit < 0; sg++) {
                        const int s0 = min(nw, sg * seg), cnt = s_segcnt[sg];
                        for (int j = 0; j < cnt; j++) {
                            const int a = X.sa[s0 + j];
                            if (same && a == cura) continue;
                            const float dsq = f_ddist(tx, ty, tz, wx[W0 + a], wy[W0 + a], wz[W0 + a]);
                            const double rr = (double)wr[WQ0 + a] + probe;
                            if ((double)dsq < rr * rr) { hit = (double)we[WQ0 + a] < 2.8 ? 1 : 0; break; }
                        }
                    }
                }
                if (hit < 0) atomicAdd(&acc[4 * u + pass], 1);
                else if (pass == 0) atomicAdd(&acc[4 * u + (hit ? 2 : 3)], 1);
            }
            __syncwarp();
        };
        auto mc_samples = [&]() {
            const float xmin = s_mcbox[0], xmax = s_mcbox[1], ymin = s_mcbox[2], ymax = s_mcbox[3], zmin = s_mcbox[4], zmax = s_mcbox[5];
            const int mode = s_mcmode;
            const int dx = s_mcdim[0], dy = s_mcdim[1], dz = s_mcdim[2];
            const float hx = s_mch[0], hy = s_mch[1], hz = s_mch[2];
            const int niter = prm.nb_mcv_iter;
            const unsigned k0 = (unsigned)prm.mc_seed, k1 = (unsigned)(prm.mc_seed >> 32);
            const unsigned mcid = Pk.mc_id ? (unsigned)Pk.mc_id[k] : (unsigned)c;
            const int total = 100 * niter;
            for (;;) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_mcnext, 256);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= total) break;
                const int stop = min(total, base + 256);
                for (int s = base + lane; s < stop; s += 32) {
                    const int li = s / niter, i = s - li * niter;
                    unsigned o[4];
                    philox4x32((unsigned)i, (unsigned)li, mcid, 0u, k0, k1, o);
                    const float xr = unif_from_bits(o[0] >> 1, xmin, xmax), yr = unif_from_bits(o[1] >> 1, ymin, ymax),
                                zr = unif_from_bits(o[2] >> 1, zmin, zmax);
                    bool in = false;
                    if (mode == 1) {
                        const int a = max(0, min(dx - 1, (int)floorf((xr - xmin) * hx))), b = max(0, min(dy - 1, (int)floorf((yr - ymin) * hy))),
                                  d = max(0, min(dz - 1, (int)floorf((zr - zmin) * hz)));
                        const int cid = (a * dy + b) * dz + d;
                        const int e0 = cid ? s_cell[cid - 1] : 0, e1 = s_cell[cid];
                        for (int j = e0; j < e1 && !in; j++) {
                            const float4 v = s_sph[s_ent[j]];
                            const float xt = v.x - xr, yt = v.y - yr, zt = v.z - zr;
                            in = v.w > (xt * xt + yt * yt + zt * zt);
                        }
                    } else if (mode == 2) {           // a large pocket: the same cell lists, entries (and beyond 512 spheres the tile) in the block's global scratch
                        const int a = max(0, min(dx - 1, (int)floorf((xr - xmin) * hx))), b = max(0, min(dy - 1, (int)floorf((yr - ymin) * hy))),
                                  d = max(0, min(dz - 1, (int)floorf((zr - zmin) * hz)));
                        const int cid = (a * dy + b) * dz + d;
                        const int e0 = cid ? s_cell[cid - 1] : 0, e1 = s_cell[cid];
                        const unsigned short *entg = X.mcent;
                        if (insm) {
                            for (int j = e0; j < e1 && !in; j++) {
                                const float4 v = s_sph[entg[j]];
                                const float xt = v.x - xr, yt = v.y - yr, zt = v.z - zr;
                                in = v.w > (xt * xt + yt * yt + zt * zt);
                            }
                        } else {
                            const float4 *tileg = X.sphmc;
                            for (int j = e0; j < e1 && !in; j++) {
                                const float4 v = tileg[entg[j]];
                                const float xt = v.x - xr, yt = v.y - yr, zt = v.z - zr;
                                in = v.w > (xt * xt + yt * yt + zt * zt);
                            }
                        }
                    } else if (insm) {
                        for (int j = 0; j < n && !in; j++) {
                            const float4 v = s_sph[j];
                            const float xt = v.x - xr, yt = v.y - yr, zt = v.z - zr;
                            in = v.w > (xt * xt + yt * yt + zt * zt);
                        }
                    } else {
                        for (int j = 0; j < n && !in; j++) {
                            const float4 v = sx[L[j]];
                            const float xt = v.x - xr, yt = v.y - yr, zt = v.z - zr;
                            in = ((correct + v.w) * (v.w + correct)) > (xt * xt + yt * yt + zt * zt);
                        }
                    }
                    if (in) atomicAdd(&s_cnt[li], 1);
                }
            }
        };
        if (wid == 0) {
            // ---- 3a. warp 0: the sequential float sums, in the reference's order ---------------------------------------------------
            // sphere pair loop (descriptors.c:186-236): as_density accumulates dist(i,j), i<j, row-major, in ONE float
            float as_density = 0.0f, as_max_dst = -1.0f, lane_max = -1.0f;
            int mlhd_cnt = 0, nAlphaApol = 0;
            for (int i = 0; i < n; i++) {
                const float4 vi = stg ? Xs[i] : sx[L[i]];
                const int key_i = stg ? Xk[i] : 0;
                const bool apol_i = stg ? !(key_i & 1) : sty[L[i]] == 0;
                int napol = 0;
                if (apol_i && stg) {
                    // two spheres are "the same" iff they carry the same identity (vc != vcur, descriptors.c:199): i itself, or in a zone batch another copy of it
#pragma unroll 4
                    for (int j = lane; j < n; j += 32) {          // integer count: order-free, so the loads of several rounds may be in flight together
                        const int key_j = Xk[j];
                        const float4 vj = Xs[j];
                        if (!((key_j & 1) || key_j == key_i) && (double)(f_dist(vi.x, vi.y, vi.z, vj.x, vj.y, vj.z) - (vj.w + vi.w)) <= 0.) napol++;
                    }
                } else if (apol_i) {
                    const int uid_i = Pk.sph_uid ? Pk.sph_uid[S.sph_off + L[i]] : 0;
                    for (int j = lane; j < n; j += 32) {
                        if (j == i || sty[L[j]] != 0) continue;
                        if (Pk.sph_uid && Pk.sph_uid[S.sph_off + L[j]] == uid_i) continue;          // another copy of the same sphere (vc != vcur)
                        float4 vj = sx[L[j]];
                        if ((double)(f_dist(vi.x, vi.y, vi.z, vj.x, vj.y, vj.z) - (vj.w + vi.w)) <= 0.) napol++;
                    }
                }
                if (apol_i) {
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) napol += __shfl_xor_sync(0xffffffffu, napol, d);
                    mlhd_cnt += napol;
                    nAlphaApol++;
                }
                float4 vnext = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i + 1 + lane < n) vnext = stg ? Xs[i + 1 + lane] : sx[L[i + 1 + lane]];
                for (int base = i + 1; base < n; base += 32) {
                    int j = base + lane;
                    float dj = 0.0f;
                    const float4 vj = vnext;
                    if (j + 32 < n) vnext = stg ? Xs[j + 32] : sx[L[j + 32]];      // requested before this chunk's 32 dependent additions
                    if (j < n) {
                        dj = f_dist(vi.x, vi.y, vi.z, vj.x, vj.y, vj.z);
                        lane_max = fmaxf(lane_max, dj);       // the maximum does not depend on the order: kept per lane, reduced once
                    }
                    // only the SUM is order-bound: one float, pairs in row-major order (descriptors.c:214)
                    if (n - base >= 32) {
#pragma unroll
                        for (int q = 0; q < 32; q++) as_density += __shfl_sync(0xffffffffu, dj, q);
                    } else {
                        const int cntj = n - base;
                        for (int q = 0; q < cntj; q++) as_density += __shfl_sync(0xffffffffu, dj, q);
                    }
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) lane_max = fmaxf(lane_max, __shfl_xor_sync(0xffffffffu, lane_max, d));
            if (lane_max > as_max_dst) as_max_dst = lane_max;
            // atoms of the atom-based descriptors: the contacted atoms; for dpocket's explicit pocket the interface atoms (dpocket.c:350)
            const int *alist = explicit_pocket ? Pk.xp_ilist + S.atom_off + k : uat;
            const int UA = explicit_pocket ? Pk.xp_ni[k] : U;
            // first occurrence of every residue id among them (descriptors.c:408 in_tab), one lane per atom
            if (UA <= 16 * n) {
                // the residue ids are fetched once (fx_val is idle until 6a) and the search runs over that list instead of chasing atom -> residue per step
                int *rids = reinterpret_cast<int *>(fxv);
                for (int u = lane; u < UA; u += 32) rids[u] = B.ares[Q0 + alist[u]];
                __syncwarp();
                for (int u = lane; u < UA; u += 32) {
                    const int rid = rids[u];
                    bool seen = false;
                    int q = 0;
                    for (; q + 4 <= u && !seen; q += 4) seen = rids[q] == rid || rids[q + 1] == rid || rids[q + 2] == rid || rids[q + 3] == rid;
                    for (; q < u && !seen; q++) seen = rids[q] == rid;
                    entu[u] = seen ? 0 : 1;                 // ent_u is free once the CSR is built
                }
            } else
                for (int u = lane; u < UA; u += 32) {
                    const int rid = B.ares[Q0 + alist[u]];
                    bool seen = false;
                    for (int q = 0; q < u && !seen; q++) seen = B.ares[Q0 + alist[q]] == rid;
                    entu[u] = seen ? 0 : 1;
                }
            __syncwarp();
            {
                // per-sphere and per-atom sums: every lane fetches one element of a chunk of 32 (the loads of a chunk are in flight together), then all
                // lanes add the 32 values in list order - the same additions in the same order as the reference's loops (descriptors.c:238-246, :355-436)
                float mean_r = 0.0f, masph = 0.0f, as_max_r = -1.0f, ps0 = 0.0f, ps1 = 0.0f, ps2 = 0.0f;
                int napolc = 0;
                for (int base = 0; base < n; base += 32) {
                    const int i = base + lane;
                    float4 v = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
                    float q = 0.0f;
                    bool ap = false;
                    if (i < n) {
                        const int si = L[i];
                        v = stg ? Xs[i] : sx[si];
                        const float dd = f_dist(v.x, v.y, v.z, sb[3 * si], sb[3 * si + 1], sb[3 * si + 2]);
                        q = dd / v.w;
                        ap = sty[si] == 0;
                    }
                    napolc += __popc(__ballot_sync(0xffffffffu, ap));
                    const int cnt = min(32, n - base);
                    for (int t = 0; t < cnt; t++) {
                        const float w = __shfl_sync(0xffffffffu, v.w, t);
                        if (w > as_max_r) as_max_r = w;
                        mean_r += w;
                        masph += __shfl_sync(0xffffffffu, q, t);
                        ps0 += __shfl_sync(0xffffffffu, v.x, t); ps1 += __shfl_sync(0xffffffffu, v.y, t); ps2 += __shfl_sync(0xffffffffu, v.z, t);
                    }
                }
                for (int w = lane; w < (int)(sizeof(fpk_desc) / sizeof(int)); w += 32) reinterpret_cast<int *>(&s_desc)[w] = 0;
                __syncwarp();
                int nb_res = 0, nb_polar = 0, pol_sc = 0, chg_sc = 0;
                float flex = 0.0f, hyd = 0.0f, vol = 0.0f;
                // atom based (descriptors.c:355-436, :453-513)
                if (UA > 1) {
                    for (int base = 0; base < UA; base += 32) {
                        const int u = base + lane;
                        int aa = -1;
                        bool fl = false, pl = false;
                        float bf = 0.0f;
                        if (u < UA) {
                            const int a = Q0 + alist[u];
                            fl = entu[u] != 0;
                            if (fl) aa = B.aaa[a];
                            bf = B.abf[a];
                            pl = (double)B.aen[a] > 2.7;
                        }
                        nb_res += __popc(__ballot_sync(0xffffffffu, fl));
                        nb_polar += __popc(__ballot_sync(0xffffffffu, pl));
                        const int cnt = min(32, UA - base);
                        for (int t = 0; t < cnt; t++) {
                            const int aat = __shfl_sync(0xffffffffu, aa, t);        // < 0: not the first atom of its residue, or no standard residue
                            if (aat >= 0) {
                                if (lane == 0) s_desc.aa_compo[aat]++;
                                hyd += c_aa_hyd[aat]; pol_sc += c_aa_pol[aat]; vol += c_aa_vol[aat]; chg_sc += c_aa_chg[aat];
                            }
                            flex += __shfl_sync(0xffffffffu, bf, t);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    fpk_desc &d = s_desc;
                    d.first_sphere = L[0];
                    d.nAlphaApol = napolc; d.nAlphaPol = n - napolc;
                    d.bary[0] = ps0 / (float)n; d.bary[1] = ps1 / (float)n; d.bary[2] = ps2 / (float)n;      // refine.c:192-240
                    d.n_atoms = U;
                    if (UA > 1) {
                        d.polarity_score = pol_sc; d.charge_score = chg_sc;
                        d.hydrophobicity_score = hyd / (float)nb_res;
                        d.volume_score = vol / (float)nb_res;
                        d.flex = flex / (float)UA;
                        d.prop_polar_atm = (float)((double)(((float)nb_polar) / ((float)UA)) * 100.0);
                    }
                    d.mean_loc_hyd_dens = nAlphaApol > 0 ? (float)mlhd_cnt / (float)nAlphaApol : 0.0f;
                    d.as_max_dst = as_max_dst;
                    d.apolar_asphere_prop = (float)nAlphaApol / (float)n;
                    d.masph_sacc = masph / (float)n;
                    d.mean_asph_ray = mean_r / (float)n;
                    d.nb_asph = n;
                    d.as_density = (float)((double)as_density / ((n * n - n) * 0.5));
                    d.as_max_r = as_max_r;
                }
            }
#ifdef FPK_K4_PROFILE
            if (lane == 0) s_prof[0] = (unsigned long long)clock64();
#endif
            if (full) {
                // the workers have ARRIVED at barrier 2 once the surrounding atoms (X.sa) and the Monte-Carlo cell grid are in place; warp 0 waits for
                // that only now, then takes samples (observable clusters) or contacted atoms (the rest: s_ent is idle there and holds its candidates)
                asm volatile("bar.sync 2, 128;" ::: "memory");
            }
        } else if (full) {
            const int wt = tid - 32, ww = wid - 1;
            // ---- 5a. Monte-Carlo cell grid first (voronoi.c:1150-1233 samples come later): warp 0 may start sampling while the workers are in the SASA pass ----
            if (observable) {
                const float xmin = s_mcbox[0], xmax = s_mcbox[1], ymin = s_mcbox[2], ymax = s_mcbox[3], zmin = s_mcbox[4], zmax = s_mcbox[5];
                // cell grid over the box: a sample only meets the spheres registered in its cell (the test is an OR: order-free)
                if (wt == 0) {
                    int mode = 0;
                    if ((insm || stg) && n >= FPK_K4_MC_GRID_MIN) {
                        const float ex = xmax - xmin, ey = ymax - ymin, ez = zmax - zmin;
                        const int dx = max(1, min(8, (int)(ex / 4.5f))), dy = max(1, min(8, (int)(ey / 4.5f))), dz = max(1, min(8, (int)(ez / 4.5f)));
                        s_mcdim[0] = dx; s_mcdim[1] = dy; s_mcdim[2] = dz;
                        s_mch[0] = (float)dx / ex; s_mch[1] = (float)dy / ey; s_mch[2] = (float)dz / ez;
                        mode = (ex > 0 && ey > 0 && ez > 0) ? 1 : 0;
                    }
                    s_mcmode = mode;
                }
                workers_sync();
                if (s_mcmode) {
                    const int dx = s_mcdim[0], dy = s_mcdim[1], dz = s_mcdim[2], ncell = dx * dy * dz;
                    const float hx = s_mch[0], hy = s_mch[1], hz = s_mch[2];
                    for (int i = wt; i <= ncell; i += K4_WORKERS) s_cell[i] = 0;
                    workers_sync();
                    for (int pass = 0; pass < 2; pass++) {
                        const bool gent = s_mcmode == 2;          // decided after the counting pass
                        for (int i = wt; i < n; i += K4_WORKERS) {
                            const float4 v = insm ? s_sph[i] : X.sphmc[i];
                            const float r = sqrtf(fmaxf(v.w, 0.0f)) + 2e-3f;
                            const int x0 = max(0, min(dx - 1, (int)floorf((v.x - r - xmin) * hx))), x1 = max(0, min(dx - 1, (int)floorf((v.x + r - xmin) * hx)));
                            const int y0 = max(0, min(dy - 1, (int)floorf((v.y - r - ymin) * hy))), y1 = max(0, min(dy - 1, (int)floorf((v.y + r - ymin) * hy)));
                            const int z0 = max(0, min(dz - 1, (int)floorf((v.z - r - zmin) * hz))), z1 = max(0, min(dz - 1, (int)floorf((v.z + r - zmin) * hz)));
                            for (int a = x0; a <= x1; a++) for (int b = y0; b <= y1; b++) for (int d = z0; d <= z1; d++) {
                                const int cid = (a * dy + b) * dz + d;
                                if (pass == 0) atomicAdd(&s_cell[cid], 1);
                                else {
                                    const int slot = atomicAdd(&s_cell[cid], 1);
                                    if (gent) { if (slot < K4_MC_GENT) X.mcent[slot] = (unsigned short)i; }
                                    else if (slot < MC_ENTRIES) s_ent[slot] = (unsigned short)i;
                                }
                            }
                        }
                        workers_sync();
                        if (pass == 0) {
                            if (ww == 0) {          // exclusive scan of the cell counts by one warp
                                int run = 0;
                                for (int base = 0; base <= ncell; base += 32) {
                                    const int i = base + lane;
                                    const int v = i <= ncell ? s_cell[i] : 0;
                                    const int inc = warp_incl_scan(v);
                                    if (i <= ncell) s_cell[i] = run + inc - v;
                                    run += __shfl_sync(0xffffffffu, inc, 31);
                                }
                                // lists that do not fit the shared array go to the block's global scratch; beyond that every sample meets every sphere
                                if (lane == 0) { if (run > K4_MC_GENT) s_mcmode = 0; else if (run > MC_ENTRIES || !insm) s_mcmode = 2; }
                            }
                            workers_sync();
                            if (!s_mcmode) break;
                        }
                    }
                    // after the fill, s_cell[c] = END of cell c; start = end of cell c-1
                }
                workers_sync();
            }
            // ---- 3b. warps 1..3: surrounding atoms (asa.c:84-112): heavy atoms of pdb_w_lig within ray+1 of a sphere, atom order ----
            float bmn[3] = {3e38f, 3e38f, 3e38f}, bmx[3] = {-3e38f, -3e38f, -3e38f};
            for (int i = wt; i < n; i += K4_WORKERS) {
                float4 v = sx[L[i]];
                float rp = v.w + 1.01f;
                bmn[0] = fminf(bmn[0], v.x - rp); bmn[1] = fminf(bmn[1], v.y - rp); bmn[2] = fminf(bmn[2], v.z - rp);
                bmx[0] = fmaxf(bmx[0], v.x + rp); bmx[1] = fmaxf(bmx[1], v.y + rp); bmx[2] = fmaxf(bmx[2], v.z + rp);
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1)
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    bmn[q] = fminf(bmn[q], __shfl_xor_sync(0xffffffffu, bmn[q], d));
                    bmx[q] = fmaxf(bmx[q], __shfl_xor_sync(0xffffffffu, bmx[q], d));
                }
            if (lane == 0) for (int q = 0; q < 3; q++) { s_bb[ww][q] = bmn[q]; s_bb[ww][3 + q] = bmx[q]; }
            workers_sync();
            for (int w = 0; w < 3; w++)
                for (int q = 0; q < 3; q++) { bmn[q] = fminf(bmn[q], s_bb[w][q]); bmx[q] = fmaxf(bmx[q], s_bb[w][3 + q]); }
            if (nw <= K4_GRID_MIN_ATOMS) {
            // every worker warp scans one contiguous third of the atoms and appends, in order, to its own segment of X.sa (measured on the
            // 2k-5k-atom workload: 110 ms per 3,000 structures, against 119 ms through the cell grid below - its sort costs more than the scan saves)
            const int a_lo = min(nw, ww * seg), a_hi = min(nw, a_lo + seg);
            int nseg = 0;
            for (int base = a_lo; base < a_hi; base += 32) {
                const int i = base + lane;
                bool in = false;
                if (i < a_hi && !(wf[WQ0 + i] & whmask)) {
                    const float x = wx[W0 + i], y = wy[W0 + i], z = wz[W0 + i];
                    if (x >= bmn[0] && x <= bmx[0] && y >= bmn[1] && y <= bmx[1] && z >= bmn[2] && z <= bmx[2]) {
                        for (int q = 0; q < n && !in; q++) {
                            const float4 v = sx[L[q]];
                            const double rp = (double)v.w + 1.0;
                            in = (double)f_ddist(x, y, z, v.x, v.y, v.z) < rp * rp;
                        }
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, in);
                if (in) X.sa[a_lo + nseg + __popc(bal & ((1u << lane) - 1u))] = i;
                nseg += __popc(bal);
            }
            if (lane == 0) s_segcnt[ww] = nseg;
            } else
            // large structures: every sphere marks, in a bit map over the atoms (X.bits), the atoms of the grid cells within its reach (AtomGrid) that
            // lie within ray + 1.0 of it; the three worker warps then read their contiguous third of the map in order, so the list comes out in atom
            // order with no sort (the FIRST atom that buries a point decides the apolar / polar tally, asa.c:313-325). Cost: spheres x ~60 nearby
            // atoms, not (atoms in the pocket's bounding box) x spheres: the 941-sphere pocket of the 150,000-atom complex dropped from 38 to ~15 ms.
            {
                const float4 gg = AG.geom[k];
                const int4 gd = AG.dims[k];
                const int *cellend = AG.cellend + AG.cbase[k];
                const int *cellatom = AG.cellatom + (S.natoms_w ? S.w_off : AG.same_base + S.atom_off);
                unsigned *bits = X.bits;
                const int nwords = (nw + 31) >> 5;
                for (int i = wt; i < nwords; i += K4_WORKERS) bits[i] = 0u;
                workers_sync();
                for (int s2 = wt; s2 < n; s2 += K4_WORKERS) {
                    const float4 v = sx[L[s2]];
                    const float rp1 = v.w + 1.01f;
                    const double rp = (double)v.w + 1.0;
                    const int x0 = max(0, min(gd.x - 1, (int)floorf((v.x - rp1 - gg.x) / gg.w))), x1 = max(0, min(gd.x - 1, (int)floorf((v.x + rp1 - gg.x) / gg.w)));
                    const int y0 = max(0, min(gd.y - 1, (int)floorf((v.y - rp1 - gg.y) / gg.w))), y1 = max(0, min(gd.y - 1, (int)floorf((v.y + rp1 - gg.y) / gg.w)));
                    const int z0 = max(0, min(gd.z - 1, (int)floorf((v.z - rp1 - gg.z) / gg.w))), z1 = max(0, min(gd.z - 1, (int)floorf((v.z + rp1 - gg.z) / gg.w)));
                    for (int cx = x0; cx <= x1; cx++) for (int cy = y0; cy <= y1; cy++) for (int cz = z0; cz <= z1; cz++) {
                        const int cid = (cx * gd.y + cy) * gd.z + cz;
                        const int lo = cid ? cellend[cid - 1] : 0, hi = cellend[cid];
                        for (int q = lo; q < hi; q++) {
                            const int i = cellatom[q];
                            if ((double)f_ddist(wx[W0 + i], wy[W0 + i], wz[W0 + i], v.x, v.y, v.z) < rp * rp) atomicOr(&bits[i >> 5], 1u << (i & 31));
                        }
                    }
                }
                workers_sync();
                // worker warp ww reads the bits of ITS third of the atoms, [a_lo, a_hi), as the direct scan does: same segments of X.sa for asa_atom
                const int a_lo = min(nw, ww * seg), a_hi = min(nw, a_lo + seg);
                int nseg = 0;
                for (int base = a_lo >> 5; base < ((a_hi + 31) >> 5); base += 32) {
                    const int wi = base + lane;
                    unsigned word = wi < ((a_hi + 31) >> 5) ? bits[wi] : 0u;
                    if ((wi << 5) < a_lo) word &= ~0u << (a_lo & 31);                           // bits below the segment's first atom
                    if ((wi << 5) + 32 > a_hi) word &= (a_hi & 31) ? ~(~0u << (a_hi & 31)) : ((wi << 5) < a_hi ? ~0u : 0u);      // bits from its end on
                    const int cnt = __popc(word);
                    const int inc = warp_incl_scan(cnt);
                    int pos = a_lo + nseg + inc - cnt;
                    for (unsigned m = word; m; m &= m - 1) X.sa[pos++] = (wi << 5) + __ffs(m) - 1;
                    nseg += __shfl_sync(0xffffffffu, inc, 31);
                }
                if (lane == 0) s_segcnt[ww] = nseg;
            }
            workers_sync();
#ifdef FPK_K4_PROFILE
            if (wt == 0) s_prof[1] = (unsigned long long)clock64();
#endif
            __threadfence_block();
            asm volatile("bar.arrive 2, 128;" ::: "memory");          // lets warp 0 in (it may still be busy with its sums: nobody waits for it here)
        }
        if (full) {
            // all four warps from here (warp 0 as soon as its sums are done and the workers have arrived at barrier 2)
            // ---- 4. 100-point spiral SASA, probes 1.4 and 2.2 (asa.c:240-419): one warp per contacted atom ----------------------------
            // warp 0 keeps its candidates in s_ent, which is idle in clusters without Monte-Carlo samples; in the others it goes straight to the samples
            float4 *cand = wid ? s_cand[wid - 1] : reinterpret_cast<float4 *>(s_ent);
            unsigned char *candapol = wid ? s_candapol[wid - 1] : reinterpret_cast<unsigned char *>(s_ent) + sizeof(float4) * ASA_CAND;
            if (wid || !observable)
                for (;;) {
                    int u = 0;
                    if (lane == 0) u = atomicAdd(&s_asanext, 1);
                    u = __shfl_sync(0xffffffffu, u, 0);
                    if (u >= U) break;
                    asa_atom(u, cand, candapol);
                }
#ifdef FPK_K4_PROFILE
            if (lane == 0) atomicMax(&s_prof[2], (unsigned long long)clock64());
#endif
            // ---- 5. Monte-Carlo volume (voronoi.c:1150-1233), radius ray-1.6, 100 x nb_mcv_iter samples ------------------------------
            if (observable) mc_samples();
        }
        __syncthreads();
#ifdef FPK_K4_PROFILE
        if (tid == 0) { prof_p[2] = clock64(); }
#endif

        // ---- 6. the ASA / volume part of the record (needs the counts of all four warps) ----------------------------------------------
        // 6a. per contacted atom, all threads: the two areas (asa.c:330,384: an f64 expression rounded to f32) and the per-atom side effects
        //     (abpa_sourrounding_prob, a0, dA, abpa: asa.c:337,392-396). The areas replace the counts in acc[4u], acc[4u+1] (as float bits);
        //     acc[4u+2] keeps the polarity of the atom for 6b.
        if (full) {
            for (int u = tid; u < U; u += nt) {
                const int a = Q0 + uat[u];
                const size_t ga = (size_t)A0 + uat[u];       // per-structure slot for the side effects
                const float crad = B.arad[a];
                const bool apolar_atom = (double)B.aen[a] < 2.8;
                const float a14 = (float)((4 * 3.1415926535897932384626433832795 * ((double)crad + 1.4) * ((double)crad + 1.4) / (double)(float)100) * (double)acc[4 * u]);
                const float a22 = (float)((4 * 3.1415926535897932384626433832795 * ((double)crad + 2.2) * ((double)crad + 2.2) / (double)(float)100) * (double)acc[4 * u + 1]);
                const float na = (float)acc[4 * u + 2], np = (float)acc[4 * u + 3];
                fxv[4 * u + 0] = na / (na + np);             // abpa_sourrounding_prob
                fxv[4 * u + 1] = a14;                        // abpatmp
                float flag = 0.0f;
                if (!explicit_pocket) atomicMax(&Pk.fx_who[2 * ga], c);
                if (!apolar_atom && a22 < a14 && (double)a14 <= 10.0 && (double)a22 >= 0.0) {
                    fxv[4 * u + 2] = a22 - a14;              // dA ; a0 = a14
                    flag = 1.0f;
                    if (!explicit_pocket) atomicMax(&Pk.fx_who[2 * ga + 1], c);
                }
                fxv[4 * u + 3] = flag;
                acc[4 * u] = __float_as_int(a14); acc[4 * u + 1] = __float_as_int(a22); acc[4 * u + 2] = apolar_atom ? 1 : 0; acc[4 * u + 3] = (int)flag;
            }
        }
        __syncthreads();
#ifdef FPK_K4_PROFILE
        if (tid == 0) { prof_p[3] = clock64(); }
#endif
        // 6b. only the float sums whose order is the reference's (atoms in first-seen order, probe 1.4 then 2.2). The rest of the block waits for
        //     this, so the per-atom values are fetched 32 atoms at a time by the lanes of warp 0 (one 16-byte load each) and handed round by shuffles:
        //     every lane adds the same values in the same order; the two passes of the reference touch disjoint sums and are fused.
        if (wid == 0) {
            float apol14 = s_desc.surf_apol_vdw14, pol14 = s_desc.surf_pol_vdw14, v14 = s_desc.surf_vdw14;
            float apol22 = s_desc.surf_apol_vdw22, pol22 = s_desc.surf_pol_vdw22, v22 = s_desc.surf_vdw22;
            int nabpa = s_desc.n_abpa;
            if (full) {
                for (int base = 0; base < U; base += 32) {
                    const int u = base + lane;
                    int4 a = make_int4(0, 0, 0, 0);
                    if (u < U) a = *reinterpret_cast<const int4 *>(acc + 4 * (size_t)u);
                    const int cnt = min(32, U - base);
                    for (int q = 0; q < cnt; q++) {
                        const float a14 = __int_as_float(__shfl_sync(0xffffffffu, a.x, q)), a22 = __int_as_float(__shfl_sync(0xffffffffu, a.y, q));
                        const int apolar_atom = __shfl_sync(0xffffffffu, a.z, q), flag = __shfl_sync(0xffffffffu, a.w, q);
                        if (apolar_atom) { apol14 = apol14 + a14; apol22 = apol22 + a22; }
                        else { pol14 = pol14 + a14; nabpa += flag; pol22 = pol22 + a22; }
                        v14 = v14 + a14; v22 = v22 + a22;
                    }
                }
            }
            if (lane == 0) {
                fpk_desc d = s_desc;
                if (full) {
                    d.surf_apol_vdw14 = apol14; d.surf_pol_vdw14 = pol14; d.surf_vdw14 = v14;
                    d.surf_apol_vdw22 = apol22; d.surf_pol_vdw22 = pol22; d.surf_vdw22 = v22; d.n_abpa = nabpa;
                    d.volume = 0.0f;
                    if (observable) {
                        const float vbox = (s_mcbox[1] - s_mcbox[0]) * (s_mcbox[3] - s_mcbox[2]) * (s_mcbox[5] - s_mcbox[4]);
                        float sum_volume = 0.0f;
                        for (int li = 0; li < 100; li++) sum_volume += ((float)s_cnt[li]) / ((float)prm.nb_mcv_iter) * vbox;
                        d.volume = sum_volume / (float)100;
                    }
                }
                d.convex_hull_volume = D->convex_hull_volume;       // written by k_hull_warp / k_hull_volume before this kernel
                *D = d;
            }
        }
    }
}

// ============================================= K4s: descriptors of the small clusters, one WARP each ==============================
// Nine clusters in ten have fewer than min_pock_nb_asph spheres: refine.c:124 always drops them, yet all of them enter the min-max
// normalisation (pocket.c:664-785) and their ASA pass leaves its marks on the atoms (asa.c:337,392-396), so they must be described.
// k_descriptors gives each of them a block of four warps that mostly wait for one another (round 1: 38 % of K4's time at 127 k cycles per
// small cluster). Here a warp does one such cluster from start to end with no block barrier - same loops, same operation order, so the
// same bits: contacted atoms, the reference-order sums, surrounding atoms, the two SASA passes, the record. Spheres, the touching-sphere
// masks (<= 32 spheres: one word per atom), the surrounding atoms and the burial candidates sit in 5.5 KB of shared memory per warp.
// A cluster that does not fit (more than K4S_SA surrounding atoms, more than ASA_CAND candidates) is left to k_descriptors: `done` tells.
enum { K4S_WARPS = 4, K4S_THREADS = 32 * K4S_WARPS, K4S_SA = 512, K4S_NMAX = 32 };
struct SmallWarp {
    float4 cand[ASA_CAND];
    float4 sph[K4S_NMAX];
    int sa[K4S_SA];
    unsigned tmask[4 * K4S_NMAX];
    unsigned char candapol[ASA_CAND];
    fpk_desc d;
    int cursor;
    int entu[4 * K4S_NMAX];             // entry (sphere, slot) -> position of its atom; then: first occurrence of the atom's residue id (descriptors.c:408)
    unsigned cnt4[4 * K4S_NMAX];        // per contacted atom the four SASA tallies (each <= 100), one byte each: free points at 1.4 / 2.2, apolar / polar burying atoms
    unsigned char plist[208];           // the test points of the current atom (0..199: probe 1.4 then 2.2) that lie inside a touching sphere
};
__global__ void __launch_bounds__(K4S_THREADS, 8) k_descriptors_small(BatchDev B, PocketDev Pk, int *chunk_counter, AtomGrid AG, unsigned char *done)
{
    __shared__ SmallWarp s_w[K4S_WARPS];
    __shared__ float s_spiral[300];          // the lanes of a warp read 32 DIFFERENT points at once: from constant memory that is 32 serialised reads
    for (int i = threadIdx.x; i < 300; i += blockDim.x) s_spiral[i] = c_spiral[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    SmallWarp &W = s_w[threadIdx.x >> 5];
    const fpk_params &prm = B.prm;
    const bool full = prm.do_asa_and_volume != 0;
    const int nlimit = min(prm.min_pock_nb_asph, K4S_NMAX + 1);
    for (;;) {
        int base = 0;
        if (lane == 0) base = atomicAdd(chunk_counter, 32);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= Pk.total_clusters) return;
        int my_k = 0, my_off = 0, my_n = 0;
        bool want = false;
        if (base + lane < Pk.total_clusters) {
            const int item = base + lane;
            int lo = 0, hi = B.nstruct;
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (Pk.clbase[mid] <= item) lo = mid; else hi = mid; }
            my_k = lo;
            const int c = item - Pk.clbase[lo];
            const int *cl_off = B.cl_off + B.S[lo].sph_off + lo;
            my_off = cl_off[c]; my_n = cl_off[c + 1] - my_off;
            const int *kc = B.counts + (size_t)lo * CNT_STRIDE;
            const bool xp = kc[CNT_XP] == 1 && c == kc[CNT_CLUSTERS] - 1;
            want = my_n < nlimit && my_n > 0 && !xp;
        }
        unsigned todo = __ballot_sync(0xffffffffu, want);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int item = base + src, k = __shfl_sync(0xffffffffu, my_k, src), off = __shfl_sync(0xffffffffu, my_off, src),
                      n = __shfl_sync(0xffffffffu, my_n, src);
            const int c = item - Pk.clbase[k];
            const StructDev S = B.S[k];
            const int R = Pk.sbase[k] + off;
            const int *L = B.cl_sph + S.sph_off + off;
            const float4 *sx = B.s_xyzr + S.sph_off;
            const int4 *sn = B.s_neigh + S.sph_off;
            const uint8_t *sty = B.s_type + S.sph_off;
            const float *sb = B.s_bary + 3 * (size_t)S.sph_off;
            int *uat = Pk.cl_atoms + 4 * (size_t)R;
            int *const entu = W.entu;                 // this kernel keeps its per-cluster scratch in shared memory (k_descriptors: global, Pk.ent_u / Pk.acc)
            float *fxv = Pk.fx_val + 16 * (size_t)R;
            fpk_desc *D = Pk.desc + item;
            const int A0 = S.atom_off, Q0 = S.prop_off;
            const bool same = S.same_pdb != 0;
            const int W0 = S.natoms_w ? S.w_off : S.atom_off, WQ0 = S.natoms_w ? S.w_off : S.prop_off, nw = S.natoms_w ? S.natoms_w : S.natoms;
            const unsigned whmask = S.natoms_w ? FPK_ATOM_H : FPK_ATOM_H1;
            const float *wx = S.natoms_w ? B.wx : B.ax, *wy = S.natoms_w ? B.wy : B.ay, *wz = S.natoms_w ? B.wz : B.az;
            const float *wr = S.natoms_w ? B.wrad : B.arad, *we = S.natoms_w ? B.wen : B.aen;
            const uint8_t *wf = S.natoms_w ? B.wfl : B.afl;
            __syncwarp();
            if (lane < n) W.sph[lane] = sx[L[lane]];
            for (int u = lane; u < 4 * K4S_NMAX; u += 32) W.tmask[u] = 0u;
            __syncwarp();
            // ---- 1. unique contacted atoms by atom id, first-seen order (pocket.c:1280-1320) -----------------------------------------
            int U = 0;
            for (int b0 = 0; b0 < 4 * n; b0 += 32) {
                const int e = b0 + lane;
                const bool valid = e < 4 * n;
                int a = -1, id = -2 - lane;
                if (valid) {
                    const int4 q = sn[L[e >> 2]];
                    const int j = e & 3;
                    a = j == 0 ? q.x : (j == 1 ? q.y : (j == 2 ? q.z : q.w));
                    id = B.aid[Q0 + a];
                }
                int dupidx = -1;
                for (int u = 0; u < U; u++)
                    if (B.aid[Q0 + uat[u]] == id) dupidx = u;
                const unsigned m = __match_any_sync(0xffffffffu, id);
                const int leader = __ffs(m) - 1;
                const bool isnew = valid && dupidx < 0 && lane == leader;
                const unsigned bal = __ballot_sync(0xffffffffu, isnew);
                const int pos = U + __popc(bal & ((1u << lane) - 1u));
                if (isnew) uat[pos] = a;
                const int lp = __shfl_sync(0xffffffffu, pos, leader);
                if (valid) { const int uu = dupidx >= 0 ? dupidx : lp; entu[e] = uu; atomicOr(&W.tmask[uu], 1u << (e >> 2)); }
                U += __popc(bal);
                __syncwarp();
            }
            // ---- 3a. the sequential float sums, in the reference's order (descriptors.c:186-236, same code as k_descriptors' warp 0) ---
            float as_density = 0.0f, as_max_dst = -1.0f, lane_max = -1.0f;
            int mlhd_cnt = 0, nAlphaApol = 0;
            for (int i = 0; i < n; i++) {
                const float4 vi = W.sph[i];
                const bool apol_i = sty[L[i]] == 0;
                int napol = 0;
                if (apol_i) {
                    for (int j = lane; j < n; j += 32) {
                        if (j == i || sty[L[j]] != 0) continue;
                        const float4 vj = W.sph[j];
                        if ((double)(f_dist(vi.x, vi.y, vi.z, vj.x, vj.y, vj.z) - (vj.w + vi.w)) <= 0.) napol++;
                    }
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) napol += __shfl_xor_sync(0xffffffffu, napol, d);
                    mlhd_cnt += napol;
                    nAlphaApol++;
                }
                for (int b0 = i + 1; b0 < n; b0 += 32) {
                    const int j = b0 + lane;
                    float dj = 0.0f;
                    if (j < n) {
                        const float4 vj = W.sph[j];
                        dj = f_dist(vi.x, vi.y, vi.z, vj.x, vj.y, vj.z);
                        lane_max = fmaxf(lane_max, dj);
                    }
                    const int cntj = min(32, n - b0);
                    for (int q = 0; q < cntj; q++) as_density += __shfl_sync(0xffffffffu, dj, q);
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) lane_max = fmaxf(lane_max, __shfl_xor_sync(0xffffffffu, lane_max, d));
            if (lane_max > as_max_dst) as_max_dst = lane_max;
            __syncwarp();
            for (int u = lane; u < U; u += 32) {        // first occurrence of every residue id (descriptors.c:408 in_tab)
                const int rid = B.ares[Q0 + uat[u]];
                bool seen = false;
                for (int q = 0; q < u && !seen; q++) seen = B.ares[Q0 + uat[q]] == rid;
                entu[u] = seen ? 0 : 1;
            }
            __syncwarp();
            if (lane == 0) {
                float mean_r = 0.0f, masph = 0.0f, as_max_r = -1.0f, ps0 = 0.0f, ps1 = 0.0f, ps2 = 0.0f;
                int napolc = 0;
                for (int i = 0; i < n; i++) {
                    const int si = L[i];
                    const float4 v = W.sph[i];
                    if (v.w > as_max_r) as_max_r = v.w;
                    mean_r += v.w;
                    const float dd = f_dist(v.x, v.y, v.z, sb[3 * si], sb[3 * si + 1], sb[3 * si + 2]);
                    masph += dd / v.w;
                    ps0 += v.x; ps1 += v.y; ps2 += v.z;
                    if (sty[si] == 0) napolc++;
                }
                fpk_desc d;
                memset(&d, 0, sizeof d);
                d.first_sphere = L[0];
                d.nAlphaApol = napolc; d.nAlphaPol = n - napolc;
                d.bary[0] = ps0 / (float)n; d.bary[1] = ps1 / (float)n; d.bary[2] = ps2 / (float)n;
                d.n_atoms = U;
                if (U > 1) {
                    int nb_res = 0, nb_polar = 0;
                    float flex = 0.0f, hyd = 0.0f, vol = 0.0f;
                    for (int u = 0; u < U; u++) {
                        const int a = Q0 + uat[u];
                        if (entu[u]) {
                            const int aa = B.aaa[a];
                            if (aa >= 0) {
                                d.aa_compo[aa]++;
                                hyd += c_aa_hyd[aa]; d.polarity_score += c_aa_pol[aa]; vol += c_aa_vol[aa]; d.charge_score += c_aa_chg[aa];
                            }
                            nb_res++;
                        }
                        flex += B.abf[a];
                        if ((double)B.aen[a] > 2.7) nb_polar++;
                    }
                    d.hydrophobicity_score = hyd / (float)nb_res;
                    d.volume_score = vol / (float)nb_res;
                    d.flex = flex / (float)U;
                    d.prop_polar_atm = (float)((double)(((float)nb_polar) / ((float)U)) * 100.0);
                }
                d.mean_loc_hyd_dens = nAlphaApol > 0 ? (float)mlhd_cnt / (float)nAlphaApol : 0.0f;
                d.as_max_dst = as_max_dst;
                d.apolar_asphere_prop = (float)nAlphaApol / (float)n;
                d.masph_sacc = masph / (float)n;
                d.mean_asph_ray = mean_r / (float)n;
                d.nb_asph = n;
                d.as_density = (float)((double)as_density / ((n * n - n) * 0.5));
                d.as_max_r = as_max_r;
                W.d = d;
            }
            __syncwarp();
            bool fits = true;
            if (full) {
                // ---- 3b. surrounding atoms (asa.c:84-112), in atom order ---------------------------------------------------------------
                float bmn[3] = {3e38f, 3e38f, 3e38f}, bmx[3] = {-3e38f, -3e38f, -3e38f};
                if (lane < n) {
                    const float4 v = W.sph[lane];
                    const float rp = v.w + 1.01f;
                    bmn[0] = v.x - rp; bmn[1] = v.y - rp; bmn[2] = v.z - rp; bmx[0] = v.x + rp; bmx[1] = v.y + rp; bmx[2] = v.z + rp;
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1)
#pragma unroll
                    for (int q = 0; q < 3; q++) {
                        bmn[q] = fminf(bmn[q], __shfl_xor_sync(0xffffffffu, bmn[q], d));
                        bmx[q] = fmaxf(bmx[q], __shfl_xor_sync(0xffffffffu, bmx[q], d));
                    }
                int nsa = 0;
                if (nw <= K4_GRID_MIN_ATOMS) {
                    // runs of 32 consecutive atoms whose box meets the cluster's, in ascending order (AtomGrid.cbox): same atoms, same order as the full scan
                    const float4 *cbx = AG.cbox + 2 * chunk_base(AG, S, k);
                    const int nch = (nw + 31) >> 5;
                    for (int c0 = 0; c0 < nch && fits; c0 += 32) {
                        const int ci = c0 + lane;
                        bool ov = false;
                        if (ci < nch) {
                            const float4 lo = cbx[2 * ci], hi = cbx[2 * ci + 1];
                            ov = lo.x <= bmx[0] && hi.x >= bmn[0] && lo.y <= bmx[1] && hi.y >= bmn[1] && lo.z <= bmx[2] && hi.z >= bmn[2];
                        }
                        unsigned runs = __ballot_sync(0xffffffffu, ov);
                        while (runs && fits) {
                            const int i = 32 * (c0 + __ffs(runs) - 1) + lane;
                            runs &= runs - 1;
                            bool in = false;
                            if (i < nw && !(wf[WQ0 + i] & whmask)) {
                                const float x = wx[W0 + i], y = wy[W0 + i], z = wz[W0 + i];
                                if (x >= bmn[0] && x <= bmx[0] && y >= bmn[1] && y <= bmx[1] && z >= bmn[2] && z <= bmx[2]) {
                                    for (int q = 0; q < n && !in; q++) {
                                        const float4 v = W.sph[q];
                                        const double rp = (double)v.w + 1.0;
                                        in = (double)f_ddist(x, y, z, v.x, v.y, v.z) < rp * rp;
                                    }
                                }
                            }
                            const unsigned bal = __ballot_sync(0xffffffffu, in);
                            if (nsa + __popc(bal) > K4S_SA) { fits = false; break; }
                            if (in) W.sa[nsa + __popc(bal & ((1u << lane) - 1u))] = i;
                            nsa += __popc(bal);
                        }
                    }
                } else {                   // large structure: the cells of the atom grid the box overlaps, then a sort (atom order matters, asa.c:313-325)
                    const float4 gg = AG.geom[k];
                    const int4 gd = AG.dims[k];
                    const int *cellend = AG.cellend + AG.cbase[k];
                    const int *cellatom = AG.cellatom + (S.natoms_w ? S.w_off : AG.same_base + S.atom_off);
                    const int x0 = max(0, min(gd.x - 1, (int)floorf((bmn[0] - gg.x) / gg.w))), x1 = max(0, min(gd.x - 1, (int)floorf((bmx[0] - gg.x) / gg.w)));
                    const int y0 = max(0, min(gd.y - 1, (int)floorf((bmn[1] - gg.y) / gg.w))), y1 = max(0, min(gd.y - 1, (int)floorf((bmx[1] - gg.y) / gg.w)));
                    const int z0 = max(0, min(gd.z - 1, (int)floorf((bmn[2] - gg.z) / gg.w))), z1 = max(0, min(gd.z - 1, (int)floorf((bmx[2] - gg.z) / gg.w)));
                    const int ny = y1 - y0 + 1, nz = z1 - z0 + 1, ncb = (x1 - x0 + 1) * ny * nz;
                    int *cursor = &W.cursor;
                    if (lane == 0) *cursor = 0;
                    __syncwarp();
                    for (int ci = lane; ci < ncb; ci += 32) {
                        const int cz = z0 + ci % nz, cy = y0 + (ci / nz) % ny, cx = x0 + ci / (nz * ny);
                        const int cid = (cx * gd.y + cy) * gd.z + cz;
                        const int lo = cid ? cellend[cid - 1] : 0, hi = cellend[cid];
                        for (int q2 = lo; q2 < hi; q2++) {
                            const int i = cellatom[q2];
                            const float x = wx[W0 + i], y = wy[W0 + i], z = wz[W0 + i];
                            if (!(x >= bmn[0] && x <= bmx[0] && y >= bmn[1] && y <= bmx[1] && z >= bmn[2] && z <= bmx[2])) continue;
                            bool in = false;
                            for (int q = 0; q < n && !in; q++) {
                                const float4 v = W.sph[q];
                                const double rp = (double)v.w + 1.0;
                                in = (double)f_ddist(x, y, z, v.x, v.y, v.z) < rp * rp;
                            }
                            if (in) { const int pos = atomicAdd(cursor, 1); if (pos < K4S_SA) W.sa[pos] = i; }
                        }
                    }
                    __syncwarp();
                    nsa = *cursor;
                    __syncwarp();
                    if (lane == 0) *cursor = 0;
                    if (nsa > K4S_SA) fits = false;
                    else {
                        int npad = 32;
                        while (npad < nsa) npad <<= 1;
                        for (int i = nsa + lane; i < npad; i += 32) W.sa[i] = 0x7fffffff;
                        __syncwarp();
                        for (int kk = 2; kk <= npad; kk <<= 1)
                            for (int j = kk >> 1; j > 0; j >>= 1) {
                                for (int i = lane; i < npad; i += 32) {
                                    const int ixj = i ^ j;
                                    if (ixj > i) {
                                        const int a = W.sa[i], b2 = W.sa[ixj];
                                        if ((a > b2) == ((i & kk) == 0)) { W.sa[i] = b2; W.sa[ixj] = a; }
                                    }
                                }
                                __syncwarp();
                            }
                    }
                }
                __syncwarp();
                // ---- 4. 100-point spiral SASA, probes 1.4 and 2.2 (asa.c:240-419), one contacted atom after the other --------------------
                for (int u = 0; u < U && fits; u++) {
                    const int cura = uat[u];
                    const float cx = B.ax[A0 + cura], cy = B.ay[A0 + cura], cz = B.az[A0 + cura], crad = B.arad[Q0 + cura];
                    const float reach = crad + 2.2f + 2.2f + 0.01f;
                    int ncand = 0;
                    for (int b0 = 0; b0 < nsa; b0 += 32) {
                        const int j = b0 + lane;
                        bool take = false;
                        int a = -1;
                        float ax_ = 0, ay_ = 0, az_ = 0, ar_ = 0;
                        if (j < nsa) {
                            a = W.sa[j];
                            if (!(same && a == cura)) {                 // asa.c:315 pointer inequality
                                ax_ = wx[W0 + a]; ay_ = wy[W0 + a]; az_ = wz[W0 + a]; ar_ = wr[WQ0 + a];
                                const float dx = ax_ - cx, dy = ay_ - cy, dz = az_ - cz, rr = reach + ar_;
                                take = dx * dx + dy * dy + dz * dz <= rr * rr;
                            }
                        }
                        const unsigned bal = __ballot_sync(0xffffffffu, take);
                        if (ncand + __popc(bal) > ASA_CAND) { fits = false; break; }
                        if (take) {
                            const int pos = ncand + __popc(bal & ((1u << lane) - 1u));
                            W.cand[pos] = make_float4(ax_, ay_, az_, ar_);
                            W.candapol[pos] = (double)we[WQ0 + a] < 2.8 ? 1 : 0;
                        }
                        ncand += __popc(bal);
                    }
                    __syncwarp();
                    if (!fits) break;
                    const unsigned touching = W.tmask[u];
                    int c14 = 0, c22 = 0, cap = 0, cpo = 0;
                    // two steps, so that the burial loop runs with full warps: (a) which of the 200 test points lie inside a touching sphere (about one in
                    // five; the others count for nothing, asa.c:306-312), (b) those, 32 at a time, against the burial candidates. Only counts: order-free.
                    int npts = 0;
                    for (int r = 0; r < 7; r++) {
                        const int it = lane + 32 * r;
                        bool inside = false;
                        if (it < 200) {
                            const int pass = it >= 100 ? 1 : 0, kp = it - 100 * pass;
                            const double probe = pass == 0 ? 1.4 : 2.2;
                            const float tx = (float)((double)cx + (double)s_spiral[3 * kp] * ((double)crad + probe));
                            const float ty = (float)((double)cy + (double)s_spiral[3 * kp + 1] * ((double)crad + probe));
                            const float tz = (float)((double)cz + (double)s_spiral[3 * kp + 2] * ((double)crad + probe));
                            for (unsigned tm = touching; tm && !inside; tm &= tm - 1) {
                                const float4 v = W.sph[__ffs(tm) - 1];
                                if (f_ddist(tx, ty, tz, v.x, v.y, v.z) <= v.w * v.w) inside = true;
                            }
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, inside);
                        if (inside) W.plist[npts + __popc(m & ((1u << lane) - 1u))] = (unsigned char)it;
                        npts += __popc(m);
                    }
                    __syncwarp();
                    for (int b0 = 0; b0 < npts; b0 += 32) {
                        if (b0 + lane >= npts) continue;
                        const int it = W.plist[b0 + lane];
                        const int pass = it >= 100 ? 1 : 0, kp = it - 100 * pass;
                        const double probe = pass == 0 ? 1.4 : 2.2;
                        const float tx = (float)((double)cx + (double)s_spiral[3 * kp] * ((double)crad + probe));
                        const float ty = (float)((double)cy + (double)s_spiral[3 * kp + 1] * ((double)crad + probe));
                        const float tz = (float)((double)cz + (double)s_spiral[3 * kp + 2] * ((double)crad + probe));
                        int hit = -1;
                        for (int j = 0; j < ncand; j++) {
                            const float4 q = W.cand[j];
                            const float dsq = f_ddist(tx, ty, tz, q.x, q.y, q.z);
                            const double rr = (double)q.w + probe;
                            if ((double)dsq < rr * rr) { hit = W.candapol[j]; break; }
                        }
                        if (hit < 0) { if (pass == 0) c14++; else c22++; }
                        else if (pass == 0) { if (hit) cap++; else cpo++; }
                    }
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) {
                        c14 += __shfl_xor_sync(0xffffffffu, c14, d); c22 += __shfl_xor_sync(0xffffffffu, c22, d);
                        cap += __shfl_xor_sync(0xffffffffu, cap, d); cpo += __shfl_xor_sync(0xffffffffu, cpo, d);
                    }
                    if (lane == 0) W.cnt4[u] = (unsigned)c14 | ((unsigned)c22 << 8) | ((unsigned)cap << 16) | ((unsigned)cpo << 24);
                    __syncwarp();
                }
            }
            if (!fits) continue;                    // k_descriptors takes it (done[item] stays 0); nothing written so far is read by anyone else
            __syncwarp();
            // ---- 6. per contacted atom: the two areas and the side effects on the atom (asa.c:330-337,384-396), and - 32 atoms at a time, handed
            // round by shuffles so that every lane adds the same values in the same order - the float sums whose order is the reference's (atoms
            // in first-seen order, probe 1.4 then 2.2; the two passes of the reference touch disjoint sums and are fused)
            float apol14 = 0.0f, pol14 = 0.0f, v14 = 0.0f, apol22 = 0.0f, pol22 = 0.0f, v22 = 0.0f;
            int nabpa = 0;
            if (full) {
                for (int b0 = 0; b0 < U; b0 += 32) {
                    const int u = b0 + lane;
                    float my14 = 0.0f, my22 = 0.0f;
                    int myapol = 0, myflag = 0;
                    if (u < U) {
                        const int a = Q0 + uat[u];
                        const size_t ga = (size_t)A0 + uat[u];
                        const float crad = B.arad[a];
                        const bool apolar_atom = (double)B.aen[a] < 2.8;
                        const unsigned c4 = W.cnt4[u];
                        const float a14 = (float)((4 * 3.1415926535897932384626433832795 * ((double)crad + 1.4) * ((double)crad + 1.4) / (double)(float)100) * (double)(int)(c4 & 0xffu));
                        const float a22 = (float)((4 * 3.1415926535897932384626433832795 * ((double)crad + 2.2) * ((double)crad + 2.2) / (double)(float)100) * (double)(int)((c4 >> 8) & 0xffu));
                        const float na = (float)(int)((c4 >> 16) & 0xffu), np = (float)(int)(c4 >> 24);
                        fxv[4 * u + 0] = na / (na + np);
                        fxv[4 * u + 1] = a14;
                        float flag = 0.0f;
                        atomicMax(&Pk.fx_who[2 * ga], c);
                        if (!apolar_atom && a22 < a14 && (double)a14 <= 10.0 && (double)a22 >= 0.0) {
                            fxv[4 * u + 2] = a22 - a14;
                            flag = 1.0f;
                            atomicMax(&Pk.fx_who[2 * ga + 1], c);
                        }
                        fxv[4 * u + 3] = flag;
                        my14 = a14; my22 = a22; myapol = apolar_atom ? 1 : 0; myflag = (int)flag;
                    }
                    const int cntu = min(32, U - b0);
                    for (int q = 0; q < cntu; q++) {
                        const float a14 = __shfl_sync(0xffffffffu, my14, q), a22 = __shfl_sync(0xffffffffu, my22, q);
                        const int apolar_atom = __shfl_sync(0xffffffffu, myapol, q), flag = __shfl_sync(0xffffffffu, myflag, q);
                        if (apolar_atom) { apol14 = apol14 + a14; apol22 = apol22 + a22; }
                        else { pol14 = pol14 + a14; nabpa += flag; pol22 = pol22 + a22; }
                        v14 = v14 + a14; v22 = v22 + a22;
                    }
                }
            }
            if (lane == 0) {
                fpk_desc d = W.d;
                if (full) {
                    d.surf_apol_vdw14 = apol14; d.surf_pol_vdw14 = pol14; d.surf_vdw14 = v14;
                    d.surf_apol_vdw22 = apol22; d.surf_pol_vdw22 = pol22; d.surf_vdw22 = v22; d.n_abpa = nabpa;
                    d.volume = 0.0f;               // below min_pock_nb_asph: always dropped, no Monte-Carlo samples (see k_descriptors)
                }
                d.convex_hull_volume = D->convex_hull_volume;
                *D = d;
                done[item] = 1;
            }
            __syncwarp();
        }
    }
}

// ============================================= K4c: finalize per structure ====================================================
__global__ void __launch_bounds__(128) k_finalize(BatchDev B, PocketDev Pk)
{
    const int k = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    if (k >= B.nstruct) return;
    int *counts = B.counts + (size_t)k * CNT_STRIDE;
    const StructDev S = B.S[k];
    if (counts[CNT_STATUS] != FPK_OK) { if (tid == 0) counts[CNT_POCKETS] = 0; return; }
    const int nc = counts[CNT_CLUSTERS] - (counts[CNT_XP] == 1 ? 1 : 0);        // dpocket's explicit pocket is not one of the pockets
    fpk_desc *D = Pk.desc + Pk.clbase[k];
    const fpk_params &prm = B.prm;
    __shared__ float s_f[10];
    __shared__ int s_i[6];
    // ---- pocket.c:664-785: min/max over all clusters. The reference scans in list order: M = m = first value, then "if (v > M) M = v; else if
    // (v < m) m = v". Since m <= M always, that is the plain maximum and minimum over the values that compare at all (a NaN never does), EXCEPT
    // that a NaN FIRST value stays in both bounds for good. So: block-wide max / min that skip NaNs, then the first value's NaN rule.
    {
        __shared__ float s_rf[4][10];
        __shared__ int s_ri[4][4];
        float fM[5], fm[5];
        int iM[2], im[2];
#pragma unroll
        for (int q = 0; q < 5; q++) { fM[q] = -3.4e38f; fm[q] = 3.4e38f; }
        iM[0] = iM[1] = -2147483647 - 1; im[0] = im[1] = 2147483647;
        for (int i = tid; i < nc; i += nt) {
            const fpk_desc &d = D[i];
            const float v[5] = {d.as_max_dst, d.as_density, d.mean_loc_hyd_dens, d.flex, d.apolar_asphere_prop};
#pragma unroll
            for (int q = 0; q < 5; q++) { if (v[q] > fM[q]) fM[q] = v[q]; if (v[q] < fm[q]) fm[q] = v[q]; }
            iM[0] = max(iM[0], d.polarity_score); im[0] = min(im[0], d.polarity_score);
            iM[1] = max(iM[1], d.nb_asph); im[1] = min(im[1], d.nb_asph);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int q = 0; q < 5; q++) { fM[q] = fmaxf(fM[q], __shfl_xor_sync(0xffffffffu, fM[q], o)); fm[q] = fminf(fm[q], __shfl_xor_sync(0xffffffffu, fm[q], o)); }
#pragma unroll
            for (int q = 0; q < 2; q++) { iM[q] = max(iM[q], __shfl_xor_sync(0xffffffffu, iM[q], o)); im[q] = min(im[q], __shfl_xor_sync(0xffffffffu, im[q], o)); }
        }
        if ((tid & 31) == 0) {
            for (int q = 0; q < 5; q++) { s_rf[tid >> 5][2 * q] = fM[q]; s_rf[tid >> 5][2 * q + 1] = fm[q]; }
            for (int q = 0; q < 2; q++) { s_ri[tid >> 5][2 * q] = iM[q]; s_ri[tid >> 5][2 * q + 1] = im[q]; }
        }
        __syncthreads();
        if (tid == 0 && nc > 0) {
            const fpk_desc &d0 = D[0];
            const float v0[5] = {d0.as_max_dst, d0.as_density, d0.mean_loc_hyd_dens, d0.flex, d0.apolar_asphere_prop};
            for (int q = 0; q < 5; q++) {
                float M = s_rf[0][2 * q], m = s_rf[0][2 * q + 1];
                for (int w = 1; w < (nt >> 5); w++) { M = fmaxf(M, s_rf[w][2 * q]); m = fminf(m, s_rf[w][2 * q + 1]); }
                if (v0[q] != v0[q]) M = m = v0[q];              // the first value is a NaN: nothing compares with it afterwards
                s_f[2 * q] = M; s_f[2 * q + 1] = m;
            }
            for (int q = 0; q < 2; q++) {
                int M = s_ri[0][2 * q], m = s_ri[0][2 * q + 1];
                for (int w = 1; w < (nt >> 5); w++) { M = max(M, s_ri[w][2 * q]); m = min(m, s_ri[w][2 * q + 1]); }
                s_i[2 * q] = M; s_i[2 * q + 1] = m;
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < nc; i += nt) {
        fpk_desc d = D[i];
        if (nc > 1) {
            if ((double)(s_f[0] - s_f[1]) != 0.0) d.as_max_dst_norm = (d.as_max_dst - s_f[1]) / (s_f[0] - s_f[1]);
            if ((double)(s_f[2] - s_f[3]) != 0.0) d.as_density_norm = (d.as_density - s_f[3]) / (s_f[2] - s_f[3]);
            if ((double)(s_i[0] - s_i[1]) != 0.0) d.polarity_score_norm = (float)(d.polarity_score - s_i[1]) / (float)(s_i[0] - s_i[1]);
            if ((double)(s_f[4] - s_f[5]) != 0.0) d.mean_loc_hyd_dens_norm = (d.mean_loc_hyd_dens - s_f[5]) / (s_f[4] - s_f[5]);
            if ((double)(s_f[6] - s_f[7]) != 0.0) d.flex = (d.flex - s_f[7]) / (s_f[6] - s_f[7]);
            if ((double)(s_i[2] - s_i[3]) != 0.0) d.nas_norm = (float)(d.nb_asph - s_i[3]) / (float)(s_i[2] - s_i[3]);
            if ((double)(s_f[8] - s_f[9]) != 0.0) d.prop_asapol_norm = (d.apolar_asphere_prop - s_f[9]) / (s_f[8] - s_f[9]);
        } else {
            d.polarity_score_norm = 0.0f; d.as_density_norm = 0.0f; d.as_max_dst_norm = 0.0f;
            d.mean_loc_hyd_dens_norm = (float)(((double)d.mean_loc_hyd_dens - 8.23) / (24.20 - 8.23));
            d.flex = 0.0f; d.nas_norm = 0.0f; d.prop_asapol_norm = 0.0f;
        }
        // pscoring.c:241-247 and :325-329
        d.score = (float)(-0.03783394 + 0.48461469 * (double)d.nas_norm + 0.09093926 * (double)d.as_density +
                          0.0004155899 * (double)d.convex_hull_volume - 0.003995233 * (double)d.surf_pol_vdw14 -
                          0.004072336 * (double)d.surf_apol_vdw14);
        const float b0 = (float)-9.5698768, b1 = (float)7.479844, b2 = (float)0.3696134, b3 = (float)-0.04671833;
        d.drug_score = (float)(1.0 / (1.0 + exp((double)(-(b0 + b1 * d.mean_loc_hyd_dens_norm + b2 * d.as_max_dst + b3 * d.surf_pol_vdw22)))));
        d.final_rank = 0;
        D[i] = d;
    }
    __syncthreads();
    // ---- refine.c:114-145 drop, psorting.c:64-165 first-pivot quicksort (same partition scheme => same tie order) -- one thread ----
    if (tid == 0) {
        int *keep = Pk.rank_to_cluster + Pk.sbase[k] + k, nk = 0;
        for (int i = 0; i < nc; i++) {
            const fpk_desc &d = D[i];
            double pasph = (float)((float)d.nAlphaApol / (float)d.nb_asph);
            int rhs = (prm.refine_min_apolar_asphere_prop != 0.0f) || (d.as_density < prm.min_as_density);
            if (d.nb_asph < prm.min_pock_nb_asph || pasph < (double)rhs) continue;
            keep[nk++] = i;
        }
        int stack[64][2], sp = 0;
        if (nk > 1) { stack[0][0] = 0; stack[0][1] = nk - 1; sp = 1; }
        while (sp > 0) {
            sp--;
            int start = stack[sp][0], end = stack[sp][1];
            while (start < end) {
                int cpos = start, piv = keep[start], tmp;
                for (int i = start + 1; i <= end; i++) {
                    int cp = keep[i];
                    if (!(D[cp].score < D[piv].score)) {
                        cpos++;
                        if (cpos != i) { tmp = keep[cpos]; keep[cpos] = keep[i]; keep[i] = tmp; }
                    }
                }
                if (cpos != start) { tmp = keep[cpos]; keep[cpos] = keep[start]; keep[start] = tmp; }
                // recurse on the smaller half through the explicit stack, iterate on the other (order of the halves is immaterial)
                int l0 = start, l1 = cpos - 1, r0 = cpos + 1, r1 = end;
                if (l1 - l0 < r1 - r0) { if (l0 < l1 && sp < 64) { stack[sp][0] = l0; stack[sp][1] = l1; sp++; } start = r0; end = r1; }
                else { if (r0 < r1 && sp < 64) { stack[sp][0] = r0; stack[sp][1] = r1; sp++; } start = l0; end = l1; }
            }
        }
        for (int r = 0; r < nk; r++) D[keep[r]].final_rank = r + 1;
        counts[CNT_POCKETS] = nk;
        if (nk == 0) counts[CNT_STATUS] = FPK_ERR_NO_POCKETS;
    }
}

// per-atom ASA side effects (asa.c:337,392-396): the LAST pocket (in list order) that wrote a field wins
__global__ void k_atom_fx(BatchDev B, PocketDev Pk)
{
    const int item = blockIdx.x;
    if (item >= Pk.total_clusters) return;
    int lo = 0, hi = B.nstruct;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (Pk.clbase[mid] <= item) lo = mid; else hi = mid; }
    const int k = lo, c = item - Pk.clbase[k];
    const StructDev S = B.S[k];
    if (B.counts[(size_t)k * CNT_STRIDE + CNT_CLUSTERS] - (B.counts[(size_t)k * CNT_STRIDE + CNT_XP] == 1 ? 1 : 0) <= c) return;
    const int *cl_off = B.cl_off + S.sph_off + k;
    const int R = Pk.sbase[k] + cl_off[c];
    const int U = Pk.desc[item].n_atoms;
    const int *uat = Pk.cl_atoms + 4 * (size_t)R;
    const float *fxv = Pk.fx_val + 16 * (size_t)R;
    for (int u = threadIdx.x; u < U; u += blockDim.x) {
        const size_t a = (size_t)S.atom_off + uat[u];
        if (Pk.fx_who[2 * a] == c) Pk.atom_fx[4 * a + 3] = fxv[4 * u + 0];
        if (fxv[4 * u + 3] != 0.0f && Pk.fx_who[2 * a + 1] == c) {
            Pk.atom_fx[4 * a + 0] = 1.0f; Pk.atom_fx[4 * a + 1] = fxv[4 * u + 1]; Pk.atom_fx[4 * a + 2] = fxv[4 * u + 2];
        }
    }
}

}  // namespace fpk
