// K3L: the agglomerative linkages behind -C m | a | c (reference clusterlib.c: pmlcluster :3705 complete, palcluster :3781 average,
// pclcluster :3337 centroid), cut by cuttree_distance (:3235). One block per structure; its ragged lower-triangular f64 distance matrix
// (distancematrix :2919) lives in a pool slice the host sized from the structure's sphere count.
//
// The reference repeats n-1 times: find the FIRST strict minimum of the matrix in row-major order (find_closest_pair :285), fold row/column
// `is` into `js` with the linkage rule, move the last row into slot `is`. Ties are decided by that scan order in the permuted matrix, so the
// layout and the moves are kept exactly; only the search changes: every row carries its minimum and the first column that holds it, the pair is
// the first row with the smallest minimum (a 256-thread reduction over n numbers instead of n^2/2), and after a merge only rows whose
// remembered column was touched are scanned again, one warp per row. O(n^2) reads per structure instead of the reference's O(n^3).
//
// The cut walks the nodes from the last merge down and lets a node at or below the cut label every leaf under it that has no label yet
// (first label wins): a leaf's cluster is its highest-numbered ancestor with distance <= cut - with centroid linkage distances are not
// monotone and that ancestor may sit above nodes that exceed the cut - and a leaf with no such ancestor stays alone. Here every leaf climbs
// its own ancestor chain. The block leaves  s_cluster[i] = lowest sphere of i's cluster ; k_cluster then numbers the clusters by that
// sphere and fills the member lists exactly as for single linkage (refine.c:58-93).
#pragma once
#include "k_pockets.cuh"

namespace fpk {

struct LinkItem {
    int k;                      // structure
    int ns;                     // its kept spheres (the host read the counts back to size the pool)
    unsigned long long off;     // byte offset of its slice in the pool
};

__host__ __device__ inline size_t linkage_bytes(size_t ns)
{
    // D | pos[3 ns] rmin[ns] nd[ns] (f64) | rarg cnt cid npar lpar omin list (int)
    return 8 * (ns * (ns - 1) / 2 + 1) + 8 * 5 * ns + 4 * 7 * ns + 512;
}

__device__ __forceinline__ double metric_rows(bool cityblock, const double *a, const double *b)
{
    double result = 0.;
    if (cityblock) {                    // clusterlib.c:1022-1083 with weights 1: the MEAN of the three absolute differences
        double tweight = 0;
#pragma unroll
        for (int c = 0; c < 3; c++) { const double term = __dsub_rn(a[c], b[c]); result = __dadd_rn(result, __dmul_rn(1.0, fabs(term))); tweight = __dadd_rn(tweight, 1.0); }
        return __ddiv_rn(result, tweight);
    }
#pragma unroll
    for (int c = 0; c < 3; c++) { const double term = __dsub_rn(a[c], b[c]); result = __dadd_rn(result, __dmul_rn(term, term)); }
    return __dsqrt_rn(result);         // clusterlib.c:933-1018 (the division by tweight is commented out in the reference)
}

// minimum of row i over its first i columns and the FIRST column holding it; the whole warp calls this
__device__ __forceinline__ void link_scan_row(const double *row, int i, int lane, double &mv, int &mj)
{
    double v = 1e300;
    int c = 0x7fffffff;
    for (int j = lane; j < i; j += 32) { const double t = row[j]; if (t < v) { v = t; c = j; } }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, d);
        const int oc = __shfl_xor_sync(0xffffffffu, c, d);
        if (ov < v || (ov == v && oc < c)) { v = ov; c = oc; }
    }
    mv = v; mj = c;
}

__global__ void __launch_bounds__(256) k_linkage(BatchDev B, const LinkItem *items, unsigned char *pool)
{
    __shared__ double s_v[8];
    __shared__ int s_i[8], s_is, s_js, s_nlist;
    __shared__ double s_best;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const LinkItem it = items[blockIdx.x];
    const StructDev S = B.S[it.k];
    const int *counts = B.counts + (size_t)it.k * CNT_STRIDE;
    if (counts[CNT_STATUS] != FPK_OK) return;
    const int ns = counts[CNT_SPHERES];
    if (ns < 2 || ns != it.ns) return;
    const float4 *sx = B.s_xyzr + S.sph_off;
    int *par = B.s_cluster + S.sph_off;
    unsigned char *base = pool + it.off;
    const size_t n0 = (size_t)ns;
    double *D = (double *)base;
    double *pos = D + (n0 * (n0 - 1) / 2 + 1), *rmin = pos + 3 * n0, *nd = rmin + n0;
    int *rarg = (int *)(nd + n0), *cnt = rarg + n0, *cid = cnt + n0, *npar = cid + n0, *lpar = npar + n0, *omin = lpar + n0, *list = omin + n0;
#define DM(i, j) D[(size_t)(i) * ((size_t)(i) - 1) / 2 + (size_t)(j)]                 /* i > j */
    const bool cityblock = B.prm.distance_measure == 'b';
    const int method = B.prm.clustering_method;
    const double cut = (double)B.prm.clust_max_dist;

    for (int i = tid; i < ns; i += nt) {
        const float4 v = sx[i];
        pos[3 * i] = (double)v.x; pos[3 * i + 1] = (double)v.y; pos[3 * i + 2] = (double)v.z;
        cnt[i] = 1; cid[i] = i; npar[i] = -1; lpar[i] = -1; omin[i] = 0x7fffffff;
    }
    __syncthreads();
    for (int i = 1 + wid; i < ns; i += nw) {                 // the matrix, and every row's first minimum
        double *row = &DM(i, 0);
        const double a[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
        for (int j = lane; j < i; j += 32) row[j] = metric_rows(cityblock, a, pos + 3 * j);
        __syncwarp();
        double mv; int mj;
        link_scan_row(row, i, lane, mv, mj);
        if (lane == 0) { rmin[i] = mv; rarg[i] = mj; }
    }
    __syncthreads();

    for (int n = ns, node = 0; n > 1; n--, node++) {
        // ---- the closest pair: first row (then first column) with the smallest entry -----------------------------------------------
        double v = 1e300;
        int bi = 0x7fffffff;
        for (int i = 1 + tid; i < n; i += nt) { const double t = rmin[i]; if (t < v) { v = t; bi = i; } }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, d);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (ov < v || (ov == v && oi < bi)) { v = ov; bi = oi; }
        }
        if (lane == 0) { s_v[wid] = v; s_i[wid] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < nw; w++) if (s_v[w] < v || (s_v[w] == v && s_i[w] < bi)) { v = s_v[w]; bi = s_i[w]; }
            s_is = bi; s_js = rarg[bi]; s_best = v; s_nlist = 0;
        }
        __syncthreads();
        const int is = s_is, js = s_js, last = n - 1;
        const int ci = cnt[is], cj = cnt[js], sum = ci + cj;
        // ---- fold `is` into `js` ---------------------------------------------------------------------------------------------------------
        if (method != 'c') {
            for (int j = tid; j < n; j += nt) {
                if (j == is || j == js) continue;
                const double a = j < is ? DM(is, j) : DM(j, is);
                double *b = j < js ? &DM(js, j) : &DM(j, js);
                if (method == 'm') *b = a > *b ? a : *b;                                           // clusterlib.c:3757-3762
                else *b = __ddiv_rn(__dadd_rn(__dmul_rn(a, (double)ci), __dmul_rn(*b, (double)cj)), (double)sum);      // :3846-3860
            }
        }
        __syncthreads();
        if (tid == 0) {
            nd[node] = s_best;
            const int a = cid[is], b = cid[js];
            if (a >= 0) lpar[a] = node; else npar[-a - 1] = node;
            if (b >= 0) lpar[b] = node; else npar[-b - 1] = node;
            cid[js] = -node - 1; cid[is] = cid[last];
            cnt[js] = sum; cnt[is] = cnt[last];
            if (method == 'c') {                                                                    // clusterlib.c:3463-3471
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    double x = __dadd_rn(__dmul_rn(pos[3 * js + c], (double)cj), __dmul_rn(pos[3 * is + c], (double)ci));
                    pos[3 * js + c] = __ddiv_rn(x, (double)sum);
                }
#pragma unroll
                for (int c = 0; c < 3; c++) pos[3 * is + c] = pos[3 * last + c];
            }
        }
        // ---- the last row moves into slot `is` -----------------------------------------------------------------------------------------
        for (int j = tid; j < last; j += nt) {
            if (j == is) continue;
            const double t = DM(last, j);
            if (j < is) DM(is, j) = t; else DM(j, is) = t;
        }
        __syncthreads();
        const int m = last;                                  // rows and columns 0 .. m-1 remain
        if (method == 'c') {                                 // new centre: its distances are computed afresh (clusterlib.c:3481-3484)
            const double a[3] = {pos[3 * js], pos[3 * js + 1], pos[3 * js + 2]};
            for (int j = tid; j < m; j += nt) {
                if (j == js) continue;
                const double t = metric_rows(cityblock, a, pos + 3 * j);
                if (j < js) DM(js, j) = t; else DM(j, js) = t;
            }
            __syncthreads();
        }
        // ---- row minima: rows js and is are new; another row is scanned again only if its remembered column was js or is ------------------
        for (int i = 1 + tid; i < m; i += nt) {
            bool rescan = (i == js) || (i == is);
            if (!rescan && i > js) {
                const int ra = rarg[i];
                if (ra == js || ra == is) rescan = true;
                else {
                    double mv = rmin[i];
                    int mj = ra;
                    const double t = DM(i, js);
                    if (t < mv || (t == mv && js < mj)) { mv = t; mj = js; }
                    if (i > is) { const double u = DM(i, is); if (u < mv || (u == mv && is < mj)) { mv = u; mj = is; } }
                    rmin[i] = mv; rarg[i] = mj;
                }
            }
            if (rescan) list[atomicAdd(&s_nlist, 1)] = i;
        }
        __syncthreads();
        const int nl = s_nlist;
        for (int q = wid; q < nl; q += nw) {
            const int i = list[q];
            double mv; int mj;
            link_scan_row(&DM(i, 0), i, lane, mv, mj);
            if (lane == 0) { rmin[i] = mv; rarg[i] = mj; }
        }
        __syncthreads();
    }
#undef DM
    // ---- cuttree_distance: every leaf climbs to its highest ancestor at or below the cut --------------------------------------------------
    for (int i = tid; i < ns; i += nt) {
        int own = -1;
        for (int a = lpar[i]; a >= 0; a = npar[a]) if (!(nd[a] > cut)) own = a;
        list[i] = own;
        if (own >= 0) atomicMin(&omin[own], i);
    }
    __syncthreads();
    for (int i = tid; i < ns; i += nt) par[i] = list[i] >= 0 ? omin[list[i]] : i;
}

}  // namespace fpk
