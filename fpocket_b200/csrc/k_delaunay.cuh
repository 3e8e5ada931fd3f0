// K1: one thread block per structure (persistent blocks pulling structures from a queue):
//   sites -> exact Delaunay triangulation (delaunay.cuh) -> fused alpha-sphere filter + typing -> canonical order.
// The tetrahedra never leave the block's scratch (L2 resident); HBM sees atoms in and kept spheres out.
//
// Replaces reference src/voronoi.c:69-185 load_vvertices (text to qhull and back), :369-521 fill_vvertices,
// :969-1087 testVvertice and :892-916 set_barycenter.
#pragma once
#include "common.cuh"
#include "delaunay_warp.cuh"

#ifndef FPK_K1_PREWALK
#define FPK_K1_PREWALK 1          // use the wait for the slowest claimant of a sub-round to locate the warp's next point (delaunay_warp.cuh w2_prewalk)
#endif

namespace fpk {

struct DelScratch {
    double *P;                    // [3*nmax]
    unsigned long long *keys;     // [npad]
    int *order;                   // [nmax]
    int *tet, *claim, *mark, *free0, *free1;     // [8*cap], [2*cap], [2*cap], [cap], [cap]
    int cap;
    int *hint;                    // [32^3]
    int *cav, *newid, *tinfo, *cursor;
    int *bigcav, *bignewid;
    int *cnt;                     // [C_COUNT]
    int *rep;                     // [nmax]
    int *Pi;                      // [3*nmax] lattice offsets when they do not fit in shared memory
    int *rstart;                  // [32]
    int4 *rawkey;                 // [scap]
    float4 *rawxyzr;              // [scap]
    int *sitecnt;                 // [nmax + 1]
    int *bucket;                  // [scap]
    int scap;
};

// qhull geom2.c:2012-2093 qh_voronoi_center (Cramer relative to the first point, f64), same operation order as the oracle
__device__ __forceinline__ double det3d(const double *r0, const double *r1, const double *r2)
{
    return r0[0] * (r1[1] * r2[2] - r1[2] * r2[1]) - r1[0] * (r0[1] * r2[2] - r0[2] * r2[1]) +
           r2[0] * (r0[1] * r1[2] - r0[2] * r1[1]);
}
__device__ __forceinline__ void circumcentre(const double *p0, const double *p1, const double *p2, const double *p3, double c[3])
{
    double m[3][3], s2[3];
    const double *p[3] = {p1, p2, p3};
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int j = 0; j < 3; j++) m[k][j] = p[j][k] - p0[k];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) s += m[k][j] * m[k][j];
        s2[j] = s;
    }
    double det = det3d(m[0], m[1], m[2]);
    double factor = 0.5 / det;
    c[0] = det3d(s2, m[1], m[2]) * factor + p0[0];
    c[1] = det3d(m[0], s2, m[2]) * factor + p0[1];
    c[2] = det3d(m[0], m[1], s2) * factor + p0[2];
}

// voronoi.c:969-1087 testVvertice; ac[4] / ap[4] = indices of the 4 atoms into the packed coordinate / property arrays,
// in the tet's vertex order
__device__ float test_vvertice(const BatchDev &B, const StructDev &S, const float xyz[3], const int ac[4], const int ap[4])
{
    const fpk_params &p = B.prm;
    float min_asph_size = p.asph_min_size, max_asph_size = p.asph_max_size;
    float x = xyz[0] - 0.0f, y = xyz[1] - 0.0f, z = xyz[2] - 0.0f;
    float baryx = 0.0f, baryy = 0.0f, baryz = 0.0f, barybf = 0.0f, sdbf = 0.0f;
    if (p.xpocket_active) {
        int ok = 0;
        for (int k = 0; k < 4; k++)
            if (B.afl[ap[k]] & FPK_ATOM_XPOCKET) ok++;
        if (ok < p.min_n_explicit_pocket_atoms) return -3.0f;
    }
    int c = ac[0];
    baryx += B.ax[c]; baryy += B.ay[c]; baryz += B.az[c]; barybf += B.abf[ap[0]];
    float d1 = f_dist(x, y, z, B.ax[c], B.ay[c], B.az[c]), d2, d3, d4;
    if ((double)min_asph_size <= (double)d1 + 1e-3 && (double)d1 - 1e-3 <= (double)max_asph_size) {
        c = ac[1]; d2 = f_dist(x, y, z, B.ax[c], B.ay[c], B.az[c]);
        baryx += B.ax[c]; baryy += B.ay[c]; baryz += B.az[c]; barybf += B.abf[ap[1]];
        c = ac[2]; d3 = f_dist(x, y, z, B.ax[c], B.ay[c], B.az[c]);
        baryx += B.ax[c]; baryy += B.ay[c]; baryz += B.az[c]; barybf += B.abf[ap[2]];
        c = ac[3]; d4 = f_dist(x, y, z, B.ax[c], B.ay[c], B.az[c]);
        baryx += B.ax[c]; baryy += B.ay[c]; baryz += B.az[c]; barybf += B.abf[ap[3]];
        baryx = (float)((double)baryx * 0.25); baryy = (float)((double)baryy * 0.25);
        baryz = (float)((double)baryz * 0.25); barybf = (float)((double)barybf * 0.25);
        if (fabs((double)(d1 - d2)) < 1e-3 && fabs((double)(d1 - d3)) < 1e-3 && fabs((double)(d1 - d4)) < 1e-3) {
            if (S.n_xlig) {
                const float *xl = B.xlig + 3 * (size_t)S.xlig_off;
                for (int i = 0; i < S.n_xlig; i++)
                    if (f_dist(xl[3 * i], xl[3 * i + 1], xl[3 * i + 2], x, y, z) <= d1) return d1;
                return -1.0f;
            }
            for (int k = 0; k < 4; k++) {
                float b = B.abf[ap[k]];
                sdbf += (b - barybf) * (b - barybf);
            }
            sdbf = (float)sqrt((double)(sdbf / 3));
            if (((sdbf > S.avg_b) || (sdbf > ((S.max_b - S.min_b) / 4))) &&
                (((double)S.avg_b > 0.0) && ((double)(barybf / S.avg_b) > 1.4)))
                return -1.0f;
            if ((double)f_dist(baryx, baryy, baryz, xyz[0], xyz[1], xyz[2]) > 1.0 && (double)d1 > ((double)max_asph_size - 1.5))
                return -2.0f;
            return d1;
        }
    }
    return -3.0f;
}

__device__ __forceinline__ bool key_less(const int4 &a, const int4 &b)
{
    if (a.x != b.x) return a.x < b.x;
    if (a.y != b.y) return a.y < b.y;
    if (a.z != b.z) return a.z < b.z;
    return a.w < b.w;
}


// bounding box of X.P[0..3n) by the whole block -> s_box[0..2]=min, [3..5]=max (s_box, s_bmin, s_bmax in shared memory)
__device__ void block_bbox(const double *P, int n, double *s_box, double (*s_bmin)[3], double (*s_bmax)[3])
{
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int i = tid; i < n; i += nt)
        for (int c = 0; c < 3; c++) { double v = P[3 * (size_t)i + c]; mn[c] = fmin(mn[c], v); mx[c] = fmax(mx[c], v); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            mn[c] = fmin(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], d));
            mx[c] = fmax(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], d));
        }
    if (lane == 0) for (int c = 0; c < 3; c++) { s_bmin[wid][c] = mn[c]; s_bmax[wid][c] = mx[c]; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < ((nt + 31) >> 5); w++)
            for (int c = 0; c < 3; c++) { mn[c] = fmin(mn[c], s_bmin[w][c]); mx[c] = fmax(mx[c], s_bmax[w][c]); }
        for (int c = 0; c < 3; c++) { s_box[c] = mn[c]; s_box[3 + c] = mx[c]; }
    }
    __syncthreads();
}

// the serial path for one over-sized cavity (warp w alone, the rest of the block waits): global lists, delaunay.cuh
__device__ __forceinline__ void w_big_insert(DelaunayCtx &c, int w)
{
    const int lane = threadIdx.x & 31;
    const int popstack = c.cnt[C_SUBROUND] & 1;
    int *ti = c.tinfo + 4 * w;
    int ncav = 0;
    int st = w_locate_and_claim(c, ti[0], ti[3], c.bigcav, c.cap, &ncav);
    __syncwarp();
    w_reset_claims(c, c.bigcav, ncav);
    __syncwarp();
    int nt = -1;
    if (st == ST_CLAIMED) nt = w_commit_insertion(c, ti[0], c.bigcav, c.bignewid, ncav, popstack);
    __syncwarp();
    if (lane == 0) {
        if (st == ST_CLAIMED) ti[3] = nt;
        else if (st != ST_DUP) c.cnt[C_STATUS] = DERR_CAPACITY;
        c.cnt[C_NBIG]++;
        ti[0] = -1;
        blk().resume[w] = -1;
    }
    __syncwarp();
}

// Delaunay triangulation of the n lattice points in X.P by the whole block (blockDim.x / 32 <= FPK_MAXWARPS warps).
// S = static shared state, Wall = one WarpCav per warp (shared), Pi_sm / pi_cap = shared buffer for the sites.
// Returns 0 or a DERR_* code (uniform over the block). On return the mesh is in X.tet, S.cnt[C_NTET] slots.
__device__ int block_delaunay(int pi_cap, const DelScratch &X, int n, const double *s_box)
{
    BlockDel &S = blk();
    DelaunayCtx &C = S.C;
    int *Pi_sm = sm_sites();
    const int tid = threadIdx.x, nt = blockDim.x;
    const double bmn[3] = {s_box[0], s_box[1], s_box[2]};
    const double ext = fmax(1.0, fmax(s_box[3] - s_box[0], fmax(s_box[4] - s_box[1], s_box[5] - s_box[2])));
    const int R = brio_nrounds(n);
    int npad = 1;
    while (npad < n) npad <<= 1;
    int *Pi = n <= pi_cap ? Pi_sm : X.Pi;
    for (int i = tid; i < npad; i += nt)
        X.keys[i] = i < n ? brio_key(i, X.P + 3 * (size_t)i, bmn, 1023.999 / ext, R) : ~0ull;
    for (int i = tid; i < 3 * n; i += nt) Pi[i] = (int)(X.P[i] - bmn[i % 3]);        // exact: integers, extent checked below
    __syncthreads();
    block_bitonic_sort(X.keys, npad);
    if (tid <= R) X.rstart[tid] = n;
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        unsigned long long key = X.keys[i];
        X.order[i] = (int)(key & 0xffffffull);
        int r = (int)(key >> 54);
        if (i == 0 || (int)(X.keys[i - 1] >> 54) != r) X.rstart[r] = i;
    }
    int hgc = 2;
    while (hgc < 32 && hgc * hgc * hgc * 2 < n) hgc++;
    for (int i = tid; i < hgc * hgc * hgc; i += nt) X.hint[i] = -1;
    for (int i = tid; i < n; i += nt) X.rep[i] = i;
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    if (lane == 0) { S.tinfo[4 * wid] = -1; S.tinfo[4 * wid + 3] = -1; S.resume[wid] = -1; S.prew_rank[wid] = -1; S.prew_tet[wid] = -1; }
    __syncthreads();
    if (tid == 0) {
        S.ndone = 0;
        for (int r = R - 1; r >= 0; r--) if (X.rstart[r] > X.rstart[r + 1]) X.rstart[r] = X.rstart[r + 1];
        X.rstart[0] = 0;
        C.n = n; C.P = X.P; C.Pi = Pi; C.tet = X.tet; C.claim = X.claim; C.mark = X.mark;
        C.cap = X.cap; C.cnt = S.cnt; C.freest[0] = X.free0; C.freest[1] = X.free1; C.order = X.order;
        C.hint = X.hint;
        C.hg = hgc; C.hinv = hgc / (ext * 1.0001);
        C.hmin[0] = bmn[0]; C.hmin[1] = bmn[1]; C.hmin[2] = bmn[2];
        C.nworkers = nw; C.cav = X.cav; C.newid = X.newid; C.tinfo = S.tinfo; C.cursor = S.cursor;
        C.bigcav = X.bigcav; C.bignewid = X.bignewid; C.rep = X.rep;
        delaunay_init(C, S.first4);
        if (ext > 2.0e9) S.cnt[C_STATUS] = DERR_CAPACITY;          // lattice offsets must fit in 31 bits (> 2000 A across: not a molecule)
        S.bigw = 1 << 30;
    }
    __syncthreads();
    // rounds of concurrent insertions: one pending point per warp, four block barriers per sub-round
    const bool sm_pts = n <= pi_cap;
    int *ti = S.tinfo + 4 * wid;
    bool failed = S.cnt[C_STATUS] != 0;
    for (int r = 0; r < R && !failed; r++) {
        if (lane == 0) round_chunk(X.rstart[r], X.rstart[r + 1], wid, nw, &S.cursor[2 * wid], &S.cursor[2 * wid + 1]);
        __syncwarp();
        for (;;) {
            int work = w_phase_assign(C, wid) ? 1 : 0;
            if (S.cnt[C_STATUS]) work = 0;
            const int nactive = __syncthreads_count(work) >> 5;            // claimants of this sub-round (all 32 lanes of a warp vote alike)
            if (!nactive) break;
            // ---- claim: walk + cavity flood on a read-only mesh ----------------------------------------------------------
            const int rank = work ? ti[0] : -1;
            int st = ST_IDLE, ncav = 0;
            if (rank >= 0) {
                st = sm_pts ? w2_locate_and_claim<true>(rank, ti[3], &ncav) : w2_locate_and_claim<false>(rank, ti[3], &ncav);
                __syncwarp();
                if (lane == 0) atomicAdd(&S.ndone, 1);
                // while the slower claimants finish: locate the warp's next point (only a likely winner moves on to it)
                const int nxt = S.cursor[2 * wid];
                if (FPK_K1_PREWALK && st == ST_CLAIMED && nxt < S.cursor[2 * wid + 1]) {
                    if (sm_pts) w2_prewalk<true>(nxt, my_cav().id[0], nactive); else w2_prewalk<false>(nxt, my_cav().id[0], nactive);
                }
            }
            __syncthreads();
            if (tid == 0) S.ndone = 0;
            // ---- verify: a claimant wins iff nothing stronger touched its cavity or rim ---------------------------------
            if (st == ST_CLAIMED) st = w2_verify_claims(rank, ncav) ? ST_WIN : ST_LOST;
            __syncthreads();
            // ---- release the claims; winners rewrite the mesh concurrently --------------------------------------------
            if (rank >= 0) {
                if (st == ST_WIN || st == ST_LOST || st == ST_BIG) w2_reset_claims(ncav);
                int newtet = -1;
                if (st == ST_WIN) newtet = w2_commit_insertion(rank, ncav, S.cnt[C_SUBROUND] & 1);
                __syncwarp();
                if (lane == 0) {
                    ti[1] = st; ti[2] = ncav;
                    if (st != ST_LOST && st != ST_BIG) S.resume[wid] = -1;
                    after_attempt(C, wid, newtet);
                    if (st == ST_BIG) atomicMin(&S.bigw, wid);
                }
            }
            __syncthreads();
            if (S.bigw < nw) {                     // rare: at most one over-sized cavity per sub-round goes through the serial path
                if (wid == S.bigw) w_big_insert(C, wid);
                __syncthreads();
                if (tid == 0) S.bigw = 1 << 30;
            }
            if (tid == 0) end_subround(C);
        }
        failed = S.cnt[C_STATUS] != 0;
    }
    __syncthreads();
    return S.cnt[C_STATUS];
}

// ---- TMA (bulk async copy) staging of a structure's coordinate tile into shared memory ------------------------------------------------
// The x / y / z arrays of one structure are contiguous in the batch arena; when every atom is a site (no hydrogens) the tile is brought
// into the shared site buffer by three cp.async.bulk copies completing on an mbarrier (SASS: UBLKCP), instead of per-thread gathers.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// bounded wait (never spins forever: a failed copy is reported, not hung on)
__device__ __forceinline__ bool mbar_wait(unsigned long long *bar, unsigned phase)
{
    const unsigned a = smem_u32(bar);
    for (int it = 0; it < (1 << 22); it++) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(a), "r"(phase) : "memory");
        if (ok) return true;
    }
    return false;
}

#ifndef FPK_K1_MINBLOCKS
#define FPK_K1_MINBLOCKS 3        // 3 x 8 warps at 80 registers: more resident warps beat more registers (profiles/README.md)
#endif
#ifndef FPK_K1_THREADS
#define FPK_K1_THREADS 256
#endif
__global__ void __launch_bounds__(FPK_K1_THREADS, FPK_K1_MINBLOCKS) k_delaunay_filter(BatchDev B, const DelScratch *scratch, int pi_cap)
{
    __shared__ int s_k, s_red[40];
    __shared__ double s_box[6];
    __shared__ double s_bmin[FPK_MAXWARPS][3], s_bmax[FPK_MAXWARPS][3];
    __shared__ __align__(8) unsigned long long s_mbar;
    BlockDel &BD = blk();
    DelaunayCtx &C = BD.C;
    const int tid = threadIdx.x, nt = blockDim.x;
    const DelScratch &X = scratch[blockIdx.x];
    unsigned tma_phase = 0;
    if (tid == 0) mbar_init(&s_mbar, 1);
    __syncthreads();

    for (;;) {
        __syncthreads();
        if (tid == 0) s_k = atomicAdd(B.work_counter, 1);
        __syncthreads();
        if (s_k >= B.nstruct) return;
        const int k = B.work_order[s_k];
        const StructDev S = B.S[k];
        const int n = S.nsites;
        int *counts = B.counts + (size_t)k * CNT_STRIDE;
        if (n > B.k1_max_sites) continue;          // done by the whole GPU (k_delaunay_grid)
        if (n < 4) {
            if (tid == 0) { counts[CNT_TETS] = 0; counts[CNT_SPHERES] = 0; counts[CNT_STATUS] = FPK_ERR_DEGENERATE; }
            continue;
        }
        // ---- 1. lattice coordinates: what "%.6f" prints (voronoi.c:130), exact in double -----------------------
        const int lead = S.atom_off & 3, tile = (n + lead + 3) & ~3;             // 16-byte aligned window around the structure's atoms
        bool staged = false;
        if (pi_cap > 0 && n == S.natoms && 3 * tile <= 3 * pi_cap + 24) {       // every atom is a site: the tile is contiguous (pi_cap 0: no site buffer at all)
            float *stg = reinterpret_cast<float *>(sm_sites());
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // earlier generic-proxy accesses to the buffer come first
                mbar_expect_tx(&s_mbar, 3u * (unsigned)tile * 4u);
                tma_load_1d(stg, B.ax + S.atom_off - lead, (unsigned)tile * 4u, &s_mbar);
                tma_load_1d(stg + tile, B.ay + S.atom_off - lead, (unsigned)tile * 4u, &s_mbar);
                tma_load_1d(stg + 2 * tile, B.az + S.atom_off - lead, (unsigned)tile * 4u, &s_mbar);
            }
            staged = mbar_wait(&s_mbar, tma_phase);
            tma_phase ^= 1u;
            if (staged)
                for (int i = tid; i < n; i += nt) {
                    X.P[3 * (size_t)i] = (double)__double2ll_rn(__dmul_rn((double)stg[lead + i], 1e6));
                    X.P[3 * (size_t)i + 1] = (double)__double2ll_rn(__dmul_rn((double)stg[tile + lead + i], 1e6));
                    X.P[3 * (size_t)i + 2] = (double)__double2ll_rn(__dmul_rn((double)stg[2 * tile + lead + i], 1e6));
                }
        }
        if (!__syncthreads_and(staged ? 1 : 0)) {                                // hydrogens present (sites are a subset) or a tile too large
            for (int i = tid; i < n; i += nt) {
                int a = S.atom_off + B.site2atom[S.site_off + i];
                X.P[3 * (size_t)i] = (double)__double2ll_rn(__dmul_rn((double)B.ax[a], 1e6));
                X.P[3 * (size_t)i + 1] = (double)__double2ll_rn(__dmul_rn((double)B.ay[a], 1e6));
                X.P[3 * (size_t)i + 2] = (double)__double2ll_rn(__dmul_rn((double)B.az[a], 1e6));
            }
            __syncthreads();
        }
        block_bbox(X.P, n, s_box, s_bmin, s_bmax);
        // ---- 2.+3. insertion order, rounds of concurrent insertions --------------------------------------------------
        const int err = block_delaunay(pi_cap, X, n, s_box);
        if (err) {
            if (tid == 0) {
                counts[CNT_TETS] = 0; counts[CNT_SPHERES] = 0;
                counts[CNT_STATUS] = err == DERR_DEGENERATE ? FPK_ERR_DEGENERATE : FPK_ERR_CAPACITY;
                counts[CNT_SUBROUNDS] = C.cnt[C_NSUB]; counts[CNT_RETRIES] = C.cnt[C_NRETRY]; counts[CNT_NBIG] = 1000 * err + C.cnt[C_NBIG];
            }
            continue;
        }
        // ---- 4. fused filter over the finite tets (circumcentre f64 -> f32, testVvertice) ----------------------------
        const int ntet = C.cnt[C_NTET];
        if (tid == 0) { s_red[0] = 0; s_red[1] = 0; }     // kept spheres, finite tets
        for (int i = tid; i <= n; i += nt) X.sitecnt[i] = 0;
        __syncthreads();
        for (int t = tid; t < ntet; t += nt) {
            int4 q = reinterpret_cast<const int4 *>(X.tet)[2 * (size_t)t];
            if (q.x < 0 || q.y < 0 || q.z < 0 || q.w < 0) continue;
            int v[4] = {X.rep[q.x], X.rep[q.y], X.rep[q.z], X.rep[q.w]};       // coincident sites: lowest index
#pragma unroll
            for (int i = 1; i < 4; i++) { int kk = v[i], j = i - 1; while (j >= 0 && v[j] > kk) { v[j + 1] = v[j]; j--; } v[j + 1] = kk; }
            int ord = atomicAdd(&s_red[1], 1);
            if (B.tets_out && ord < S.tet_cap)
                reinterpret_cast<int4 *>(B.tets_out)[S.tet_off + ord] = make_int4(v[0], v[1], v[2], v[3]);
            double p[4][3];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int c = 0; c < 3; c++) p[i][c] = __ddiv_rn(X.P[3 * (size_t)v[i] + c], 1e6);   // strtod of the "%.6f" text
            double cc[3];
            circumcentre(p[0], p[1], p[2], p[3], cc);
            float xyz[3] = {(float)cc[0], (float)cc[1], (float)cc[2]};
            int ac[4], ap[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { int la = B.site2atom[S.site_off + v[i]]; ac[i] = S.atom_off + la; ap[i] = S.prop_off + la; }
            float rad = test_vvertice(B, S, xyz, ac, ap);
            if (rad > 0) {
                int pos = atomicAdd(&s_red[0], 1);
                if (pos < X.scap && pos < S.sph_cap) {
                    X.rawkey[pos] = make_int4(v[0], v[1], v[2], v[3]);
                    X.rawxyzr[pos] = make_float4(xyz[0], xyz[1], xyz[2], rad);
                    atomicAdd(&X.sitecnt[v[0]], 1);
                }
            }
        }
        __syncthreads();
        const int ns = s_red[0];
        if (tid == 0) {
            counts[CNT_TETS] = s_red[1];
            counts[CNT_SUBROUNDS] = C.cnt[C_NSUB]; counts[CNT_RETRIES] = C.cnt[C_NRETRY]; counts[CNT_NBIG] = C.cnt[C_NBIG];
            counts[CNT_WALK] = C.cnt[C_NWALK]; counts[CNT_LEVELS] = C.cnt[C_NLEVEL]; counts[CNT_ATTEMPTS] = C.cnt[C_NATTEMPT]; counts[CNT_HOPS] = C.cnt[C_NHOP];
        }
        if (ns > S.sph_cap || ns > X.scap) {
            if (tid == 0) { counts[CNT_SPHERES] = ns; counts[CNT_STATUS] = FPK_ERR_CAPACITY; }
            continue;
        }
        // ---- 5. canonical order: ascending (a,b,c,d) of the sorted site tuple. Bucket by a, rank inside the bucket -----
        block_exclusive_scan(X.sitecnt, n + 1, s_red + 2);
        // sitecnt[a] = start of bucket a; reuse order[] (no longer needed) as the per-bucket fill cursor
        for (int i = tid; i < n; i += nt) X.order[i] = 0;
        __syncthreads();
        for (int i = tid; i < ns; i += nt) {
            int a0 = X.rawkey[i].x;
            int slot = atomicAdd(&X.order[a0], 1);
            X.bucket[X.sitecnt[a0] + slot] = i;
        }
        __syncthreads();
        for (int i = tid; i < ns; i += nt) {
            int4 key = X.rawkey[i];
            int b0 = X.sitecnt[key.x], b1 = X.sitecnt[key.x + 1], rank = 0;
            for (int j = b0; j < b1; j++) {
                int o = X.bucket[j];
                if (o != i && key_less(X.rawkey[o], key)) rank++;
            }
            const int dst = S.sph_off + b0 + rank;
            float4 xr = X.rawxyzr[i];
            const int a[4] = {B.site2atom[S.site_off + key.x], B.site2atom[S.site_off + key.y],
                              B.site2atom[S.site_off + key.z], B.site2atom[S.site_off + key.w]};
            // typing (voronoi.c:467-495) and barycentre of the 4 contacted atoms (voronoi.c:892-916)
            int apol = 0;
            float xs = 0.0f, ys = 0.0f, zs = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if ((double)B.aen[S.prop_off + a[j]] < 2.8) apol++;
                xs += B.ax[S.atom_off + a[j]]; ys += B.ay[S.atom_off + a[j]]; zs += B.az[S.atom_off + a[j]];
            }
            B.s_xyzr[dst] = xr;
            B.s_key[dst] = key;
            B.s_neigh[dst] = make_int4(a[0], a[1], a[2], a[3]);
            B.s_type[dst] = (apol >= B.prm.min_apol_neigh) ? 0 : 1;
            B.s_bary[3 * (size_t)dst] = (float)((double)xs * 0.25);
            B.s_bary[3 * (size_t)dst + 1] = (float)((double)ys * 0.25);
            B.s_bary[3 * (size_t)dst + 2] = (float)((double)zs * 0.25);
        }
        if (tid == 0) { counts[CNT_SPHERES] = ns; counts[CNT_STATUS] = FPK_OK; }
    }
}

}  // namespace fpk
