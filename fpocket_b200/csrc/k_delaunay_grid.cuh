// K1g: ONE structure on the WHOLE GPU - the Delaunay triangulation of a large complex (BASELINE configs[3]: ~150,000 heavy atoms) by
// every resident block of a cooperative launch, then the same fused filter as K1.
//
// Replaces, like K1, reference src/voronoi.c:69-185 load_vvertices (qhull through text files), :369-521 fill_vvertices, :969-1087
// testVvertice and :892-916 set_barycenter; the reference's cap on kept spheres (headers/voronoi.h:41, 1e6) is far away at this size.
//
// The claim / verify / commit protocol of delaunay.cuh never needed the claimants to sit in one block: claims are atomicMin on global
// words, winners' cavities neither overlap nor touch. What a block shared was (i) its barrier between the phases and (ii) the counters
// of the triangulation. Here (i) is the grid barrier of the cooperative launch (four per sub-round) and (ii) lives in global memory
// (GridDel); per-warp state and the cavity records stay in each block's shared memory. With 3 blocks of 8 warps on each of the 148 SMs,
// 3,552 points are in flight per sub-round instead of 8. Insertion order, predicates and symbolic perturbation are K1's, so the
// tetrahedra are the same set (the triangulation is unique): tests/test_gpu_parity.py checks it against the oracle at 60,000 sites.
#pragma once
#include <cooperative_groups.h>
#include "k_delaunay.cuh"

namespace fpk {
namespace cg = cooperative_groups;

struct GridDel {                 // global state of one grid-wide triangulation (zeroed by the kernel itself)
    int cnt[C_COUNT];            // the counters block_delaunay keeps in shared memory
    int first4[4];
    int work[4];                 // rotating flags: "some warp still has a point" of sub-round it & 3
    int bigw[4];                 // rotating: lowest global warp whose cavity outgrew its shared-memory record
    long long bbmin[3], bbmax[3];
    int kept, finite;            // fused filter: kept spheres, finite tetrahedra
    int pad[2];
};

__device__ __forceinline__ int ldcg_i(const int *p) { return __ldcg(p); }

__global__ void __launch_bounds__(FPK_K1_THREADS, FPK_K1_MINBLOCKS) k_delaunay_grid(BatchDev B, DelScratch X, GridDel *G, int k)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ int s_red[40];
    BlockDel &S = blk();
    DelaunayCtx &C = S.C;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    const int gtid = blockIdx.x * nt + tid, gnt = gridDim.x * nt, gw = blockIdx.x * nw + wid, ngw = gridDim.x * nw;
    const StructDev Sd = B.S[k];
    const int n = Sd.nsites;
    int *counts = B.counts + (size_t)k * CNT_STRIDE;

    // ---- 0. state ----------------------------------------------------------------------------------------------------------------
    if (blockIdx.x == 0) {
        for (int i = tid; i < (int)(sizeof(GridDel) / sizeof(int)); i += nt) reinterpret_cast<int *>(G)[i] = 0;
        __syncthreads();
        if (tid < 3) { G->bbmin[tid] = 0x7fffffffffffffffll; G->bbmax[tid] = -0x7fffffffffffffffll; }
        if (tid < 4) G->bigw[tid] = 1 << 30;
    }
    grid.sync();
    // ---- 1. lattice coordinates: what "%.6f" prints (voronoi.c:130), exact in double; bounding box ------------------------------------
    {
        long long mn[3] = {0x7fffffffffffffffll, 0x7fffffffffffffffll, 0x7fffffffffffffffll}, mx[3] = {-0x7fffffffffffffffll, -0x7fffffffffffffffll, -0x7fffffffffffffffll};
        for (int i = gtid; i < n; i += gnt) {
            const int a = Sd.atom_off + B.site2atom[Sd.site_off + i];
            const long long x = __double2ll_rn(__dmul_rn((double)B.ax[a], 1e6)), y = __double2ll_rn(__dmul_rn((double)B.ay[a], 1e6)),
                            z = __double2ll_rn(__dmul_rn((double)B.az[a], 1e6));
            X.P[3 * (size_t)i] = (double)x; X.P[3 * (size_t)i + 1] = (double)y; X.P[3 * (size_t)i + 2] = (double)z;
            mn[0] = min(mn[0], x); mn[1] = min(mn[1], y); mn[2] = min(mn[2], z);
            mx[0] = max(mx[0], x); mx[1] = max(mx[1], y); mx[2] = max(mx[2], z);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                mn[c] = min(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], d));
                mx[c] = max(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], d));
            }
        if (lane == 0 && mn[0] <= mx[0])
            for (int c = 0; c < 3; c++) { atomicMin(&G->bbmin[c], mn[c]); atomicMax(&G->bbmax[c], mx[c]); }
    }
    grid.sync();
    const double bmn[3] = {(double)G->bbmin[0], (double)G->bbmin[1], (double)G->bbmin[2]};
    const double ext = fmax(1.0, fmax((double)(G->bbmax[0] - G->bbmin[0]), fmax((double)(G->bbmax[1] - G->bbmin[1]), (double)(G->bbmax[2] - G->bbmin[2]))));
    // ---- 2. insertion order: biased randomised rounds, Morton order inside a round (delaunay.cuh), sorted by the whole grid ----------
    const int R = brio_nrounds(n);
    int npad = 1;
    while (npad < n) npad <<= 1;
    for (int i = gtid; i < npad; i += gnt) X.keys[i] = i < n ? brio_key(i, X.P + 3 * (size_t)i, bmn, 1023.999 / ext, R) : ~0ull;
    for (int i = gtid; i < 3 * n; i += gnt) X.Pi[i] = (int)(X.P[i] - bmn[i % 3]);
    int hgc = 2;
    while (hgc < 32 && hgc * hgc * hgc * 2 < n) hgc++;
    for (int i = gtid; i < hgc * hgc * hgc; i += gnt) X.hint[i] = -1;
    for (int i = gtid; i < n; i += gnt) X.rep[i] = i;
    if (gtid <= R) X.rstart[gtid] = n;
    grid.sync();
    for (int kk = 2; kk <= npad; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = gtid; i < npad; i += gnt) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = X.keys[i], b = X.keys[ixj];
                    const bool up = ((i & kk) == 0);
                    if ((a > b) == up) { X.keys[i] = b; X.keys[ixj] = a; }
                }
            }
            grid.sync();
        }
    for (int i = gtid; i < n; i += gnt) {
        const unsigned long long key = X.keys[i];
        X.order[i] = (int)(key & 0xffffffull);
        const int r = (int)(key >> 54);
        if (i == 0 || (int)(X.keys[i - 1] >> 54) != r) X.rstart[r] = i;
    }
    if (lane == 0) { S.tinfo[4 * wid] = -1; S.tinfo[4 * wid + 3] = -1; S.resume[wid] = -1; }
    grid.sync();
    // ---- 3. every block gets its own copy of the context; block 0 builds the first simplex -----------------------------------------------
    if (tid == 0) {
        if (blockIdx.x == 0) {
            for (int r = R - 1; r >= 0; r--) if (X.rstart[r] > X.rstart[r + 1]) X.rstart[r] = X.rstart[r + 1];
            X.rstart[0] = 0;
        }
        C.n = n; C.P = X.P; C.Pi = X.Pi; C.tet = X.tet; C.claim = X.claim; C.mark = X.mark;
        C.cap = X.cap; C.cnt = G->cnt; C.freest[0] = X.free0; C.freest[1] = X.free1; C.order = X.order;
        C.hint = X.hint;
        C.hg = hgc; C.hinv = hgc / (ext * 1.0001);
        C.hmin[0] = bmn[0]; C.hmin[1] = bmn[1]; C.hmin[2] = bmn[2];
        C.nworkers = nw; C.cav = X.cav; C.newid = X.newid; C.tinfo = S.tinfo; C.cursor = S.cursor;
        C.bigcav = X.bigcav; C.bignewid = X.bignewid; C.rep = X.rep;
        if (blockIdx.x == 0) {
            delaunay_init(C, S.first4);
            for (int q = 0; q < 4; q++) G->first4[q] = S.first4[q];
            if (ext > 2.0e9) G->cnt[C_STATUS] = DERR_CAPACITY;      // lattice offsets must fit in 31 bits
        }
    }
    grid.sync();
    if (tid == 0) for (int q = 0; q < 4; q++) { S.first4[q] = G->first4[q]; C.first4[q] = G->first4[q]; }
    __syncthreads();
    // ---- 4. rounds of concurrent insertions: one pending point per warp of the GRID, four grid barriers per sub-round ---------------------
    int *ti = S.tinfo + 4 * wid;
    int it = 0;
    bool failed = ldcg_i(&G->cnt[C_STATUS]) != 0;
    for (int r = 0; r < R && !failed; r++) {
        if (lane == 0) round_chunk(X.rstart[r], X.rstart[r + 1], gw, ngw, &S.cursor[2 * wid], &S.cursor[2 * wid + 1]);
        __syncwarp();
        for (;;) {
            int work = w_phase_assign(C, wid) ? 1 : 0;
            if (ldcg_i(&G->cnt[C_STATUS])) work = 0;
            // Flags rotate over four slots: sub-round `it` raises slot it & 3 before the barrier and reads it after; thread 0 clears slot
            // (it + 2) & 3 meanwhile - last read two sub-rounds ago, next raised only after thread 0 has passed the next barrier.
            const int slot = it & 3;
            if (__syncthreads_or(work) && tid == 0) atomicOr(&G->work[slot], 1);
            grid.sync();
            if (gtid == 0) { G->work[(it + 2) & 3] = 0; G->bigw[(it + 2) & 3] = 1 << 30; }
            if (!ldcg_i(&G->work[slot])) { it++; break; }          // the next round starts on a fresh slot
            // ---- claim: walk + cavity flood on a read-only mesh ----------------------------------------------------------------------
            const int rank = work ? ti[0] : -1;
            int st = ST_IDLE, ncav = 0;
            if (rank >= 0) st = w2_locate_and_claim<false, true>(rank, ti[3], &ncav);
            grid.sync();
            // ---- verify: a claimant wins iff nothing stronger touched its cavity or rim ---------------------------------------------
            if (st == ST_CLAIMED) st = w2_verify_claims(rank, ncav) ? ST_WIN : ST_LOST;
            grid.sync();
            // ---- release the claims; winners rewrite the mesh concurrently ------------------------------------------------------------
            if (rank >= 0) {
                if (st == ST_WIN || st == ST_LOST || st == ST_BIG) w2_reset_claims(ncav);
                int newtet = -1;
                if (st == ST_WIN) newtet = w2_commit_insertion<true>(rank, ncav, ldcg_i(&G->cnt[C_SUBROUND]) & 1);
                __syncwarp();
                if (lane == 0) {
                    ti[1] = st; ti[2] = ncav;
                    if (st != ST_LOST && st != ST_BIG) S.resume[wid] = -1;
                    after_attempt(C, wid, newtet);
                    if (st == ST_BIG) atomicMin(&G->bigw[slot], gw);
                }
            }
            grid.sync();
            const int bw = ldcg_i(&G->bigw[slot]);
            if (bw < ngw) {                        // rare: at most one over-sized cavity per sub-round goes through the serial path
                if (gw == bw) w_big_insert(C, wid);
                grid.sync();
            }
            if (gtid == 0) end_subround(C);
            it++;
        }
        failed = ldcg_i(&G->cnt[C_STATUS]) != 0;
    }
    grid.sync();
    const int err = ldcg_i(&G->cnt[C_STATUS]);
    if (err) {
        if (gtid == 0) {
            counts[CNT_TETS] = 0; counts[CNT_SPHERES] = 0;
            counts[CNT_STATUS] = err == DERR_DEGENERATE ? FPK_ERR_DEGENERATE : FPK_ERR_CAPACITY;
            counts[CNT_SUBROUNDS] = G->cnt[C_NSUB]; counts[CNT_RETRIES] = G->cnt[C_NRETRY]; counts[CNT_NBIG] = 1000 * err + G->cnt[C_NBIG];
        }
        return;
    }
    // ---- 5. fused filter over the finite tets (circumcentre f64 -> f32, testVvertice), the whole grid ----------------------------------------
    const int ntet = ldcg_i(&G->cnt[C_NTET]);
    for (int i = gtid; i <= n; i += gnt) X.sitecnt[i] = 0;
    for (int i = gtid; i < n; i += gnt) X.order[i] = 0;          // the insertion order is no longer needed: per-bucket fill cursor of step 6
    grid.sync();
    for (int t = gtid; t < ntet; t += gnt) {
        int4 q = reinterpret_cast<const int4 *>(X.tet)[2 * (size_t)t];
        if (q.x < 0 || q.y < 0 || q.z < 0 || q.w < 0) continue;
        int v[4] = {X.rep[q.x], X.rep[q.y], X.rep[q.z], X.rep[q.w]};       // coincident sites: lowest index
#pragma unroll
        for (int i = 1; i < 4; i++) { int kk = v[i], j = i - 1; while (j >= 0 && v[j] > kk) { v[j + 1] = v[j]; j--; } v[j + 1] = kk; }
        const int ord = atomicAdd(&G->finite, 1);
        if (B.tets_out && ord < Sd.tet_cap)
            reinterpret_cast<int4 *>(B.tets_out)[Sd.tet_off + ord] = make_int4(v[0], v[1], v[2], v[3]);
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
        for (int i = 0; i < 4; i++) { int la = B.site2atom[Sd.site_off + v[i]]; ac[i] = Sd.atom_off + la; ap[i] = Sd.prop_off + la; }
        const float rad = test_vvertice(B, Sd, xyz, ac, ap);
        if (rad > 0) {
            const int pos = atomicAdd(&G->kept, 1);
            if (pos < X.scap && pos < Sd.sph_cap) {
                X.rawkey[pos] = make_int4(v[0], v[1], v[2], v[3]);
                X.rawxyzr[pos] = make_float4(xyz[0], xyz[1], xyz[2], rad);
                atomicAdd(&X.sitecnt[v[0]], 1);
            }
        }
    }
    grid.sync();
    const int ns = ldcg_i(&G->kept);
    if (gtid == 0) {
        counts[CNT_TETS] = G->finite;
        counts[CNT_SUBROUNDS] = G->cnt[C_NSUB]; counts[CNT_RETRIES] = G->cnt[C_NRETRY]; counts[CNT_NBIG] = G->cnt[C_NBIG];
        counts[CNT_WALK] = G->cnt[C_NWALK]; counts[CNT_LEVELS] = G->cnt[C_NLEVEL]; counts[CNT_ATTEMPTS] = G->cnt[C_NATTEMPT]; counts[CNT_HOPS] = G->cnt[C_NHOP];
    }
    if (ns > Sd.sph_cap || ns > X.scap) {
        if (gtid == 0) { counts[CNT_SPHERES] = ns; counts[CNT_STATUS] = FPK_ERR_CAPACITY; }
        return;
    }
    // ---- 6. canonical order: ascending (a,b,c,d) of the sorted site tuple. Bucket by a, rank inside the bucket ------------------------------
    if (blockIdx.x == 0) block_exclusive_scan(X.sitecnt, n + 1, s_red);
    grid.sync();
    for (int i = gtid; i < ns; i += gnt) {
        const int a0 = X.rawkey[i].x;
        const int slot = atomicAdd(&X.order[a0], 1);
        X.bucket[X.sitecnt[a0] + slot] = i;
    }
    grid.sync();
    for (int i = gtid; i < ns; i += gnt) {
        const int4 key = X.rawkey[i];
        const int b0 = X.sitecnt[key.x], b1 = X.sitecnt[key.x + 1];
        int rank = 0;
        for (int j = b0; j < b1; j++) {
            const int o = X.bucket[j];
            if (o != i && key_less(X.rawkey[o], key)) rank++;
        }
        const int dst = Sd.sph_off + b0 + rank;
        const float4 xr = X.rawxyzr[i];
        const int a[4] = {B.site2atom[Sd.site_off + key.x], B.site2atom[Sd.site_off + key.y],
                          B.site2atom[Sd.site_off + key.z], B.site2atom[Sd.site_off + key.w]};
        int apol = 0;
        float xs = 0.0f, ys = 0.0f, zs = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; j++) {       // typing (voronoi.c:467-495) and barycentre of the 4 contacted atoms (voronoi.c:892-916)
            if ((double)B.aen[Sd.prop_off + a[j]] < 2.8) apol++;
            xs += B.ax[Sd.atom_off + a[j]]; ys += B.ay[Sd.atom_off + a[j]]; zs += B.az[Sd.atom_off + a[j]];
        }
        B.s_xyzr[dst] = xr;
        B.s_key[dst] = key;
        B.s_neigh[dst] = make_int4(a[0], a[1], a[2], a[3]);
        B.s_type[dst] = (apol >= B.prm.min_apol_neigh) ? 0 : 1;
        B.s_bary[3 * (size_t)dst] = (float)((double)xs * 0.25);
        B.s_bary[3 * (size_t)dst + 1] = (float)((double)ys * 0.25);
        B.s_bary[3 * (size_t)dst + 2] = (float)((double)zs * 0.25);
    }
    if (gtid == 0) { counts[CNT_SPHERES] = ns; counts[CNT_STATUS] = FPK_OK; }
}

}  // namespace fpk
