// Warp-cooperative fast path of the block Delaunay (device only): same claim / verify / commit protocol as
// delaunay.cuh, laid out for the SM:
//   * sites live in SHARED memory as int32 lattice offsets from the bounding-box corner (12 B/site; exact: the
//     "%.6f" lattice of reference src/voronoi.c:130 in units of 1e-6 A), so a predicate costs LDS, not L2 trips;
//   * the cavity of a warp's point (tet id, 4 vertices, 4 neighbours, rim-face mask) lives in shared memory; one
//     BFS level = one round trip (claim pair + the 32-byte tet record of 8 x 4 neighbours, issued together);
//   * the winners' new tets are linked to each other through a per-warp edge hash in shared memory (each internal
//     face (p,a,b) is met by exactly two new tets) instead of rotating around rim edges through global memory;
//   * counters, per-warp state and chunk cursors live in shared memory.
// Cavities larger than WCAV tets take the serial path of delaunay.cuh (w_locate_and_claim / w_commit_insertion).
#pragma once
#include "delaunay.cuh"

namespace fpk {

enum { WCAV = 64, WHASH = 256, WALK_CAP = 20 };
#define FPK_HEMPTY 0xffffffffffffffffull

struct WarpCav {                        // one per warp, shared memory
    int4 v[WCAV];                       // vertices of cavity tet i
    int4 n[WCAV];                       // neighbours of cavity tet i; after pass A of the commit, component q of a rim face holds the NEW tet
    unsigned long long hkey[WHASH];     // edge hash: (lo+1) << 32 | (hi+1)
    int hval[WHASH];                    // 4 * new tet + slot of the face waiting for its twin
    int id[WCAV];
    unsigned char rf[4 * WCAV];         // per face 4 * i + q of cavity tet i: 1 = on the rim (its neighbour stays), 0 = interior
    unsigned char fq[4 * WCAV];         // frontier queue of the flood: faces still to be tested
    unsigned short rl[2 * WCAV + 8];    // compact list of rim faces (4 * i + q), filled by pass A of the commit
};

// Shared-memory state of one block's triangulation. Dynamic shared memory = [BlockDel][WarpCav x warps][sites]; every function
// reaches it through the accessors below so that the compiler sees SHARED addresses (LDS/STS/ATOMS), not generic ones.
#ifndef FPK_MAXWARPS_N
#define FPK_MAXWARPS_N 8
#endif
enum { FPK_MAXWARPS = FPK_MAXWARPS_N };
struct BlockDel {
    DelaunayCtx C;
    int cnt[C_COUNT];
    int tinfo[4 * FPK_MAXWARPS], cursor[2 * FPK_MAXWARPS];
    int first4[4];
    int bigw;
    int resume[FPK_MAXWARPS];      // tet at which the warp's capped walk stopped (-1: none)
    int prew_rank[FPK_MAXWARPS];   // the point a warp located AHEAD of time while it waited for the slower claimants of a sub-round ...
    int prew_tet[FPK_MAXWARPS];    // ... and where that walk got to
    int ndone;                     // claimants that have finished the claim phase of the current sub-round
};
#define FPK_BD_BYTES ((sizeof(BlockDel) + 15) & ~(size_t)15)
__device__ __forceinline__ unsigned char *dsm_base() { extern __shared__ __align__(16) unsigned char fpk_dsm[]; return fpk_dsm; }
__device__ __forceinline__ BlockDel &blk() { return *reinterpret_cast<BlockDel *>(dsm_base()); }
__device__ __forceinline__ WarpCav &my_cav() { return reinterpret_cast<WarpCav *>(dsm_base() + FPK_BD_BYTES)[threadIdx.x >> 5]; }
__device__ __forceinline__ int *sm_sites() { return reinterpret_cast<int *>(dsm_base() + FPK_BD_BYTES + (blockDim.x >> 5) * sizeof(WarpCav)); }
#define TVp(t) (tet + 8 * (size_t)(t))
#define TNp(t) (tet + 8 * (size_t)(t) + 4)
#define OWNp(t) (claim + 2 * (size_t)(t))
#define RIMp(t) (claim + 2 * (size_t)(t) + 1)

struct P3 { double x, y, z; };
__device__ __forceinline__ P3 ldp(const int *Pi, int i)
{
    // 2^52 + offset, built by placing the (non-negative) offset in the low word: no conversion instruction. Predicates only
    // use DIFFERENCES of coordinates, and (2^52 + a) - (2^52 + b) = a - b exactly.
    const int *q = Pi + 3 * (size_t)i;
    P3 r; r.x = __hiloint2double(0x43300000, q[0]); r.y = __hiloint2double(0x43300000, q[1]); r.z = __hiloint2double(0x43300000, q[2]);
    return r;
}
__device__ __forceinline__ int comp4(const int4 &a, int q) { return q == 0 ? a.x : (q == 1 ? a.y : (q == 2 ? a.z : a.w)); }

// ---- exact slow paths (rare): reload the lattice offsets, integer determinants ----------------------------------------
__device__ __noinline__ int orient3d_slow(const int *Pi, int ia, int ib, int ic, int id)
{
    const int *a = Pi + 3 * (size_t)ia, *b = Pi + 3 * (size_t)ib, *c = Pi + 3 * (size_t)ic, *d = Pi + 3 * (size_t)id;
    i128 adx = (long long)a[0] - d[0], ady = (long long)a[1] - d[1], adz = (long long)a[2] - d[2];
    i128 bdx = (long long)b[0] - d[0], bdy = (long long)b[1] - d[1], bdz = (long long)b[2] - d[2];
    i128 cdx = (long long)c[0] - d[0], cdy = (long long)c[1] - d[1], cdz = (long long)c[2] - d[2];
    i128 e = adx * (bdy * cdz - bdz * cdy) + bdx * (cdy * adz - cdz * ady) + cdx * (ady * bdz - adz * bdy);
    return e > 0 ? 1 : (e < 0 ? -1 : 0);
}
// same sign convention, filter and error bound as predicates.cuh orient3d (differences of lattice offsets are exact)
__device__ __forceinline__ int orient3d_v(const int *Pi, int ia, int ib, int ic, int id, const P3 &a, const P3 &b, const P3 &c, const P3 &d)
{
    double adx = a.x - d.x, ady = a.y - d.y, adz = a.z - d.z;
    double bdx = b.x - d.x, bdy = b.y - d.y, bdz = b.z - d.z;
    double cdx = c.x - d.x, cdy = c.y - d.y, cdz = c.z - d.z;
    double bdxcdy = bdx * cdy, cdxbdy = cdx * bdy, cdxady = cdx * ady, adxcdy = adx * cdy, adxbdy = adx * bdy, bdxady = bdx * ady;
    double det = adz * (bdxcdy - cdxbdy) + bdz * (cdxady - adxcdy) + cdz * (adxbdy - bdxady);
    double permanent = (fabs(bdxcdy) + fabs(cdxbdy)) * fabs(adz) + (fabs(cdxady) + fabs(adxcdy)) * fabs(bdz) +
                       (fabs(adxbdy) + fabs(bdxady)) * fabs(cdz);
    double errbound = 1.6e-15 * permanent;
    if (det > errbound) return 1;
    if (-det > errbound) return -1;
    return orient3d_slow(Pi, ia, ib, ic, id);
}

// exact insphere with the symbolic perturbation of delaunay.cuh insphere_sos (points ordered by site index)
__device__ __noinline__ int insphere_sos_slow(const int *Pi, int v0, int v1, int v2, int v3, int p)
{
    const int *a = Pi + 3 * (size_t)v0, *b = Pi + 3 * (size_t)v1, *c = Pi + 3 * (size_t)v2, *d = Pi + 3 * (size_t)v3, *e = Pi + 3 * (size_t)p;
    i128 aex = (long long)a[0] - e[0], aey = (long long)a[1] - e[1], aez = (long long)a[2] - e[2];
    i128 bex = (long long)b[0] - e[0], bey = (long long)b[1] - e[1], bez = (long long)b[2] - e[2];
    i128 cex = (long long)c[0] - e[0], cey = (long long)c[1] - e[1], cez = (long long)c[2] - e[2];
    i128 dex = (long long)d[0] - e[0], dey = (long long)d[1] - e[1], dez = (long long)d[2] - e[2];
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
    int r = i256_sign(acc);
    if (r) return r;
    int v[4] = {v0, v1, v2, v3};
    int idx[5] = {v0, v1, v2, v3, p};
#pragma unroll 1   // measured: the compiler's own unrolling of this loop costs the hull kernel 14 % (profiles/README.md)
    for (int i = 1; i < 5; i++) {
        int k = idx[i], j = i - 1;
#pragma unroll 1
        while (j >= 0 && idx[j] > k) { idx[j + 1] = idx[j]; j--; }
        idx[j + 1] = k;
    }
#pragma unroll 1
    for (int i = 4; i > 0; i--) {
        int q = idx[i], o;
        if (q == p) return -1;
        if (q == v[3] && (o = orient3d_slow(Pi, v[0], v[1], v[2], p)) != 0) return o;
        if (q == v[2] && (o = orient3d_slow(Pi, v[0], v[1], p, v[3])) != 0) return o;
        if (q == v[1] && (o = orient3d_slow(Pi, v[0], p, v[2], v[3])) != 0) return o;
        if (q == v[0] && (o = orient3d_slow(Pi, p, v[1], v[2], v[3])) != 0) return o;
    }
    return -1;
}

// > 0: p strictly inside the (perturbed) circumsphere of the positively oriented finite tet v
__device__ __forceinline__ int insphere_sos_v(const int *Pi, const int4 &v, int p, const P3 &e)
{
    const P3 a = ldp(Pi, v.x), b = ldp(Pi, v.y), c = ldp(Pi, v.z), d = ldp(Pi, v.w);
    double aex = a.x - e.x, aey = a.y - e.y, aez = a.z - e.z;
    double bex = b.x - e.x, bey = b.y - e.y, bez = b.z - e.z;
    double cex = c.x - e.x, cey = c.y - e.y, cez = c.z - e.z;
    double dex = d.x - e.x, dey = d.y - e.y, dez = d.z - e.z;
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
    double az = fabs(aez), bz = fabs(bez), cz = fabs(cez), dz = fabs(dez);
    double pab = fabs(aexbey) + fabs(bexaey), pbc = fabs(bexcey) + fabs(cexbey), pcd = fabs(cexdey) + fabs(dexcey);
    double pda = fabs(dexaey) + fabs(aexdey), pac = fabs(aexcey) + fabs(cexaey), pbd = fabs(bexdey) + fabs(dexbey);
    double permanent = (pcd * bz + pbd * cz + pbc * dz) * alift + (pda * cz + pac * dz + pcd * az) * blift +
                       (pab * dz + pbd * az + pda * bz) * clift + (pbc * az + pac * bz + pab * cz) * dlift;
    double errbound = 3.6e-15 * permanent;
    if (det > errbound) return 1;
    if (-det > errbound) return -1;
    return insphere_sos_slow(Pi, v.x, v.y, v.z, v.w, p);
}

// orientation of tet v with the vertex in `slot` replaced by p (all four of v finite except possibly v[slot])
__device__ __forceinline__ int orient_sub_v(const int *Pi, const int4 &v, int slot, int p, const P3 &pp)
{
    const int i0 = slot == 0 ? p : v.x, i1 = slot == 1 ? p : v.y, i2 = slot == 2 ? p : v.z, i3 = slot == 3 ? p : v.w;
    const P3 a = slot == 0 ? pp : ldp(Pi, i0), b = slot == 1 ? pp : ldp(Pi, i1), c = slot == 2 ? pp : ldp(Pi, i2), d = slot == 3 ? pp : ldp(Pi, i3);
    return orient3d_v(Pi, i0, i1, i2, i3, a, b, c, d);
}

__device__ __forceinline__ int inf_slot4(const int4 &v) { return v.x < 0 ? 0 : (v.y < 0 ? 1 : (v.z < 0 ? 2 : (v.w < 0 ? 3 : -1))); }

// in_conflict of delaunay.cuh on the shared-memory sites; xn = neighbours of the tet (needed for a ghost whose facet plane holds p)
__device__ __forceinline__ bool in_conflict_v(const int *tet, const int *Pi, const int4 &v, const int4 &xn, int p, const P3 &pp)
{
    const int k = inf_slot4(v);
    if (k < 0) return insphere_sos_v(Pi, v, p, pp) > 0;
    const int o = orient_sub_v(Pi, v, k, p, pp);
    if (o) return o > 0;
    const int4 w = *reinterpret_cast<const int4 *>(TVp(comp4(xn, k)));
    return insphere_sos_v(Pi, w, p, pp) > 0;
}

__device__ __forceinline__ int2 ldcg2(const int *p) { return __ldcg(reinterpret_cast<const int2 *>(p)); }

// ---- claim phase: locate the point, flood its cavity while claiming it (all 32 lanes, identical arguments) ------------------------
// GRID: the triangulation is shared by every block of a cooperative launch (k_delaunay_grid.cuh): the counters live in global memory
// (c.cnt) instead of the block's shared memory; everything else - per-warp state, cavity records - stays with the block.
template <bool SM, bool GRID = false> __device__ __noinline__ int w2_locate_and_claim(int rank, int own_hint, int *ncav_out)
{
    BlockDel &S = blk();
    DelaunayCtx &c = S.C;
    WarpCav &W = my_cav();
    int *const cnt = GRID ? c.cnt : S.cnt;
    if (GRID) __builtin_assume(__isGlobal(cnt));
    const int *Pi = SM ? sm_sites() : c.Pi;
    int *const tet = c.tet, *const claim = c.claim, *const hint = c.hint;
    const int *const order = c.order;
    __builtin_assume(__isGlobal(tet)); __builtin_assume(__isGlobal(claim)); __builtin_assume(__isGlobal(hint)); __builtin_assume(__isGlobal(order));
    if (!SM) __builtin_assume(__isGlobal(Pi));
    const int lane = threadIdx.x & 31;
    const int p = order[rank];
    const int prio = prio_of(rank);
    const P3 pp = ldp(Pi, p);
    *ncav_out = 0;
    const int *pe = Pi + 3 * (size_t)p;
    const int pe0 = pe[0], pe1 = pe[1], pe2 = pe[2];
    // start of the walk: where an earlier attempt on this point stopped, else the warp's last new tet (the previous point of its
    // Morton-ordered chunk), else the hint grid's tet for the point's cell, else the most recent tet of the block
    int t = S.resume[threadIdx.x >> 5];
    if (t < 0 && !GRID && S.prew_rank[threadIdx.x >> 5] == rank) t = S.prew_tet[threadIdx.x >> 5];      // located ahead of time (w2_prewalk)
    if (t < 0) t = own_hint;
    if (t < 0) {
        const int g = c.hg;
        int ix = (int)((double)pe0 * c.hinv), iy = (int)((double)pe1 * c.hinv), iz = (int)((double)pe2 * c.hinv);
        ix = ix < 0 ? 0 : (ix >= g ? g - 1 : ix); iy = iy < 0 ? 0 : (iy >= g ? g - 1 : iy); iz = iz < 0 ? 0 : (iz >= g ? g - 1 : iz);
        t = hint[(ix * g + iy) * g + iz];
    }
    if (t < 0) t = cnt[C_RECENT];
    int4 v, nb;
    int dead = 0;
#pragma unroll 1
    for (int steps = 0;; steps++) {         // visibility walk: the four face tests of a step run on four lanes
        v = *reinterpret_cast<const int4 *>(TVp(t));
        nb = *reinterpret_cast<const int4 *>(TNp(t));
        if (v.x == T_DEAD) {                // a dead tet forwards to a tet of the star that replaced it
            dead++;
            if (dead > 200) return ST_ERR;
            t = dead == 64 ? cnt[C_RECENT] : nb.x;
            continue;
        }
        if (steps - dead >= WALK_CAP) {     // bound the claim phase: go on from here in the next sub-round
            if (lane == 0) { S.resume[threadIdx.x >> 5] = t; atomicAdd(&cnt[C_NWALK], steps - dead); atomicAdd(&cnt[C_NHOP], dead); }
            return ST_LOST;
        }
        const int k = inf_slot4(v);
        if (k >= 0) {
            if (in_conflict_v(tet, Pi, v, nb, p, pp)) break;
            t = comp4(nb, k);
            continue;
        }
        int o = 1;
        bool dup = false;
        if (lane < 4) {
            // orient(v with slot q := p) = -+ orient(v[q+1], v[q+2], v[q+3], p): the cyclic shift is an even permutation for odd q
            int iq = v.x, i1 = v.y, i2 = v.z, i3 = v.w;                       // rotate left by `lane`
            if (lane & 1) { const int t0 = iq; iq = i1; i1 = i2; i2 = i3; i3 = t0; }
            if (lane & 2) { int t0 = iq; iq = i2; i2 = t0; t0 = i1; i1 = i3; i3 = t0; }
            const int *q = Pi + 3 * (size_t)iq;
            dup = q[0] == pe0 && q[1] == pe1 && q[2] == pe2;
            o = orient3d_v(Pi, i1, i2, i3, p, ldp(Pi, i1), ldp(Pi, i2), ldp(Pi, i3), pp);
            if (!(lane & 1)) o = -o;
        }
        const unsigned dupm = __ballot_sync(FPK_FULL, dup) & 0xfu, neg = __ballot_sync(FPK_FULL, o < 0) & 0xfu;
        if (dupm) { if (lane == 0) S.resume[threadIdx.x >> 5] = -1; *ncav_out = comp4(v, __ffs(dupm) - 1); return ST_DUP; }      // coincident site: reports the vertex it duplicates
        if (!neg) {
            if (lane == 0) { atomicAdd(&cnt[C_NWALK], steps - dead + 1); atomicAdd(&cnt[C_NATTEMPT], 1); atomicAdd(&cnt[C_NHOP], dead); }
            break;
        }
        t = comp4(nb, __ffs(neg) - 1);
    }
    if (lane == 0) S.resume[threadIdx.x >> 5] = t;      // a lost attempt starts again here; cleared when the point is done
    int st = ST_CLAIMED;
    if (lane == 0) {
        if (__ldcg(RIMp(t)) < prio) st = ST_LOST;
        else if (atomicMin(OWNp(t), prio) < prio) st = ST_LOST;
        else { W.id[0] = t; W.v[0] = v; W.n[0] = nb; }
    }
    st = __shfl_sync(FPK_FULL, st, 0);
    if (st == ST_LOST) return ST_LOST;
    __syncwarp();
    int ncav = 1, nlev = 0, qhead = 0, qtail = 4;
    if (lane < 4) W.fq[lane] = (unsigned char)lane;
    __syncwarp();
#pragma unroll 1
    while (qhead < qtail) {                 // breadth-first over FACES: up to 32 untested faces of the cavity per step, one per lane
        const int room = WCAV - ncav;                                   // every claimed tet must fit in the list (the reset walks it)
        if (room < 1) { *ncav_out = ncav; return ST_BIG; }
        const int nproc = min(min(32, room), qtail - qhead);
        const bool active = lane < nproc;
        bool lost = false, add = false, isrim = false;
        int x = -1, e = 0;
        int4 xv, xn;
        if (active) {
            e = W.fq[qhead + lane];
            x = reinterpret_cast<const int *>(W.n)[e];
            const int2 cl = ldcg2(OWNp(x));                       // one round trip: claim pair and the tet record together
            xv = *reinterpret_cast<const int4 *>(TVp(x));
            xn = *reinterpret_cast<const int4 *>(TNp(x));
            if (cl.x < prio) lost = true;                               // in (or next to) a stronger cavity
            else if (cl.x == prio) isrim = false;                       // already in my cavity: interior face
            else if (cl.y == prio) isrim = true;                        // already known as my rim
            else if (in_conflict_v(tet, Pi, xv, xn, p, pp)) {
                if (cl.y < prio) lost = true;                           // a stronger insertion rewrites x's adjacency
                else {
                    const int old = atomicMin(OWNp(x), prio);
                    if (old < prio) lost = true;
                    else if (old != prio) add = true;                   // exactly one lane sees the transition to prio
                }
            } else {
                atomicMin(RIMp(x), prio);
                isrim = true;
            }
            W.rf[e] = isrim ? 1 : 0;
        }
        qhead += nproc;
        const unsigned addm = __ballot_sync(FPK_FULL, add), lostm = __ballot_sync(FPK_FULL, lost);
        if (add) {
            const int rk = __popc(addm & ((1u << lane) - 1u)), pos = ncav + rk;
            W.id[pos] = x; W.v[pos] = xv; W.n[pos] = xn;
            const int cur = W.id[e >> 2];                               // the face of x back to its parent is interior; the other three join the queue
            const int r = xn.x == cur ? 0 : (xn.y == cur ? 1 : (xn.z == cur ? 2 : 3));
            W.rf[4 * pos + r] = 0;
            unsigned char *qq = W.fq + qtail + 3 * rk;
            qq[0] = (unsigned char)(4 * pos + ((r + 1) & 3)); qq[1] = (unsigned char)(4 * pos + ((r + 2) & 3)); qq[2] = (unsigned char)(4 * pos + ((r + 3) & 3));
        }
        qtail += 3 * __popc(addm);
        ncav += __popc(addm);
        nlev++;
        __syncwarp();
        if (lostm) { *ncav_out = ncav; return ST_LOST; }
    }
    if (lane == 0) atomicAdd(&cnt[C_NLEVEL], nlev);
    *ncav_out = ncav;
    return ST_CLAIMED;
}

// A warp that has finished its claim while slower claimants of the block are still at it uses the wait: it walks towards its NEXT point
// (read-only, the mesh does not change before the barrier) and stops as soon as the last claimant is done, so the barrier is never held up.
// Where it got to is only a hint for the next sub-round's walk (a tet destroyed meanwhile forwards or has been replaced in place).
template <bool SM> __device__ __noinline__ void w2_prewalk(int rank_next, int start, int nactive)
{
    BlockDel &S = blk();
    DelaunayCtx &c = S.C;
    const int *Pi = SM ? sm_sites() : c.Pi;
    int *const tet = c.tet;
    __builtin_assume(__isGlobal(tet)); __builtin_assume(__isGlobal(c.order));
    if (!SM) __builtin_assume(__isGlobal(Pi));
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int p = c.order[rank_next];
    const P3 pp = ldp(Pi, p);
    int t = start;
#pragma unroll 1
    for (int steps = 0; steps < 2 * WALK_CAP; steps++) {
        if (*reinterpret_cast<volatile int *>(&S.ndone) >= nactive) break;          // everybody is waiting for the barrier now
        const int4 v = *reinterpret_cast<const int4 *>(TVp(t));
        const int4 nb = *reinterpret_cast<const int4 *>(TNp(t));
        if (v.x == T_DEAD) { t = nb.x; continue; }
        const int k = inf_slot4(v);
        if (k >= 0) {
            if (in_conflict_v(tet, Pi, v, nb, p, pp)) break;
            t = comp4(nb, k);
            continue;
        }
        int o = 1;
        if (lane < 4) {
            int iq = v.x, i1 = v.y, i2 = v.z, i3 = v.w;
            if (lane & 1) { const int t0 = iq; iq = i1; i1 = i2; i2 = i3; i3 = t0; }
            if (lane & 2) { int t0 = iq; iq = i2; i2 = t0; t0 = i1; i1 = i3; i3 = t0; }
            o = orient3d_v(Pi, i1, i2, i3, p, ldp(Pi, i1), ldp(Pi, i2), ldp(Pi, i3), pp);
            if (!(lane & 1)) o = -o;
        }
        const unsigned neg = __ballot_sync(FPK_FULL, o < 0) & 0xfu;
        if (!neg) break;
        t = comp4(nb, __ffs(neg) - 1);
    }
    if (lane == 0) { S.prew_rank[w] = rank_next; S.prew_tet[w] = t; }
}

__device__ __forceinline__ bool w2_verify_claims(int rank, int ncav)
{
    const WarpCav &W = my_cav();
    const int *const claim = blk().C.claim;
    __builtin_assume(__isGlobal(claim));
    const int lane = threadIdx.x & 31, prio = prio_of(rank), q = lane & 3;
    bool ok = true;
#pragma unroll 1
    for (int base = 0; base < ncav; base += 8) {
        const int f = base + (lane >> 2);
        if (f < ncav) {
            if (__ldcg(OWNp(reinterpret_cast<const int *>(&W.n[f])[q])) < prio) ok = false;           // rim tet inside a stronger cavity
            if (q == 0) { const int2 cl = ldcg2(OWNp(W.id[f])); if (cl.x != prio || cl.y < prio) ok = false; }
        }
    }
    return __all_sync(FPK_FULL, ok);
}

__device__ __forceinline__ void w2_reset_claims(int ncav)
{
    const WarpCav &W = my_cav();
    int *const claim = blk().C.claim;
    __builtin_assume(__isGlobal(claim));
    const int lane = threadIdx.x & 31, q = lane & 3;
#pragma unroll 1
    for (int base = 0; base < ncav; base += 8) {
        const int f = base + (lane >> 2);
        if (f < ncav) {
            *reinterpret_cast<int2 *>(OWNp(reinterpret_cast<const int *>(&W.n[f])[q])) = make_int2(OWN_FREE, OWN_FREE);
            if (q == 0) *reinterpret_cast<int2 *>(OWNp(W.id[f])) = make_int2(OWN_FREE, OWN_FREE);
        }
    }
}

// a winner replaces its cavity (in W) by the star of p. Returns a new tet (the warp's next walk starts there) or -1.
template <bool GRID = false> __device__ __noinline__ int w2_commit_insertion(int rank, int ncav, int popstack)
{
    BlockDel &S = blk();
    DelaunayCtx &c = S.C;
    WarpCav &W = my_cav();
    int *const gc = GRID ? c.cnt : S.cnt;       // the triangulation counters (global memory in a grid-wide run)
    if (GRID) __builtin_assume(__isGlobal(gc));
    int *const tet = c.tet, *const claim = c.claim, *const mark = c.mark, *const hint = c.hint;
    int *const popst = c.freest[popstack], *const pushst = c.freest[popstack ^ 1];
    __builtin_assume(__isGlobal(tet)); __builtin_assume(__isGlobal(claim)); __builtin_assume(__isGlobal(mark)); __builtin_assume(__isGlobal(hint));
    __builtin_assume(__isGlobal(popst)); __builtin_assume(__isGlobal(pushst)); __builtin_assume(__isGlobal(c.order));
    const int lane = threadIdx.x & 31;
    const int p = c.order[rank];
#pragma unroll 1
    for (int i = lane; i < WHASH; i += 32) { W.hkey[i] = FPK_HEMPTY; W.hval[i] = -1; }
    __syncwarp();
    int mynew = 0x7fffffff, nrim = 0;
    bool bad = false;
    // ---- pass A: one new tet per rim face, outside adjacency both ways -----------------------------------------------------------------
    // Slots: the k-th new tet takes the slot of the k-th cavity tet (the cavity dies in this very step, its records are in shared memory and
    // nobody else may touch it: winners' cavities neither overlap nor touch). The star of p almost always has more tets than the cavity
    // (2 V - 4 faces around V rim vertices against V - 3 + interior edges); only the surplus comes from the free stack / fresh slots, and
    // only a cavity with more tets than rim faces leaves dead slots behind (pass C). New tets therefore land in cache lines the warp has
    // just read instead of in slots recycled from anywhere in the mesh, and tets that are neighbours in space stay neighbours in memory.
#pragma unroll 1
    for (int base = 0; base < 4 * ncav; base += 32) {
        const int e = base + lane, f = e >> 2, q = e & 3;
        const bool isrim = e < 4 * ncav && W.rf[e];
        const unsigned m = __ballot_sync(FPK_FULL, isrim);
        if (!m) continue;
        const int cnt = __popc(m), my = __popc(m & ((1u << lane) - 1u));
        const int kth = nrim + my;                                           // index of this lane's new tet among the warp's new tets
        const bool reuse = isrim && kth < ncav;
        const unsigned extra = __ballot_sync(FPK_FULL, isrim && !reuse);     // new tets beyond the cavity's own slots
        const int nextra = __popc(extra), myx = __popc(extra & ((1u << lane) - 1u));
        int fb = 0;
        if (nextra && lane == 0) fb = atomicAdd(&gc[C_NFREE0 + popstack], -nextra);
        fb = __shfl_sync(FPK_FULL, fb, 0);
        const int fi = nextra ? fb - 1 - myx : -1;                           // index into the free stack, < 0: take a fresh slot
        const unsigned fresh = __ballot_sync(FPK_FULL, isrim && !reuse && fi < 0);
        int nbase = 0;
        if (fresh && lane == 0) nbase = atomicAdd(&gc[C_NTET], __popc(fresh));
        nbase = __shfl_sync(FPK_FULL, nbase, 0);
        if (isrim) {
            int nt = reuse ? W.id[kth] : (fi >= 0 ? popst[fi] : nbase + __popc(fresh & ((1u << lane) - 1u)));
            if (nt >= c.cap) { bad = true; nt = -1; }
            if (nt >= 0) {
                const int cur = W.id[f];
                const int4 v = W.v[f];
                const int o = reinterpret_cast<const int *>(&W.n[f])[q];
                int4 nv = v;
                if (q == 0) nv.x = p; else if (q == 1) nv.y = p; else if (q == 2) nv.z = p; else nv.w = p;
                *reinterpret_cast<int4 *>(TVp(nt)) = nv;
                TNp(nt)[q] = o;
                *reinterpret_cast<int2 *>(OWNp(nt)) = make_int2(OWN_FREE, OWN_FREE);
                mark[2 * (size_t)nt] = -1;                       // the serial path (delaunay.cuh) tests marks of live tets
                const int4 ov = *reinterpret_cast<const int4 *>(TVp(o)), on = *reinterpret_cast<const int4 *>(TNp(o));
                // the slot of o that faced cur through this face: its vertex there is the one of o that is not on the face
                const int vq = reinterpret_cast<const int *>(&W.v[f])[q];
                int r = -1;
                if (on.x == cur && ov.x != v.x && ov.x != v.y && ov.x != v.z && ov.x != v.w) r = 0;
                else if (on.y == cur && ov.y != v.x && ov.y != v.y && ov.y != v.z && ov.y != v.w) r = 1;
                else if (on.z == cur && ov.z != v.x && ov.z != v.y && ov.z != v.z && ov.z != v.w) r = 2;
                else if (on.w == cur && ov.w != v.x && ov.w != v.y && ov.w != v.z && ov.w != v.w) r = 3;
                if (r < 0) {                                                    // o's apex coincides with v[q] only in the 5-tet start complex
                    if (on.x == cur && ov.x == vq) r = 0; else if (on.y == cur && ov.y == vq) r = 1;
                    else if (on.z == cur && ov.z == vq) r = 2; else if (on.w == cur && ov.w == vq) r = 3;
                }
                if (r >= 0) TNp(o)[r] = nt; else bad = true;
                reinterpret_cast<int *>(&W.n[f])[q] = nt;
                W.rl[nrim + my] = (unsigned short)e;
                mynew = min(mynew, nt);
            }
        }
        nrim += cnt;
    }
    if (__any_sync(FPK_FULL, bad)) { if (lane == 0) gc[C_STATUS] = DERR_CAPACITY; return -1; }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mynew = min(mynew, __shfl_xor_sync(FPK_FULL, mynew, d));
    const int firstnew = mynew;
    __syncwarp();
    // ---- pass B: one lane per (new tet, face through p): twins meet through the edge hash ---------------------------------------------
#pragma unroll 1
    for (int it = lane; it < 3 * nrim; it += 32) {
        const int ri = it / 3, jj = it - 3 * ri;
        const int e = W.rl[ri], f = e >> 2, q = e & 3;
        const int *vv = reinterpret_cast<const int *>(&W.v[f]);
        const int j = (q + 1 + jj) & 3;                                         // slot of the new tet whose opposite face holds p
        const int a = vv[(q + 1 + ((jj + 1) % 3)) & 3], b = vv[(q + 1 + ((jj + 2) % 3)) & 3];      // the rim edge of that face
        const int nt = reinterpret_cast<const int *>(&W.n[f])[q];
        const unsigned lo = (unsigned)(min(a, b) + 1), hi = (unsigned)(max(a, b) + 1);
        const unsigned long long key = ((unsigned long long)lo << 32) | hi;
        unsigned slot = (lo * 0x9E3779B1u ^ hi * 0x85EBCA77u) >> 11;
#pragma unroll 1
        for (int probe = 0; probe < WHASH; probe++, slot++) {
            slot &= WHASH - 1;
            const unsigned long long old = atomicCAS(&W.hkey[slot], FPK_HEMPTY, key);
            if (old == FPK_HEMPTY || old == key) break;
        }
        const int prev = atomicExch(&W.hval[slot], 4 * nt + j);
        if (prev >= 0) { TNp(nt)[j] = prev >> 2; TNp(prev >> 2)[prev & 3] = nt; }
    }
    __syncwarp();
    // ---- pass C: cavity tets whose slot no new tet took (a cavity with more tets than rim faces: rare) die; their slots go to the stack the
    // NEXT sub-round pops ------------------------------------------------------------------------------------------------------------------
    if (nrim < ncav) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&gc[C_NFREE0 + (popstack ^ 1)], ncav - nrim);
        base = __shfl_sync(FPK_FULL, base, 0);
#pragma unroll 1
        for (int i = nrim + lane; i < ncav; i += 32) {
            const int cur = W.id[i];
            TVp(cur)[0] = T_DEAD;
            TNp(cur)[0] = firstnew;
            pushst[base + i - nrim] = cur;
        }
    }
    if (lane == 0) {
        const int g = c.hg;
        const int *e = c.Pi + 3 * (size_t)p;
        int ix = (int)((double)e[0] * c.hinv), iy = (int)((double)e[1] * c.hinv), iz = (int)((double)e[2] * c.hinv);
        ix = ix < 0 ? 0 : (ix >= g ? g - 1 : ix); iy = iy < 0 ? 0 : (iy >= g ? g - 1 : iy); iz = iz < 0 ? 0 : (iz >= g ? g - 1 : iz);
        hint[(ix * g + iy) * g + iz] = firstnew;
        gc[C_RECENT] = firstnew;
    }
    return firstnew;
}

}  // namespace fpk
