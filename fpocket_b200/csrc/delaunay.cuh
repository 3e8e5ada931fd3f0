// Parallel exact 3-D Delaunay triangulation of one point set by one thread block.
//
// Replaces the reference's call into the vendored qhull (reference src/voronoi.c:111-138 ->
// src/qhull/src/qvoronoi/qvoronoi.c:236 run_qvoronoi -> libqhull.c:60 qh_qhull): instead of a
// sequential lifted convex hull with tolerance-based facet merging, points are inserted by
// Bowyer-Watson cavity retriangulation, SEVERAL POINTS PER STEP: every warp owns one pending point,
// finds its cavity on a read-only mesh - 32 lanes test 8 frontier tetrahedra x 4 faces at a time -
// while claiming the cavity and its rim with atomicMin on per-tet claim words (best priority wins),
// and the winners of a step, whose cavities neither overlap nor touch, rewrite the mesh concurrently.
// Predicates are exact (predicates.cuh) with Devillers-Teillaud symbolic perturbation, so the result is
// THE Delaunay triangulation of the lattice points, independent of insertion order and thread timing.
//
// Two forms of the same phases live here:
//   * scalar (one thread per point): plain C++; stepped sequentially by tests/hostemu (a test-only host
//     build, never shipped) to check the algorithm without a GPU;
//   * warp-cooperative (w_* functions, device only): what the kernels run.
#pragma once
#include "predicates.cuh"

namespace fpk {

enum { T_DEAD = -2, V_INF = -1, OWN_FREE = 0x7fffffff, CAVMAX = 192 };
enum { ST_IDLE = 0, ST_CLAIMED = 1, ST_LOST = 2, ST_BIG = 3, ST_DUP = 4, ST_WIN = 5, ST_ERR = 6 };
enum { C_NTET = 0, C_NFREE0 = 1, C_NFREE1 = 2, C_RECENT = 3, C_STATUS = 4, C_SUBROUND = 5, C_NDUP = 6, C_NBIG = 7,
       C_NSUB = 8, C_NRETRY = 9, C_NEXACT = 10, C_NWALK = 11, C_NLEVEL = 12, C_NATTEMPT = 13, C_NHOP = 14, C_COUNT = 16 };
enum { DERR_NONE = 0, DERR_DEGENERATE = 1, DERR_CAPACITY = 2, DERR_WALK = 3, DERR_ROTATE = 4 };

struct DelaunayCtx {
    int n;              // sites
    const double *P;    // [3n] lattice coordinates (integers stored in doubles)
    int *tet;           // [8*cap] per tet: 4 vertices then 4 neighbours (neighbour k is opposite vertex k): one 32-byte sector
    int *claim;         // [2*cap] per tet: own (cavity claim) and rim (rim claim): priority of the claimant, OWN_FREE if none
    int *mark;          // [2*cap] per tet: rank of the winner whose cavity holds it, and its position in that winner's list
    int cap;
    int *cnt;           // [C_COUNT] counters (C_*)
    int *freest[2];     // two free-slot stacks [cap] used alternately: pop from one, push to the other
    const int *order;   // [n] insertion order (site ids)
    int *hint;          // [hg^3] a recently created tet near each cell
    int hg;
    double hmin[3], hinv;
    int nworkers;       // scalar form: threads; warp form: warps
    int *cav;           // [nworkers*CAVMAX] cavity tets of the worker's point
    int *newid;         // [nworkers*CAVMAX*4] new tet on face q of cavity tet i (-1: interior face)
    int *tinfo;         // [nworkers*4] {rank of current point or -1, state, ncav, last tet created by this worker}
    int *cursor;        // [nworkers*2] {next, end} of the worker's chunk of `order`
    int *bigcav, *bignewid;   // [cap], [4*cap]: serial path for cavities > CAVMAX
    int first4[4];      // sites of the initial simplex (skipped by phase_assign)
    int *rep;           // [n] representative of every site: lowest index among coincident sites (identity when all distinct)
    const int *Pi;      // [3n] lattice offsets from the bounding-box corner (shared memory when they fit): fast path of delaunay_warp.cuh
};

#define FPK_TV(c, t) ((c).tet + 8 * (size_t)(t))
#define FPK_TN(c, t) ((c).tet + 8 * (size_t)(t) + 4)
#define FPK_OWN(c, t) ((c).claim + 2 * (size_t)(t))
#define FPK_RIM(c, t) ((c).claim + 2 * (size_t)(t) + 1)
#define FPK_MARK(c, t) ((c).mark[2 * (size_t)(t)])
#define FPK_CAVPOS(c, t) ((c).mark[2 * (size_t)(t) + 1])

// ---- insertion order: biased randomised rounds (BRIO), Morton order inside a round --------------------------
FPK_HD uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
FPK_HD int brio_nrounds(int n)
{
    int lg = 0;
    while ((1 << (lg + 1)) <= n) lg++;
    int r = lg - 3;
    return r < 1 ? 1 : (r > 24 ? 24 : r);
}
FPK_HD uint64_t spread10(uint32_t v)
{
    uint64_t x = v & 0x3ffu;
    x = (x | (x << 16)) & 0x30000ffull;
    x = (x | (x << 8)) & 0x300f00full;
    x = (x | (x << 4)) & 0x30c30c3ull;
    x = (x | (x << 2)) & 0x9249249ull;
    return x;
}
// key = round(8 bits) | morton(30 bits) | site(24 bits); the last round holds half of the points
FPK_HD uint64_t brio_key(int site_id, const double *p, const double mn[3], double inv1024, int nrounds)
{
    uint32_t h = hash32((uint32_t)site_id * 2654435761u + 12345u) | (1u << (nrounds - 1));
    int k = 0;
    while (!((h >> k) & 1u)) k++;
    int round = nrounds - 1 - k;
    int q[3];
    for (int a = 0; a < 3; a++) {
        int v = (int)((p[a] - mn[a]) * inv1024);
        q[a] = v < 0 ? 0 : (v > 1023 ? 1023 : v);
    }
#ifdef FPK_BRIO_MORTON
    uint64_t m = spread10((uint32_t)q[0]) | (spread10((uint32_t)q[1]) << 1) | (spread10((uint32_t)q[2]) << 2);
#else
    // Hilbert order inside a round (Skilling's transpose form, 10 bits per axis): unlike the Z curve it never jumps, so the walk from a
    // warp's previous point to its next one stays short (the order changes nothing in the result: the triangulation is unique)
    uint32_t X[3] = {(uint32_t)q[0], (uint32_t)q[1], (uint32_t)q[2]};
    for (uint32_t Q = 1u << 9; Q > 1; Q >>= 1) {
        const uint32_t P = Q - 1;
        for (int i = 0; i < 3; i++) {
            if (X[i] & Q) X[0] ^= P;
            else { const uint32_t t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
        }
    }
    X[1] ^= X[0]; X[2] ^= X[1];
    uint32_t t = 0;
    for (uint32_t Q = 1u << 9; Q > 1; Q >>= 1) if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t; X[1] ^= t; X[2] ^= t;
    uint64_t m = (spread10(X[0]) << 2) | (spread10(X[1]) << 1) | spread10(X[2]);
#endif
    return ((uint64_t)round << 54) | (m << 24) | (uint64_t)(uint32_t)site_id;
}

FPK_HD void ld4(const int *p, int v[4])
{
#if defined(__CUDA_ARCH__)
    int4 q = *reinterpret_cast<const int4 *>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
#else
    v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3];
#endif
}
FPK_HD void st4(int *p, const int v[4])
{
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<int4 *>(p) = make_int4(v[0], v[1], v[2], v[3]);
#else
    p[0] = v[0]; p[1] = v[1]; p[2] = v[2]; p[3] = v[3];
#endif
}
FPK_HD int atomic_min_i(int *p, int v)
{
#if defined(__CUDA_ARCH__)
    return atomicMin(p, v);
#else
    int o = *p; if (v < o) *p = v; return o;
#endif
}
FPK_HD int atomic_add_i(int *p, int v)
{
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, v);
#else
    int o = *p; *p = o + v; return o;
#endif
}
FPK_HD int ld_cg(const int *p)
{
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}

FPK_HD const double *site(const DelaunayCtx &c, int i) { return c.P + 3 * (size_t)i; }

// insphere with symbolic perturbation (Devillers & Teillaud): points ordered by site index
FPK_HD int insphere_sos(const DelaunayCtx &c, const int v[4], int p)
{
    int r = insphere(site(c, v[0]), site(c, v[1]), site(c, v[2]), site(c, v[3]), site(c, p));
    if (r) return r;
    int idx[5] = {v[0], v[1], v[2], v[3], p};
    for (int i = 1; i < 5; i++) {
        int k = idx[i], j = i - 1;
        while (j >= 0 && idx[j] > k) { idx[j + 1] = idx[j]; j--; }
        idx[j + 1] = k;
    }
    for (int i = 4; i > 0; i--) {
        int q = idx[i], o;
        if (q == p) return -1;
        if (q == v[3] && (o = orient3d(site(c, v[0]), site(c, v[1]), site(c, v[2]), site(c, p))) != 0) return o;
        if (q == v[2] && (o = orient3d(site(c, v[0]), site(c, v[1]), site(c, p), site(c, v[3]))) != 0) return o;
        if (q == v[1] && (o = orient3d(site(c, v[0]), site(c, p), site(c, v[2]), site(c, v[3]))) != 0) return o;
        if (q == v[0] && (o = orient3d(site(c, p), site(c, v[1]), site(c, v[2]), site(c, v[3]))) != 0) return o;
    }
    return -1;
}

// orientation of tet v with the vertex in `slot` replaced by point p
FPK_HD int orient_sub(const DelaunayCtx &c, const int v[4], int slot, int p)
{
    const double *q0 = site(c, slot == 0 ? p : v[0]), *q1 = site(c, slot == 1 ? p : v[1]);
    const double *q2 = site(c, slot == 2 ? p : v[2]), *q3 = site(c, slot == 3 ? p : v[3]);
    return orient3d(q0, q1, q2, q3);
}

FPK_HD int inf_slot(const int v[4])
{
    return v[0] < 0 ? 0 : (v[1] < 0 ? 1 : (v[2] < 0 ? 2 : (v[3] < 0 ? 3 : -1)));
}

// is point p inside the (perturbed) circumsphere of tet t?  Ghost tets (one vertex at infinity) are in
// conflict when p is beyond their hull facet, or on its plane and inside the facet's circumcircle.
FPK_HD bool in_conflict(const DelaunayCtx &c, int t, const int v[4], int p)
{
    int k = inf_slot(v);
    if (k < 0) return insphere_sos(c, v, p) > 0;
    int o = orient_sub(c, v, k, p);
    if (o) return o > 0;
    int nb = FPK_TN(c, t)[k], w[4];
    ld4(FPK_TV(c, nb), w);
    return insphere_sos(c, w, p) > 0;
}

FPK_HD int hint_cell(const DelaunayCtx &c, const double *p)
{
    int g = c.hg, ix = (int)((p[0] - c.hmin[0]) * c.hinv), iy = (int)((p[1] - c.hmin[1]) * c.hinv), iz = (int)((p[2] - c.hmin[2]) * c.hinv);
    ix = ix < 0 ? 0 : (ix >= g ? g - 1 : ix); iy = iy < 0 ? 0 : (iy >= g ? g - 1 : iy); iz = iz < 0 ? 0 : (iz >= g ? g - 1 : iz);
    return (ix * g + iy) * g + iz;
}

// ---- serial set-up by one thread: first simplex + its four ghosts -------------------------------------
FPK_HD bool pts_equal(const double *a, const double *b) { return a[0] == b[0] && a[1] == b[1] && a[2] == b[2]; }
FPK_HD bool pts_collinear(const double *a, const double *b, const double *c)
{
    i128 ux = (long long)(b[0] - a[0]), uy = (long long)(b[1] - a[1]), uz = (long long)(b[2] - a[2]);
    i128 vx = (long long)(c[0] - a[0]), vy = (long long)(c[1] - a[1]), vz = (long long)(c[2] - a[2]);
    return (uy * vz - uz * vy) == 0 && (uz * vx - ux * vz) == 0 && (ux * vy - uy * vx) == 0;
}

FPK_HDN void delaunay_init(DelaunayCtx &c, int first4[4])
{
    for (int k = 0; k < C_COUNT; k++) c.cnt[k] = 0;
    first4[0] = first4[1] = first4[2] = first4[3] = -1;
    for (int k = 0; k < 4; k++) c.first4[k] = -1;
    if (c.n < 4) { c.cnt[C_STATUS] = DERR_DEGENERATE; return; }
    int f[4] = {c.order[0], -1, -1, -1};
    for (int i = 1; i < c.n && f[1] < 0; i++) if (!pts_equal(site(c, c.order[i]), site(c, f[0]))) f[1] = c.order[i];
    if (f[1] >= 0) for (int i = 1; i < c.n && f[2] < 0; i++) if (!pts_collinear(site(c, f[0]), site(c, f[1]), site(c, c.order[i]))) f[2] = c.order[i];
    if (f[2] >= 0) for (int i = 1; i < c.n && f[3] < 0; i++) if (orient3d(site(c, f[0]), site(c, f[1]), site(c, f[2]), site(c, c.order[i])) != 0) f[3] = c.order[i];
    if (f[3] < 0) { c.cnt[C_STATUS] = DERR_DEGENERATE; return; }
    if (orient3d(site(c, f[0]), site(c, f[1]), site(c, f[2]), site(c, f[3])) < 0) { int t = f[0]; f[0] = f[1]; f[1] = t; }
    // tet 0 = finite; tets 1..4 = ghosts g_i = tet 0 with v_i := infinity and two other slots swapped (odd permutation)
    int T[5][4];
    for (int k = 0; k < 4; k++) T[0][k] = f[k];
    for (int i = 0; i < 4; i++) {
        for (int k = 0; k < 4; k++) T[1 + i][k] = f[k];
        T[1 + i][i] = V_INF;
        int a = (i + 1) & 3, b = (i + 2) & 3, tmp = T[1 + i][a];
        T[1 + i][a] = T[1 + i][b]; T[1 + i][b] = tmp;
    }
    for (int x = 0; x < 5; x++) {
        int nb[4] = {-1, -1, -1, -1};
        for (int k = 0; k < 4; k++)
            for (int y = 0; y < 5 && nb[k] < 0; y++) {
                if (y == x) continue;
                int shared = 0;
                for (int q = 0; q < 4; q++) { if (q == k) continue; for (int r = 0; r < 4; r++) if (T[y][r] == T[x][q]) { shared++; break; } }
                if (shared == 3) nb[k] = y;
            }
        st4(FPK_TV(c, x), T[x]); st4(FPK_TN(c, x), nb);
        *FPK_OWN(c, x) = OWN_FREE; *FPK_RIM(c, x) = OWN_FREE; FPK_MARK(c, x) = -1;
    }
    c.cnt[C_NTET] = 5;
    c.cnt[C_RECENT] = 0;
    for (int k = 0; k < 4; k++) { first4[k] = f[k]; c.first4[k] = f[k]; }
}

// Two claim words per tet, both lowered with atomicMin to the claimant's priority:
//   own  - the tet is in the claimant's CAVITY (will be destroyed)
//   rim  - the tet is face-adjacent to the claimant's cavity (its adjacency will be rewritten)
// Two insertions may proceed together iff no tet is in one cavity and in the other's cavity or rim
// (cavities neither overlap nor touch through a face: then neither new star can conflict with the other
// point); sharing a rim tet is harmless. The claimant with the smaller priority value wins a contested tet.
FPK_HD int prio_of(int rank) { return (int)(((hash32((uint32_t)rank ^ 0x9e3779b9u) & 0x7fu) << 24) | (uint32_t)rank); }

// starting tet of a walk: the worker's own last tet, else the hint grid, else the most recent tet; dead tets forward
FPK_HD int start_tet(const DelaunayCtx &c, int own_hint, const double *pp)
{
    int t = own_hint;
    if (t < 0) t = c.hint[hint_cell(c, pp)];
    if (t < 0) t = c.cnt[C_RECENT];
    for (int hop = 0; hop < 64 && FPK_TV(c, t)[0] == T_DEAD; hop++) t = FPK_TN(c, t)[0];
    if (FPK_TV(c, t)[0] == T_DEAD) {
        t = c.cnt[C_RECENT];
        for (int hop = 0; hop < 64 && FPK_TV(c, t)[0] == T_DEAD; hop++) t = FPK_TN(c, t)[0];
        if (FPK_TV(c, t)[0] == T_DEAD) return -1;
    }
    return t;
}

FPK_HD int alloc_tet(DelaunayCtx &c, int popstack)
{
    int i = atomic_add_i(&c.cnt[C_NFREE0 + popstack], -1) - 1;
    if (i >= 0) return c.freest[popstack][i];
    int t = atomic_add_i(&c.cnt[C_NTET], 1);
    if (t >= c.cap) { c.cnt[C_STATUS] = DERR_CAPACITY; return -1; }
    return t;
}

// new tet on rim face q of cavity tet `cur` (vertices v, neighbours nb): v with v[q] := p, outside neighbour kept, back pointer set
FPK_HD int make_rim_tet(DelaunayCtx &c, int cur, const int v[4], const int nb[4], int q, int p, int popstack)
{
    int nt = alloc_tet(c, popstack);
    if (nt < 0) return -1;
    int nv[4] = {v[0], v[1], v[2], v[3]};
    nv[q] = p;
    st4(FPK_TV(c, nt), nv);
    const int o = nb[q];
    FPK_TN(c, nt)[q] = o;
    *FPK_OWN(c, nt) = OWN_FREE; *FPK_RIM(c, nt) = OWN_FREE; FPK_MARK(c, nt) = -1;
    int w[4], wn[4];
    ld4(FPK_TV(c, o), w); ld4(FPK_TN(c, o), wn);
    for (int r = 0; r < 4; r++) {          // the slot of o that faced cur through this face
        if (wn[r] != cur) continue;
        bool inface = false;
        for (int z = 0; z < 4; z++) inface |= (z != q && v[z] == w[r]);
        if (!inface) { FPK_TN(c, o)[r] = nt; break; }
    }
    return nt;
}

// neighbour of the new tet on (cur0, q) across its face opposite v[j]: rotate around the rim edge through the cavity
FPK_HD int rotate_to_rim(const DelaunayCtx &c, int rank, int cur0, const int v[4], int q, int j, const int *newid, int ncav)
{
    int e0 = -9, e1 = -9;               // the rim edge: vertices of cur0 other than v[q], v[j]
    for (int z = 0; z < 4; z++) if (z != q && z != j) { if (e0 == -9) e0 = v[z]; else e1 = v[z]; }
    int cur = cur0, wv = v[j], other = v[q];
    int w[4], wn[4];
    for (int guard = 0; guard <= ncav + 1; guard++) {
        ld4(FPK_TV(c, cur), w); ld4(FPK_TN(c, cur), wn);
        int ws = (w[0] == wv) ? 0 : (w[1] == wv) ? 1 : (w[2] == wv) ? 2 : 3;
        int nx = wn[ws];
        if (FPK_MARK(c, nx) != rank) return newid[4 * FPK_CAVPOS(c, cur) + ws];
        ld4(FPK_TV(c, nx), w);
        int apex = -9;
        for (int z = 0; z < 4; z++) if (w[z] != other && w[z] != e0 && w[z] != e1) apex = w[z];
        wv = other; other = apex; cur = nx;
    }
    return -1;
}

// ================================================ scalar form =================================================
FPK_HDN int locate_and_claim(DelaunayCtx &c, int rank, int own_hint, int *list, int cap, int *ncav_out)
{
    const int p = c.order[rank];
    const int prio = prio_of(rank);
    const double *pp = site(c, p);
    *ncav_out = 0;
    int t = start_tet(c, own_hint, pp);
    if (t < 0) return ST_ERR;
    int v[4], nb[4];
    const int maxsteps = c.cnt[C_NTET] + 64;
    for (int steps = 0;; steps++) {         // visibility walk
        if (steps > maxsteps) return ST_ERR;
        ld4(FPK_TV(c, t), v);
        int k = inf_slot(v);
        if (k >= 0) {
            if (in_conflict(c, t, v, p)) break;
            t = FPK_TN(c, t)[k];
            continue;
        }
        for (int q = 0; q < 4; q++)
            if (pts_equal(site(c, v[q]), pp)) { *ncav_out = v[q]; return ST_DUP; }      // coincident site: reports the vertex it duplicates
        int go = -1;
        for (int q = 0; q < 4 && go < 0; q++)
            if (orient_sub(c, v, q, p) < 0) go = q;
        if (go < 0) break;
        t = FPK_TN(c, t)[go];
    }
    // breadth-first cavity; own==prio doubles as the visited flag of cavity tets, rim==prio of rim tets
    if (ld_cg(FPK_RIM(c, t)) < prio) return ST_LOST;
    if (atomic_min_i(FPK_OWN(c, t), prio) < prio) return ST_LOST;
    list[0] = t;
    int ncav = 1;
    for (int qi = 0; qi < ncav; qi++) {
        int cur = list[qi];
        ld4(FPK_TN(c, cur), nb);
        for (int q = 0; q < 4; q++) {
            int x = nb[q];
            int oc = ld_cg(FPK_OWN(c, x));
            if (oc == prio) continue;                                   // already in my cavity
            if (oc < prio) { *ncav_out = ncav; return ST_LOST; }        // in (or next to) a stronger cavity
            int rr = ld_cg(FPK_RIM(c, x));
            if (rr == prio) continue;                                   // already known as my rim
            ld4(FPK_TV(c, x), v);
            if (in_conflict(c, x, v, p)) {
                if (rr < prio) { *ncav_out = ncav; return ST_LOST; }    // a stronger insertion rewrites x's adjacency
                if (ncav >= cap) { *ncav_out = ncav; return ST_BIG; }
                list[ncav++] = x;                                       // recorded first so that a reset finds it
                if (atomic_min_i(FPK_OWN(c, x), prio) < prio) { *ncav_out = ncav; return ST_LOST; }
            } else {
                atomic_min_i(FPK_RIM(c, x), prio);
            }
        }
    }
    *ncav_out = ncav;
    return ST_CLAIMED;
}

FPK_HD bool verify_claims(const DelaunayCtx &c, int rank, const int *list, int ncav)
{
    const int prio = prio_of(rank);
    for (int i = 0; i < ncav; i++) {
        int cur = list[i], nb[4];
        if (ld_cg(FPK_OWN(c, cur)) != prio || ld_cg(FPK_RIM(c, cur)) < prio) return false;
        ld4(FPK_TN(c, cur), nb);
        for (int q = 0; q < 4; q++)
            if (ld_cg(FPK_OWN(c, nb[q])) < prio) return false;               // rim tet inside a stronger cavity
    }
    return true;
}

FPK_HD void reset_claims(DelaunayCtx &c, const int *list, int ncav)
{
    for (int i = 0; i < ncav; i++) {
        int cur = list[i], nb[4];
        ld4(FPK_TN(c, cur), nb);
        *FPK_OWN(c, cur) = OWN_FREE; *FPK_RIM(c, cur) = OWN_FREE;
        for (int q = 0; q < 4; q++) { *FPK_OWN(c, nb[q]) = OWN_FREE; *FPK_RIM(c, nb[q]) = OWN_FREE; }
    }
}

// a winner replaces its cavity by the star of p; returns a new tet (hint for the worker's next point) or -1
FPK_HDN int commit_insertion(DelaunayCtx &c, int rank, const int *list, int *newid, int ncav, int popstack)
{
    const int p = c.order[rank];
    for (int i = 0; i < ncav; i++) { FPK_MARK(c, list[i]) = rank; FPK_CAVPOS(c, list[i]) = i; }
    int firstnew = -1;
    int v[4], nb[4];
    for (int i = 0; i < ncav; i++) {
        int cur = list[i];
        ld4(FPK_TV(c, cur), v); ld4(FPK_TN(c, cur), nb);
        for (int q = 0; q < 4; q++) {
            if (FPK_MARK(c, nb[q]) == rank) { newid[4 * i + q] = -1; continue; }
            int nt = make_rim_tet(c, cur, v, nb, q, p, popstack);
            if (nt < 0) return -1;
            newid[4 * i + q] = nt;
            if (firstnew < 0) firstnew = nt;
        }
    }
    for (int i = 0; i < ncav; i++) {
        int cur0 = list[i];
        ld4(FPK_TV(c, cur0), v);
        for (int q = 0; q < 4; q++) {
            int nt = newid[4 * i + q];
            if (nt < 0) continue;
            for (int j = 0; j < 4; j++) {
                if (j == q) continue;
                int found = rotate_to_rim(c, rank, cur0, v, q, j, newid, ncav);
                if (found < 0) { c.cnt[C_STATUS] = DERR_ROTATE; return -1; }
                FPK_TN(c, nt)[j] = found;
            }
        }
    }
    int *pushst = c.freest[popstack ^ 1];
    int base = atomic_add_i(&c.cnt[C_NFREE0 + (popstack ^ 1)], ncav);
    for (int i = 0; i < ncav; i++) {
        int cur = list[i];
        FPK_TV(c, cur)[0] = T_DEAD;
        FPK_TN(c, cur)[0] = firstnew;
        pushst[base + i] = cur;
    }
    c.hint[hint_cell(c, site(c, p))] = firstnew;
    c.cnt[C_RECENT] = firstnew;
    return firstnew;
}

// P0: take the next point of my chunk if I have none pending. Returns true if this worker has work.
FPK_HD bool phase_assign(DelaunayCtx &c, int w)
{
    int *ti = c.tinfo + 4 * w, *cu = c.cursor + 2 * w;
    while (ti[0] < 0 && cu[0] < cu[1]) {
        int r = cu[0]++, s = c.order[r];
        if (s != c.first4[0] && s != c.first4[1] && s != c.first4[2] && s != c.first4[3]) ti[0] = r;
    }
    ti[1] = ST_IDLE;
    return ti[0] >= 0;
}
FPK_HD void phase_claim(DelaunayCtx &c, int w)
{
    int *ti = c.tinfo + 4 * w;
    if (ti[0] < 0) return;
    int ncav = 0;
    ti[1] = locate_and_claim(c, ti[0], ti[3], c.cav + (size_t)w * CAVMAX, CAVMAX, &ncav);
    ti[2] = ncav;
}
FPK_HD void phase_verify(DelaunayCtx &c, int w)
{
    int *ti = c.tinfo + 4 * w;
    if (ti[0] < 0 || ti[1] != ST_CLAIMED) return;
    ti[1] = verify_claims(c, ti[0], c.cav + (size_t)w * CAVMAX, ti[2]) ? ST_WIN : ST_LOST;
}
FPK_HD void phase_reset(DelaunayCtx &c, int w)
{
    int *ti = c.tinfo + 4 * w;
    if (ti[0] < 0) return;
    if (ti[1] == ST_WIN || ti[1] == ST_LOST || ti[1] == ST_BIG || ti[1] == ST_CLAIMED)
        reset_claims(c, c.cav + (size_t)w * CAVMAX, ti[2]);
}
// bookkeeping shared by the scalar and the warp form once a worker's attempt is decided (one thread per worker)
FPK_HD void after_attempt(DelaunayCtx &c, int w, int newtet)
{
    int *ti = c.tinfo + 4 * w;
    if (ti[1] == ST_WIN) { ti[0] = -1; ti[3] = newtet; }
    else if (ti[1] == ST_DUP) {      // ti[2] = the vertex this site coincides with: the lower index of the two represents both
        atomic_add_i(&c.cnt[C_NDUP], 1);
        if (c.rep) atomic_min_i(&c.rep[ti[2]], c.order[ti[0]]);
        ti[0] = -1;
    }
    else if (ti[1] == ST_ERR) { c.cnt[C_STATUS] = DERR_WALK; ti[0] = -1; }
    else if (ti[1] == ST_LOST) atomic_add_i(&c.cnt[C_NRETRY], 1);
}
FPK_HD void phase_commit(DelaunayCtx &c, int w)
{
    int *ti = c.tinfo + 4 * w;
    if (ti[0] < 0) return;
    int nt = -1;
    if (ti[1] == ST_WIN)
        nt = commit_insertion(c, ti[0], c.cav + (size_t)w * CAVMAX, c.newid + (size_t)w * CAVMAX * 4, ti[2], c.cnt[C_SUBROUND] & 1);
    after_attempt(c, w, nt);
}
FPK_HD void end_subround(DelaunayCtx &c)
{
    int popstack = c.cnt[C_SUBROUND] & 1;
    if (c.cnt[C_NFREE0 + popstack] < 0) c.cnt[C_NFREE0 + popstack] = 0;
    c.cnt[C_SUBROUND]++;
    c.cnt[C_NSUB]++;
    if (c.cnt[C_NSUB] > 64 * c.n + 4096) c.cnt[C_STATUS] = DERR_WALK;     // cannot happen (the best-ranked point always wins); never spin
}
// serial tail of a sub-round, one thread: insert at most one over-sized cavity, then flip the free stacks
FPK_HDN void phase_serial(DelaunayCtx &c)
{
    int popstack = c.cnt[C_SUBROUND] & 1;
    for (int t = 0; t < c.nworkers; t++) {
        int *ti = c.tinfo + 4 * t;
        if (ti[0] >= 0 && ti[1] == ST_BIG) {
            int ncav = 0;
            int st = locate_and_claim(c, ti[0], ti[3], c.bigcav, c.cap, &ncav);
            reset_claims(c, c.bigcav, ncav);
            if (st == ST_CLAIMED) ti[3] = commit_insertion(c, ti[0], c.bigcav, c.bignewid, ncav, popstack);
            else if (st != ST_DUP) c.cnt[C_STATUS] = DERR_CAPACITY;
            c.cnt[C_NBIG]++;
            ti[0] = -1;
            break;
        }
    }
    end_subround(c);
}

// chunk [lo,hi) of `order` handed to worker w for a round [r0,r1)
FPK_HD void round_chunk(int r0, int r1, int w, int nworkers, int *lo, int *hi)
{
    int len = r1 - r0, cs = (len + nworkers - 1) / nworkers;
    int a = r0 + w * cs, b = a + cs;
    if (a > r1) a = r1;
    if (b > r1) b = r1;
    *lo = a; *hi = b;
}

// ============================================ warp-cooperative form (device) ======================================
#if defined(__CUDACC__)
#define FPK_FULL 0xffffffffu

// All 32 lanes call these with identical arguments. Lists live in global scratch (L1/L2 resident).
__device__ __noinline__ int w_locate_and_claim(DelaunayCtx &c, int rank, int own_hint, int *list, int cap, int *ncav_out)
{
    const int lane = threadIdx.x & 31;
    const int p = c.order[rank];
    const int prio = prio_of(rank);
    const double *pp = site(c, p);
    *ncav_out = 0;
    int t = start_tet(c, own_hint, pp);
    if (t < 0) return ST_ERR;
    int v[4];
    const int maxsteps = c.cnt[C_NTET] + 64;
    for (int steps = 0;; steps++) {         // visibility walk: the four face tests of a step run on four lanes
        if (steps > maxsteps) return ST_ERR;
        ld4(FPK_TV(c, t), v);
        int k = inf_slot(v);
        if (k >= 0) {
            if (in_conflict(c, t, v, p)) break;
            t = FPK_TN(c, t)[k];
            continue;
        }
        for (int q = 0; q < 4; q++)
            if (pts_equal(site(c, v[q]), pp)) { *ncav_out = v[q]; return ST_DUP; }
        int o = 1;
        if (lane < 4) o = orient_sub(c, v, lane, p);
        unsigned neg = __ballot_sync(FPK_FULL, o < 0) & 0xfu;
        if (!neg) break;
        t = FPK_TN(c, t)[__ffs(neg) - 1];
    }
    int st = ST_CLAIMED;
    if (lane == 0) {
        if (ld_cg(FPK_RIM(c, t)) < prio) st = ST_LOST;
        else if (atomicMin(FPK_OWN(c, t), prio) < prio) st = ST_LOST;
        else list[0] = t;
    }
    st = __shfl_sync(FPK_FULL, st, 0);
    if (st == ST_LOST) return ST_LOST;
    __syncwarp();
    int ncav = 1;
    // breadth-first cavity: 8 frontier tets x 4 faces per step
    for (int head = 0; head < ncav;) {
        if (ncav + 32 > cap) { *ncav_out = ncav; return ST_BIG; }       // every claimed tet must fit in the list (the reset walks it)
        const int f = head + (lane >> 2), q = lane & 3;
        head = min(head + 8, ncav);                                     // tets appended by this step are expanded by the next ones
        bool lost = false, add = false;
        int x = -1;
        if (f < head) {
            x = FPK_TN(c, list[f])[q];
            int oc = ld_cg(FPK_OWN(c, x));
            if (oc < prio) lost = true;                                 // in (or next to) a stronger cavity
            else if (oc != prio) {
                int rr = ld_cg(FPK_RIM(c, x));
                if (rr != prio) {
                    ld4(FPK_TV(c, x), v);
                    if (in_conflict(c, x, v, p)) {
                        if (rr < prio) lost = true;                     // a stronger insertion rewrites x's adjacency
                        else {
                            int old = atomicMin(FPK_OWN(c, x), prio);
                            if (old < prio) lost = true;
                            else if (old != prio) add = true;           // exactly one lane sees the transition to prio
                        }
                    } else atomicMin(FPK_RIM(c, x), prio);
                }
            }
        }
        const unsigned addm = __ballot_sync(FPK_FULL, add), lostm = __ballot_sync(FPK_FULL, lost);
        if (add) list[ncav + __popc(addm & ((1u << lane) - 1u))] = x;
        ncav += __popc(addm);
        __syncwarp();
        if (lostm) { *ncav_out = ncav; return ST_LOST; }
    }
    *ncav_out = ncav;
    return ST_CLAIMED;
}

__device__ __forceinline__ bool w_verify_claims(const DelaunayCtx &c, int rank, const int *list, int ncav)
{
    const int lane = threadIdx.x & 31, prio = prio_of(rank);
    bool ok = true;
    for (int i = lane; i < ncav; i += 32) {
        int cur = list[i], nb[4];
        if (ld_cg(FPK_OWN(c, cur)) != prio || ld_cg(FPK_RIM(c, cur)) < prio) ok = false;
        ld4(FPK_TN(c, cur), nb);
        for (int q = 0; q < 4; q++)
            if (ld_cg(FPK_OWN(c, nb[q])) < prio) ok = false;
    }
    return __all_sync(FPK_FULL, ok);
}

__device__ __forceinline__ void w_reset_claims(DelaunayCtx &c, const int *list, int ncav)
{
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < ncav; i += 32) {
        int cur = list[i], nb[4];
        ld4(FPK_TN(c, cur), nb);
        *reinterpret_cast<int2 *>(FPK_OWN(c, cur)) = make_int2(OWN_FREE, OWN_FREE);
        for (int q = 0; q < 4; q++) *reinterpret_cast<int2 *>(FPK_OWN(c, nb[q])) = make_int2(OWN_FREE, OWN_FREE);
    }
}

__device__ __noinline__ int w_commit_insertion(DelaunayCtx &c, int rank, const int *list, int *newid, int ncav, int popstack)
{
    const int lane = threadIdx.x & 31;
    const int p = c.order[rank];
    for (int i = lane; i < ncav; i += 32) { FPK_MARK(c, list[i]) = rank; FPK_CAVPOS(c, list[i]) = i; }
    __syncwarp();
    int mynew = 0x7fffffff;
    bool bad = false;
    int v[4], nb[4];
    for (int e = lane; e < 4 * ncav; e += 32) {                     // one lane per (cavity tet, face)
        const int i = e >> 2, q = e & 3, cur = list[i];
        ld4(FPK_TV(c, cur), v); ld4(FPK_TN(c, cur), nb);
        if (FPK_MARK(c, nb[q]) == rank) { newid[e] = -1; continue; }
        int nt = make_rim_tet(c, cur, v, nb, q, p, popstack);
        if (nt < 0) { bad = true; newid[e] = -1; continue; }
        newid[e] = nt;
        if (nt < mynew) mynew = nt;
    }
    if (__any_sync(FPK_FULL, bad)) return -1;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mynew = min(mynew, __shfl_xor_sync(FPK_FULL, mynew, d));
    const int firstnew = mynew;
    __syncwarp();
    for (int e = lane; e < 4 * ncav; e += 32) {                     // adjacency between the new tets
        const int nt = newid[e];
        if (nt < 0) continue;
        const int i = e >> 2, q = e & 3, cur0 = list[i];
        ld4(FPK_TV(c, cur0), v);
        for (int j = 0; j < 4; j++) {
            if (j == q) continue;
            int found = rotate_to_rim(c, rank, cur0, v, q, j, newid, ncav);
            if (found < 0) { bad = true; break; }
            FPK_TN(c, nt)[j] = found;
        }
    }
    if (__any_sync(FPK_FULL, bad)) { if (lane == 0) c.cnt[C_STATUS] = DERR_ROTATE; return -1; }
    __syncwarp();
    int base = 0;
    if (lane == 0) base = atomicAdd(&c.cnt[C_NFREE0 + (popstack ^ 1)], ncav);
    base = __shfl_sync(FPK_FULL, base, 0);
    int *pushst = c.freest[popstack ^ 1];
    for (int i = lane; i < ncav; i += 32) {
        int cur = list[i];
        FPK_TV(c, cur)[0] = T_DEAD;
        FPK_TN(c, cur)[0] = firstnew;
        pushst[base + i] = cur;
    }
    if (lane == 0) { c.hint[hint_cell(c, site(c, p))] = firstnew; c.cnt[C_RECENT] = firstnew; }
    return firstnew;
}

// one sub-round for warp w of the block (all lanes). The caller separates the phases with __syncthreads().
__device__ __forceinline__ bool w_phase_assign(DelaunayCtx &c, int w)
{
    int has = 0;
    if ((threadIdx.x & 31) == 0) has = phase_assign(c, w) ? 1 : 0;
    return __shfl_sync(FPK_FULL, has, 0) != 0;
}
__device__ __forceinline__ void w_phase_claim(DelaunayCtx &c, int w)
{
    int *ti = c.tinfo + 4 * w;
    const int rank = ti[0];
    if (rank < 0) return;
    int ncav = 0;
    int st = w_locate_and_claim(c, rank, ti[3], c.cav + (size_t)w * CAVMAX, CAVMAX, &ncav);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) { ti[1] = st; ti[2] = ncav; }
    __syncwarp();
}
__device__ __forceinline__ void w_phase_verify(DelaunayCtx &c, int w)
{
    int *ti = c.tinfo + 4 * w;
    if (ti[0] < 0 || ti[1] != ST_CLAIMED) return;
    bool ok = w_verify_claims(c, ti[0], c.cav + (size_t)w * CAVMAX, ti[2]);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) ti[1] = ok ? ST_WIN : ST_LOST;
    __syncwarp();
}
__device__ __forceinline__ void w_phase_reset(DelaunayCtx &c, int w)
{
    int *ti = c.tinfo + 4 * w;
    if (ti[0] < 0) return;
    const int st = ti[1];
    if (st == ST_WIN || st == ST_LOST || st == ST_BIG || st == ST_CLAIMED) w_reset_claims(c, c.cav + (size_t)w * CAVMAX, ti[2]);
}
__device__ __forceinline__ void w_phase_commit(DelaunayCtx &c, int w)
{
    int *ti = c.tinfo + 4 * w;
    if (ti[0] < 0) return;
    int nt = -1;
    if (ti[1] == ST_WIN)
        nt = w_commit_insertion(c, ti[0], c.cav + (size_t)w * CAVMAX, c.newid + (size_t)w * CAVMAX * 4, ti[2], c.cnt[C_SUBROUND] & 1);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) after_attempt(c, w, nt);
    __syncwarp();
}
// tail of a sub-round, executed by warp 0 only: at most one over-sized cavity through the big lists, then flip the stacks
__device__ __forceinline__ void w_phase_serial(DelaunayCtx &c)
{
    const int lane = threadIdx.x & 31;
    const int popstack = c.cnt[C_SUBROUND] & 1;
    int big = -1;
    for (int t = 0; t < c.nworkers && big < 0; t++)
        if (c.tinfo[4 * t] >= 0 && c.tinfo[4 * t + 1] == ST_BIG) big = t;
    if (big >= 0) {
        int *ti = c.tinfo + 4 * big;
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
        }
    }
    if (lane == 0) end_subround(c);
}
#endif  // __CUDACC__

}  // namespace fpk
