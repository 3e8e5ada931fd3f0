// TEST-ONLY host build of fpocket_b200/csrc/delaunay.cuh: steps the block-synchronous phases of the
// CUDA Delaunay kernel sequentially (thread by thread, phase by phase) so its logic can be checked
// against the oracle in a container without a GPU. It is not a fallback: nothing in the package
// loads it, and it is built only by tests/test_hostemu_delaunay.py into tests/hostemu/_build/.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../fpocket_b200/csrc/delaunay.cuh"

using namespace fpk;

extern "C" int emu_delaunay(int n, const long long *ki, int nthreads, int cavmax_unused, int **tets_out, int *ntets,
                            long long *stats /* [8] */)
{
    (void)cavmax_unused;
    std::vector<double> P(3 * (size_t)n);
    for (size_t i = 0; i < 3 * (size_t)n; i++) P[i] = (double)ki[i];
    int cap = 10 * n + 256;
    std::vector<int> tet(8 * (size_t)cap, 0), claim(2 * (size_t)cap, OWN_FREE), mark(2 * (size_t)cap, -1);
    std::vector<int> f0(cap), f1(cap), cnt(C_COUNT, 0), order(n);
    std::vector<int> cav((size_t)nthreads * CAVMAX), newid((size_t)nthreads * CAVMAX * 4), tinfo(4 * (size_t)nthreads, 0),
        cursor(2 * (size_t)nthreads, 0), bigcav(cap), bignewid(4 * (size_t)cap);
    DelaunayCtx c;
    c.n = n; c.P = P.data(); c.tet = tet.data(); c.claim = claim.data(); c.mark = mark.data();
    c.cap = cap; c.cnt = cnt.data(); c.freest[0] = f0.data(); c.freest[1] = f1.data(); c.order = order.data();
    c.nworkers = nthreads; c.cav = cav.data(); c.newid = newid.data(); c.tinfo = tinfo.data(); c.cursor = cursor.data();
    c.bigcav = bigcav.data(); c.bignewid = bignewid.data();
    std::vector<int> rep(n);
    for (int i = 0; i < n; i++) rep[i] = i;
    c.rep = rep.data();
    // bounding box, keys, order
    double mn[3] = {P[0], P[1], P[2]}, mx[3] = {P[0], P[1], P[2]};
    for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) { mn[k] = std::min(mn[k], P[3 * i + k]); mx[k] = std::max(mx[k], P[3 * i + k]); }
    double ext = std::max(1.0, std::max(mx[0] - mn[0], std::max(mx[1] - mn[1], mx[2] - mn[2])));
    int R = brio_nrounds(n);
    std::vector<uint64_t> keys(n);
    for (int i = 0; i < n; i++) keys[i] = brio_key(i, &P[3 * i], mn, 1023.999 / ext, R);
    std::sort(keys.begin(), keys.end());
    std::vector<int> rstart(R + 1, n);
    for (int i = n - 1; i >= 0; i--) { order[i] = (int)(keys[i] & 0xffffffu); rstart[(int)(keys[i] >> 54)] = i; }
    for (int r = R - 1; r >= 0; r--) if (rstart[r] > rstart[r + 1]) rstart[r] = rstart[r + 1];
    rstart[0] = 0;
    int hg = 2;
    while (hg < 32 && hg * hg * hg * 2 < n) hg++;
    std::vector<int> hint((size_t)hg * hg * hg, -1);
    c.hint = hint.data(); c.hg = hg; c.hinv = hg / (ext * 1.0001);
    for (int k = 0; k < 3; k++) c.hmin[k] = mn[k];
    int first4[4];
    delaunay_init(c, first4);
    if (cnt[C_STATUS]) { *ntets = 0; *tets_out = nullptr; return cnt[C_STATUS]; }
    for (int t = 0; t < nthreads; t++) { tinfo[4 * t] = -1; tinfo[4 * t + 3] = -1; }
    for (int r = 0; r < R && !cnt[C_STATUS]; r++) {
        for (int t = 0; t < nthreads; t++) round_chunk(rstart[r], rstart[r + 1], t, nthreads, &cursor[2 * t], &cursor[2 * t + 1]);
        for (;;) {
            bool any = false;
            for (int t = 0; t < nthreads; t++) any |= phase_assign(c, t);
            if (!any || cnt[C_STATUS]) break;
            for (int t = 0; t < nthreads; t++) phase_claim(c, t);
            for (int t = 0; t < nthreads; t++) phase_verify(c, t);
            for (int t = 0; t < nthreads; t++) phase_reset(c, t);
            for (int t = nthreads - 1; t >= 0; t--) phase_commit(c, t);     // reversed on purpose: order must not matter
            phase_serial(c);
            for (int q = 0; q < cnt[C_NTET]; q++) if (claim[2 * q] != OWN_FREE || claim[2 * q + 1] != OWN_FREE) { fprintf(stderr, "emu: stale claim on tet %d\n", q); return 99; }
        }
    }
    int nt = 0;
    int *out = (int *)malloc(sizeof(int) * 4 * (size_t)(cnt[C_NTET] + 1));
    for (int t = 0; t < cnt[C_NTET]; t++) {
        const int *v = &tet[8 * (size_t)t];
        if (v[0] >= 0 && v[1] >= 0 && v[2] >= 0 && v[3] >= 0) { for (int k = 0; k < 4; k++) out[4 * nt + k] = rep[v[k]]; nt++; }
    }
    *tets_out = out; *ntets = nt;
    stats[0] = cnt[C_NSUB]; stats[1] = cnt[C_NRETRY]; stats[2] = cnt[C_NBIG]; stats[3] = cnt[C_NDUP]; stats[4] = cnt[C_NTET];
    stats[5] = R;
    return cnt[C_STATUS];
}
