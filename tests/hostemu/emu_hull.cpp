// TEST-ONLY host build of fpocket_b200/csrc/hull_warp.cuh: the warp-per-pocket convex-hull routine of K4b run as a "warp"
// of one lane, so its logic (first tetrahedron, strict visibility, horizon, face bookkeeping, exact volume) can be checked
// against the oracle's Delaunay volume in a container without a GPU. It is not a fallback: nothing in the package loads it;
// tests/test_hostemu_hull.py builds it into tests/hostemu/_build/.
#include <cstdlib>
#include "../../fpocket_b200/csrc/hull_warp.cuh"

using namespace fpk;

// ki: n x 3 lattice coordinates. Returns 0 and *vol (float, the value the kernel stores), 1 = outside the fast path
// (the kernel would hand the pocket to the Delaunay route), 2 = the kernel's pre-checks (n, extent) send it there.
extern "C" int emu_hull(int n, const long long *ki, float *vol, double *vol_exact)
{
    if (n < 4 || n > HW_NCAP) return 2;
    HullWarp *H = (HullWarp *)malloc(sizeof(HullWarp));
    for (int i = 0; i < n; i++) {
        long long rx = ki[3 * i] - ki[0], ry = ki[3 * i + 1] - ki[1], rz = ki[3 * i + 2] - ki[2];
        if (llabs(rx) >= HW_MAXREL || llabs(ry) >= HW_MAXREL || llabs(rz) >= HW_MAXREL) { free(H); return 2; }
        H->px[i] = (int)rx; H->py[i] = (int)ry; H->pz[i] = (int)rz;
    }
    i128 tot = 0;
    int rc = hull_warp_run(*H, n, 0, &tot);
    free(H);
    if (rc) return 1;
    *vol = hull_total_to_volume(tot);
    long long h = (long long)(tot >> 64);
    unsigned long long l = (unsigned long long)tot;
    *vol_exact = ((double)h * 18446744073709551616.0 + (double)l) / 6.0 * 1e-18;
    return 0;
}
