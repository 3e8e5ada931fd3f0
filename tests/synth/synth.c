/*
 * Synthetic protein-like structures for the benchmark configurations of BASELINE.json:
 * a compact self-avoiding C-alpha walk confined to a sphere, one residue per step drawn uniformly from the
 * 20 amino acids with their real heavy-atom counts and elements, side chains grown as 1.5 A branches with a
 * non-bonded exclusion distance, coordinates on the PDB's 0.001 A lattice (jittered so that no lattice
 * degeneracy survives), crystal-like smooth B-factors. Deterministic in (seed, natoms).
 * Bench / test infrastructure only.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rnd64(rng_t *r) { uint64_t z = (r->s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static inline double rndu(rng_t *r) { return (double)(rnd64(r) >> 11) * (1.0 / 9007199254740992.0); }
static double rndn(rng_t *r) { double u = rndu(r) + 1e-300, v = rndu(r); return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v); }
static void rnddir(rng_t *r, double d[3]) { double z = 2 * rndu(r) - 1, t = 6.283185307179586 * rndu(r), s = sqrt(1 - z * z); d[0] = s * cos(t); d[1] = s * sin(t); d[2] = z; }

/* heavy atoms per residue, ALA ARG ASN ASP CYS GLN GLU GLY HIS ILE LEU LYS MET PHE PRO SER THR TRP TYR VAL (aa.h order) */
static const int AA_N[20] = {5, 11, 8, 8, 6, 9, 9, 4, 10, 8, 8, 9, 8, 11, 7, 6, 7, 14, 12, 7};
static const char *AA_NAME[20] = {"ALA", "ARG", "ASN", "ASP", "CYS", "GLN", "GLU", "GLY", "HIS", "ILE", "LEU", "LYS", "MET", "PHE", "PRO", "SER", "THR", "TRP", "TYR", "VAL"};
/* side-chain elements after N CA C O (0 C, 1 N, 2 O, 3 S) */
static const char *AA_SC[20] = {"0", "0001011", "0021", "0022", "03", "00021", "00022", "", "001001", "0000", "0000", "00001", "0030", "0000000", "000", "02", "020",
                                "0000100000", "00000002", "000"};

#define GRIDN 64
typedef struct { int *head, *next; double org[3], cell; int n; } grid_t;
static int gcell(const grid_t *g, const double p[3]) {
    int c[3];
    for (int k = 0; k < 3; k++) { int v = (int)((p[k] - g->org[k]) / g->cell); c[k] = v < 0 ? 0 : (v >= GRIDN ? GRIDN - 1 : v); }
    return (c[0] * GRIDN + c[1]) * GRIDN + c[2];
}
/* minimum distance from p to placed atoms, ignoring atoms `ex0`, `ex1` (bonded) */
static double mindist(const grid_t *g, const double *xyz, const double p[3], int ex0, int ex1, double stop) {
    int c[3];
    double best = 1e30;
    for (int k = 0; k < 3; k++) { int v = (int)((p[k] - g->org[k]) / g->cell); c[k] = v < 0 ? 0 : (v >= GRIDN ? GRIDN - 1 : v); }
    for (int a = c[0] - 1; a <= c[0] + 1; a++) for (int b = c[1] - 1; b <= c[1] + 1; b++) for (int d = c[2] - 1; d <= c[2] + 1; d++) {
        if (a < 0 || b < 0 || d < 0 || a >= GRIDN || b >= GRIDN || d >= GRIDN) continue;
        for (int i = g->head[(a * GRIDN + b) * GRIDN + d]; i >= 0; i = g->next[i]) {
            if (i == ex0 || i == ex1) continue;
            double dx = xyz[3 * i] - p[0], dy = xyz[3 * i + 1] - p[1], dz = xyz[3 * i + 2] - p[2], q = dx * dx + dy * dy + dz * dz;
            if (q < best) { best = q; if (best < stop * stop) return sqrt(best); }
        }
    }
    return sqrt(best);
}
static void gadd(grid_t *g, const double *xyz, int i) { int c = gcell(g, xyz + 3 * i); g->next[i] = g->head[c]; g->head[c] = i; }

/* Returns the number of atoms written (<= cap). Outputs per atom: xyz (float, 3-decimal lattice), bfactor, element (0..3),
 * residue number (1-based), aa index, atom name slot (0 N,1 CA,2 C,3 O,4.. side chain). density = atoms / A^3 of the confining sphere. */
int synth_protein(uint64_t seed, int natoms_target, double density, int cap, float *oxyz, float *obf, signed char *oelem, int *ores, signed char *oaa, signed char *oslot)
{
    rng_t R = {seed * 0x9E3779B97F4A7C15ull + 0x1234567ull};
    const double Rc = cbrt(3.0 * natoms_target / (4.0 * 3.141592653589793 * density));
    double *xyz = malloc(sizeof(double) * 3 * (size_t)(natoms_target + 32));
    grid_t G;
    G.head = malloc(sizeof(int) * GRIDN * GRIDN * GRIDN); G.next = malloc(sizeof(int) * (size_t)(natoms_target + 32));
    for (int i = 0; i < GRIDN * GRIDN * GRIDN; i++) G.head[i] = -1;
    G.cell = fmax(3.2, (2 * Rc + 12) / GRIDN);
    for (int k = 0; k < 3; k++) G.org[k] = -(Rc + 6);
    int n = 0, res = 0;
    double ca[3] = {0, 0, 0}, prev_dir[3] = {1, 0, 0};
    /* a second, coarse grid of C-alphas is not needed: C-alphas are atoms too */
    while (n + 14 <= natoms_target && n + 14 <= cap) {
        /* ---- next C-alpha: 3.8 A step, inside the sphere, away from everything placed so far ---- */
        double best[3] = {0, 0, 0}, bestscore = -1e30;
        if (res == 0) { rnddir(&R, best); for (int k = 0; k < 3; k++) best[k] *= 0.3 * Rc * rndu(&R); }
        else {
            for (int t = 0; t < 48; t++) {
                double d[3], p[3];
                rnddir(&R, d);
                /* persistence: bias towards the previous direction so that helices/strands-like stretches appear */
                for (int k = 0; k < 3; k++) d[k] = d[k] + 0.6 * prev_dir[k];
                double nd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) + 1e-12;
                for (int k = 0; k < 3; k++) { d[k] /= nd; p[k] = ca[k] + 3.8 * d[k]; }
                double rr = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
                double md = mindist(&G, xyz, p, -1, -1, 0.0);
                double score = fmin(md, 4.4) - (rr > Rc ? 4.0 * (rr - Rc) : 0.0) + 0.05 * rndu(&R);
                if (score > bestscore) { bestscore = score; memcpy(best, p, sizeof best); }
                if (md >= 4.4 && rr <= Rc) break;
            }
            double nd = 0;
            for (int k = 0; k < 3; k++) { prev_dir[k] = best[k] - ca[k]; nd += prev_dir[k] * prev_dir[k]; }
            nd = sqrt(nd) + 1e-12;
            for (int k = 0; k < 3; k++) prev_dir[k] /= nd;
        }
        memcpy(ca, best, sizeof ca);
        const int aa = (int)(rnd64(&R) % 20);
        const char *sc = AA_SC[aa];
        const int nsc = (int)strlen(sc);
        res++;
        /* backbone: N, CA, C, O then the side chain as a branch from CA */
        int idx_ca = n + 1, parent = -1, first = n;
        for (int j = 0; j < 4 + nsc; j++) {
            double p[3];
            int elem, par;
            double bond;
            if (j == 1) { memcpy(p, ca, sizeof p); elem = 0; par = -1; bond = 0; }
            else {
                if (j == 0) { elem = 1; par = -2; bond = 1.46; }        /* N bonded to CA (placed next) */
                else if (j == 2) { elem = 0; par = idx_ca; bond = 1.52; }
                else if (j == 3) { elem = 2; par = n - 1; bond = 1.23; }
                else { elem = sc[j - 4] - '0'; par = (j == 4) ? idx_ca : parent; bond = 1.52; if (aa == 17 || aa == 13 || aa == 18 || aa == 8) bond = 1.40; }
                const double *pp = (par == -2) ? ca : xyz + 3 * par;
                double bs = -1e30, bp[3] = {0, 0, 0};
                for (int t = 0; t < 24; t++) {
                    double d[3], q[3];
                    rnddir(&R, d);
                    for (int k = 0; k < 3; k++) q[k] = pp[k] + bond * d[k];
                    double md = mindist(&G, xyz, q, par, (par >= 0 && j >= 5) ? -1 : idx_ca, 0.0);
                    /* the not-yet-placed CA counts for N */
                    if (par == -2) { double dx = q[0] - ca[0], dy = q[1] - ca[1], dz = q[2] - ca[2]; (void)dx; (void)dy; (void)dz; }
                    double score = fmin(md, 3.1) + 0.02 * rndu(&R);
                    if (score > bs) { bs = score; memcpy(bp, q, sizeof bp); }
                    if (md >= 3.1) break;
                }
                memcpy(p, bp, sizeof p);
            }
            memcpy(xyz + 3 * n, p, sizeof p);
            oelem[n] = (signed char)elem; ores[n] = res; oaa[n] = (signed char)aa; oslot[n] = (signed char)j;
            gadd(&G, xyz, n);
            if (j >= 4) parent = n;
            n++;
        }
        (void)first;
    }
    /* PDB lattice: 3 decimals, jitter +-0.002, re-round; B-factor smooth in the distance from the centroid */
    double c[3] = {0, 0, 0}, rmax = 1e-9;
    for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) c[k] += xyz[3 * i + k] / n;
    for (int i = 0; i < n; i++) { double d = 0; for (int k = 0; k < 3; k++) d += (xyz[3 * i + k] - c[k]) * (xyz[3 * i + k] - c[k]); if (sqrt(d) > rmax) rmax = sqrt(d); }
    for (int i = 0; i < n; i++) {
        double d = 0;
        for (int k = 0; k < 3; k++) {
            double v = floor(xyz[3 * i + k] * 1000.0 + 0.5) / 1000.0 + 0.002 * (2 * rndu(&R) - 1) + 50.0;   /* +50: positive PDB coordinates */
            v = floor(v * 1000.0 + 0.5) / 1000.0;
            oxyz[3 * i + k] = (float)v;
            d += (xyz[3 * i + k] - c[k]) * (xyz[3 * i + k] - c[k]);
        }
        double b = 15.0 + 35.0 * d / (rmax * rmax) + 2.0 * rndn(&R);
        if (b < 2.0) b = 2.0;
        obf[i] = (float)(floor(b * 100.0 + 0.5) / 100.0);
    }
    free(xyz); free(G.head); free(G.next);
    (void)AA_N; (void)AA_NAME;
    return n;
}
