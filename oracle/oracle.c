/*
 * TEST INFRASTRUCTURE — CPU oracle (see oracle.h). Never linked into or called by the product.
 *
 * Every function cites the reference lines it restates (paths relative to the reference root).
 * Arithmetic discipline: the reference is gcc -O2 x86-64, scalar SSE2, no FMA
 * contraction, float/double promotions exactly as C99 dictates. This file is compiled with
 * -ffp-contract=off and spells every promotion out.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

#include "oracle_delaunay.inc"   /* exact predicates + Bowyer-Watson (same translation unit) */

/* ------------------------------------------------------------------------------------------------ */
/* calc.c:51-58 dist(), calc.c:76-83 ddist(): float differences/products/sums, sqrt through double    */
static float f_dist(float x1, float y1, float z1, float x2, float y2, float z2)
{
    float xdif = x1 - x2, ydif = y1 - y2, zdif = z1 - z2;
    float s = (xdif * xdif) + (ydif * ydif) + (zdif * zdif);
    return (float)sqrt((double)s);
}
static float f_ddist(float x1, float y1, float z1, float x2, float y2, float z2)
{
    float xdif = x1 - x2, ydif = y1 - y2, zdif = z1 - z2;
    return (xdif * xdif) + (ydif * ydif) + (zdif * zdif);
}

/* voronoi.c:94-107 (h_tr) and :114-132 ("%.6f" text that qhull parses as double) */
int orc_sites(const fpk_structure *in, int *site2atom, long long *ki, double *xd)
{
    int n = 0;
    for (int i = 0; i < in->natoms; i++) {
        if (in->flags[i] & FPK_ATOM_H) continue;
        const float c[3] = {in->x[i], in->y[i], in->z[i]};
        for (int k = 0; k < 3; k++) {
            /* float * 1e6 is exact in double (24-bit x 14-bit significands); rint = round-half-even = printf */
            long long q = llrint((double)c[k] * 1e6);
            ki[3 * n + k] = q;
            xd[3 * n + k] = (double)q / 1e6;   /* correctly rounded quotient == strtod of the decimal */
        }
        site2atom[n++] = i;
    }
    return n;
}

/* qhull geom2.c:2012-2093 qh_voronoi_center: Cramer relative to the simplex' first point, f64 */
static double det3(const double *r0, const double *r1, const double *r2)
{
    /* qhull geom.c det3_ macro order */
    return r0[0] * (r1[1] * r2[2] - r1[2] * r2[1]) - r1[0] * (r0[1] * r2[2] - r0[2] * r2[1]) +
           r2[0] * (r0[1] * r1[2] - r0[2] * r1[1]);
}
void orc_circumcentre(const double *p0, const double *p1, const double *p2, const double *p3, double c[3])
{
    /* gm_row[k][j] = p_{j+1}[k] - p0[k]; sum2[j] = |p_{j+1}-p0|^2 */
    double m[3][3], s2[3];
    const double *p[3] = {p1, p2, p3};
    for (int k = 0; k < 3; k++)
        for (int j = 0; j < 3; j++) m[k][j] = p[j][k] - p0[k];
    for (int j = 0; j < 3; j++) {
        double s = 0.0;
        for (int k = 0; k < 3; k++) s += m[k][j] * m[k][j];
        s2[j] = s;
    }
    double det = det3(m[0], m[1], m[2]);
    double factor = 0.5 / det;
    for (int i = 0; i < 3; i++) {
        const double *r[3];
        for (int k = 0; k < 3; k++) r[k] = (k == i) ? s2 : m[k];
        c[i] = det3(r[0], r[1], r[2]) * factor + p0[i];
    }
}

/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    float x, y, z, ray, bary[3];
    int neigh[4];
    int type;
    int cluster;
} sph_t;

/* voronoi.c:969-1087 testVvertice. a[4] = atom indices in the tet's vertex order. */
static float test_vvertice(const fpk_structure *in, const fpk_params *p, const float xyz[3], const int a[4])
{
    float min_asph_size = p->asph_min_size, max_asph_size = p->asph_max_size;
    float x = xyz[0] - 0.0f, y = xyz[1] - 0.0f, z = xyz[2] - 0.0f;
    float baryx = 0.0f, baryy = 0.0f, baryz = 0.0f, barybf = 0.0f, sdbf = 0.0f;
    if (p->xpocket_active) {
        int ok = 0;
        for (int k = 0; k < 4; k++)
            if (in->flags[a[k]] & FPK_ATOM_XPOCKET) ok++;
        if (ok < p->min_n_explicit_pocket_atoms) return -3.0f;
    }
    int c = a[0];
    baryx += in->x[c]; baryy += in->y[c]; baryz += in->z[c]; barybf += in->bfactor[c];
    float d1 = f_dist(x, y, z, in->x[c], in->y[c], in->z[c]), d2, d3, d4;
    if ((double)min_asph_size <= (double)d1 + 1e-3 && (double)d1 - 1e-3 <= (double)max_asph_size) {
        c = a[1]; d2 = f_dist(x, y, z, in->x[c], in->y[c], in->z[c]);
        baryx += in->x[c]; baryy += in->y[c]; baryz += in->z[c]; barybf += in->bfactor[c];
        c = a[2]; d3 = f_dist(x, y, z, in->x[c], in->y[c], in->z[c]);
        baryx += in->x[c]; baryy += in->y[c]; baryz += in->z[c]; barybf += in->bfactor[c];
        c = a[3]; d4 = f_dist(x, y, z, in->x[c], in->y[c], in->z[c]);
        baryx += in->x[c]; baryy += in->y[c]; baryz += in->z[c]; barybf += in->bfactor[c];
        baryx = (float)((double)baryx * 0.25); baryy = (float)((double)baryy * 0.25);
        baryz = (float)((double)baryz * 0.25); barybf = (float)((double)barybf * 0.25);
        if (fabs((double)(d1 - d2)) < 1e-3 && fabs((double)(d1 - d3)) < 1e-3 && fabs((double)(d1 - d4)) < 1e-3) {
            if (in->n_xlig) {
                for (int i = 0; i < in->n_xlig; i++)
                    if (f_dist(in->xlig[3 * i], in->xlig[3 * i + 1], in->xlig[3 * i + 2], x, y, z) <= d1) return d1;
                return -1.0f;
            }
            for (int k = 0; k < 4; k++) {
                float b = in->bfactor[a[k]];
                sdbf += (b - barybf) * (b - barybf);
            }
            sdbf = (float)sqrt((double)(sdbf / 3));
            if (((sdbf > in->avg_bfactor) || (sdbf > ((in->max_bfactor - in->min_bfactor) / 4))) &&
                (((double)in->avg_bfactor > 0.0) && ((double)(barybf / in->avg_bfactor) > 1.4)))
                return -1.0f;
            if ((double)f_dist(baryx, baryy, baryz, xyz[0], xyz[1], xyz[2]) > 1.0 &&
                (double)d1 > ((double)max_asph_size - 1.5))
                return -2.0f;
            return d1;
        }
    }
    return -3.0f;
}

/* voronoi.c:369-521 fill_vvertices (+ :892-916 set_barycenter) over the tet list */
static int filter_tets(const fpk_structure *in, const fpk_params *p, const int *site2atom, const double *xd,
                       const int *tets, int ntets, sph_t **out)
{
    sph_t *S = malloc(sizeof(sph_t) * (size_t)(ntets > 0 ? ntets : 1));
    int ns = 0;
    for (int t = 0; t < ntets; t++) {
        const int *s = tets + 4 * t;
        double c[3];
        orc_circumcentre(xd + 3 * s[0], xd + 3 * s[1], xd + 3 * s[2], xd + 3 * s[3], c);
        float xyz[3] = {(float)c[0], (float)c[1], (float)c[2]};   /* sscanf("%f") of "%6.16g" text, voronoi.c:443 */
        int a[4] = {site2atom[s[0]], site2atom[s[1]], site2atom[s[2]], site2atom[s[3]]};
        float r = test_vvertice(in, p, xyz, a);
        if (r > 0) {
            sph_t *v = S + ns++;
            v->x = xyz[0] - 0.0f; v->y = xyz[1] - 0.0f; v->z = xyz[2] - 0.0f; v->ray = r;
            int apol = 0;
            for (int j = 0; j < 4; j++) {
                if ((double)in->electroneg[a[j]] < 2.8) apol++;
                v->neigh[j] = a[j];
            }
            v->type = (apol >= p->min_apol_neigh) ? 0 : 1;
            v->cluster = -1;
            float xs = 0.0f, ys = 0.0f, zs = 0.0f;
            for (int j = 0; j < 4; j++) { xs += in->x[a[j]]; ys += in->y[a[j]]; zs += in->z[a[j]]; }
            v->bary[0] = (float)((double)xs * 0.25); v->bary[1] = (float)((double)ys * 0.25); v->bary[2] = (float)((double)zs * 0.25);
        }
    }
    *out = S;
    return ns;
}

/* clusterlib.c:933-1018 euclid + :3514 pslcluster + :3235 cuttree_distance + refine.c:58 apply_clustering:
 * connected components of the graph  sqrt(dx^2+dy^2+dz^2) (f64 of f32 centres) <= (double)clust_max_dist,
 * numbered by lowest sphere index (SURVEY Appendix B check). */
static int uf_find(int *par, int i) { while (par[i] != i) { par[i] = par[par[i]]; i = par[i]; } return i; }
/* -e (fparams.c: distance_measure): 'e' euclid (the default, clusterlib.c:933-1018) or 'b' city block (clusterlib.c:1022-1083: the MEAN of the three absolute
 * differences, weights 1). Single linkage cut at clust_max_dist is the connected components of "distance <= cut" under either. fpk_params has no field
 * for it yet (the library refuses anything but 'e'): the oracle takes it through this switch (oracle first, SURVEY par. 8 row f4). */
static int g_distance_measure = 'e';
void orc_set_distance_measure(int c) { g_distance_measure = c; }

/* clusterlib.c:933 / :1022 on two rows of f64 data (weights 1, every mask set): what setmetric() hands to distancematrix and pclcluster */
static double metric3(const double *a, const double *b)
{
    double result = 0.;
    if (g_distance_measure == 'b') {
        double tweight = 0;
        for (int k = 0; k < 3; k++) { double term = a[k] - b[k]; result = result + 1.0 * fabs(term); tweight += 1.0; }
        result /= tweight;
    } else {
        for (int k = 0; k < 3; k++) { double term = a[k] - b[k]; result += term * term; }
        result = sqrt(result);
    }
    return result;
}

/* -C m | a | c: the agglomerative linkages of clusterlib.c (pmlcluster :3705 complete, palcluster :3781 average, pclcluster :3337 centroid) followed by
 * cuttree_distance (:3235). All three walk a ragged lower-triangular f64 matrix (distancematrix :2919), take the FIRST strict minimum in row-major order
 * (find_closest_pair :285), fold row/column `is` into `js` and move the last row into slot `is`; ties therefore depend on that layout, which is kept here
 * exactly. The cut walks the nodes from the last merge down: a node at or below the cut labels every leaf under it that has no label yet, so a leaf's
 * cluster is its highest-numbered ancestor with distance <= cut (with centroid linkage distances are not monotone, and such an ancestor may sit above
 * nodes that exceed the cut); a leaf with no such ancestor stays alone. */
static void linkage_labels(const sph_t *S, int ns, double cut, int method, int *label)
{
    const size_t n0 = (size_t)ns;
    double *D = malloc(sizeof(double) * (n0 * (n0 - 1) / 2 + 1));
#define DM(i, j) D[(size_t)(i) * ((size_t)(i) - 1) / 2 + (size_t)(j)]                 /* i > j */
    double *pos = malloc(sizeof(double) * 3 * n0);
    int *cnt = malloc(sizeof(int) * n0), *cid = malloc(sizeof(int) * n0);
    int *nl = malloc(sizeof(int) * n0), *nr = malloc(sizeof(int) * n0);
    double *nd = malloc(sizeof(double) * n0);
    for (int i = 0; i < ns; i++) { pos[3 * i] = (double)S[i].x; pos[3 * i + 1] = (double)S[i].y; pos[3 * i + 2] = (double)S[i].z; cnt[i] = 1; cid[i] = i; }
    for (int i = 1; i < ns; i++) for (int j = 0; j < i; j++) DM(i, j) = metric3(pos + 3 * i, pos + 3 * j);
    for (int n = ns, node = 0; n > 1; n--, node++) {
        int is = 1, js = 0;
        double best = DM(1, 0);
        for (int i = 1; i < n; i++) for (int j = 0; j < i; j++) { double t = DM(i, j); if (t < best) { best = t; is = i; js = j; } }
        nd[node] = best; nl[node] = cid[is]; nr[node] = cid[js];
        const int last = n - 1;
        if (method == 'c') {
            for (int k = 0; k < 3; k++) {
                pos[3 * js + k] = pos[3 * js + k] * cnt[js] + pos[3 * is + k] * cnt[is];
                pos[3 * js + k] /= (cnt[js] + cnt[is]);
            }
            cnt[js] += cnt[is];
            for (int k = 0; k < 3; k++) pos[3 * is + k] = pos[3 * last + k];
            cnt[is] = cnt[last];
            for (int j = 0; j < is; j++) DM(is, j) = DM(last, j);
            for (int j = is + 1; j < last; j++) DM(j, is) = DM(last, j);
            for (int j = 0; j < js; j++) DM(js, j) = metric3(pos + 3 * js, pos + 3 * j);
            for (int j = js + 1; j < last; j++) DM(j, js) = metric3(pos + 3 * js, pos + 3 * j);
        } else {
            const int sum = cnt[is] + cnt[js];
            for (int j = 0; j < n; j++) {
                if (j == js || j == is) continue;
                double a = j < is ? DM(is, j) : DM(j, is);
                double *b = j < js ? &DM(js, j) : &DM(j, js);
                if (method == 'm') *b = a > *b ? a : *b;
                else { *b = a * cnt[is] + *b * cnt[js]; *b /= sum; }
            }
            for (int j = 0; j < is; j++) DM(is, j) = DM(last, j);
            for (int j = is + 1; j < last; j++) DM(j, is) = DM(last, j);
            cnt[js] = sum; cnt[is] = cnt[last];
        }
        cid[js] = -node - 1; cid[is] = cid[last];
    }
#undef DM
    /* cuttree_distance: top-most (= highest-numbered) node <= cut owns its leaves */
    for (int i = 0; i < ns; i++) label[i] = -1;
    int next = 0, *stack = malloc(sizeof(int) * (n0 + 2));
    for (int node = ns - 2; node >= 0; node--) {
        if (nd[node] > cut) {
            if (nl[node] >= 0 && label[nl[node]] < 0) label[nl[node]] = next;
            next++;
            if (nr[node] >= 0 && label[nr[node]] < 0) label[nr[node]] = next;
            next++;
        } else {
            next++;
            int sp = 0;
            stack[sp++] = -node - 1;
            while (sp) {
                int x = stack[--sp];
                if (x >= 0) { if (label[x] < 0) label[x] = next; }
                else { stack[sp++] = nr[-x - 1]; stack[sp++] = nl[-x - 1]; }
            }
        }
    }
    free(stack); free(D); free(pos); free(cnt); free(cid); free(nl); free(nr); free(nd);
}

static int cluster_spheres(sph_t *S, int ns, float clust_max_dist, int skip, int measure, int method)
{
    if (measure == 'b' || measure == 'e') g_distance_measure = measure;        /* fpk_params.distance_measure; 0 keeps what orc_set_distance_measure chose */
    if (skip) { for (int i = 0; i < ns; i++) S[i].cluster = 0; return ns ? 1 : 0; }
    if ((method == 'm' || method == 'a' || method == 'c') && ns >= 2) {
        int *raw = malloc(sizeof(int) * (size_t)(ns + 1)), *first = malloc(sizeof(int) * 2 * (size_t)(ns + 1)), nc = 0;
        linkage_labels(S, ns, (double)clust_max_dist, method, raw);
        for (int i = 0; i < 2 * ns + 2; i++) first[i] = -1;
        for (int i = 0; i < ns; i++) {                                        /* refine.c:58 apply_clustering: clusters in order of their lowest sphere */
            if (first[raw[i]] < 0) first[raw[i]] = nc++;
            S[i].cluster = first[raw[i]];
        }
        free(raw); free(first);
        return nc;
    }
    int *par = malloc(sizeof(int) * (size_t)(ns + 1));
    for (int i = 0; i < ns; i++) par[i] = i;
    double cut = (double)clust_max_dist;
    for (int i = 0; i < ns; i++) {
        double xi = (double)S[i].x, yi = (double)S[i].y, zi = (double)S[i].z;
        for (int j = i + 1; j < ns; j++) {
            double result = 0.;
            if (g_distance_measure == 'b') {
                double tweight = 0;
                double term = xi - (double)S[j].x; result = result + 1.0 * fabs(term); tweight += 1.0;
                term = yi - (double)S[j].y; result = result + 1.0 * fabs(term); tweight += 1.0;
                term = zi - (double)S[j].z; result = result + 1.0 * fabs(term); tweight += 1.0;
                result /= tweight;
            } else {
                double term = xi - (double)S[j].x; result += term * term;
                term = yi - (double)S[j].y; result += term * term;
                term = zi - (double)S[j].z; result += term * term;
                result = sqrt(result);
            }
            if (!(result > cut)) {
                int a = uf_find(par, i), b = uf_find(par, j);
                if (a != b) { if (a < b) par[b] = a; else par[a] = b; }
            }
        }
    }
    /* number clusters by order of first (lowest) member */
    int nc = 0;
    int *label = malloc(sizeof(int) * (size_t)(ns + 1));
    for (int i = 0; i < ns; i++) label[i] = -1;
    for (int i = 0; i < ns; i++) {
        int r = uf_find(par, i);
        if (label[r] < 0) label[r] = nc++;
        S[i].cluster = label[r];
    }
    free(par); free(label);
    return nc;
}

/* test entry: the clustering step alone on n centres (x, y, z per point), any method / measure; label[i] = cluster of point i, numbered by lowest member */
int orc_cluster_points(const float *xyz, int n, float clust_max_dist, int measure, int method, int *label)
{
    sph_t *S = calloc((size_t)n + 1, sizeof(sph_t));
    for (int i = 0; i < n; i++) { S[i].x = xyz[3 * i]; S[i].y = xyz[3 * i + 1]; S[i].z = xyz[3 * i + 2]; }
    const int saved = g_distance_measure;
    const int nc = cluster_spheres(S, n, clust_max_dist, 0, measure, method);
    g_distance_measure = saved;
    for (int i = 0; i < n; i++) label[i] = S[i].cluster;
    free(S);
    return nc;
}

/* aa.c:60-80 property table: volume score, hydrophobicity, charge, polarity */
static const float AA_VOL[20] = {2, 7, 3, 3, 3, 4, 4, 1, 4, 5, 5, 6, 5, 6, 3, 2, 3, 8, 7, 4};
static const float AA_HYD[20] = {41, -14, -28, -55, 49, -10, -31, 0, 8, 99, 97, -23, 74, 100, -46, -5, 13, 97, 63, 76};
static const int AA_CHG[20] = {0, 1, 0, -1, 0, 0, -1, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0};
static const int AA_POL[20] = {0, 1, 1, 1, 0, 1, 1, 0, 1, 0, 0, 1, 0, 0, 0, 1, 1, 1, 1, 0};

/* asa.c:438-454 get_points_on_sphere(100) */
void orc_spiral_points(float *pts)
{
    const int nop = 100;
    float inc = (float)(3.1415926535897932384626433832795 * (3.0 - sqrt(5.0)));
    float off = (float)(2.0 / (double)(float)nop);
    for (int k = 0; k < nop; k++) {
        float y = (float)((double)(((float)k) * off) - 1.0 + ((double)off / 2.0));
        float r = (float)sqrt(1.0 - (double)(y * y));
        float phi = (float)k * inc;
        pts[3 * k] = (float)(cos((double)phi) * (double)r);
        pts[3 * k + 1] = y;
        pts[3 * k + 2] = (float)(sin((double)phi) * (double)r);
    }
}

/* Philox4x32-10 (Salmon et al., SC'11), the counter-based generator the CUDA path uses */
void orc_philox4x32(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned out[4])
{
    for (int r = 0; r < 10; r++) {
        unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
        unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* utils.c:113-128 rand_uniform: ((float)rand()/(float)RAND_MAX)*(max-min) + min, all float */
static float unif_from_bits(unsigned r31, float mn, float mx)
{
    float rnd = ((float)(int)r31 / (float)2147483647) * (mx - mn);
    return mn + rnd;
}

/* voronoi.c:1150-1233 get_verts_volume_ptr(verts, n, niter, -1.6) */
static float mc_volume(const sph_t *S, const int *idx, int n, int niter, float correct, int rng_mode,
                       unsigned long long seed, int cluster)
{
    float xmin = 0, xmax = 0, ymin = 0, ymax = 0, zmin = 0, zmax = 0, xtmp, ytmp, ztmp;
    for (int i = 0; i < n; i++) {
        const sph_t *v = S + idx[i];
        if (i == 0) {
            xmin = v->x - v->ray + correct; xmax = v->x + v->ray + correct;
            ymin = v->y - v->ray + correct; ymax = v->y + v->ray + correct;
            zmin = v->z - v->ray + correct; zmax = v->z + v->ray + correct;
        } else {
            if (xmin > (xtmp = v->x - v->ray + correct)) xmin = xtmp;
            else if (xmax < (xtmp = v->x + v->ray + correct)) xmax = xtmp;
            if (ymin > (ytmp = v->y - v->ray + correct)) ymin = ytmp;
            else if (ymax < (ytmp = v->y + v->ray + correct)) ymax = ytmp;
            if (zmin > (ztmp = v->z - v->ray + correct)) zmin = ztmp;
            else if (zmax < (ztmp = v->z + v->ray + correct)) zmax = ztmp;
        }
    }
    float vbox = (xmax - xmin) * (ymax - ymin) * (zmax - zmin);
    float sum_volume = 0.0f;
    const int nreps = 100;
    for (int li = 0; li < nreps; li++) {
        int nb_in = 0;
        for (int i = 0; i < niter; i++) {
            float xr, yr, zr;
            if (rng_mode == ORC_RNG_LIBC) {
                xr = unif_from_bits((unsigned)rand(), xmin, xmax);
                yr = unif_from_bits((unsigned)rand(), ymin, ymax);
                zr = unif_from_bits((unsigned)rand(), zmin, zmax);
            } else {
                unsigned o[4];
                orc_philox4x32((unsigned)i, (unsigned)li, (unsigned)cluster, 0u, (unsigned)seed, (unsigned)(seed >> 32), o);
                xr = unif_from_bits(o[0] >> 1, xmin, xmax);
                yr = unif_from_bits(o[1] >> 1, ymin, ymax);
                zr = unif_from_bits(o[2] >> 1, zmin, zmax);
            }
            for (int j = 0; j < n; j++) {
                const sph_t *v = S + idx[j];
                xtmp = v->x - xr; ytmp = v->y - yr; ztmp = v->z - zr;
                if (((correct + v->ray) * (v->ray + correct)) > (xtmp * xtmp + ytmp * ytmp + ztmp * ztmp)) { nb_in++; break; }
            }
        }
        sum_volume += ((float)nb_in) / ((float)niter) * vbox;
    }
    return sum_volume / (float)nreps;
}

/* voronoi.c:1235-1291 get_convex_hull_volume: qconvex "FS" on the "%.6f" text of the centres.
 * Restated as: volume of the convex hull = sum of the volumes of the Delaunay tets of those points. */
static float hull_volume(const sph_t *S, const int *idx, int n)
{
    if (n < 10) return 0.0f;
    long long *ki = malloc(sizeof(long long) * 3 * (size_t)n);
    for (int i = 0; i < n; i++) {
        const sph_t *v = S + idx[i];
        ki[3 * i] = llrint((double)v->x * 1e6); ki[3 * i + 1] = llrint((double)v->y * 1e6); ki[3 * i + 2] = llrint((double)v->z * 1e6);
    }
    int *tets = NULL, nt = 0;
    double vol = 0.0;
    int rc = orc_delaunay(n, ki, &tets, &nt, &vol);
    free(ki); free(tets);
    if (rc) return -1.0f;            /* voronoi.c:1276-1279 */
    return (float)vol;               /* sscanf("%f") of "%6.16g" */
}

/* asa.c:240-419 set_ASA, with :84-112 get_surrounding_atoms_idx and :183-216 get_unique_atoms_DEPRECATED */
static void set_asa(const fpk_structure *in, const sph_t *S, const int *idx, int n, fpk_desc *d, float *atom_fx,
                    const float *pts)
{
    const int same = in->same_pdb;
    const int nw = same ? in->natoms : in->natoms_w;
    const float *wx = same ? in->x : in->wx, *wy = same ? in->y : in->wy, *wz = same ? in->z : in->wz;
    const float *wr = same ? in->radius : in->wradius, *we = same ? in->electroneg : in->welectroneg;
    const uint8_t *wf = same ? in->flags : in->wflags;
    d->surf_pol_vdw14 = d->surf_apol_vdw14 = d->surf_vdw14 = d->surf_vdw22 = d->surf_pol_vdw22 = d->surf_apol_vdw22 = 0.0f;
    d->n_abpa = 0;
    int *sa = malloc(sizeof(int) * (size_t)(nw + 1)), n_sa = 0;
    for (int i = 0; i < nw; i++) {
        if (wf[i] & (same ? FPK_ATOM_H1 : FPK_ATOM_H)) continue;
        for (int z = 0; z < n; z++) {
            const sph_t *v = S + idx[z];
            double rp = (double)v->ray + 1.0;
            if ((double)f_ddist(wx[i], wy[i], wz[i], v->x, v->y, v->z) < rp * rp) { sa[n_sa++] = i; break; }
        }
    }
    int *ua = malloc(sizeof(int) * (size_t)(4 * n + 1)), n_ua = 0;
    for (int z = 0; z < n; z++)
        for (int j = 0; j < 4; j++) {
            int a = S[idx[z]].neigh[j], seen = 0;
            for (int k = 0; k < n_ua; k++) if (ua[k] == a) { seen = 1; break; }
            if (!seen) ua[n_ua++] = a;
        }
    float *abpatmp = malloc(sizeof(float) * (size_t)(n_ua + 1));
    for (int pass = 0; pass < 2; pass++) {
        const double probe = pass == 0 ? 1.4 : 2.2;
        for (int i = 0; i < n_ua; i++) {
            int cura = ua[i];
            float n_apol = 0, n_pol = 0;
            int nnburried = 0;
            float cx = in->x[cura], cy = in->y[cura], cz = in->z[cura], crad = in->radius[cura];
            for (int k = 0; k < 100; k++) {
                int burried = 0, j = 0, vrefburried = 1;
                float tx = (float)((double)cx + (double)pts[3 * k] * ((double)crad + probe));
                float ty = (float)((double)cy + (double)pts[3 * k + 1] * ((double)crad + probe));
                float tz = (float)((double)cz + (double)pts[3 * k + 2] * ((double)crad + probe));
                for (int iv = 0; iv < n; iv++) {
                    const sph_t *v = S + idx[iv];
                    for (int q = 0; q < 4; q++)
                        if (v->neigh[q] == cura)
                            if (f_ddist(tx, ty, tz, v->x, v->y, v->z) <= v->ray * v->ray) vrefburried = 0;
                }
                while (!burried && j < n_sa && !vrefburried) {
                    int a = sa[j];
                    if (!(same && a == cura)) {               /* asa.c:315 pointer inequality */
                        float dsq = f_ddist(tx, ty, tz, wx[a], wy[a], wz[a]);
                        double rr = (double)wr[a] + probe;
                        if ((double)dsq < rr * rr) {
                            burried = 1;
                            if (pass == 0) { if ((double)we[a] < 2.8) n_apol++; else n_pol++; }
                        }
                    }
                    j++;
                }
                if (!burried && !vrefburried) nnburried++;
            }
            float area = (float)((4 * 3.1415926535897932384626433832795 * ((double)crad + probe) * ((double)crad + probe) / (double)(float)100) * (double)nnburried);
            if (pass == 0) {
                abpatmp[i] = area;
                if ((double)in->electroneg[cura] < 2.8) d->surf_apol_vdw14 = d->surf_apol_vdw14 + area;
                else d->surf_pol_vdw14 = d->surf_pol_vdw14 + area;
                d->surf_vdw14 = d->surf_vdw14 + area;
                atom_fx[4 * cura + 3] = n_apol / (n_apol + n_pol);
            } else {
                if ((double)in->electroneg[cura] < 2.8) d->surf_apol_vdw22 = d->surf_apol_vdw22 + area;
                else {
                    if (area < abpatmp[i] && (double)abpatmp[i] <= 10.0 && (double)area >= 0.0) {
                        d->n_abpa += 1;
                        atom_fx[4 * cura + 0] = 1.0f;
                        atom_fx[4 * cura + 1] = abpatmp[i];
                        atom_fx[4 * cura + 2] = area - abpatmp[i];
                    }
                    d->surf_pol_vdw22 = d->surf_pol_vdw22 + area;
                }
                d->surf_vdw22 = d->surf_vdw22 + area;
            }
        }
    }
    free(sa); free(ua); free(abpatmp);
}

static int g_force_observable = 0;      /* dpocket's explicit pocket always gets its volume and hull (dpocket.c:350) */
/* descriptors.c:158-265 set_descriptors + :355-436 set_atom_based_descriptors + :453-513 set_aa_desc */
static void set_descriptors(const fpk_structure *in, const fpk_params *p, const sph_t *S, const int *idx, int nvert,
                            const int *atoms, int natoms, fpk_desc *d, float *atom_fx, const float *pts,
                            int rng_mode, int cluster)
{
    /* atom based */
    if (natoms > 1) {
        int *res_ids = malloc(sizeof(int) * (size_t)natoms), nb_res_ids = 0, nb_polar_atm = 0;
        for (int i = 0; i < natoms; i++) {
            int a = atoms[i], seen = 0;
            for (int k = 0; k < nb_res_ids; k++) if (res_ids[k] == in->res_id[a]) { seen = 1; break; }
            if (!seen) {
                int aa = in->aa_idx[a];
                if (aa >= 0) {
                    d->aa_compo[aa]++;
                    d->hydrophobicity_score += AA_HYD[aa];
                    d->polarity_score += AA_POL[aa];
                    d->volume_score += AA_VOL[aa];
                    d->charge_score += AA_CHG[aa];
                }
                res_ids[nb_res_ids++] = in->res_id[a];
            }
            d->flex += in->bfactor[a];
            if ((double)in->electroneg[a] > 2.7) nb_polar_atm += 1;
        }
        d->hydrophobicity_score = d->hydrophobicity_score / (float)nb_res_ids;
        d->volume_score = d->volume_score / (float)nb_res_ids;
        d->flex /= natoms;
        d->prop_polar_atm = (float)((double)(((float)nb_polar_atm) / ((float)natoms)) * 100.0);
        free(res_ids);
    } else {
        d->hydrophobicity_score = 0.0f; d->volume_score = 0.0f; d->flex = 0.0f; d->prop_polar_atm = 0.0f;
    }
    /* sphere based */
    float dd, vx, vy, vz, vrad, masph_sacc = 0.0f, mean_ashape_radius = 0.0f, as_density = 0.0f, as_max_dst = -1.0f, dtmp;
    int napol_neigh, nAlphaApol = 0;
    float as_max_r = -1.0f;
    d->mean_loc_hyd_dens = 0.0f;
    for (int i = 0; i < nvert; i++) {
        const sph_t *vcur = S + idx[i];
        if (vcur->ray > as_max_r) as_max_r = vcur->ray;
        vx = vcur->x; vy = vcur->y; vz = vcur->z; vrad = vcur->ray;
        if (vcur->type == 0) {
            napol_neigh = 0;
            for (int j = 0; j < nvert; j++) {
                const sph_t *vc = S + idx[j];
                if (vc != vcur && vc->type == 0 &&
                    (double)(f_dist(vx, vy, vz, vc->x, vc->y, vc->z) - (vc->ray + vrad)) <= 0.) napol_neigh += 1;
                if (j > i) {
                    dtmp = f_dist(vcur->x, vcur->y, vcur->z, vc->x, vc->y, vc->z);
                    if (dtmp > as_max_dst) as_max_dst = dtmp;
                    as_density += dtmp;
                }
            }
            d->mean_loc_hyd_dens += (float)napol_neigh;
            nAlphaApol += 1;
        } else {
            for (int j = i + 1; j < nvert; j++) {
                const sph_t *vc = S + idx[j];
                dtmp = f_dist(vcur->x, vcur->y, vcur->z, vc->x, vc->y, vc->z);
                if (dtmp > as_max_dst) as_max_dst = dtmp;
                as_density += dtmp;
            }
        }
        mean_ashape_radius += vcur->ray;
        dd = f_dist(vcur->x, vcur->y, vcur->z, vcur->bary[0], vcur->bary[1], vcur->bary[2]);
        masph_sacc += dd / vcur->ray;
    }
    if (nAlphaApol > 0) d->mean_loc_hyd_dens /= (float)nAlphaApol;
    else d->mean_loc_hyd_dens = 0.0f;
    if (p->do_asa_and_volume) {
        set_asa(in, S, idx, nvert, d, atom_fx, pts);
        /* A cluster with fewer than min_pock_nb_asph spheres is always dropped (refine.c:124) and nothing of it is ever
         * written, so the product (ORC_RNG_PHILOX semantics) does not spend 30,000 Monte-Carlo samples and a hull on it:
         * volume = convex_hull_volume = 0 there. ORC_RNG_LIBC keeps the reference's behaviour exactly (it must: the libc
         * rand() stream is consumed by every cluster, fpocket.c:183), which is what pins this file to the reference. */
        int observable = g_force_observable || rng_mode == ORC_RNG_LIBC || nvert >= p->min_pock_nb_asph;
        d->volume = observable ? mc_volume(S, idx, nvert, p->nb_mcv_iter, -1.6f, rng_mode, p->mc_seed, cluster) : 0.0f;
        d->convex_hull_volume = observable ? hull_volume(S, idx, nvert) : 0.0f;
    }
    d->as_max_dst = as_max_dst;
    d->apolar_asphere_prop = (float)nAlphaApol / (float)nvert;
    d->masph_sacc = masph_sacc / nvert;
    d->mean_asph_ray = mean_ashape_radius / (float)nvert;
    d->nb_asph = nvert;
    d->as_density = (float)((double)as_density / ((nvert * nvert - nvert) * 0.5));
    d->as_max_r = as_max_r;
}

/* pocket.c:664-785 set_normalized_descriptors */
static void normalize(fpk_desc *D, int np)
{
    if (np <= 0) return;
    float flex_M = 0, flex_m = 1, nas_apolp_M = 0, nas_apolp_m = 1, density_M = 0, density_m = 100, mlhd_M = 0, mlhd_m = 1000,
          as_max_dst_M = 0, as_max_dst_m = 1000;
    int nas_M = 0, nas_m = 1000, polarity_M = -1, polarity_m = 10000;
    for (int i = 0; i < np; i++) {
        fpk_desc *d = D + i;
        if (i == 0) {
            as_max_dst_M = as_max_dst_m = d->as_max_dst; density_M = density_m = d->as_density;
            polarity_M = polarity_m = d->polarity_score; flex_M = flex_m = d->flex;
            nas_apolp_M = nas_apolp_m = d->apolar_asphere_prop; mlhd_M = mlhd_m = d->mean_loc_hyd_dens; nas_M = nas_m = d->nb_asph;
        } else {
            if (d->as_max_dst > as_max_dst_M) as_max_dst_M = d->as_max_dst; else if (d->as_max_dst < as_max_dst_m) as_max_dst_m = d->as_max_dst;
            if (d->as_density > density_M) density_M = d->as_density; else if (d->as_density < density_m) density_m = d->as_density;
            if (d->polarity_score > polarity_M) polarity_M = d->polarity_score; else if (d->polarity_score < polarity_m) polarity_m = d->polarity_score;
            if (d->mean_loc_hyd_dens > mlhd_M) mlhd_M = d->mean_loc_hyd_dens; else if (d->mean_loc_hyd_dens < mlhd_m) mlhd_m = d->mean_loc_hyd_dens;
            if (d->flex > flex_M) flex_M = d->flex; else if (d->flex < flex_m) flex_m = d->flex;
            if (d->nb_asph > nas_M) nas_M = d->nb_asph; else if (d->nb_asph < nas_m) nas_m = d->nb_asph;
            if (d->apolar_asphere_prop > nas_apolp_M) nas_apolp_M = d->apolar_asphere_prop; else if (d->apolar_asphere_prop < nas_apolp_m) nas_apolp_m = d->apolar_asphere_prop;
        }
    }
    if (np > 1) {
        for (int i = 0; i < np; i++) {
            fpk_desc *d = D + i;
            if ((double)(as_max_dst_M - as_max_dst_m) != 0.0) d->as_max_dst_norm = (d->as_max_dst - as_max_dst_m) / (as_max_dst_M - as_max_dst_m);
            if ((double)(density_M - density_m) != 0.0) d->as_density_norm = (d->as_density - density_m) / (density_M - density_m);
            if ((double)(polarity_M - polarity_m) != 0.0) d->polarity_score_norm = (float)(d->polarity_score - polarity_m) / (float)(polarity_M - polarity_m);
            if ((double)(mlhd_M - mlhd_m) != 0.0) d->mean_loc_hyd_dens_norm = (d->mean_loc_hyd_dens - mlhd_m) / (mlhd_M - mlhd_m);
            if ((double)(flex_M - flex_m) != 0.0) d->flex = (d->flex - flex_m) / (flex_M - flex_m);
            if ((double)(nas_M - nas_m) != 0.0) d->nas_norm = (float)(d->nb_asph - nas_m) / (float)(nas_M - nas_m);
            if ((double)(nas_apolp_M - nas_apolp_m) != 0.0) d->prop_asapol_norm = (d->apolar_asphere_prop - nas_apolp_m) / (nas_apolp_M - nas_apolp_m);
        }
    } else {
        fpk_desc *d = D;
        d->polarity_score_norm = 0.0f; d->as_density_norm = 0.0f; d->as_max_dst_norm = 0.0f;
        d->mean_loc_hyd_dens_norm = (float)(((double)d->mean_loc_hyd_dens - 8.23) / (24.20 - 8.23));
        d->flex = 0.0f; d->nas_norm = 0.0f; d->prop_asapol_norm = 0.0f;
    }
}

/* pscoring.c:241-247 and :325-329 */
static float score_pocket(const fpk_desc *d)
{
    return (float)(-0.03783394 + 0.48461469 * (double)d->nas_norm + 0.09093926 * (double)d->as_density +
                   0.0004155899 * (double)d->convex_hull_volume - 0.003995233 * (double)d->surf_pol_vdw14 -
                   0.004072336 * (double)d->surf_apol_vdw14);
}
static float drug_score_pocket(const fpk_desc *d)
{
    float b0 = (float)-9.5698768, b1 = (float)7.479844, b2 = (float)0.3696134, b3 = (float)-0.04671833;
    return (float)(1.0 / (1.0 + exp((double)(-(b0 + b1 * d->mean_loc_hyd_dens_norm + b2 * d->as_max_dst + b3 * d->surf_pol_vdw22)))));
}

/* psorting.c:64-165: first-element-pivot quicksort; fcmp(c,piv)==-1 <=> !(c.score < piv.score) */
static int q_partition(int *pocks, const fpk_desc *D, int start, int end)
{
    int c = start, piv = pocks[start], tmp;
    for (int i = start + 1; i <= end; i++) {
        int cp = pocks[i];
        if (!(D[cp].score < D[piv].score)) {
            c++;
            if (c != i) { tmp = pocks[c]; pocks[c] = pocks[i]; pocks[i] = tmp; }
        }
    }
    if (c != start) { tmp = pocks[c]; pocks[c] = pocks[start]; pocks[start] = tmp; }
    return c;
}
static void q_sort(int *pocks, const fpk_desc *D, int start, int end)
{
    if (start < end) {
        int piv = q_partition(pocks, D, start, end);
        q_sort(pocks, D, start, piv - 1);
        q_sort(pocks, D, piv + 1, end);
    }
}

void orc_result_free(fpk_result *r)
{
    free(r->tets); free(r->sph_xyzr); free(r->sph_bary); free(r->sph_neigh); free(r->sph_type); free(r->sph_cluster);
    free(r->cl_off); free(r->cl_sph); free(r->cl_atom_off); free(r->cl_atoms); free(r->desc); free(r->rank_to_cluster); free(r->atom_fx);
    free(r->xspheres); free(r->iface_atoms); free(r->pk_ovlp); free(r->xdesc);
    memset(r, 0, sizeof *r);
}

/* ---------------------------------------- mdpocket characterize mode ---------------------------------------- */
/* mdpocket.c:436-460 (one snapshot of mdpocket_characterize): after search_pocket, extract_wanted_vertices (mdpocket.c:695-745) gathers the spheres of the
 * SURVIVING pockets, in rank order and pocket order, once for EVERY point of the wanted zone within (float)M_MDP_CUBE_SIDE / 2.0 = 1.0 of the centre
 * (dist() <= 1.0: a sphere near three points is listed three times); get_pocket_contacted_atms + set_descriptors then describe that list as one pocket
 * (no normalisation, no score), with the side effects of ASA on the snapshot's atoms. The zone is handed in before the call; the record is read back after. */
static const float *g_want_xyz = NULL;
static int g_want_n = 0;
static int *g_want_idx = NULL, g_want_nidx = 0, *g_want_atoms = NULL, g_want_natoms = 0;
static fpk_desc g_want_desc;
void orc_set_wanted_zone(const float *xyz, int n) { g_want_xyz = xyz; g_want_n = xyz ? n : 0; }
int orc_get_wanted(const int **idx, int *nidx, const int **atoms, int *natoms, fpk_desc *d)
{
    *idx = g_want_idx; *nidx = g_want_nidx; *atoms = g_want_atoms; *natoms = g_want_natoms; *d = g_want_desc;
    return g_want_nidx;
}

int orc_search_pocket(const fpk_structure *in, const fpk_params *p, const int *tets_in, int ntets, int rng_mode, fpk_result *out)
{
    memset(out, 0, sizeof *out);
    int natoms = in->natoms;
    int *site2atom = malloc(sizeof(int) * (size_t)(natoms + 1));
    long long *ki = malloc(sizeof(long long) * 3 * (size_t)(natoms + 1));
    double *xd = malloc(sizeof(double) * 3 * (size_t)(natoms + 1));
    int ns_sites = orc_sites(in, site2atom, ki, xd);
    out->n_sites = ns_sites;
    int *own_tets = NULL;
    const int *tets = tets_in;
    if (!tets) {
        int rc = orc_delaunay(ns_sites, ki, &own_tets, &ntets, NULL);
        if (rc) { out->status = rc; free(site2atom); free(ki); free(xd); return rc; }
        tets = own_tets;
    }
    out->n_tets = ntets;
    if (p->want_tets) { out->tets = malloc(sizeof(int) * 4 * (size_t)(ntets + 1)); memcpy(out->tets, tets, sizeof(int) * 4 * (size_t)ntets); }

    sph_t *S = NULL;
    int ns = filter_tets(in, p, site2atom, xd, tets, ntets, &S);
    free(own_tets); free(ki); free(xd); free(site2atom);
    out->n_spheres = ns;
    out->atom_fx = calloc((size_t)natoms * 4 + 4, sizeof(float));
    if ((!p->skip_clustering && ns < 2) || ns == 0) {          /* clusterlib.c:3969 -> fpocket.c:135-139; no sphere at all with -r / -P: assign_pockets
                                                                   returns NULL (pocket.c:77), search_pocket returns NULL and nothing is written */
        free(S);
        out->status = FPK_ERR_NO_SPHERES;
        return out->status;
    }
    const int saved_measure = g_distance_measure;
    int nc = cluster_spheres(S, ns, p->clust_max_dist, p->skip_clustering, p->distance_measure, p->clustering_method);
    g_distance_measure = saved_measure;
    out->n_clusters = nc;
    out->sph_xyzr = malloc(sizeof(float) * 4 * (size_t)(ns + 1)); out->sph_bary = malloc(sizeof(float) * 3 * (size_t)(ns + 1));
    out->sph_neigh = malloc(sizeof(int) * 4 * (size_t)(ns + 1)); out->sph_type = malloc((size_t)ns + 1); out->sph_cluster = malloc(sizeof(int) * (size_t)(ns + 1));
    for (int i = 0; i < ns; i++) {
        out->sph_xyzr[4 * i] = S[i].x; out->sph_xyzr[4 * i + 1] = S[i].y; out->sph_xyzr[4 * i + 2] = S[i].z; out->sph_xyzr[4 * i + 3] = S[i].ray;
        for (int k = 0; k < 3; k++) out->sph_bary[3 * i + k] = S[i].bary[k];
        for (int k = 0; k < 4; k++) out->sph_neigh[4 * i + k] = S[i].neigh[k];
        out->sph_type[i] = (uint8_t)S[i].type; out->sph_cluster[i] = S[i].cluster;
    }
    /* pocket.c:77 assign_pockets + refine.c:58 apply_clustering: members ascending, clusters by first member */
    out->cl_off = calloc((size_t)nc + 2, sizeof(int)); out->cl_sph = malloc(sizeof(int) * (size_t)(ns + 1));
    for (int i = 0; i < ns; i++) out->cl_off[S[i].cluster + 1]++;
    for (int c = 0; c < nc; c++) out->cl_off[c + 1] += out->cl_off[c];
    int *fill = calloc((size_t)nc + 1, sizeof(int));
    for (int i = 0; i < ns; i++) { int c = S[i].cluster; out->cl_sph[out->cl_off[c] + fill[c]++] = i; }
    free(fill);
    out->desc = calloc((size_t)nc + 1, sizeof(fpk_desc));
    out->cl_atom_off = calloc((size_t)nc + 2, sizeof(int)); out->cl_atoms = malloc(sizeof(int) * 4 * (size_t)(ns + 1));
    float pts[300];
    orc_spiral_points(pts);
    if (rng_mode == ORC_RNG_LIBC) srand((unsigned)p->mc_seed);
    int na_tot = 0;
    for (int c = 0; c < nc; c++) {
        fpk_desc *d = out->desc + c;
        const int *idx = out->cl_sph + out->cl_off[c];
        int n = out->cl_off[c + 1] - out->cl_off[c];
        d->first_sphere = idx[0];
        /* refine.c:192-240 reIndexPockets barycentre; pocket.c:1071 apol/pol counts */
        float ps[3] = {0, 0, 0};
        for (int i = 0; i < n; i++) {
            ps[0] += S[idx[i]].x; ps[1] += S[idx[i]].y; ps[2] += S[idx[i]].z;
            if (S[idx[i]].type == 0) d->nAlphaApol++; else d->nAlphaPol++;
        }
        d->bary[0] = ps[0] / (float)n; d->bary[1] = ps[1] / (float)n; d->bary[2] = ps[2] / (float)n;
        /* pocket.c:1280-1320 unique contacted atoms by id, first-seen order */
        int *atoms = out->cl_atoms + na_tot, na = 0;
        for (int i = 0; i < n; i++)
            for (int k = 0; k < 4; k++) {
                int a = S[idx[i]].neigh[k], seen = 0;
                for (int q = 0; q < na; q++) if (in->atom_id[atoms[q]] == in->atom_id[a]) { seen = 1; break; }
                if (!seen) atoms[na++] = a;
            }
        d->n_atoms = na;
        na_tot += na;
        out->cl_atom_off[c + 1] = na_tot;
        set_descriptors(in, p, S, idx, n, atoms, na, d, out->atom_fx, pts, rng_mode, c);
    }
    normalize(out->desc, nc);
    for (int c = 0; c < nc; c++) {
        out->desc[c].score = score_pocket(out->desc + c);
        out->desc[c].drug_score = drug_score_pocket(out->desc + c);
    }
    /* refine.c:114-145 dropSmallNpolarPockets (note the C-boolean right-hand side) */
    int *keep = malloc(sizeof(int) * (size_t)(nc + 1)), nk = 0;
    for (int c = 0; c < nc; c++) {
        fpk_desc *d = out->desc + c;
        double pasph = (float)((float)d->nAlphaApol / (float)d->nb_asph);
        int rhs = (p->refine_min_apolar_asphere_prop != 0.0f) || (d->as_density < p->min_as_density);
        if (d->nb_asph < p->min_pock_nb_asph || pasph < (double)rhs) continue;
        keep[nk++] = c;
    }
    if (nk > 0) q_sort(keep, out->desc, 0, nk - 1);
    out->n_pockets = nk;
    out->rank_to_cluster = keep;
    for (int k = 0; k < nk; k++) out->desc[keep[k]].final_rank = k + 1;
    out->status = nk > 0 ? FPK_OK : FPK_ERR_NO_POCKETS;
    if (in->n_lig > 0 && !p->skip_clustering) {
        orc_ligand_queries(in, p, out);
        if (out->n_xspheres > 0) {
            /* dpocket.c:341-350 get_explicit_desc: set_descriptors(interface atoms, explicit-pocket spheres) - one more record, computed
             * like a pocket's (canonical = ascending order of both lists) but from the interface atoms, never normalised or scored */
            const int *idx = out->xspheres, n = out->n_xspheres;
            fpk_desc *d = out->xdesc = calloc(1, sizeof(fpk_desc));
            d->first_sphere = idx[0];
            float ps[3] = {0, 0, 0};
            for (int i = 0; i < n; i++) {
                ps[0] += S[idx[i]].x; ps[1] += S[idx[i]].y; ps[2] += S[idx[i]].z;
                if (S[idx[i]].type == 0) d->nAlphaApol++; else d->nAlphaPol++;
            }
            d->bary[0] = ps[0] / (float)n; d->bary[1] = ps[1] / (float)n; d->bary[2] = ps[2] / (float)n;
            int *atoms = malloc(sizeof(int) * 4 * (size_t)(n + 1)), na = 0;
            for (int i = 0; i < n; i++)
                for (int k = 0; k < 4; k++) {
                    int a = S[idx[i]].neigh[k], seen = 0;
                    for (int q = 0; q < na; q++) if (in->atom_id[atoms[q]] == in->atom_id[a]) { seen = 1; break; }
                    if (!seen) atoms[na++] = a;
                }
            d->n_atoms = na;
            free(atoms);
            float *fx = calloc((size_t)natoms * 4 + 4, sizeof(float));         /* the explicit call's side effects reach no output */
            g_force_observable = 1;
            set_descriptors(in, p, S, idx, n, out->iface_atoms, out->n_iface, d, fx, pts, rng_mode, nc);
            g_force_observable = 0;
            free(fx);
        }
    }
    free(g_want_idx); free(g_want_atoms); g_want_idx = g_want_atoms = NULL; g_want_nidx = g_want_natoms = 0;
    memset(&g_want_desc, 0, sizeof g_want_desc);
    if (g_want_n > 0 && nk > 0) {
        int cap = 64, n = 0;
        int *idx = malloc(sizeof(int) * (size_t)cap);
        for (int k = 0; k < nk; k++) {
            const int c = keep[k];
            for (int i = out->cl_off[c]; i < out->cl_off[c + 1]; i++) {
                const sph_t *v = S + out->cl_sph[i];
                for (int z = 0; z < g_want_n; z++)
                    if ((double)f_dist(g_want_xyz[3 * z], g_want_xyz[3 * z + 1], g_want_xyz[3 * z + 2], v->x, v->y, v->z) <= (double)(float)2.0 / 2.0) {
                        if (n == cap) { cap *= 2; idx = realloc(idx, sizeof(int) * (size_t)cap); }
                        idx[n++] = out->cl_sph[i];
                    }
            }
        }
        g_want_idx = idx; g_want_nidx = n;
        if (n > 0) {
            int *atoms = malloc(sizeof(int) * 4 * (size_t)(n + 1)), na = 0;
            for (int i = 0; i < n; i++)
                for (int k = 0; k < 4; k++) {
                    int a = S[idx[i]].neigh[k], seen = 0;
                    for (int q = 0; q < na; q++) if (in->atom_id[atoms[q]] == in->atom_id[a]) { seen = 1; break; }
                    if (!seen) atoms[na++] = a;
                }
            g_want_atoms = atoms; g_want_natoms = na;
            g_want_desc.n_atoms = na;
            g_want_desc.first_sphere = idx[0];
            g_force_observable = 1;                   /* the record of the zone is written for every snapshot: always with volume and hull */
            set_descriptors(in, p, S, idx, n, atoms, na, &g_want_desc, out->atom_fx, pts, rng_mode, nc + 1);
            g_force_observable = 0;
        }
    }
    free(S);
    return out->status;
}

/* ---------------------------------------- dpocket ligand queries ---------------------------------------- */
/* neighbor.c:192 get_mol_vert_neigh, :313 get_mol_ctd_atm_neigh(M_INTERFACE_SEARCH), :490 count_pocket_lig_vert_ovlp,
 * :598 count_atm_prop_vert_neigh, atom.c:264 atm_corsp, dpocket.c:246-270. The reference finds the pairs by sweeping an x-sorted
 * merged array outwards from every ligand atom until the x-gap exceeds the criterion; what it returns are the SETS / counts of the
 * strict relation dist() < d, restated here directly (pinned against the reference's own functions by tests/golden, the _dpocket fixture). */
static int g_interface_method = 1;       /* dparams.h:41-44; fpk_params has no field for it yet: oracle first, like the distance measure */
void orc_set_interface_method(int m) { g_interface_method = m; }
void orc_ligand_queries(const fpk_structure *in, const fpk_params *p, fpk_result *out)
{
    const int nl = in->n_lig, ns = out->n_spheres, na = in->natoms, np = out->n_pockets;
    const float *wx = in->natoms_w ? in->wx : in->x, *wy = in->natoms_w ? in->wy : in->y, *wz = in->natoms_w ? in->wz : in->z;
    const float dcrit = p->lig_interface_dist, ccrit = p->lig_crit_dist;
    unsigned char *xf = calloc((size_t)ns + 1, 1), *af = calloc((size_t)na + 1, 1);
    for (int s = 0; s < ns; s++) {
        const float *v = out->sph_xyzr + 4 * s;
        for (int l = 0; l < nl; l++) {
            int a = in->lig_atoms[l];
            if (f_dist(v[0], v[1], v[2], wx[a], wy[a], wz[a]) < dcrit) {
                xf[s] = 1;
                for (int j = 0; j < 4; j++) {
                    int c = out->sph_neigh[4 * s + j];
                    if ((double)f_dist(in->x[c], in->y[c], in->z[c], wx[a], wy[a], wz[a]) < 8.0) af[c] = 1;     /* neighbor.h:30 */
                }
            }
        }
    }
    if (g_interface_method == 2 || p->lig_interface_method == 2) {
        /* dparams.h:44 M_INTERFACE_METHOD2 (dpocket.c:331-336, neighbor.c:67 get_mol_atm_neigh): the atoms of the complex with dist() < dcrit to a ligand
         * atom, the ligand's own atoms excluded. pdb is pdb_w_lig without the ligand, in the same order: index in pdb = index in pdb_w_lig minus the
         * ligand atoms before it. (The reference excludes and de-duplicates by atom ID, neighbor.c:118: with repeated ids in a file its result is an
         * order-dependent subset; this is the relation itself, as for the spheres above.) */
        const int nw = in->natoms_w ? in->natoms_w : in->natoms;
        unsigned char *isl = calloc((size_t)nw + 1, 1);
        for (int l = 0; l < nl; l++) isl[in->lig_atoms[l]] = 1;
        memset(af, 0, (size_t)na + 1);
        int pi = 0;
        for (int a = 0; a < nw; a++) {
            if (isl[a]) continue;
            if (pi < na)
                for (int l = 0; l < nl; l++) {
                    int b = in->lig_atoms[l];
                    if (f_dist(wx[a], wy[a], wz[a], wx[b], wy[b], wz[b]) < dcrit) { af[pi] = 1; break; }
                }
            pi++;
        }
        free(isl);
    }
    out->xspheres = malloc(sizeof(int) * ((size_t)ns + 1));
    out->iface_atoms = malloc(sizeof(int) * ((size_t)na + 1));
    out->n_xspheres = out->n_iface = 0;
    for (int s = 0; s < ns; s++) if (xf[s]) out->xspheres[out->n_xspheres++] = s;
    for (int a = 0; a < na; a++) if (af[a]) out->iface_atoms[out->n_iface++] = a;
    free(xf); free(af);
    float *o = malloc(sizeof(float) * (4 * (size_t)np + 4));
    out->pk_ovlp = o; out->pk_dst = o + np; out->pk_c4 = o + 2 * np; out->pk_c5 = o + 3 * np;
    int nmol = 0;
    for (int l = 0; l < nl; l++) if (in->lig_mol[l] + 1 > nmol) nmol = in->lig_mol[l] + 1;
    int *hit = malloc(sizeof(int) * ((size_t)nmol + 1)), *tot = malloc(sizeof(int) * ((size_t)nmol + 1));
    for (int r = 0; r < np; r++) {
        const int c = out->rank_to_cluster[r], off = out->cl_off[c], n = out->cl_off[c + 1] - off;
        const fpk_desc *d = out->desc + c;
        int c5n = 0;
        for (int i = 0; i < n; i++) {
            const float *v = out->sph_xyzr + 4 * out->cl_sph[off + i];
            for (int l = 0; l < nl; l++) {
                int a = in->lig_atoms[l];
                if (f_dist(v[0], v[1], v[2], wx[a], wy[a], wz[a]) < ccrit) { c5n++; break; }
            }
        }
        for (int m = 0; m < nmol; m++) hit[m] = tot[m] = 0;
        float dst = 0.0f;
        for (int l = 0; l < nl; l++) {
            int a = in->lig_atoms[l], m = in->lig_mol[l];
            tot[m]++;
            for (int i = 0; i < n; i++) {
                const float *v = out->sph_xyzr + 4 * out->cl_sph[off + i];
                if (f_dist(v[0], v[1], v[2], wx[a], wy[a], wz[a]) < ccrit) { hit[m]++; break; }
            }
            float t = f_dist(wx[a], wy[a], wz[a], d->bary[0], d->bary[1], d->bary[2]);
            if (l == 0) dst = t;
            else if (t < dst) dst = t;
        }
        float c4 = 0.0f;
        for (int m = 0; m < nmol; m++) { float t = (float)hit[m] / (float)tot[m]; if (t > c4) c4 = t; }
        const int *pa = out->cl_atoms + out->cl_atom_off[c], npa = out->cl_atom_off[c + 1] - out->cl_atom_off[c];
        int found = 0;
        for (int i = 0; i < out->n_iface; i++) {
            int a = out->iface_atoms[i];
            for (int j = 0; j < npa; j++) {
                int b = pa[j];
                if (in->res_id[b] == in->res_id[a] && (in->atom_id[b] == in->atom_id[a] || (in->atom_name_key && in->atom_name_key[b] == in->atom_name_key[a]))) { found++; break; }
            }
        }
        out->pk_ovlp[r] = (out->n_iface <= 0 || npa <= 0) ? 0.0f : (float)((double)(((float)found) / ((float)out->n_iface)) * 100.0);
        out->pk_dst[r] = dst; out->pk_c4[r] = c4; out->pk_c5[r] = (float)c5n / (float)n;
    }
    free(hit); free(tot);
    /* atom.c:156 get_mol_volume_ptr(ligand atoms, nb_mcv_iter) (dpocket.c:245): box with the else-if updates, first-hit loop; the reference's
     * draws come from time-seeded rand(), here from Philox: counter (sample, 0x11A, 0, 0), key mc_seed (the library's stream) */
    {
        float xmin = 0, xmax = 0, ymin = 0, ymax = 0, zmin = 0, zmax = 0, t;
        const float *wr = in->natoms_w ? in->wradius : in->radius;
        for (int l = 0; l < nl; l++) {
            const int a = in->lig_atoms[l];
            if (l == 0) { xmin = wx[a] - wr[a]; xmax = wx[a] + wr[a]; ymin = wy[a] - wr[a]; ymax = wy[a] + wr[a]; zmin = wz[a] - wr[a]; zmax = wz[a] + wr[a]; }
            else {
                if (xmin > (t = wx[a] - wr[a])) xmin = t; else if (xmax < (t = wx[a] + wr[a])) xmax = t;
                if (ymin > (t = wy[a] - wr[a])) ymin = t; else if (ymax < (t = wy[a] + wr[a])) ymax = t;
                if (zmin > (t = wz[a] - wr[a])) zmin = t; else if (zmax < (t = wz[a] + wr[a])) zmax = t;
            }
        }
        const float vbox = (xmax - xmin) * (ymax - ymin) * (zmax - zmin);
        int nb_in = 0;
        for (int i = 0; i < p->nb_mcv_iter; i++) {
            unsigned o[4];
            orc_philox4x32((unsigned)i, 0x11Au, 0u, 0u, (unsigned)p->mc_seed, (unsigned)(p->mc_seed >> 32), o);
            const float xr = unif_from_bits(o[0] >> 1, xmin, xmax), yr = unif_from_bits(o[1] >> 1, ymin, ymax), zr = unif_from_bits(o[2] >> 1, zmin, zmax);
            for (int l = 0; l < nl; l++) {
                const int a = in->lig_atoms[l];
                const float xt = wx[a] - xr, yt = wy[a] - yr, zt = wz[a] - zr;
                if ((wr[a] * wr[a]) > (xt * xt + yt * yt + zt * zt)) { nb_in++; break; }
            }
        }
        out->lig_volume = (nl > 0 && p->nb_mcv_iter > 0) ? ((float)nb_in) / ((float)p->nb_mcv_iter) * vbox : 0.0f;
    }
}

/* ---------------------------------------- mdpocket grids ---------------------------------------- */
/* mdpbase.c:403-416 float_get_min_max_from_pockets over ALL kept spheres + :444-502 init_md_grid */
void orc_md_grid_geometry(const fpk_result *r, float origin[3], int dims[3])
{
    float mn[3] = {50000.f, 50000.f, 50000.f}, mx[3] = {-50000.f, -50000.f, -50000.f};
    for (int z = 0; z < r->n_spheres; z++)
        for (int k = 0; k < 3; k++) {
            float v = r->sph_xyzr[4 * z + k];
            if (mn[k] > v) mn[k] = v;
            else if (mx[k] < v) mx[k] = v;
        }
    float span = (float)(2.0 / 2.0), res = (float)1.0;
    for (int k = 0; k < 3; k++) {
        dims[k] = (int)((float)1 + (float)((int)((double)mx[k] + 30. * (double)span - (double)mn[k])) / res);   /* cast binds before the division, mdpbase.c:475-477 */
        origin[k] = (float)((double)mn[k] - 15. * (double)span);
    }
}

static int md_idx(int o, float v, float org, float res, int ctr)
{
    if (o < 0) return (int)ceil((double)(((float)o + v - org) / res));
    if (o > 0) return (int)floor((double)(((float)o + v - org) / res));
    return ctr;
}

/* mdpbase.c:108-169 calculate_md_dens_grid and :188-262 update_md_grid (refg cleared first, :275) */
void orc_md_scatter(const fpk_result *r, int use_drug_score, const float origin[3], const int dims[3], float *dens, float *freq, float *refg)
{
    const int nx = dims[0], ny = dims[1], nz = dims[2];
    const float res = 1.0f;
    memset(refg, 0, sizeof(float) * (size_t)nx * ny * nz);
    for (int pass = 0; pass < 2; pass++) {
        for (int k = 0; k < r->n_pockets; k++) {
            int c = r->rank_to_cluster[k];
            float incre = use_drug_score ? r->desc[c].drug_score : 1.0f;
            int xidx = 0, yidx = 0, zidx = 0;
            for (int q = r->cl_off[c]; q < r->cl_off[c + 1]; q++) {
                int si = r->cl_sph[q];
                float vx = r->sph_xyzr[4 * si], vy = r->sph_xyzr[4 * si + 1], vz = r->sph_xyzr[4 * si + 2], ray = r->sph_xyzr[4 * si + 3];
                int s = pass == 0 ? (int)(((float)2.0 / 2.0) / res) : (int)(((double)ray / 2.0) / (double)res);
                xidx = (int)roundf((vx - origin[0]) / res); yidx = (int)roundf((vy - origin[1]) / res); zidx = (int)roundf((vz - origin[2]) / res);
                for (int sxi = -s; sxi <= s; sxi++) for (int syi = -s; syi <= s; syi++) for (int szi = -s; szi <= s; szi++) {
                    if ((sxi * sxi + syi * syi + szi * szi) == 0) continue;
                    int sx = md_idx(sxi, vx, origin[0], res, xidx), sy = md_idx(syi, vy, origin[1], res, yidx), sz = md_idx(szi, vz, origin[2], res, zidx);
                    if (((sx != xidx) || (sy != yidx) || (sz != zidx)) && (sx >= 0 && sy >= 0 && sz >= 0 && sx < nx && sy < ny && sz < nz)) {
                        size_t g = ((size_t)sx * ny + sy) * nz + sz;
                        if (pass == 0) dens[g] += incre;
                        else if (refg[g] < 1.0f) { freq[g] += incre; refg[g] = 1.0f; }
                    }
                }
            }
            if (xidx < nx && yidx < ny && zidx < nz && xidx >= 0 && yidx >= 0 && zidx >= 0) {
                size_t g = ((size_t)xidx * ny + yidx) * nz + zidx;
                if (pass == 0) dens[g] += incre;
                else if (refg[g] < 1.0f) { freq[g] += incre; refg[g] = 1.0f; }
            }
        }
    }
}
