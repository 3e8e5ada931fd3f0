/*
 * TEST INFRASTRUCTURE — not part of the product.
 *
 * ref_dump: a driver of OUR OWN that is linked against the UNMODIFIED reference objects
 * (see oracle/Makefile) and records what every stage of the reference hot path produces,
 * so that oracle.c (the restatement) and the CUDA path can be pinned to it.
 *
 * It calls the reference's public functions in the order search_pocket() does
 * (reference src/fpocket.c:58-225) but stops between the stages to write arrays:
 *
 *   atoms of pdb / pdb_w_lig as parsed by rpdb.c / read_mmcif.c       (kernel inputs)
 *   qvoronoi output (Voronoi centres + tets, qhull order)              (voronoi.c:111-138)
 *   kept alpha spheres after load_vvertices                            (voronoi.c:369-521)
 *   cluster id per sphere after treecluster/cuttree_distance           (fpocket.c:114-144)
 *   pockets + full s_desc before dropSmallNpolarPockets                (fpocket.c:163-183)
 *   final ranked pockets                                               (fpocket.c:191-213)
 *   per-atom ASA side effects                                          (asa.c:337,392-396)
 *   optional: mdpocket grids for this single frame                     (mdpbase.c:108-262,444-502)
 *
 *   optional: dpocket's ligand queries (--lig NAME: the structure is loaded as dpocket.c:186-203 does, and the
 *             reference's own neighbor.c / atom.c functions are called on the final pockets)            (dpocket.c:228-290)
 *
 * Usage:  ref_dump <out.dump> [--md] [--seed S] [--lig NAME] [--want ZONE.pdb] <fpocket options, e.g. -f X.pdb -m 3.0>
 *   --want  one snapshot of mdpocket's characterize mode (mdpocket.c:436-460) on the final pockets: spheres near the points of ZONE.pdb
 *   --md    mimic mdpocket's call (mdpocket.c:833-846): search_pocket(pdb,params,pdb),
 *           ASA/volume/hull off (mdpocket.c:89), then the three grids of frame 1.
 *
 * File format ("FPKDUMP1\n" then records): name[32] | dtype char | ndim int32 | dims int32[ndim] | raw data
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "fpocket.h"
#include "fpout.h"
#include "mdpbase.h"
#include "mdparams.h"
#include "mdpocket.h"
#include "read_mmcif.h"
#include "neighbor.h"
#include "dparams.h"
#include "tpocket.h"

extern int run_qvoronoi(FILE *fin, FILE *fout);

static FILE *OUT;

static void rec(const char *name, char dtype, int ndim, const int *dims, const void *data)
{
    char nm[32];
    memset(nm, 0, sizeof nm);
    strncpy(nm, name, 31);
    fwrite(nm, 1, 32, OUT);
    fwrite(&dtype, 1, 1, OUT);
    fwrite(&ndim, 4, 1, OUT);
    fwrite(dims, 4, ndim, OUT);
    size_t n = 1;
    for (int i = 0; i < ndim; i++) n *= (size_t)dims[i];
    size_t es = (dtype == 'd') ? 8 : (dtype == 'b') ? 1 : 4;
    if (n) fwrite(data, es, n, OUT);
}
static void rec1(const char *name, char dtype, int n, const void *data) { rec(name, dtype, 1, &n, data); }
static void rec_i(const char *name, int v) { rec1(name, 'i', 1, &v); }
static void rec_f(const char *name, float v) { rec1(name, 'f', 1, &v); }

static void dump_atoms(const char *pfx, s_pdb *pdb)
{
    int n = pdb->natoms, i;
    char nm[64];
    float *f = malloc(sizeof(float) * n * 6);
    int *iv = malloc(sizeof(int) * n * 4);
    unsigned char *sym = calloc(n, 4), *rn = calloc(n, 8), *ch = calloc(n, 4), *an = calloc(n, 8), *ty = calloc(n, 8), *ia = calloc(n, 2);
    for (i = 0; i < n; i++) {
        s_atm *a = pdb->latoms + i;
        f[i * 6 + 0] = a->x; f[i * 6 + 1] = a->y; f[i * 6 + 2] = a->z;
        f[i * 6 + 3] = a->radius; f[i * 6 + 4] = a->electroneg; f[i * 6 + 5] = a->bfactor;
        iv[i * 4 + 0] = a->id; iv[i * 4 + 1] = a->res_id; iv[i * 4 + 2] = a->charge; iv[i * 4 + 3] = a->atomic_num;
        memcpy(sym + i * 4, a->symbol, 3);
        memcpy(rn + i * 8, a->res_name, 8);
        memcpy(ch + i * 4, a->chain, 4);
        memcpy(an + i * 8, a->name, 5);
        memcpy(ty + i * 8, a->type, 7);
        ia[i * 2] = a->pdb_insert; ia[i * 2 + 1] = a->pdb_aloc;
    }
    int d2[2];
    d2[0] = n; d2[1] = 6; snprintf(nm, 64, "%s.f", pfx); rec(nm, 'f', 2, d2, f);        /* x y z radius electroneg bfactor */
    d2[1] = 4; snprintf(nm, 64, "%s.i", pfx); rec(nm, 'i', 2, d2, iv);                   /* id res_id charge atomic_num */
    d2[1] = 4; snprintf(nm, 64, "%s.symbol", pfx); rec(nm, 'b', 2, d2, sym);
    d2[1] = 8; snprintf(nm, 64, "%s.res_name", pfx); rec(nm, 'b', 2, d2, rn);
    d2[1] = 4; snprintf(nm, 64, "%s.chain", pfx); rec(nm, 'b', 2, d2, ch);
    d2[1] = 8; snprintf(nm, 64, "%s.name", pfx); rec(nm, 'b', 2, d2, an);
    d2[1] = 8; snprintf(nm, 64, "%s.type", pfx); rec(nm, 'b', 2, d2, ty);
    d2[1] = 2; snprintf(nm, 64, "%s.ins_aloc", pfx); rec(nm, 'b', 2, d2, ia);
    float bf[3] = {pdb->avg_bfactor, pdb->min_bfactor, pdb->max_bfactor};
    snprintf(nm, 64, "%s.bfstat", pfx); rec1(nm, 'f', 3, bf);
    snprintf(nm, 64, "%s.natoms_h", pfx); rec_i(nm, pdb->natoms_h);
    /* hetatm list (indices into latoms) — used by set_pocket_contacted_lig_name (pocket.c:535) */
    int *het = malloc(sizeof(int) * (pdb->nhetatm + 1));
    for (i = 0; i < pdb->nhetatm; i++) het[i] = (int)(pdb->lhetatm[i] - pdb->latoms);
    snprintf(nm, 64, "%s.hetatm", pfx); rec1(nm, 'i', pdb->nhetatm, het);
    if (pdb->n_xlig_atoms > 0) {
        float *xl = malloc(sizeof(float) * 3 * pdb->n_xlig_atoms);
        for (i = 0; i < pdb->n_xlig_atoms; i++) { xl[3 * i] = pdb->xlig_x[i]; xl[3 * i + 1] = pdb->xlig_y[i]; xl[3 * i + 2] = pdb->xlig_z[i]; }
        d2[0] = pdb->n_xlig_atoms; d2[1] = 3; snprintf(nm, 64, "%s.xlig", pfx); rec(nm, 'f', 2, d2, xl);
        free(xl);
    }
    free(f); free(iv); free(sym); free(rn); free(ch); free(an); free(ty); free(ia); free(het);
}

/* run qvoronoi on the heavy atoms exactly as load_vvertices feeds it (voronoi.c:111-138) and keep its raw answer */
static void dump_qvoronoi(s_pdb *pdb)
{
    char fin[256], fout[256], line[256];
    snprintf(fin, 256, "/tmp/refdump_in_%d.dat", (int)getpid());
    snprintf(fout, 256, "/tmp/refdump_out_%d.dat", (int)getpid());
    FILE *fo = fopen(fout, "w"), *fi = fopen(fin, "w+");
    int i, nheavy = 0;
    for (i = 0; i < pdb->natoms; i++) if (strcmp(pdb->latoms[i].symbol, "H") != 0) nheavy++;
    fprintf(fi, "3 rbox D3\n%d\n", nheavy);
    for (i = 0; i < pdb->natoms; i++) {
        s_atm *ca = pdb->latoms + i;
        if (strcmp(ca->symbol, "H") != 0) fprintf(fi, "%.6f %.6f %.6f\n", ca->x + 0.0f, ca->y + 0.0f, ca->z + 0.0f);
    }
    fflush(fi); rewind(fi);
    run_qvoronoi(fi, fo);
    fclose(fi); fclose(fo);
    fo = fopen(fout, "r");
    int dim = 0, nv = 0, nt = 0;
    if (!fgets(line, 256, fo)) return;
    sscanf(line, "%d", &dim);
    if (!fgets(line, 256, fo)) return;
    sscanf(line, "%d", &nv);
    double *c = malloc(sizeof(double) * 3 * (nv + 1));
    for (i = 0; i < nv; i++) {
        if (!fgets(line, 256, fo)) break;
        sscanf(line, "%lf %lf %lf", c + 3 * i, c + 3 * i + 1, c + 3 * i + 2);
    }
    if (!fgets(line, 256, fo)) return;
    sscanf(line, "%d", &nt);
    int *t = malloc(sizeof(int) * 4 * (nt + 1));
    for (i = 0; i < nt; i++) {
        if (!fgets(line, 256, fo)) break;
        sscanf(line, "%d %d %d %d", t + 4 * i, t + 4 * i + 1, t + 4 * i + 2, t + 4 * i + 3);
    }
    long flen = ftell(fo);
    fseek(fo, 0, SEEK_END);
    flen = ftell(fo);
    fclose(fo);
    int d2[2] = {nv, 3};
    rec("qh.centres", 'd', 2, d2, c);
    d2[0] = nt; d2[1] = 4;
    rec("qh.tets", 'i', 2, d2, t);

    /* What the reference actually READS. load_vvertices() calls fill_vvertices() on the output path
     * while the FILE* qhull wrote through is still open and unflushed (voronoi.c:138-146: fclose(ftmp)
     * comes after fill_vvertices), so only the whole stdio blocks (st_blksize, 4096 B) that were
     * already written out are visible: the tail of qhull's tet list is silently lost, and the last
     * visible line may be cut in the middle. Re-enact fill_vvertices' reading loop (voronoi.c:395-445)
     * on a copy truncated to that size. */
    {
        long vis = (flen / 4096) * 4096;
        char ftr[256];
        snprintf(ftr, 256, "/tmp/refdump_trunc_%d.dat", (int)getpid());
        FILE *src = fopen(fout, "r"), *dst = fopen(ftr, "w");
        char *buf = malloc((size_t)vis + 1);
        size_t got = fread(buf, 1, (size_t)vis, src);
        fwrite(buf, 1, got, dst);
        fclose(src); fclose(dst); free(buf);
        FILE *f = fopen(ftr, "r"), *fNb = fopen(ftr, "r");
        char cline[255], nbline[255], s_nvert[255];
        char *st;
        st = fgets(cline, 255, f); st = fgets(cline, 255, f);
        st = fgets(nbline, 255, fNb); st = fgets(nbline, 255, fNb);
        int nseen = 0, *seen = malloc(sizeof(int) * 4 * (nt + 2));
        if (st) {
            int nvert_hdr = 0;
            sscanf(cline, "%d", &nvert_hdr);
            sprintf(s_nvert, "%d", nvert_hdr);
            strcat(s_nvert, "\n");
            while (fgets(nbline, 255, fNb) != NULL && strcmp(s_nvert, nbline) != 0)
                ;
            int cur[4] = {0, 0, 0, 0};     /* curNbIdx is uninitialised in the reference; first line always has 4 ints */
            float xyz[3];
            while (fgets(cline, 255, f) != NULL) {
                if (fgets(nbline, 255, fNb) != NULL) {
                    if (strcmp("\n", cline) != 0 && strcmp("\n", nbline) != 0) {
                        sscanf(cline, "%f %f %f", &xyz[0], &xyz[1], &xyz[2]);
                        sscanf(nbline, "%d %d %d %d", &cur[0], &cur[1], &cur[2], &cur[3]);
                        memcpy(seen + 4 * nseen, cur, sizeof cur);
                        nseen++;
                    }
                }
            }
        }
        fclose(f); fclose(fNb);
        remove(ftr);
        d2[0] = nseen; d2[1] = 4;
        rec("qh.tets_seen", 'i', 2, d2, seen);
        int fl[2] = {(int)flen, (int)vis};
        rec1("qh.out_len", 'i', 2, fl);
        free(seen);
    }
    remove(fin); remove(fout);
    free(c); free(t);
}

static void dump_spheres(const char *pfx, s_lst_vvertice *lv, s_pdb *pdb)
{
    int n = lv->nvert, i, j;
    char nm[64];
    float *f = malloc(sizeof(float) * (n + 1) * 7);
    int *iv = malloc(sizeof(int) * (n + 1) * 8);
    for (i = 0; i < n; i++) {
        s_vvertice *v = lv->vertices + i;
        f[7 * i] = v->x; f[7 * i + 1] = v->y; f[7 * i + 2] = v->z; f[7 * i + 3] = v->ray;
        f[7 * i + 4] = v->bary[0]; f[7 * i + 5] = v->bary[1]; f[7 * i + 6] = v->bary[2];
        for (j = 0; j < 4; j++) iv[8 * i + j] = (int)(v->neigh[j] - pdb->latoms);
        iv[8 * i + 4] = v->type; iv[8 * i + 5] = v->qhullId; iv[8 * i + 6] = v->resid; iv[8 * i + 7] = v->id;
    }
    int d2[2] = {n, 7};
    snprintf(nm, 64, "%s.f", pfx); rec(nm, 'f', 2, d2, f);      /* x y z ray bary[3] */
    d2[1] = 8;
    snprintf(nm, 64, "%s.i", pfx); rec(nm, 'i', 2, d2, iv);     /* neigh[4] (index into pdb atoms) type qhullId resid id */
    free(f); free(iv);
}

#define NDESC_F 30
#define NDESC_I 31
static void dump_pockets(const char *pfx, c_lst_pockets *pockets, s_pdb *pdb)
{
    char nm[64];
    int np = pockets ? (int)pockets->n_pockets : 0, p = 0, tot = 0, k;
    node_pocket *cur;
    for (cur = pockets ? pockets->first : NULL; cur; cur = cur->next) tot += (int)cur->pocket->v_lst->n_vertices;
    int *off = malloc(sizeof(int) * (np + 1)), *idx = malloc(sizeof(int) * (tot + 1));
    float *df = calloc((np + 1) * NDESC_F, sizeof(float));
    int *di = calloc((np + 1) * NDESC_I, sizeof(int));
    unsigned char *tags = calloc((np + 1), 16);
    off[0] = 0;
    for (cur = pockets ? pockets->first : NULL; cur; cur = cur->next, p++) {
        s_pocket *pk = cur->pocket;
        s_desc *d = pk->pdesc;
        node_vertice *nv;
        k = off[p];
        for (nv = pk->v_lst->first; nv; nv = nv->next) idx[k++] = (int)(nv->vertice - pockets->vertices->vertices);
        off[p + 1] = k;
        float *F = df + p * NDESC_F;
        F[0] = pk->score; F[1] = d->drug_score; F[2] = d->hydrophobicity_score; F[3] = d->volume_score;
        F[4] = d->volume; F[5] = d->convex_hull_volume; F[6] = d->prop_polar_atm; F[7] = d->mean_asph_ray;
        F[8] = d->masph_sacc; F[9] = d->apolar_asphere_prop; F[10] = d->mean_loc_hyd_dens; F[11] = d->as_density;
        F[12] = d->as_max_dst; F[13] = d->flex; F[14] = d->nas_norm; F[15] = d->polarity_score_norm;
        F[16] = d->mean_loc_hyd_dens_norm; F[17] = d->prop_asapol_norm; F[18] = d->as_density_norm; F[19] = d->as_max_dst_norm;
        F[20] = d->surf_vdw14; F[21] = d->surf_vdw22; F[22] = d->surf_pol_vdw14; F[23] = d->surf_pol_vdw22;
        F[24] = d->surf_apol_vdw14; F[25] = d->surf_apol_vdw22; F[26] = d->as_max_r;
        F[27] = pk->bary[0]; F[28] = pk->bary[1]; F[29] = pk->bary[2];
        int *I = di + p * NDESC_I;
        I[0] = d->nb_asph; I[1] = d->polarity_score; I[2] = d->charge_score; I[3] = d->n_abpa;
        I[4] = pk->nAlphaApol; I[5] = pk->nAlphaPol; I[6] = d->interChain; I[7] = d->characterChain1;
        I[8] = d->characterChain2; I[9] = d->numResChain1; I[10] = d->numResChain2;
        for (k = 0; k < 20; k++) I[11 + k] = d->aa_compo[k];
        memcpy(tags + p * 16, d->ligTag, 8);
        tags[p * 16 + 8] = d->nameChain1[0]; tags[p * 16 + 9] = d->nameChain1[1];
        tags[p * 16 + 10] = d->nameChain2[0]; tags[p * 16 + 11] = d->nameChain2[1];
    }
    (void)pdb;
    snprintf(nm, 64, "%s.off", pfx); rec1(nm, 'i', np + 1, off);
    snprintf(nm, 64, "%s.sph", pfx); rec1(nm, 'i', tot, idx);
    int d2[2] = {np, NDESC_F};
    snprintf(nm, 64, "%s.desc_f", pfx); rec(nm, 'f', 2, d2, df);
    d2[1] = NDESC_I;
    snprintf(nm, 64, "%s.desc_i", pfx); rec(nm, 'i', 2, d2, di);
    d2[1] = 16;
    snprintf(nm, 64, "%s.tags", pfx); rec(nm, 'b', 2, d2, tags);
    free(off); free(idx); free(df); free(di); free(tags);
}

static void dump_atom_effects(const char *name, s_pdb *pdb)
{
    int n = pdb->natoms, i;
    float *f = malloc(sizeof(float) * n * 4);
    for (i = 0; i < n; i++) {
        s_atm *a = pdb->latoms + i;
        f[4 * i] = (float)a->abpa; f[4 * i + 1] = a->a0; f[4 * i + 2] = a->dA; f[4 * i + 3] = a->abpa_sourrounding_prob;
    }
    int d2[2] = {n, 4};
    rec(name, 'f', 2, d2, f);
    free(f);
}

static void dump_grid(const char *name, s_mdgrid *g)
{
    int d3[3] = {g->nx, g->ny, g->nz}, x, y;
    float *buf = malloc(sizeof(float) * (size_t)g->nx * g->ny * g->nz);
    for (x = 0; x < g->nx; x++) for (y = 0; y < g->ny; y++)
        memcpy(buf + ((size_t)x * g->ny + y) * g->nz, g->gridvalues[x][y], sizeof(float) * g->nz);
    rec(name, 'f', 3, d3, buf);
    free(buf);
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: ref_dump out.dump [--md] [--score] [--seed S] <fpocket args>\n"); return 2; }
    const char *outp = argv[1];
    int md = 0, md_score = 0, a = 2;
    long seed = -1;
    char *ligname = NULL, *wantname = NULL;
    for (;;) {
        if (a + 1 < argc && !strcmp(argv[a], "--lig")) { ligname = argv[a + 1]; a += 2; continue; }
        if (a + 1 < argc && !strcmp(argv[a], "--want")) { wantname = argv[a + 1]; a += 2; continue; }
        if (a < argc && !strcmp(argv[a], "--md")) { md = 1; a++; }
        else if (a < argc && !strcmp(argv[a], "--score")) { md_score = 1; a++; }
        else if (a + 1 < argc && !strcmp(argv[a], "--seed")) { seed = atol(argv[a + 1]); a += 2; }
        else break;
    }
    int nargs = argc - a + 1;
    char **args = malloc(sizeof(char *) * (nargs + 1));
    args[0] = argv[0];
    for (int i = a; i < argc; i++) args[i - a + 1] = argv[i];
    args[nargs] = NULL;
    s_fparams *params = get_fpocket_args(nargs, args);
    if (!params) { fprintf(stderr, "bad fpocket args\n"); return 2; }
    params->fpocket_running = md ? 0 : 1;
    if (md && !wantname) params->flag_do_asa_and_volume_calculations = 0;   /* mdpocket.c:89: mdpocket_detect only; characterize keeps the default */
    char *pdbname = params->pdb_path;
    int cif = strstr(pdbname, ".cif") != NULL;
    s_pdb *pdb, *pdb_w_lig;
    if (cif) {
        pdb = open_mmcif(pdbname, NULL, M_DONT_KEEP_LIG, params->model_number, params);
        pdb_w_lig = md ? pdb : open_mmcif(pdbname, NULL, M_KEEP_LIG, params->model_number, params);
        if (!pdb) return 3;
        read_mmcif(pdb, NULL, M_DONT_KEEP_LIG, params->model_number, params);
        if (!md) read_mmcif(pdb_w_lig, NULL, M_KEEP_LIG, params->model_number, params);
    } else {
        /* with --lig: dpocket.c:186-203 (pdb_cplx_nl / pdb_cplx_l) */
        pdb = rpdb_open(pdbname, ligname, M_DONT_KEEP_LIG, ligname ? 0 : params->model_number, params);
        pdb_w_lig = md ? pdb : rpdb_open(pdbname, ligname, M_KEEP_LIG, ligname ? 0 : params->model_number, params);
        if (!pdb) return 3;
        rpdb_read(pdb, ligname, M_DONT_KEEP_LIG, ligname ? 0 : params->model_number, params);
        if (!md) rpdb_read(pdb_w_lig, ligname, M_KEEP_LIG, ligname ? 0 : params->model_number, params);
    }

    OUT = fopen(outp, "wb");
    if (!OUT) { perror(outp); return 4; }
    fputs("FPKDUMP1\n", OUT);
    rec_i("md", md);
    {
        float pf[6] = {params->asph_min_size, params->asph_max_size, params->clust_max_dist,
                       params->refine_min_apolar_asphere_prop, params->min_as_density, 0.f};
        int pi[10] = {params->min_apol_neigh, params->nb_mcv_iter, params->min_pock_nb_asph,
                      params->flag_do_asa_and_volume_calculations, params->xlig_resnumber, params->xpocket_n,
                      params->min_n_explicit_pocket_atoms, params->clustering_method, params->distance_measure, 0};
        rec1("params.f", 'f', 6, pf);
        rec1("params.i", 'i', 10, pi);
        if (params->xpocket_n > 0) {
            int *xp = malloc(sizeof(int) * 3 * params->xpocket_n);
            for (int i = 0; i < params->xpocket_n; i++) {
                xp[3 * i] = params->xpocket_residue_number[i];
                xp[3 * i + 1] = params->xpocket_chain_code[i];
                xp[3 * i + 2] = params->xpocket_insertion_code[i];
            }
            int d2[2] = {params->xpocket_n, 3};
            rec("params.xpocket", 'i', 2, d2, xp);
            free(xp);
        }
    }
    dump_atoms("pdb", pdb);
    if (!md) dump_atoms("pdbw", pdb_w_lig);
    dump_qvoronoi(pdb);

    /* ---- stage by stage, the call sequence of search_pocket (fpocket.c:58-225) ---- */
    s_lst_vvertice *lvert = load_vvertices(pdb, params, 0.0, 0.0, 0.0);
    /* load_vvertices did srand(time(NULL)) (voronoi.c:87); the first rand_uniform() would re-seed
     * with time() again (utils.c:71-91) unless the generator is already marked initialised */
    if (seed >= 0) { start_rand_generator(); srand((unsigned)seed); }
    dump_spheres("sph0", lvert, pdb);
    int clustered = 0;
    if (params->xlig_resnumber == -1 && params->xpocket_n == 0) {
        s_clusterlib_vertices *cv = prepare_vertices_for_cluster_lib(lvert, params->clustering_method, params->distance_measure);
        Node *tree = treecluster(lvert->nvert, 3, cv->pos, cv->mask, cv->weight, cv->transpose, cv->dist, cv->method, NULL);
        if (!tree) { rec_i("status", -1); fclose(OUT); return 0; }
        int **ids = cuttree_distance(lvert->nvert, tree, params->clust_max_dist);
        transferClustersToVertices(ids, lvert);
        clustered = 1;
    }
    rec_i("clustered", clustered);
    {
        int *r = malloc(sizeof(int) * (lvert->nvert + 1));
        for (int i = 0; i < lvert->nvert; i++) r[i] = lvert->vertices[i].resid;
        rec1("cluster_id", 'i', lvert->nvert, r);
        free(r);
    }
    c_lst_pockets *pockets = assign_pockets(lvert);
    if (!pockets) { rec_i("status", -2); fclose(OUT); return 0; }
    apply_clustering(pockets);
    reIndexPockets(pockets);
    clock_t t0 = clock();
    set_pockets_descriptors(pockets, pdb, params, pdb_w_lig);
    rec_f("t_desc_s", (float)(clock() - t0) / CLOCKS_PER_SEC);
    dump_pockets("pre", pockets, pdb);
    dump_atom_effects("atom_fx", pdb);
    dropSmallNpolarPockets(pockets, params);
    if (pockets->n_pockets > 0) {
        reIndexPockets(pockets);
        sort_pockets(pockets, M_SCORE_SORT_FUNCT);
        reIndexPockets(pockets);
    }
    dump_pockets("fin", pockets, pdb);
    dump_spheres("sph1", lvert, pdb);     /* resid now = final rank (stale for dropped) */

    if (md && pockets->n_pockets > 0) {
        /* one frame of mdpocket_detect (mdpocket.c:142-175): grids defined and filled by this frame */
        s_mdparams *mp = init_def_mdparams();
        mp->fpar = params;
        mp->flag_scoring = md_score;
        s_mdgrid *freq = init_md_grid(pockets, mp), *dens = init_md_grid(pockets, mp), *refg = init_md_grid(pockets, mp);
        reset_grid(refg);
        calculate_md_dens_grid(dens, pockets, mp);
        update_md_grid(freq, refg, pockets, mp);
        float org[4] = {dens->origin[0], dens->origin[1], dens->origin[2], dens->resolution};
        rec1("grid.origin", 'f', 4, org);
        dump_grid("grid.dens", dens);
        dump_grid("grid.freq", freq);
    }
    if (wantname && pockets->n_pockets > 0) {
        /* one snapshot of mdpocket_characterize (mdpocket.c:436-460): the spheres of the surviving pockets within M_MDP_CUBE_SIDE / 2 of a point of
         * the wanted zone, once per point that reaches them (extract_wanted_vertices, mdpocket.c:695-745), described as ONE pocket by set_descriptors */
        s_pdb *wp = rpdb_open(wantname, NULL, M_DONT_KEEP_LIG, 0, params);
        if (!wp) { fprintf(stderr, "cannot read %s\n", wantname); return 5; }
        rpdb_read(wp, NULL, M_DONT_KEEP_LIG, 0, params);
        float *wxyz = malloc(sizeof(float) * 3 * (wp->natoms + 1));
        for (int j = 0; j < wp->natoms; j++) { wxyz[3 * j] = wp->latoms_p[j]->x; wxyz[3 * j + 1] = wp->latoms_p[j]->y; wxyz[3 * j + 2] = wp->latoms_p[j]->z; }
        int d2w[2] = {wp->natoms, 3};
        rec("want.points", 'f', 2, d2w, wxyz);
        s_pocket *cp = extract_wanted_vertices(pockets, wp);
        int nv = (int)cp->v_lst->n_vertices, j = 0, natms = 0;
        s_vvertice **tab = malloc(sizeof(s_vvertice *) * (nv + 1));
        int *ids = malloc(sizeof(int) * (nv + 1));
        for (node_vertice *c = cp->v_lst->first; c; c = c->next) { tab[j] = c->vertice; ids[j] = (int)(c->vertice - lvert->vertices); j++; }
        rec1("want.sph", 'i', nv, ids);
        if (nv > 0) {
            s_atm **patoms = get_pocket_contacted_atms(cp, &natms);
            int *pa = malloc(sizeof(int) * (natms + 1));
            for (j = 0; j < natms; j++) pa[j] = (int)(patoms[j] - pdb->latoms);
            rec1("want.atoms", 'i', natms, pa);
            set_descriptors(patoms, natms, tab, nv, cp->pdesc, params->nb_mcv_iter, pdb, params->flag_do_asa_and_volume_calculations);
            s_desc *d = cp->pdesc;
            float F[NDESC_F] = {0};
            F[1] = d->drug_score; F[2] = d->hydrophobicity_score; F[3] = d->volume_score;
            F[4] = d->volume; F[5] = d->convex_hull_volume; F[6] = d->prop_polar_atm; F[7] = d->mean_asph_ray;
            F[8] = d->masph_sacc; F[9] = d->apolar_asphere_prop; F[10] = d->mean_loc_hyd_dens; F[11] = d->as_density;
            F[12] = d->as_max_dst; F[13] = d->flex; F[14] = d->nas_norm; F[15] = d->polarity_score_norm;
            F[16] = d->mean_loc_hyd_dens_norm; F[17] = d->prop_asapol_norm; F[18] = d->as_density_norm; F[19] = d->as_max_dst_norm;
            F[20] = d->surf_vdw14; F[21] = d->surf_vdw22; F[22] = d->surf_pol_vdw14; F[23] = d->surf_pol_vdw22;
            F[24] = d->surf_apol_vdw14; F[25] = d->surf_apol_vdw22; F[26] = d->as_max_r;
            int I[NDESC_I] = {0};
            I[0] = d->nb_asph; I[1] = d->polarity_score; I[2] = d->charge_score; I[3] = d->n_abpa;
            for (j = 0; j < 20; j++) I[11 + j] = d->aa_compo[j];
            rec1("want.desc_f", 'f', NDESC_F, F);
            rec1("want.desc_i", 'i', NDESC_I, I);
            dump_atom_effects("want.atom_fx", pdb);
        }
    }
    if (ligname && pdb_w_lig->natm_lig > 0 && pockets->n_pockets > 0) {
        /* dpocket.c:204-290 on the final pockets, with the reference's own functions */
        s_atm **lig = pdb_w_lig->latm_lig;
        int nal = pdb_w_lig->natm_lig, j, nvn = 0, nai = 0, nmol = 1;
        int *la = malloc(sizeof(int) * nal), *lm = malloc(sizeof(int) * nal);
        char chain_tmp[2];
        int resnumber_tmp = lig[0]->res_id;
        strncpy(chain_tmp, lig[0]->chain, 2);
        for (j = 0; j < nal; j++) {
            if (strncmp(chain_tmp, lig[j]->chain, 2) != 0 || resnumber_tmp != lig[j]->res_id) {
                nmol++; strncpy(chain_tmp, lig[j]->chain, 2); resnumber_tmp = lig[j]->res_id;
            }
            la[j] = (int)(lig[j] - pdb_w_lig->latoms); lm[j] = nmol - 1;
        }
        rec1("lig.atoms", 'i', nal, la);
        rec1("lig.mol", 'i', nal, lm);
        s_atm **interface = get_mol_ctd_atm_neigh(lig, nal, lvert->pvertices, lvert->nvert, M_VERT_LIG_NEIG_DIST, M_INTERFACE_SEARCH, &nai);
        s_vvertice **tpverts = get_mol_vert_neigh(lig, nal, lvert->pvertices, lvert->nvert, M_VERT_LIG_NEIG_DIST, &nvn);
        int *ia = malloc(sizeof(int) * (nai + 1)), *xs = malloc(sizeof(int) * (nvn + 1));
        for (j = 0; j < nai; j++) ia[j] = (int)(interface[j] - pdb->latoms);       /* discovery order of the sweep */
        for (j = 0; j < nvn; j++) xs[j] = (int)(tpverts[j] - lvert->vertices);
        rec1("lig.iface", 'i', nai, ia);
        rec1("lig.xspheres", 'i', nvn, xs);
        float *crit = malloc(sizeof(float) * 4 * pockets->n_pockets);
        node_pocket *cur = pockets->first;
        int q = 0;
        while (cur) {
            int nbpa = 0;
            s_atm **patoms = get_vert_contacted_atms(cur->pocket->v_lst, &nbpa);
            float ovlp = atm_corsp(interface, nai, patoms, nbpa), dst = 0, tmp;
            for (j = 0; j < nal; j++) {
                tmp = dist(lig[j]->x, lig[j]->y, lig[j]->z, cur->pocket->bary[0], cur->pocket->bary[1], cur->pocket->bary[2]);
                if (j == 0) dst = tmp; else if (tmp < dst) dst = tmp;
            }
            s_vvertice **pvert = get_pocket_pvertices(cur->pocket);
            float c4 = count_atm_prop_vert_neigh(lig, nal, pvert, cur->pocket->size, M_CRIT4_D, nmol);
            float c5 = count_pocket_lig_vert_ovlp(lig, nal, pvert, cur->pocket->size, M_CRIT5_D);
            crit[4 * q] = ovlp; crit[4 * q + 1] = dst; crit[4 * q + 2] = c4; crit[4 * q + 3] = c5;
            q++;
            cur = cur->next;
        }
        int d2[2] = {q, 4};
        rec("lig.crit", 'f', 2, d2, crit);
        /* dpocket's other interface definition (dparams.h:44 M_INTERFACE_METHOD2, dpocket.c:331-336): the atoms of the complex within M_LIG_NEIG_DIST of a
         * ligand atom, ligand excluded; indices into pdb_w_lig. Last, because get_sorted_list re-sorts the atoms it is given. */
        {
            int nai2 = 0;
            s_atm **if2 = get_mol_atm_neigh(lig, nal, pdb_w_lig->latoms_p, pdb_w_lig->natoms, M_LIG_NEIG_DIST, &nai2);
            int *i2 = malloc(sizeof(int) * (nai2 + 1));
            for (j = 0; j < nai2; j++) i2[j] = (int)(if2[j] - pdb_w_lig->latoms);
            rec1("lig.iface2", 'i', nai2, i2);
            float *ov2 = malloc(sizeof(float) * (pockets->n_pockets + 1));
            q = 0;
            for (cur = pockets->first; cur; cur = cur->next) {
                int nbpa = 0;
                s_atm **patoms = get_vert_contacted_atms(cur->pocket->v_lst, &nbpa);
                ov2[q++] = atm_corsp(if2, nai2, patoms, nbpa);
            }
            rec1("lig.ovlp2", 'f', q, ov2);
        }
    }
    rec_i("status", 0);
    fclose(OUT);
    return 0;
}
