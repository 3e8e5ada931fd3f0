/*
 * TEST INFRASTRUCTURE — not part of the product.
 *
 * ref_io: the reference's own reader and writers around a result that comes from a file, so that the native host I/O
 * (fpocket_b200/hostio) can be pinned to them on the CPU. Linked against the UNMODIFIED reference objects and the C shim
 * (fpocket_b200/shim/search_pocket_cuda.c: shim_job_prepare / shim_rebuild), with the fpk_* entry points the shim calls
 * stubbed here (no GPU, no library).
 *
 *   ref_io atoms <out.dump> [--lig NAME] <fpocket options, -f X.pdb ...>
 *        (--lig: the structure is loaded as dpocket does, dpocket.c:186-203: ligan = NAME, model 0; the ligand atom list is dumped too)
 *        rpdb_open + rpdb_read for pdb and pdb_w_lig exactly as process_pdb does (src/fpmain.c:143-164), then the shim's flattening;
 *        dumps the flat arrays of fpk_structure and the text fields of every atom (FPKDUMP1 records, tests/dumpio.py)
 *   ref_io write <result.bin> <fpocket options, -f X.pdb ...>
 *        same reading, then shim_rebuild() on the fpk_result stored in result.bin and the reference's write_out_fpocket()
 *        (src/fpout.c:55): the files the reference's writers produce for THAT result
 *
 *   ref_io mdwrite <grids.bin> <prefix> <fpocket options, -f top.pdb ...>
 *        what mdpocket_detect does after its frame loop (mdpocket.c:276-330) with the reference's own code: normalize_grid on both grids,
 *        project_grid_on_atoms, write_first_bfactor_density, write_md_grid x 2 -> <prefix>_{freq,dens}.dx, _{freq_iso_0_5,dens_iso_8}.pdb,
 *        _all_atom_pdensities.pdb. grids.bin: int32 {nx,ny,nz,n_snapshots,use_drug_score}, float origin[3], dens[ncell], freq[ncell] (sums)
 *
 * result.bin (written by tests/test_hostio.py): int32 header {n_sites,n_tets,n_spheres,n_clusters,n_pockets,natoms,sizeof(fpk_desc)},
 * then sph_xyzr, sph_bary, sph_neigh, sph_type, sph_cluster, cl_off, cl_sph, desc, rank_to_cluster, atom_fx as raw arrays.
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fpocket.h"
#include "fpout.h"
#include "mdpbase.h"
#include "mdparams.h"
#include "mdpout.h"
#include "fpocket_cuda.h"
#include "shim.h"

/* ---- stubs of the library entry points the shim object references ---- */
int fpk_detect(const fpk_structure *in, const fpk_params *p, fpk_result *out) { (void)in; (void)p; (void)out; return FPK_ERR_NO_DEVICE; }
const char *fpk_last_error(void) { return "ref_io: no library"; }
void fpk_default_params(fpk_params *p) { memset(p, 0, sizeof *p); }
void fpk_result_free(fpk_result *r)
{
    free(r->sph_xyzr); free(r->sph_bary); free(r->sph_neigh); free(r->sph_type); free(r->sph_cluster); free(r->cl_off); free(r->cl_sph);
    free(r->desc); free(r->rank_to_cluster); free(r->atom_fx);
    memset(r, 0, sizeof *r);
}

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
    size_t es = (dtype == 'b') ? 1 : 4;
    if (n) fwrite(data, es, n, OUT);
}
static void rec1(const char *name, char dtype, int n, const void *data) { rec(name, dtype, 1, &n, data); }

static void dump_text(const char *pfx, s_pdb *pdb)
{
    int n = pdb->natoms;
    unsigned char *t = calloc((size_t)n + 1, 40);
    int *iv = calloc((size_t)n + 1, 3 * sizeof(int));
    for (int i = 0; i < n; i++) {
        s_atm *a = pdb->latoms + i;
        unsigned char *q = t + (size_t)i * 40;
        memcpy(q, a->type, 7); memcpy(q + 8, a->name, 5); memcpy(q + 16, a->res_name, 8); memcpy(q + 24, a->chain, 4); memcpy(q + 28, a->symbol, 3);
        q[32] = (unsigned char)a->pdb_insert; q[33] = (unsigned char)a->pdb_aloc;
        iv[3 * i] = a->id; iv[3 * i + 1] = a->res_id; iv[3 * i + 2] = a->charge;
    }
    char nm[64];
    int d2[2] = {n, 40};
    snprintf(nm, 64, "%s.text", pfx); rec(nm, 'b', 2, d2, t);
    d2[1] = 3;
    snprintf(nm, 64, "%s.irc", pfx); rec(nm, 'i', 2, d2, iv);
    int nh = pdb->nhetatm;
    snprintf(nm, 64, "%s.nhetatm", pfx); rec1(nm, 'i', 1, &nh);
    free(t); free(iv);
}

static void *slurp(FILE *f, size_t bytes)
{
    void *p = malloc(bytes ? bytes : 1);
    if (bytes && fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "ref_io: short result file\n"); exit(5); }
    return p;
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: ref_io atoms out.dump <fpocket args> | ref_io write result.bin <fpocket args>\n"); return 2; }
    const int do_write = !strcmp(argv[1], "write");
    if (!strcmp(argv[1], "mdwrite")) {
        if (argc < 6) return 2;
        int na = argc - 3;
        char **av = malloc(sizeof(char *) * (na + 1));
        av[0] = argv[0];
        for (int i = 4; i < argc; i++) av[i - 3] = argv[i];
        av[na] = NULL;
        s_fparams *fp = get_fpocket_args(na, av);
        if (!fp) return 2;
        FILE *g = fopen(argv[2], "rb");
        int h[5];
        float org[3];
        if (!g || fread(h, 4, 5, g) != 5 || fread(org, 4, 3, g) != 3) return 5;
        const size_t nc = (size_t)h[0] * h[1] * h[2];
        float *grid[2];
        s_mdgrid *G[2];
        for (int q = 0; q < 2; q++) {           /* 0 = density, 1 = frequency; the allocation pattern of init_md_grid (mdpbase.c:446-502) */
            grid[q] = malloc(sizeof(float) * nc);
            if (fread(grid[q], 4, nc, g) != nc) return 5;
            G[q] = malloc(sizeof(s_mdgrid));
            G[q]->resolution = M_MDP_GRID_RESOLUTION; G[q]->nx = h[0]; G[q]->ny = h[1]; G[q]->nz = h[2]; G[q]->n_snapshots = h[3];
            G[q]->origin = malloc(sizeof(float) * 3);
            for (int k = 0; k < 3; k++) G[q]->origin[k] = org[k];
            G[q]->gridvalues = malloc(sizeof(float **) * h[0]);
            for (int x = 0; x < h[0]; x++) {
                G[q]->gridvalues[x] = malloc(sizeof(float *) * h[1]);
                for (int y = 0; y < h[1]; y++) G[q]->gridvalues[x][y] = grid[q] + ((size_t)x * h[1] + y) * h[2];
            }
        }
        fclose(g);
        s_mdparams par;
        memset(&par, 0, sizeof par);
        par.fpar = fp; par.flag_scoring = h[4];
        normalize_grid(G[1], h[3]);
        normalize_grid(G[0], h[3]);
        s_pdb *cp = rpdb_open(fp->pdb_path, NULL, M_DONT_KEEP_LIG, 0, fp);
        if (!cp) return 3;
        rpdb_read(cp, NULL, M_DONT_KEEP_LIG, 0, fp);
        project_grid_on_atoms(G[0], cp);
        char nm[512];
        FILE *f1, *f2;
        snprintf(nm, sizeof nm, "%s_all_atom_pdensities.pdb", argv[3]); f1 = fopen(nm, "w"); write_first_bfactor_density(f1, cp); fclose(f1);
        snprintf(nm, sizeof nm, "%s_freq.dx", argv[3]); f1 = fopen(nm, "w");
        snprintf(nm, sizeof nm, "%s_freq_iso_0_5.pdb", argv[3]); f2 = fopen(nm, "w");
        write_md_grid(G[1], f1, f2, &par, M_MDP_DEFAULT_ISO_VALUE_FREQ); fclose(f1); fclose(f2);
        snprintf(nm, sizeof nm, "%s_dens.dx", argv[3]); f1 = fopen(nm, "w");
        snprintf(nm, sizeof nm, "%s_dens_iso_8.pdb", argv[3]); f2 = fopen(nm, "w");
        write_md_grid(G[0], f1, f2, &par, M_MDP_DEFAULT_ISO_VALUE_DENS); fclose(f1); fclose(f2);
        return 0;
    }
    const char *ligname = NULL;
    int first = 3;
    if (argc > 5 && !strcmp(argv[3], "--lig")) { ligname = argv[4]; first = 5; }
    int nargs = argc - first + 1;
    char **args = malloc(sizeof(char *) * (nargs + 1));
    args[0] = argv[0];
    for (int i = first; i < argc; i++) args[i - first + 1] = argv[i];
    args[nargs] = NULL;
    s_fparams *params = get_fpocket_args(nargs, args);
    if (!params) { fprintf(stderr, "bad fpocket args\n"); return 2; }
    params->fpocket_running = 1;
    char *pdbname = params->pdb_path;
    /* process_pdb, fpmain.c:143-164 (PDB route) */
    const int model = ligname ? 0 : params->model_number;
    s_pdb *pdbw = rpdb_open(pdbname, ligname, M_KEEP_LIG, model, params);
    s_pdb *pdb = rpdb_open(pdbname, ligname, M_DONT_KEEP_LIG, model, params);
    if (!pdb || !pdbw) { fprintf(stderr, "! Structure reading failed!\n"); return 3; }
    if (ligname && pdbw->natm_lig <= 0) { fprintf(stdout, "ERROR - No ligand %s found in %s.\n", ligname, pdbname); return 6; }
    rpdb_read(pdbw, ligname, M_KEEP_LIG, model, params);
    rpdb_read(pdb, ligname, M_DONT_KEEP_LIG, model, params);
    shim_job job;
    if (shim_job_prepare(&job, pdb, params, pdbw)) return 4;
    if (!do_write) {
        OUT = fopen(argv[2], "wb");
        if (!OUT) { perror(argv[2]); return 4; }
        fputs("FPKDUMP1\n", OUT);
        const fpk_structure *s = &job.st;
        const int n = s->natoms, nw = s->natoms_w;
        rec1("n", 'i', 1, &n); rec1("nw", 'i', 1, &nw);
        rec1("x", 'f', n, s->x); rec1("y", 'f', n, s->y); rec1("z", 'f', n, s->z);
        rec1("radius", 'f', n, s->radius); rec1("electroneg", 'f', n, s->electroneg); rec1("bfactor", 'f', n, s->bfactor);
        rec1("atom_id", 'i', n, s->atom_id); rec1("res_id", 'i', n, s->res_id);
        rec1("aa_idx", 'b', n, s->aa_idx); rec1("flags", 'b', n, s->flags);
        float bf[3] = {s->avg_bfactor, s->min_bfactor, s->max_bfactor};
        rec1("bfstat", 'f', 3, bf);
        rec1("wx", 'f', nw, s->wx); rec1("wy", 'f', nw, s->wy); rec1("wz", 'f', nw, s->wz);
        rec1("wradius", 'f', nw, s->wradius); rec1("welectroneg", 'f', nw, s->welectroneg); rec1("wflags", 'b', nw, s->wflags);
        int sp = s->same_pdb;
        rec1("same_pdb", 'i', 1, &sp);
        rec1("xlig", 'f', 3 * s->n_xlig, s->xlig);
        dump_text("pdb", pdb);
        dump_text("pdbw", pdbw);
        if (ligname) {
            int nl = pdbw->natm_lig, *li = malloc(sizeof(int) * (nl + 1));
            for (int i = 0; i < nl; i++) li[i] = (int)(pdbw->latm_lig[i] - pdbw->latoms);
            rec1("lig", 'i', nl, li);
            free(li);
        }
        fclose(OUT);
        return 0;
    }
    FILE *f = fopen(argv[2], "rb");
    if (!f) { perror(argv[2]); return 4; }
    int h[7];
    if (fread(h, 4, 7, f) != 7 || h[6] != (int)sizeof(fpk_desc) || h[5] != pdb->natoms) { fprintf(stderr, "ref_io: result header does not fit this structure\n"); return 5; }
    fpk_result R;
    memset(&R, 0, sizeof R);
    R.status = FPK_OK;
    R.n_sites = h[0]; R.n_tets = h[1]; R.n_spheres = h[2]; R.n_clusters = h[3]; R.n_pockets = h[4];
    const size_t ns = (size_t)h[2], nc = (size_t)h[3], np = (size_t)h[4], na = (size_t)h[5];
    R.sph_xyzr = slurp(f, ns * 16); R.sph_bary = slurp(f, ns * 12); R.sph_neigh = slurp(f, ns * 16); R.sph_type = slurp(f, ns);
    R.sph_cluster = slurp(f, ns * 4); R.cl_off = slurp(f, (nc + 1) * 4); R.cl_sph = slurp(f, ns * 4); R.desc = slurp(f, nc * sizeof(fpk_desc));
    R.rank_to_cluster = slurp(f, np * 4); R.atom_fx = slurp(f, na * 16);
    fclose(f);
    if (getenv("REF_IO_WRITE_MODE")) { extern char write_mode[10]; strncpy(write_mode, getenv("REF_IO_WRITE_MODE"), 9); }   /* dpocket leaves it at "d" */
    c_lst_pockets *pockets = shim_rebuild(pdb, params, pdbw, &R, np ? FPK_OK : FPK_ERR_NO_POCKETS, 1);
    if (pockets) {                                         /* a list with no pocket left is still written (fpocket.c:225 returns it) */
        write_out_fpocket(pockets, pdb, pdbname);          /* fpmain.c:195 */
        c_lst_pocket_free(pockets);
    } else printf("no pockets found\n");
    return 0;
}
