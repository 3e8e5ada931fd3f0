/*
 * Reference-side binding of libfpocket_cuda: a drop-in
 *
 *     c_lst_pockets *search_pocket(s_pdb *pdb, s_fparams *params, s_pdb *pdb_w_lig)
 *
 * with the signature and ownership rules of reference src/fpocket.c:58-225. It replaces that one translation unit
 * (src/fpocket.c holds nothing else): fpmain.c:182, dpocket.c:228, mdpocket.c:844 and tpocket.c:467,542 call it unchanged.
 * It is compiled against the reference's own headers and linked with the reference's objects (shim/Makefile); it
 *   1. flattens s_pdb / s_fparams into fpk_structure / fpk_params                (include/fpocket_cuda.h)
 *   2. calls fpk_detect()                                                        (every stage runs on the GPU)
 *   3. rebuilds s_lst_vvertice + c_lst_pockets with the reference's own allocators (my_malloc, alloc_pocket,
 *      c_lst_vertices_add_last, c_lst_pockets_add_last) so that fpout.c, writepocket.c, dpocket.c, mdpbase.c and
 *      c_lst_pocket_free() work on the result as they do today, and writes the per-atom ASA side effects back.
 * The string bookkeeping of s_desc that is not arithmetic (chain names, residue counts, ligand tag) is done here, on the
 * few surviving pockets, with the reference's helpers. There is no CPU fallback: any failure returns NULL with a message on
 * stderr, as search_pocket does (fpocket.c:135-139,152-156).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "fpocket.h"            /* the reference's header: s_pdb, s_fparams, c_lst_pockets, allocators */
#include "fpocket_cuda.h"
#include "shim.h"

/* The reference's writers shell out for "mkdir [-p] <dir>" and "chmod +x <file>" (fpout.c:83,127, write_visu.c:115,185,266). A fork()
 * of a process that holds a CUDA context (pinned buffers, GPU mappings) costs ~100 ms, five times per structure, so every program linked with this shim
 * answers those three commands in-process; anything else still goes to the shell. */
#include <sys/stat.h>
#include <sys/types.h>
#include <errno.h>
extern int __libc_system(const char *cmd);
static int mkdir_p(const char *path)
{
    char tmp[4096];
    snprintf(tmp, sizeof tmp, "%s", path);
    for (char *q = tmp + 1; *q; q++)
        if (*q == '/') { *q = 0; if (mkdir(tmp, 0777) && errno != EEXIST) return 1; *q = '/'; }
    return (mkdir(tmp, 0777) && errno != EEXIST) ? 1 : 0;
}
int system(const char *cmd)
{
    if (!cmd) return 1;
    if (!strncmp(cmd, "mkdir -p ", 9) && !strpbrk(cmd + 9, " ;&|<>$`\"'*?")) return mkdir_p(cmd + 9) ? 256 : 0;
    if (!strncmp(cmd, "mkdir ", 6) && !strpbrk(cmd + 6, " ;&|<>$`\"'*?")) return (mkdir(cmd + 6, 0777) && errno != EEXIST) ? 256 : 0;
    if (!strncmp(cmd, "chmod +x ", 9) && !strpbrk(cmd + 9, " ;&|<>$`\"'*?")) {
        struct stat st;
        if (stat(cmd + 9, &st)) return 256;
        return chmod(cmd + 9, st.st_mode | S_IXUSR | S_IXGRP | S_IXOTH) ? 256 : 0;
    }
    return __libc_system(cmd);
}


/* set_aa_desc's decision on the first three letters of the residue name (descriptors.c:453-507), as an index 0..19 in the
 * order of aa.h (ALA ARG ASN ASP CYS GLN GLU GLY HIS ILE LEU LYS MET PHE PRO SER THR TRP TYR VAL), -1 when it updates nothing */
static int shim_aa_index(const char *n)
{
    const char a = n[0], b = n[0] ? n[1] : 0, c = (n[0] && n[1]) ? n[2] : 0;
    switch (a) {
    case 'A': return b == 'L' ? 0 : b == 'R' ? 1 : (b == 'S' && c == 'P') ? 3 : 2;
    case 'C': return 4;
    case 'G': return c == 'U' ? 6 : c == 'Y' ? 7 : 5;
    case 'H': return 8;
    case 'I': return 9;
    case 'L': return b == 'Y' ? 11 : 10;
    case 'M': return 12;
    case 'P': return b == 'H' ? 13 : 14;
    case 'S': return 15;
    case 'T': return b == 'H' ? 16 : b == 'R' ? 17 : 18;
    case 'V': return 19;
    }
    return -1;
}

static int shim_flatten(const s_pdb *pdb, const s_fparams *params, int with_props, shim_atoms *o)
{
    const int n = pdb->natoms;
    memset(o, 0, sizeof *o);
    o->f = (float *)malloc(sizeof(float) * 6 * (size_t)(n + 1));
    o->i = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(n + 1));
    o->aa = (int8_t *)malloc((size_t)n + 1);
    o->fl = (uint8_t *)malloc((size_t)n + 1);
    if (!o->f || !o->i || !o->aa || !o->fl) return -1;
    for (int k = 0; k < n; k++) {
        s_atm *a = pdb->latoms + k;
        o->f[k] = a->x; o->f[n + k] = a->y; o->f[2 * n + k] = a->z;
        o->f[3 * n + k] = a->radius; o->f[4 * n + k] = a->electroneg; o->f[5 * n + k] = a->bfactor;
        o->i[k] = a->id; o->i[n + k] = a->res_id;
        o->aa[k] = (int8_t)shim_aa_index(a->res_name);
        uint8_t fl = 0;
        if (with_props) {
            if (strcmp(a->symbol, "H") == 0) fl |= FPK_ATOM_H;                       /* voronoi.c:100,128 */
            if (a->symbol[0] == 'H') fl |= FPK_ATOM_H1;                              /* asa.c:89 when pdb_w_lig is pdb */
            if (params->xpocket_n > 0 && atom_in_explicit_pocket(a, (s_fparams *)params)) fl |= FPK_ATOM_XPOCKET;   /* voronoi.c:934-949 */
        } else if (a->symbol[0] == 'H') fl |= FPK_ATOM_H;                            /* asa.c:89 */
        o->fl[k] = fl;
    }
    return 0;
}
static void shim_free_atoms(shim_atoms *o) { free(o->f); free(o->i); free(o->aa); free(o->fl); }

/* chain bookkeeping of set_atom_based_descriptors (descriptors.c:355-436): not arithmetic, done here on the surviving pockets */
static void shim_chain_info(s_pocket *pocket, s_pdb *all)
{
    int natoms = 0;
    s_desc *d = pocket->pdesc;
    s_atm **atoms = get_pocket_contacted_atms(pocket, &natoms);
    d->interChain = 0; d->characterChain1 = 0; d->characterChain2 = 0; d->numResChain1 = 0; d->numResChain2 = 0;
    d->nameChain1[0] = 0; d->nameChain2[0] = 0;
    if (atoms && natoms > 1) {
        s_atm *first = atoms[0], *cur = NULL;
        char curChainName[255];
        int curChar = 0;
        if (element_in_std_res(first->res_name)) d->characterChain1 = 0;
        else if (element_in_nucl_acid(first->res_name)) d->characterChain1 = 1;
        else if (element_in_kept_res(first->res_name)) d->characterChain1 = 2;
        strcpy(d->nameChain1, first->chain);
        strcpy(curChainName, first->chain);
        d->numResChain1 = countResidues(all->latoms, all->natoms, first->chain);
        for (int i = 0; i < natoms; i++) {
            cur = atoms[i];
            if (cur == first) continue;
            if (strcmp(cur->chain, first->chain) != 0 && d->interChain < 1) {
                d->interChain = 1;
                if (!d->numResChain2) {
                    d->numResChain2 = countResidues(all->latoms, all->natoms, cur->chain);
                    strcpy(curChainName, cur->chain);
                }
            }
            if (element_in_std_res(cur->res_name)) curChar = 0;
            else if (element_in_nucl_acid(cur->res_name)) curChar = 1;
            else if (element_in_kept_res(cur->res_name)) curChar = 2;
            if (curChar != d->characterChain1) d->characterChain2 = curChar;
        }
        if (!d->numResChain2) d->numResChain2 = countResidues(all->latoms, all->natoms, cur->chain);
        strncpy(d->nameChain2, curChainName, 2);
    }
    if (atoms) my_free(atoms);
}

static void shim_fill_desc(s_desc *d, const fpk_desc *g)
{
    d->hydrophobicity_score = g->hydrophobicity_score; d->volume_score = g->volume_score; d->volume = g->volume;
    d->convex_hull_volume = g->convex_hull_volume; d->prop_polar_atm = g->prop_polar_atm; d->mean_asph_ray = g->mean_asph_ray;
    d->masph_sacc = g->masph_sacc; d->apolar_asphere_prop = g->apolar_asphere_prop; d->mean_loc_hyd_dens = g->mean_loc_hyd_dens;
    d->as_density = g->as_density; d->as_max_dst = g->as_max_dst; d->flex = g->flex; d->nas_norm = g->nas_norm;
    d->polarity_score_norm = g->polarity_score_norm; d->mean_loc_hyd_dens_norm = g->mean_loc_hyd_dens_norm;
    d->prop_asapol_norm = g->prop_asapol_norm; d->as_density_norm = g->as_density_norm; d->as_max_dst_norm = g->as_max_dst_norm;
    d->surf_vdw14 = g->surf_vdw14; d->surf_vdw22 = g->surf_vdw22; d->surf_pol_vdw14 = g->surf_pol_vdw14;
    d->surf_pol_vdw22 = g->surf_pol_vdw22; d->surf_apol_vdw14 = g->surf_apol_vdw14; d->surf_apol_vdw22 = g->surf_apol_vdw22;
    d->n_abpa = g->n_abpa;
    for (int k = 0; k < 20; k++) d->aa_compo[k] = g->aa_compo[k];
    d->nb_asph = g->nb_asph; d->polarity_score = g->polarity_score; d->charge_score = g->charge_score;
    d->as_max_r = g->as_max_r; d->drug_score = g->drug_score;
}

/* ---- the three steps, exposed so that a batch driver (fpocket_batch.c) can put many structures into ONE fpk_detect_batch call ---- */
void shim_params(const s_fparams *params, fpk_params *P)
{
    fpk_default_params(P);
    P->asph_min_size = params->asph_min_size; P->asph_max_size = params->asph_max_size; P->clust_max_dist = params->clust_max_dist;
    P->refine_min_apolar_asphere_prop = params->refine_min_apolar_asphere_prop; P->min_as_density = params->min_as_density;
    P->min_apol_neigh = params->min_apol_neigh; P->nb_mcv_iter = params->nb_mcv_iter; P->min_pock_nb_asph = params->min_pock_nb_asph;
    P->do_asa_and_volume = params->flag_do_asa_and_volume_calculations;
    P->skip_clustering = (params->xlig_resnumber != -1 || params->xpocket_n > 0);      /* fpocket.c:114 */
    P->xpocket_active = params->xpocket_n > 0;
    P->min_n_explicit_pocket_atoms = params->min_n_explicit_pocket_atoms;
    P->distance_measure = params->distance_measure;                                      /* -e e | b (clusterlib.c:933,1022) */
    P->clustering_method = params->clustering_method;                                    /* -C s | m | a | c (clusterlib.c:3886) */
    P->mc_seed = (uint64_t)time(NULL);                                                    /* voronoi.c:87 srand(time(NULL)) */
    if (getenv("FPOCKET_MC_SEED")) P->mc_seed = strtoull(getenv("FPOCKET_MC_SEED"), NULL, 10);
}

int shim_supported(const s_fparams *params)
{
    const char c = params->clustering_method, e = params->distance_measure;
    if ((c != 's' && c != 'm' && c != 'a' && c != 'c') || (e != 'e' && e != 'b')) {
        fprintf(stderr, "! libfpocket_cuda implements clusterlib's four linkages on euclidean or city-block distances (-C s|m|a|c, -e e|b); got -C %c -e %c\n", c, e);
        return 0;
    }
    return 1;
}

/* step 1: flatten (the arrays stay alive inside `job` until shim_job_release) */
int shim_job_prepare(shim_job *job, s_pdb *pdb, s_fparams *params, s_pdb *pdb_w_lig)
{
    memset(job, 0, sizeof *job);
    const int same = (pdb_w_lig == pdb || pdb_w_lig == NULL);
    if (shim_flatten(pdb, params, 1, &job->A) || (!same && shim_flatten(pdb_w_lig, params, 0, &job->W))) {
        fprintf(stderr, "! search_pocket: out of memory\n");
        return -1;
    }
    const int n = pdb->natoms;
    shim_atoms *A = &job->A, *W = &job->W;
    fpk_structure *st = &job->st;
    st->natoms = n;
    st->x = A->f; st->y = A->f + n; st->z = A->f + 2 * n; st->radius = A->f + 3 * n; st->electroneg = A->f + 4 * n; st->bfactor = A->f + 5 * n;
    st->atom_id = A->i; st->res_id = A->i + n; st->aa_idx = A->aa; st->flags = A->fl;
    st->avg_bfactor = pdb->avg_bfactor; st->min_bfactor = pdb->min_bfactor; st->max_bfactor = pdb->max_bfactor;
    st->same_pdb = pdb_w_lig == pdb;                  /* asa.c:315 compares pointers: effective only when both are one object */
    if (!same) {
        const int nw = pdb_w_lig->natoms;
        st->natoms_w = nw;
        st->wx = W->f; st->wy = W->f + nw; st->wz = W->f + 2 * nw; st->wradius = W->f + 3 * nw; st->welectroneg = W->f + 4 * nw; st->wflags = W->fl;
    } else if (!st->same_pdb) st->same_pdb = 1;
    if (pdb->n_xlig_atoms > 0) {
        job->xl = (float *)malloc(sizeof(float) * 3 * (size_t)pdb->n_xlig_atoms);
        for (int k = 0; k < pdb->n_xlig_atoms; k++) { job->xl[3 * k] = pdb->xlig_x[k]; job->xl[3 * k + 1] = pdb->xlig_y[k]; job->xl[3 * k + 2] = pdb->xlig_z[k]; }
        st->n_xlig = pdb->n_xlig_atoms; st->xlig = job->xl;
    }
    return 0;
}
void shim_job_release(shim_job *job) { shim_free_atoms(&job->A); shim_free_atoms(&job->W); free(job->xl); memset(job, 0, sizeof *job); }

/* step 3: rebuild the reference's lists from a finished result (R is consumed) */
c_lst_pockets *shim_rebuild(s_pdb *pdb, s_fparams *params, s_pdb *pdb_w_lig, fpk_result *Rp, int rc, int do_asa)
{
    fpk_result R = *Rp;
    const int n = pdb->natoms;
    memset(Rp, 0, sizeof *Rp);
    if (rc == FPK_ERR_NO_SPHERES) {
        if (params->xlig_resnumber == -1 && params->xpocket_n == 0)        /* with -r / -P the reference returns NULL without a word (pocket.c:77) */
            fprintf(stderr, "Error in creating clustering tree, return NULL pointer...breaking up");      /* fpocket.c:137 */
        fpk_result_free(&R);
        return NULL;
    }
    if (rc != FPK_OK && rc != FPK_ERR_NO_POCKETS) {
        fprintf(stderr, "! Vertice calculation failed! (libfpocket_cuda %d: %s)\n", rc, fpk_last_error());
        fpk_result_free(&R);
        return NULL;
    }
    /* ---- s_lst_vvertice: ALL kept spheres, also those of dropped pockets (mdpbase.c:403-416, dpocket.c:239 rely on it) ---------- */
    s_lst_vvertice *lv = (s_lst_vvertice *)my_malloc(sizeof(s_lst_vvertice));
    const int ns = R.n_spheres;
    lv->nvert = ns; lv->qhullSize = R.n_tets;
    lv->vertices = (s_vvertice *)my_calloc(ns + 1, sizeof(s_vvertice));
    lv->pvertices = (s_vvertice **)my_calloc(ns + 1, sizeof(s_vvertice *));
    lv->tr = (int *)my_malloc((ns + 1) * sizeof(int));
    lv->n_h_tr = R.n_sites;
    lv->h_tr = (int *)my_malloc((R.n_sites + 1) * sizeof(int));
    for (int k = 0, j = 0; k < n; k++) if (strcmp(pdb->latoms[k].symbol, "H") != 0) lv->h_tr[j++] = k;
    for (int s = 0; s < ns; s++) {
        s_vvertice *v = lv->vertices + s;
        v->x = R.sph_xyzr[4 * s]; v->y = R.sph_xyzr[4 * s + 1]; v->z = R.sph_xyzr[4 * s + 2]; v->ray = R.sph_xyzr[4 * s + 3];
        for (int j = 0; j < 4; j++) v->neigh[j] = pdb->latoms + R.sph_neigh[4 * s + j];
        v->bary[0] = R.sph_bary[3 * s]; v->bary[1] = R.sph_bary[3 * s + 1]; v->bary[2] = R.sph_bary[3 * s + 2];
        v->type = R.sph_type[s];
        v->sort_x = -1; v->seen = 0; v->electrostatic_energy = 0.0f; v->apol_neighbours = 0;
        v->qhullId = s; v->id = n + 1 + s;           /* the reference derives id from qhull's facet index (voronoi.c:488); ours is canonical */
        v->resid = pdb->n_xlig_atoms > 0 ? 1 : -1;
        lv->tr[s] = s;
        lv->pvertices[s] = v;
    }
    /* ---- pockets in final rank order (fpocket.c:191-213) ------------------------------------------------------------------------- */
    c_lst_pockets *pockets = c_lst_pockets_alloc();
    pockets->vertices = lv;
    for (int r = 0; r < R.n_pockets; r++) {
        const int c = R.rank_to_cluster[r];
        const fpk_desc *g = R.desc + c;
        s_pocket *p = alloc_pocket();
        p->v_lst = c_lst_vertices_alloc();
        for (int q = R.cl_off[c]; q < R.cl_off[c + 1]; q++) {
            s_vvertice *v = lv->vertices + R.cl_sph[q];
            v->resid = r + 1;                               /* reIndexPockets, refine.c:223 */
            c_lst_vertices_add_last(p->v_lst, v);
        }
        p->size = (int)p->v_lst->n_vertices;
        p->rank = r + 1;
        p->score = g->score;
        p->bary[0] = g->bary[0]; p->bary[1] = g->bary[1]; p->bary[2] = g->bary[2];
        shim_fill_desc(p->pdesc, g);
        c_lst_pockets_add_last(pockets, p, g->nAlphaApol, g->nAlphaPol);
        if (pdb_w_lig) set_pocket_contacted_lig_name(p, pdb_w_lig);     /* pocket.c:535-560 */
        shim_chain_info(p, pdb_w_lig ? pdb_w_lig : pdb);                /* pocket.c:617 hands pdb_w_lig to set_descriptors */
    }
    /* ---- per-atom ASA side effects (asa.c:337,392-396), printed by writepocket.c:664-669 ------------------------------------------- */
    if (do_asa && R.atom_fx)
        for (int k = 0; k < n; k++) {
            s_atm *a = pdb->latoms + k;
            a->abpa = (int)R.atom_fx[4 * k]; a->a0 = R.atom_fx[4 * k + 1]; a->dA = R.atom_fx[4 * k + 2];
            a->abpa_sourrounding_prob = R.atom_fx[4 * k + 3];
        }
    fpk_result_free(&R);
    if (params->fpocket_running && params->flag_do_grid_calculations && params->topology_path[0])
        calculate_pocket_energy_grids(pockets, params, pdb);               /* fpocket.c:219: stays the reference's code */
    return pockets;
}

c_lst_pockets *search_pocket(s_pdb *pdb, s_fparams *params, s_pdb *pdb_w_lig)
{
    if (!pdb || !params) return NULL;
    if (!shim_supported(params)) return NULL;
    shim_job job;
    if (shim_job_prepare(&job, pdb, params, pdb_w_lig)) return NULL;
    fpk_params P;
    shim_params(params, &P);
    fpk_result R;
    memset(&R, 0, sizeof R);
    int rc = fpk_detect(&job.st, &P, &R);
    shim_job_release(&job);
    return shim_rebuild(pdb, params, pdb_w_lig, &R, rc, P.do_asa_and_volume);
}
