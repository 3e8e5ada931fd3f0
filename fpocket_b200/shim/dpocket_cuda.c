/*
 * dpocket on libfpocket_cuda: a drop-in
 *
 *     void desc_pocket(char fcomplexe[], const char ligname[], s_dparams *par, FILE *f[3])      (reference src/dpocket.c:165-300)
 *
 * The reference runs search_pocket, then walks x-sorted arrays on the host for the ligand's neighbourhood (neighbor.c), computes a
 * full descriptor record for the explicit pocket (set_descriptors: ASA, Monte-Carlo volume, hull - dpocket.c:341-350) and four
 * criteria per pocket. Here ONE fpk_detect() call with the ligand attached (fpk_structure.n_lig) returns all of it: pockets,
 * explicit-pocket spheres + record (xdesc), interface atoms (either definition, dparams.h:41-44), ovlp / dst / c4 / c5 per pocket and the
 * ligand's own volume (get_mol_volume_ptr). Loading the complex, write_out_fpocket and write_pocket_desc stay the reference's code.
 * The rest of dpocket.c is linked as it is (shim/Makefile weakens the reference object's own desc_pocket so that this one is called).
 * Interface method 2 (-e, atoms within a distance of the ligand) is not supported: refused with a message.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dpocket.h"
#include "fpocket_cuda.h"
#include "shim.h"

static int name_key(const char *name)
{
    unsigned char b[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4 && name[i]; i++) b[i] = (unsigned char)name[i];
    return (int)((unsigned)b[0] | ((unsigned)b[1] << 8) | ((unsigned)b[2] << 16) | ((unsigned)b[3] << 24));
}

/* numeric fields of fpk_desc -> s_desc (the record dpocket prints, dpocket.h:44-70) */
static void fill_desc(s_desc *d, const fpk_desc *g)
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

void desc_pocket(char fcomplexe[], const char ligname[], s_dparams *par, FILE *f[3])
{
    if (!shim_supported(par->fpar)) return;
    /* dpocket.c:186-203 */
    s_pdb *pdb_l = rpdb_open(fcomplexe, ligname, M_KEEP_LIG, 0, par->fpar);
    s_pdb *pdb_nl = rpdb_open(fcomplexe, ligname, M_DONT_KEEP_LIG, 0, par->fpar);
    if (!pdb_l || !pdb_nl || pdb_l->natm_lig <= 0) {
        if (pdb_l && pdb_l->natm_lig <= 0) fprintf(stdout, "ERROR - No ligand %s found in %s.\n", ligname, fcomplexe);
        else fprintf(stdout, "ERROR - PDB file %s could not be opened\n", fcomplexe);
        return;
    }
    rpdb_read(pdb_l, ligname, M_KEEP_LIG, 0, par->fpar);
    rpdb_read(pdb_nl, ligname, M_DONT_KEEP_LIG, 0, par->fpar);
    s_atm **lig = pdb_l->latm_lig;
    const int nal = pdb_l->natm_lig;

    shim_job job;
    if (shim_job_prepare(&job, pdb_nl, par->fpar, pdb_l)) return;
    int *la = (int *)malloc(sizeof(int) * nal), *lm = (int *)malloc(sizeof(int) * nal), *nk = (int *)malloc(sizeof(int) * (pdb_nl->natoms + 1));
    char chain_tmp[2];
    int resnumber_tmp = lig[0]->res_id, nmol = 1;
    strncpy(chain_tmp, lig[0]->chain, 2);
    for (int j = 0; j < nal; j++) {                                  /* dpocket.c:211-222 / neighbor.c:619-637 */
        if (strncmp(chain_tmp, lig[j]->chain, 2) != 0 || resnumber_tmp != lig[j]->res_id) {
            nmol++; strncpy(chain_tmp, lig[j]->chain, 2); resnumber_tmp = lig[j]->res_id;
        }
        la[j] = (int)(lig[j] - pdb_l->latoms); lm[j] = nmol - 1;
    }
    for (int k = 0; k < pdb_nl->natoms; k++) nk[k] = name_key(pdb_nl->latoms[k].name);
    job.st.n_lig = nal; job.st.lig_atoms = la; job.st.lig_mol = lm; job.st.atom_name_key = nk;
    fpk_params P;
    shim_params(par->fpar, &P);
    P.lig_interface_dist = par->interface_dist_crit;
    P.lig_interface_method = par->interface_method;                                 /* dparams.h:41-44, dpocket.c:331-336 */
    fpk_result R;
    memset(&R, 0, sizeof R);
    const int rc = fpk_detect(&job.st, &P, &R);
    /* keep what the pocket list cannot carry, then let the shim rebuild the reference's lists (consumes R) */
    const int np = (rc == FPK_OK) ? R.n_pockets : 0;
    float *crit = (float *)malloc(sizeof(float) * (4 * (size_t)np + 4));
    for (int q = 0; q < np; q++) {
        crit[4 * q] = R.pk_ovlp ? R.pk_ovlp[q] : 0.0f; crit[4 * q + 1] = R.pk_dst ? R.pk_dst[q] : 0.0f;
        crit[4 * q + 2] = R.pk_c4 ? R.pk_c4[q] : 0.0f; crit[4 * q + 3] = R.pk_c5 ? R.pk_c5[q] : 0.0f;
    }
    s_desc *edesc = allocate_s_desc();
    if (R.xdesc) fill_desc(edesc, R.xdesc);
    const float lig_volume = R.lig_volume;
    shim_job_release(&job);
    free(la); free(lm); free(nk);
    c_lst_pockets *pockets = shim_rebuild(pdb_nl, par->fpar, pdb_l, &R, rc, P.do_asa_and_volume);
    if (pockets == NULL) {
        fprintf(stdout, "ERROR - No pocket found for %s\n", fcomplexe);
        return;
    }
    write_out_fpocket(pockets, pdb_nl, fcomplexe);                  /* dpocket.c:233 */
    const float vol = lig_volume;                                                   /* get_mol_volume_ptr (atom.c:156) on the GPU (K6) */
    write_pocket_desc(fcomplexe, ligname, edesc, vol, 100.0, 0.0, 1.0, 1.0, f[0]);
    node_pocket *cur = pockets->first;
    int q = 0;
    while (cur && q < np) {
        const float ovlp = crit[4 * q], dst = crit[4 * q + 1], c4 = crit[4 * q + 2], c5 = crit[4 * q + 3];
        write_pocket_desc(fcomplexe, ligname, cur->pocket->pdesc, vol, ovlp, dst, c4, c5, (ovlp > 40.0 || dst < 4.0) ? f[1] : f[2]);
        cur = cur->next;
        q++;
    }
    free(crit);
    c_lst_pocket_free(pockets);
    my_free(edesc);
    free_pdb_atoms(pdb_l);
    free_pdb_atoms(pdb_nl);
}
