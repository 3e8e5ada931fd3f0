/*
 * mdpocket on libfpocket_cuda: a drop-in
 *
 *     void mdpocket_detect(s_mdparams *par)                      (reference src/mdpocket.c:71-340)
 *
 * The reference detects pockets frame by frame and accumulates two grids on the host (search_pocket + calculate_md_dens_grid +
 * update_md_grid per frame, mdpocket.c:142-175,235-271). Here the frames are read the same way (a list of PDB snapshots, or a
 * trajectory through the molfile plugins) but handed to the GPU in blocks: fpk_md_grid_from_frame (init_md_grid from frame 1,
 * mdpbase.c:444-502), fpk_md_push_frames (detection + both grid updates for every frame), fpk_md_fetch_grids. What follows the
 * frame loop - normalize_grid, project_grid_on_atoms, write_first_bfactor_density, write_md_grid (mdpbase.c:301-354, mdpout.c) -
 * is the reference's code, called unchanged on s_mdgrid objects filled from the fetched grids.
 * The rest of mdpocket.c (characterize mode, mdprocess_pdb, extract_wanted_vertices) is linked as it is: shim/Makefile compiles the
 * reference file with -Dmdpocket_detect=mdpocket_detect_reference.
 * Not supported here (refused with a message, never a silent CPU path): user-defined grids and grid spacings other than 1 A.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "mdpocket.h"
#include "fpocket_cuda.h"
#include "shim.h"

static molfile_plugin_t *md_api;
static int md_register_cb(void *v, vmdplugin_t *p) { (void)v; md_api = (molfile_plugin_t *)p; return 0; }

#define MD_BLOCK 2368       /* frames handed to the library at a time */

static s_mdgrid *grid_from_array(const float *vals, const float origin[3], const int dims[3])
{
    /* same allocation pattern as init_md_grid (mdpbase.c:446-502) so that free_md_grid() releases it */
    s_mdgrid *g = (s_mdgrid *)malloc(sizeof(s_mdgrid));
    g->resolution = M_MDP_GRID_RESOLUTION;
    g->nx = dims[0]; g->ny = dims[1]; g->nz = dims[2];
    g->n_snapshots = 0;
    g->gridvalues = (float ***)malloc(sizeof(float **) * g->nx);
    for (int x = 0; x < g->nx; x++) {
        g->gridvalues[x] = (float **)malloc(sizeof(float *) * g->ny);
        for (int y = 0; y < g->ny; y++) {
            g->gridvalues[x][y] = (float *)malloc(sizeof(float) * g->nz);
            memcpy(g->gridvalues[x][y], vals + ((size_t)x * g->ny + y) * g->nz, sizeof(float) * g->nz);
        }
    }
    g->origin = (float *)malloc(sizeof(float) * 3);
    for (int k = 0; k < 3; k++) g->origin[k] = origin[k];
    return g;
}

/* Two frame buffers: while the library works on one block (fpk_md_push_frames blocks until the GPU is done), the reader fills the other. */
typedef struct { fpk_md *md; int natoms, have_grid, nframes; float origin[3]; int dims[3]; float *buf, *bufs[2]; int cur, nbuf; const int *slot;
                 pthread_t th; int th_live, th_rc, push_n; float *push_buf; } md_run;

static void *md_push_thread(void *arg)
{
    md_run *r = (md_run *)arg;
    r->th_rc = fpk_md_push_frames(r->md, r->push_buf, r->push_n);
    return NULL;
}
static int md_wait(md_run *r)
{
    if (!r->th_live) return 0;
    pthread_join(r->th, NULL);
    r->th_live = 0;
    if (r->th_rc) fprintf(stderr, "! libfpocket_cuda (%d): %s\n", r->th_rc, fpk_last_error());
    return r->th_rc;
}
static int md_flush(md_run *r)
{
    if (!r->nbuf) return md_wait(r);
    int rc = md_wait(r);                     /* the previous block is done: its buffer is the one the reader gets next */
    if (!rc && !r->have_grid) {              /* frame 1 defines the grid (mdpocket.c:151-154) */
        rc = fpk_md_grid_from_frame(r->md, r->buf, r->origin, r->dims);
        if (!rc) rc = fpk_md_set_grid(r->md, r->origin, r->dims);
        r->have_grid = 1;
        if (rc) fprintf(stderr, "! libfpocket_cuda (%d): %s\n", rc, fpk_last_error());
    }
    if (!rc) {
        r->push_buf = r->buf; r->push_n = r->nbuf; r->th_rc = 0;
        if (pthread_create(&r->th, NULL, md_push_thread, r) == 0) r->th_live = 1;
        else rc = fpk_md_push_frames(r->md, r->buf, r->nbuf);
    }
    r->nframes += r->nbuf;
    r->nbuf = 0;
    r->cur ^= 1;
    r->buf = r->bufs[r->cur];                /* the reader goes on in the other buffer */
    return rc;
}

void mdpocket_detect(s_mdparams *par)
{
    if (!par) return;
    if (!grid_user_data_missing(par) || par->grid_spacing > 0.0) {
        fprintf(stderr, "! mdpocket (CUDA): user-defined grids / grid spacings are not supported by libfpocket_cuda\n");
        return;
    }
    par->fpar->flag_do_asa_and_volume_calculations = 0;          /* mdpocket.c:89 */
    if (!shim_supported(par->fpar)) return;
    FILE *fout1 = fopen(par->f_freqdx, "w"), *fout2 = fopen(par->f_densdx, "w"), *fout3 = fopen(par->f_freqiso, "w"),
         *fout4 = fopen(par->f_densiso, "w"), *fout5 = fopen(par->f_appdb, "w");
    if (!(fout1 && fout2 && fout3 && fout4 && fout5)) { fprintf(stdout, "! Output file couldn't be opened.\n"); return; }

    md_run R;
    memset(&R, 0, sizeof R);
    s_pdb *topo = NULL;
    shim_job job;
    fpk_params P;
    void *h_in = NULL;
    molfile_timestep_t ts_in;
    ts_in.coords = NULL;
    int traj = par->nfiles ? 0 : 1, inatoms = 0;

    /* ---- the topology: frame-independent atom properties --------------------------------------------------------------------- */
    if (!traj) {
        topo = open_pdb_file(par->fsnapshot[0], par);
        if (topo) rpdb_read(topo, NULL, M_DONT_KEEP_LIG, 0, par->fpar);              /* mdprocess_pdb, mdpocket.c:839 */
    } else {
        if (!strncmp(par->traj_format, "dcd", 3)) { molfile_dcdplugin_init(); molfile_dcdplugin_register(NULL, md_register_cb); }
        if (!strncmp(par->traj_format, "crd", 3)) { molfile_crdplugin_init(); molfile_crdplugin_register(NULL, md_register_cb); }
        if (!strncmp(par->traj_format, "dtr", 3)) { molfile_dtrplugin_init(); molfile_dtrplugin_register(NULL, md_register_cb); }
        if (!strncmp(par->traj_format, "xtc", 3) || !strncmp(par->traj_format, "trr", 3)) { molfile_gromacsplugin_init(); molfile_gromacsplugin_register(NULL, md_register_cb); }
        topo = open_pdb_file(par->fpar->pdb_path, par);
        if (topo) rpdb_read(topo, NULL, M_KEEP_LIG, 0, par->fpar);                    /* mdpocket.c:193 */
        if (!md_api || !topo) { fprintf(stderr, "Error: trajectory format / topology could not be set up\n"); return; }
        h_in = md_api->open_file_read(par->f_traj, par->traj_format, &inatoms);
        if (!h_in) { fprintf(stderr, "Error: could not open file '%s' for reading.\n", par->f_traj); return; }
        if (inatoms != topo->natoms) {
            if (strncmp(par->traj_format, "crd", 3)) fprintf(stderr, "Number of atoms in topology and MD trajectory do not coincide : %d vs %d\n", inatoms, topo->natoms);
            inatoms = topo->natoms;
        }
        ts_in.coords = (float *)malloc(3 * (size_t)inatoms * sizeof(float));
    }
    if (!topo) { fprintf(stderr, "! PDB reading failed on snapshot 1\n"); return; }
    const int na = topo->natoms;
    if (shim_job_prepare(&job, topo, par->fpar, topo)) return;                          /* search_pocket(pdb, params, pdb), mdpocket.c:844 */
    shim_params(par->fpar, &P);
    R.md = fpk_md_begin(&job.st, &P, par->flag_scoring);
    if (!R.md) { fprintf(stderr, "! libfpocket_cuda: %s\n", fpk_last_error()); return; }
    R.natoms = na;
    R.bufs[0] = (float *)malloc(sizeof(float) * 3 * (size_t)na * MD_BLOCK);
    R.bufs[1] = (float *)malloc(sizeof(float) * 3 * (size_t)na * MD_BLOCK);
    R.buf = R.bufs[0];
    int failed = 0;

    /* ---- the frames ------------------------------------------------------------------------------------------------------------- */
    if (!traj) {
        for (int i = 0; i < par->nfiles && !failed; i++) {
            s_pdb *cp = i == 0 ? topo : open_pdb_file(par->fsnapshot[i], par);
            if (cp && i > 0) rpdb_read(cp, NULL, M_DONT_KEEP_LIG, 0, par->fpar);
            if (!cp || cp->natoms != na) { fprintf(stderr, "! PDB reading failed on snapshot %d (or its atom count differs from snapshot 1)\n", i + 1); failed = 1; break; }
            float *dst = R.buf + 3 * (size_t)na * R.nbuf;
            for (int j = 0; j < na; j++) { dst[3 * j] = cp->latoms[j].x; dst[3 * j + 1] = cp->latoms[j].y; dst[3 * j + 2] = cp->latoms[j].z; }
            if (i > 0) free_pdb_atoms(cp);
            if (++R.nbuf == MD_BLOCK || R.nframes + R.nbuf == 1) failed = md_flush(&R);
        }
    } else {
        int rc = md_api->read_next_timestep(h_in, inatoms, &ts_in);
        while (rc != -1 && !failed) {
            /* the reference writes timestep atom j to latoms_p[j] (mdpocket.c:238-243); the library wants latoms order */
            float *dst = R.buf + 3 * (size_t)na * R.nbuf;
            for (int j = 0; j < na; j++) {
                const size_t a = (size_t)(topo->latoms_p[j] - topo->latoms);
                dst[3 * a] = ts_in.coords[3 * j]; dst[3 * a + 1] = ts_in.coords[3 * j + 1]; dst[3 * a + 2] = ts_in.coords[3 * j + 2];
            }
            if (++R.nbuf == MD_BLOCK || R.nframes + R.nbuf == 1) failed = md_flush(&R);
            rc = md_api->read_next_timestep(h_in, inatoms, &ts_in);
        }
        md_api->close_file_read(h_in);
    }
    if (!failed) failed = md_flush(&R);
    if (md_wait(&R)) failed = 1;
    const int n_snapshots = R.nframes;
    fprintf(stdout, "<mdpocket> %d snapshots analysed on the GPU\n", n_snapshots);
    if (failed || !R.have_grid || n_snapshots == 0) { fpk_md_end(R.md); return; }

    /* ---- grids back into the reference's objects; everything after the frame loop is the reference's (mdpocket.c:276-330) ----------- */
    const size_t ncell = (size_t)R.dims[0] * R.dims[1] * R.dims[2];
    float *dens = (float *)malloc(sizeof(float) * ncell), *freq = (float *)malloc(sizeof(float) * ncell);
    if (fpk_md_fetch_grids(R.md, dens, freq)) { fprintf(stderr, "! libfpocket_cuda: %s\n", fpk_last_error()); return; }
    fpk_md_end(R.md);
    s_mdgrid *freqgrid = grid_from_array(freq, R.origin, R.dims), *densgrid = grid_from_array(dens, R.origin, R.dims);
    free(dens); free(freq); free(R.bufs[0]); free(R.bufs[1]);
    freqgrid->n_snapshots = n_snapshots;
    densgrid->n_snapshots = n_snapshots;
    normalize_grid(freqgrid, n_snapshots);
    normalize_grid(densgrid, n_snapshots);
    s_pdb *cpdb = open_pdb_file(traj ? par->fpar->pdb_path : par->fsnapshot[0], par);
    rpdb_read(cpdb, NULL, M_DONT_KEEP_LIG, 0, par->fpar);
    project_grid_on_atoms(densgrid, cpdb);
    write_first_bfactor_density(fout5, cpdb);
    free_pdb_atoms(cpdb);
    write_md_grid(freqgrid, fout1, fout3, par, M_MDP_DEFAULT_ISO_VALUE_FREQ);
    write_md_grid(densgrid, fout2, fout4, par, M_MDP_DEFAULT_ISO_VALUE_DENS);
    fclose(fout1); fclose(fout2); fclose(fout3); fclose(fout4); fclose(fout5);
    free_md_grid(freqgrid);
    free_md_grid(densgrid);
    shim_job_release(&job);
    free_pdb_atoms(topo);
    if (ts_in.coords) free(ts_in.coords);
}
