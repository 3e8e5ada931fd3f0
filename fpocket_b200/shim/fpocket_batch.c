/*
 * fpocket_cuda_batch: the "-F list" loop of the reference (src/fpmain.c:61-76: one process_pdb() per file) as ONE
 * fpk_detect_batch() call per group of files. Reader (rpdb.c / read_mmcif.c), option parser (fparams.c) and writers
 * (fpout.c, writepocket.c, write_visu.c) are the reference's objects, unchanged; only the detection between them is the
 * CUDA library. Built by fpocket_b200/shim/Makefile into build/fpocket_cuda_batch.
 *
 *   fpocket_cuda_batch list.txt [any fpocket option except -f / -F, e.g. -m 3.0 -M 6.0 -D 2.0]
 *
 * list.txt: one PDB / mmCIF path per line. Output: <name>_out/ next to every input, exactly what `fpocket -f <name>` writes.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fpocket.h"
#include "fpout.h"
#include "read_mmcif.h"
#include "fpocket_cuda.h"
#include "shim.h"

#define GROUP 256       /* files parsed, detected and written together (bounds the host memory held by the s_pdb copies) */

static s_pdb *load(const char *path, int keep_lig, s_fparams *params)
{
    /* process_pdb (fpmain.c:143-173) */
    s_pdb *pdb;
    if (strstr(path, ".cif")) {
        pdb = open_mmcif((char *)path, NULL, keep_lig, params->model_number, params);
        if (pdb) read_mmcif(pdb, NULL, keep_lig, params->model_number, params);
    } else {
        pdb = rpdb_open((char *)path, NULL, keep_lig, params->model_number, params);
        if (pdb) rpdb_read(pdb, NULL, keep_lig, params->model_number, params);
    }
    return pdb;
}

int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: %s list.txt [fpocket options]\n", argv[0]); return 2; }
    FILE *fl = fopen(argv[1], "r");
    if (!fl) { perror(argv[1]); return 2; }
    char **files = NULL, line[4096];
    int nfiles = 0;
    while (fgets(line, sizeof line, fl)) {
        size_t L = strlen(line);
        while (L && (line[L - 1] == '\n' || line[L - 1] == '\r' || line[L - 1] == ' ')) line[--L] = 0;
        if (!L) continue;
        files = (char **)realloc(files, sizeof(char *) * (nfiles + 1));
        files[nfiles++] = strdup(line);
    }
    fclose(fl);
    if (!nfiles) { fprintf(stderr, "empty list\n"); return 2; }
    /* the reference's own option parser, fed "-f <first file>" + the user's options */
    char **args = (char **)malloc(sizeof(char *) * (argc + 3));
    int na = 0;
    args[na++] = argv[0]; args[na++] = "-f"; args[na++] = files[0];
    for (int i = 2; i < argc; i++) args[na++] = argv[i];
    args[na] = NULL;
    s_fparams *params = get_fpocket_args(na, args);
    if (!params) { fprintf(stderr, "! bad fpocket options\n"); return 2; }
    params->fpocket_running = 1;
    if (!shim_supported(params)) return 3;
    if (fpk_init(NULL, 0) != FPK_OK) { fprintf(stderr, "! %s\n", fpk_last_error()); return 4; }
    fpk_params P;
    shim_params(params, &P);
    int done = 0, npockets = 0, failed = 0;
    for (int g0 = 0; g0 < nfiles; g0 += GROUP) {
        const int ng = nfiles - g0 < GROUP ? nfiles - g0 : GROUP;
        s_pdb **pdb = (s_pdb **)calloc(ng, sizeof(s_pdb *)), **pdbw = (s_pdb **)calloc(ng, sizeof(s_pdb *));
        shim_job *jobs = (shim_job *)calloc(ng, sizeof(shim_job));
        fpk_structure *in = (fpk_structure *)calloc(ng, sizeof(fpk_structure));
        fpk_result *out = (fpk_result *)calloc(ng, sizeof(fpk_result));
        int *idx = (int *)malloc(sizeof(int) * ng), nb = 0;
        for (int k = 0; k < ng; k++) {
            const char *path = files[g0 + k];
            strncpy(params->pdb_path, path, M_MAX_PDB_NAME_LEN - 1);
            pdb[k] = load(path, M_DONT_KEEP_LIG, params);
            pdbw[k] = pdb[k] ? load(path, M_KEEP_LIG, params) : NULL;
            if (!pdb[k] || !pdbw[k]) { fprintf(stderr, "! Structure reading failed: %s\n", path); failed++; continue; }
            /* create_coord_grid (fpmain.c:173) only serves the energy grids (energy.c:164); it makes ~1e5 my_malloc blocks per structure and
             * the reference's allocator frees by linear search of ALL live blocks (memhandler.c remove_bloc), so it is built on demand only */
            if (params->flag_do_grid_calculations && params->topology_path[0]) create_coord_grid(pdb[k]);
            if (shim_job_prepare(&jobs[k], pdb[k], params, pdbw[k])) { failed++; continue; }
            in[nb] = jobs[k].st;
            idx[nb++] = k;
        }
        if (nb > 0 && fpk_detect_batch(in, nb, &P, out) != FPK_OK) { fprintf(stderr, "! fpk_detect_batch: %s\n", fpk_last_error()); return 5; }
        for (int q = 0; q < nb; q++) {
            const int k = idx[q];
            strncpy(params->pdb_path, files[g0 + k], M_MAX_PDB_NAME_LEN - 1);
            c_lst_pockets *pockets = shim_rebuild(pdb[k], params, pdbw[k], &out[q], out[q].status, P.do_asa_and_volume);
            if (pockets) {
                npockets += (int)pockets->n_pockets;
                write_out_fpocket(pockets, pdb[k], files[g0 + k]);         /* fpmain.c:195 */
                c_lst_pocket_free(pockets);
            } else printf("no pockets found: %s\n", files[g0 + k]);
            done++;
        }
        for (int k = 0; k < ng; k++) {
            shim_job_release(&jobs[k]);
            if (pdb[k]) free_pdb_atoms(pdb[k]);
            if (pdbw[k]) free_pdb_atoms(pdbw[k]);
        }
        free(pdb); free(pdbw); free(jobs); free(in); free(out); free(idx);
    }
    printf("fpocket_cuda_batch: %d structures, %d pockets written, %d failed\n", done, npockets, failed);
    fpk_shutdown();
    return failed ? 1 : 0;
}
