/* Shared by search_pocket_cuda.c (the drop-in search_pocket) and fpocket_batch.c (many structures per fpk_detect_batch call). */
#ifndef FPK_SHIM_H
#define FPK_SHIM_H
#include "fpocket.h"
#include "fpocket_cuda.h"

typedef struct {
    float *f;        /* 6 arrays of n floats: x y z radius electroneg bfactor */
    int32_t *i;      /* 2 arrays of n ints: id res_id */
    int8_t *aa;
    uint8_t *fl;
} shim_atoms;

typedef struct shim_job {       /* one structure flattened for the library; the arrays live until shim_job_release */
    shim_atoms A, W;
    float *xl;
    fpk_structure st;
} shim_job;

int shim_supported(const s_fparams *params);                                   /* -C s -e e only; prints why otherwise */
void shim_params(const s_fparams *params, fpk_params *P);                      /* s_fparams -> fpk_params */
int shim_job_prepare(shim_job *job, s_pdb *pdb, s_fparams *params, s_pdb *pdb_w_lig);
void shim_job_release(shim_job *job);
/* rebuilds s_lst_vvertice + c_lst_pockets from a finished result (consumed) and writes the per-atom ASA side effects back */
c_lst_pockets *shim_rebuild(s_pdb *pdb, s_fparams *params, s_pdb *pdb_w_lig, fpk_result *R, int rc, int do_asa);
#endif
