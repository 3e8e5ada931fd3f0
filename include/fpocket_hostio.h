/*
 * libfpocket_hostio — native host I/O either side of the CUDA hot path (SURVEY.md par. 8 row f1: "batched host I/O").
 *
 * The reference reads every input four times (rpdb_open + rpdb_read for pdb and for pdb_w_lig, src/fpmain.c:143-164) and
 * writes ~2P+9 small files per structure through stdio and system() (src/fpout.c:55-134); neither is thread-safe
 * (globals in memhandler.c, fparams.c:60). Once detection costs ~0.1 ms per structure that host code is the whole
 * wall time (21 structures/s per process with the reference's reader and writers around libfpocket_cuda).
 * This library restates the two ends for the default PDB route, re-entrantly, so that a pool of host threads keeps a GPU fed:
 *
 *   fpkio_read_pdb      = rpdb_open + rpdb_read, for BOTH atom lists in one pass over the file (src/rpdb.c:785-1478),
 *                         + the flattening of s_pdb into fpk_structure (fpocket_b200/shim/search_pocket_cuda.c shim_job_prepare)
 *   fpkio_write_output  = write_out_fpocket (src/fpout.c:55-134) with everything it calls: write_visualization
 *                         (src/write_visu.c:56-290), write_pockets_single_pdb / _pqr, write_each_pocket, write_pocket_pqr / _pdb
 *                         (src/writepocket.c:262-305,381-416,475-640), write_out_fpocket_info_file (src/fpout.c:151-213),
 *                         the record formats of src/writepdb.c:78-216 and src/voronoi.c:1437-1557
 *   fpkio_run_list      = the "-F list" loop (src/fpmain.c:61-76) as a three-stage pipeline: reader pool -> one
 *                         fpk_detect_batch call per group of files -> writer pool
 *
 * Scope: PDB input, PDB- and mmCIF-mode output (-w; PDB is the reference's default for a .pdb input, fparams.c:397-407), chain filters -c / -k, explicit ligand -r,
 * explicit pocket -P / -u, chain as ligand -a, model -l, dpocket's ligand mode. Not restated (use build/fpocket_cuda_batch, which keeps the reference's reader and writers):
 * mmCIF input, -d, -x. fpkio_* refuse those explicitly (FPKIO_ERR_UNSUPPORTED); nothing falls back silently.
 *
 * The library has no CUDA dependency of its own: the detection call is handed in (fpk_detect_batch in the product,
 * build/fpocket_fast), so the CPU tests can drive the same pipeline with the oracle as the detector.
 */
#ifndef FPOCKET_HOSTIO_H
#define FPOCKET_HOSTIO_H
#include "fpocket_cuda.h"
#ifdef __cplusplus
extern "C" {
#endif

#define FPKIO_OK               0
#define FPKIO_ERR_OPEN        (-1)   /* "! File %s does not exist" (rpdb.c:813) */
#define FPKIO_ERR_NO_ATOMS    (-2)   /* "! File '%s' contains no atoms..." (rpdb.c:983) */
#define FPKIO_ERR_UNSUPPORTED (-3)   /* input or option outside the restated route (see above) */
#define FPKIO_ERR_WRITE       (-4)   /* an output directory or file could not be created */
#define FPKIO_ERR_BAD_ARG     (-5)
#define FPKIO_ERR_NO_LIGAND   (-6)   /* dpocket: "ERROR - No ligand %s found in %s." (dpocket.c:189) */

/* what the reference's s_fparams carries for the reader (fparams.h:182-227); zero-initialise for the defaults */
typedef struct fpkio_read_opts {
    int32_t n_chains;            /* -c / -k: number of chain names (<= 20)                     */
    char chains[20][21];         /* chain names (compared with the one-character chain column) */
    int32_t chains_are_kept;     /* 0: -c (drop these chains), 1: -k (keep only these)         */
    char ligand[8];              /* dpocket: residue name of the known ligand (rpdb_open/rpdb_read with ligan != NULL, dpocket.c:186-203): its atoms
                                    are kept in pdb_w_lig only and flagged as the ligand; other HETATM groups count only when listed. "" = fpocket */
    /* -r resnumber:resname:chain (fparams.c:241-262): the coordinates of that residue's records become pdb->xlig_* (rpdb.c:893-903,962-973):
     * testVvertice then keeps only spheres that reach one of them (voronoi.c:1047-1066) and clustering is skipped (fpocket.c:114) */
    int32_t has_xlig, xlig_resnumber;
    char xlig_resname[8], xlig_chain;
    /* -P res:ins:chain.res:ins:chain... (fparams.c:263-300): atoms of these residues get FPK_ATOM_XPOCKET (voronoi.c:934-949) */
    int32_t n_xpocket;
    uint16_t xp_res[64];
    char xp_ins[64], xp_chain[64];
    int32_t n_lig_chains;        /* -a: chains treated as the ligand (fparams.c:186-201): their records become pdb->xlig_* and leave pdb; the reference's
                                    two passes disagree at chain boundaries (rpdb_read tests the PREVIOUS record's chain, rpdb.c:1079,1225) - reproduced */
    char lig_chains[20][21];
    char write_mode;             /* -w for fpkio_run_list: 0 or 'p' = PDB set, 'm' = mmCIF set, 'b' = both (fparams.c:167-183) */
    int32_t model_number;        /* -l: > 0 reads only the records of that MODEL ... ENDMDL block when the file has MODEL records (rpdb.c:817-840,972-976) */
} fpkio_read_opts;

/* the reference's parsing of the -r and -P arguments, quirks included (strtok skips empty fields; the switch of fparams.c:285-293 falls through) */
int  fpkio_parse_custom_ligand(const char *arg, fpkio_read_opts *o);
int  fpkio_parse_custom_pocket(const char *arg, fpkio_read_opts *o);

typedef struct fpkio_file fpkio_file;      /* one parsed input: both atom lists + the flat arrays handed to the library */

int  fpkio_read_pdb(const char *path, const fpkio_read_opts *opts, fpkio_file **out);
/* same, from a buffer holding the file's bytes (the path is only kept for messages) */
int  fpkio_read_pdb_mem(const char *path, const char *data, size_t size, const fpkio_read_opts *opts, fpkio_file **out);
void fpkio_file_free(fpkio_file *f);
const fpk_structure *fpkio_structure(const fpkio_file *f);       /* valid until fpkio_file_free */
int  fpkio_natoms(const fpkio_file *f, int with_lig);            /* pdb->natoms / pdb_w_lig->natoms */
int  fpkio_nhetatm(const fpkio_file *f, int with_lig);           /* pdb->nhetatm */
/* text fields of atom k of pdb (with_lig = 0) or pdb_w_lig (1), as the reference's s_atm holds them; for tests.
 * type[8] name[8] res_name[8] chain[4] symbol[4]; ints: id res_id charge; chars: insert aloc */
int  fpkio_atom_text(const fpkio_file *f, int with_lig, int k, char *type, char *name, char *res_name, char *chain, char *symbol,
                     int32_t *id_resid_charge, char *insert_aloc);

/* write_out_fpocket(pockets, pdb, pdbname): <dir>/<code>_out/{<code>_out.pdb, _info.txt, _pockets.pqr, _VMD.sh, .tcl, _PYMOL.sh,
 * .pml, pockets/pocketK_vert.pqr, pockets/pocketK_atm.pdb}. A result with no pocket left (FPK_ERR_NO_POCKETS) is written too, as the
 * reference does with the empty list search_pocket returns (fpocket.c:225); only a failed clustering tree (FPK_ERR_NO_SPHERES) makes the
 * caller print "no pockets found" (fpmain.c:204). Returns the number of pockets written or FPKIO_ERR_*. bytes_out (may be NULL) += bytes. */
int  fpkio_write_output(const fpkio_file *f, const char *pdb_path, const fpk_result *r, long long *bytes_out);
/* the reference's global write_mode (fparams.c:60,397-407): 'p' = the PDB set above (what fpocket chooses for a .pdb input); 'd' = undecided:
 * only _info.txt, _pockets.pqr and pockets/pocketK_vert.pqr are written - this is what the reference's dpocket leaves, because there the "input
 * file" that decides the mode is the LIST file (dparams.c:98), whose name normally contains neither ".pdb" nor ".cif" */
int  fpkio_write_output_mode(const fpkio_file *f, const char *pdb_path, const fpk_result *r, char write_mode, long long *bytes_out);

/* ---- dpocket (dpocket.c:166-300) ------------------------------------------------------------------------------------------ */
int  fpkio_nligand(const fpkio_file *f);                          /* pdb_cplx_l->natm_lig (the structure then carries n_lig / lig_atoms / lig_mol / atom_name_key) */
/* get_mol_volume_ptr (atom.c:156-217) on the ligand atoms: niter samples in the bounding box, counter-based stream keyed by seed (the reference
 * draws from libc rand() seeded with the time: no stream to copy) */
float fpkio_ligand_volume(const fpkio_file *f, int niter, uint64_t seed);
/* the rows desc_pocket appends to dpout_explicitp / dpout_fpocketp / dpout_fpocketnp for one complex (write_pocket_desc, dpocket.c:387-402);
 * rows[k] is malloc'd (release with fpkio_free_rows), "" when the complex has no row for that table */
int  fpkio_dpocket_rows(const fpkio_file *f, const char *complex_path, const char *ligand, const fpk_result *r, float lig_vol, char *rows[3]);
void fpkio_free_rows(char *rows[3]);
/* the whole loop of dpocket() (dpocket.c:83-130): paths[k] + ligands[k]; writes every <code>_out/ directory and the three tables
 * (header + rows in list order) to out_paths[0..2]; detection, ligand queries and the explicit pocket come from ONE detect call per group */
typedef int  (*fpkio_detect_fn)(const fpk_structure *in, int n, const fpk_params *p, fpk_result *out);
typedef void (*fpkio_free_fn)(fpk_result *r);
struct fpkio_stats;
int  fpkio_run_dpocket(const char *const *paths, const char *const *ligands, int n, const fpk_params *p, fpkio_detect_fn detect,
                       fpkio_free_fn free_result, const char *const out_paths[3], char write_mode /* 'd' (dpocket's usual) or 'p' */,
                       int group, int io_threads, int quiet, struct fpkio_stats *stats);

/* ---- mdpocket (mdpocket.c:71-340, mdpbase.c:301-371, mdpout.c:71-122,177-191) ---------------------------------------------- */
/* the structure mdpocket hands to search_pocket(pdb, params, pdb) (mdpocket.c:844): ONE atom list, pdb_w_lig is pdb itself. st must stay alive
 * as long as f; returns FPKIO_ERR_UNSUPPORTED when the file has HETATM groups the two reference passes of trajectory mode disagree on
 * (rpdb_open counts with keep_lig = 0, rpdb_read stores with keep_lig = 1: mdpocket.c:810,193) */
int  fpkio_md_topology(const fpkio_file *f, fpk_structure *st);
/* CHARMM / NAMD DCD trajectory (the format of BASELINE configs[2]; little or big endian, optional unit-cell records): header, then frames
 * de-interleaved into xyz[nframes][natoms][3] */
typedef struct fpkio_dcd fpkio_dcd;
int  fpkio_dcd_open(const char *path, fpkio_dcd **out, int *natoms, int *nframes);
int  fpkio_dcd_read(fpkio_dcd *d, float *xyz, int max_frames);              /* returns the number of frames read (0 at the end) or FPKIO_ERR_* */
void fpkio_dcd_close(fpkio_dcd *d);
/* what follows the frame loop: normalize_grid on both grids, project_grid_on_atoms(density grid) into the B-factor column of the topology's atoms,
 * write_first_bfactor_density, write_md_grid for both (DX + iso-surface PDB). dens / freq: un-normalised sums [nx][ny][nz] as fpk_md_fetch_grids
 * returns them. paths: freq.dx, dens.dx, freq_iso.pdb, dens_iso.pdb, all_atom_pdensities.pdb */
int  fpkio_md_write_outputs(const fpkio_file *topology, const float *dens, const float *freq, const float origin[3], const int dims[3], int n_snapshots,
                            const fpk_params *p, int use_drug_score, const char *const paths[5], long long *bytes_out);

/* ---- the -F loop ------------------------------------------------------------------------------------------------------- */
typedef struct fpkio_stats {
    int32_t n_files, n_written, n_read_failed, n_no_pockets, n_detect_failed;
    long long n_pockets, n_atoms, bytes_in, bytes_out;
    double wall_s;                 /* whole call */
    double read_cpu_s, write_cpu_s;/* summed over the pool's threads */
    double detect_s;               /* wall time inside the detect callback, summed over groups */
} fpkio_stats;

/* paths[n]: PDB files. group: files per detect call (0 = 1024). io_threads: pool size (0 = hardware threads).
 * quiet = 0 prints the reference's per-file messages. Per-file failures never abort the list. */
int  fpkio_run_list(const char *const *paths, int n, const fpkio_read_opts *ropts, const fpk_params *p,
                    fpkio_detect_fn detect, fpkio_free_fn free_result, int group, int io_threads, int quiet, fpkio_stats *stats);

const char *fpkio_version(void);

#ifdef __cplusplus
}
#endif
#endif
