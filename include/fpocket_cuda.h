/*
 * libfpocket_cuda — C ABI of the B200-native fpocket hot path.
 *
 * This is the boundary a maintainer of Discngine/fpocket binds to (plain C, POD structs, raw
 * pointers and sizes; no C++ or torch types).  It replaces, for the alpha-sphere pocket
 * detection path only, the body of
 *
 *     c_lst_pockets *search_pocket(s_pdb *pdb, s_fparams *params, s_pdb *pdb_w_lig)
 *                                                        (reference src/fpocket.c:58-225)
 *
 * and the per-frame grid accumulation of mdpocket
 *
 *     calculate_md_dens_grid / update_md_grid / init_md_grid   (reference src/mdpbase.c:108-262,444-502)
 *
 * The reference's host code (CLI, PDB/mmCIF readers, writers) stays as it is; INTEGRATION.md
 * shows the shim that flattens s_pdb into fpk_structure and rebuilds c_lst_pockets from
 * fpk_result.  There is no CPU fallback behind any of these calls: if no sm_100 device (or no
 * CUDA device at all) is usable they fail with FPK_ERR_NO_DEVICE.
 */
#ifndef FPOCKET_CUDA_H
#define FPOCKET_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPK_OK 0
#define FPK_ERR_NO_DEVICE   (-1)  /* no CUDA device / not sm_100 / driver error at init                      */
#define FPK_ERR_BAD_ARG     (-2)  /* NULL pointer, negative size, unsupported parameter combination          */
#define FPK_ERR_CUDA        (-3)  /* a CUDA runtime call failed, see fpk_last_error()                         */
#define FPK_ERR_NO_SPHERES  (-4)  /* < 2 kept alpha spheres: the reference prints "Error in creating
                                     clustering tree" and returns NULL (fpocket.c:135-139, clusterlib.c:3969) */
#define FPK_ERR_DEGENERATE  (-5)  /* fewer than 4 non-coplanar heavy atoms: no Delaunay triangulation         */
#define FPK_ERR_CAPACITY    (-6)  /* an internal per-structure buffer overflowed (reported, never silent)     */
#define FPK_ERR_NO_POCKETS  (-7)  /* detection ran, every pocket was dropped (fpmain.c:204 "no pockets found")*/

/* atom flag bits (fpk_structure.flags / .wflags) */
#define FPK_ATOM_H          1u    /* pdb atoms:  strcmp(symbol,"H")==0  -> not a Delaunay site (voronoi.c:100,128)
                                     pdb_w_lig:  symbol[0]=='H'         -> not a "surrounding atom" (asa.c:89)   */
#define FPK_ATOM_XPOCKET    2u    /* atom_in_explicit_pocket() is true (voronoi.c:934-949), -P option only      */
#define FPK_ATOM_H1         4u    /* pdb atoms: symbol[0]=='H' (Hg, Ho, ...): the asa.c:89 test when pdb_w_lig is pdb itself */

/* numeric subset of s_fparams (reference headers/fparams.h:182-227), defaults fparams.h:33-63 */
typedef struct fpk_params {
    float asph_min_size;                  /* -m, 3.4  */
    float asph_max_size;                  /* -M, 6.2  */
    float clust_max_dist;                 /* -D, 2.4  */
    float refine_min_apolar_asphere_prop; /* -p, 0.0  */
    float min_as_density;                 /* 0.7      */
    int32_t min_apol_neigh;               /* -A, 3    */
    int32_t nb_mcv_iter;                  /* -v, 300  */
    int32_t min_pock_nb_asph;             /* -i, 15   */
    int32_t do_asa_and_volume;            /* flag_do_asa_and_volume_calculations; mdpocket sets 0 (mdpocket.c:89)      */
    int32_t skip_clustering;              /* xlig_resnumber != -1 || xpocket_n > 0  (fpocket.c:114)                    */
    int32_t xpocket_active;               /* xpocket_n > 0: spheres need >= min_n_explicit_pocket_atoms flagged atoms  */
    int32_t min_n_explicit_pocket_atoms;  /* -u, 4    */
    int32_t want_tets;                    /* also return the Delaunay tetrahedra (parity checks; costs a D2H copy)     */
    uint64_t mc_seed;                     /* Monte-Carlo volume: Philox key. The reference seeds libc rand() with
                                             time(NULL) (voronoi.c:87), so there is no reference stream to copy.       */
    /* dpocket ligand queries (only read when a structure carries ligand atoms, fpk_structure.n_lig > 0) */
    float lig_interface_dist;             /* dparams.h:42 M_VERT_LIG_NEIG_DIST 4.0: spheres within it of a ligand atom define the explicit pocket */
    float lig_crit_dist;                  /* tpocket.h:32-33 M_CRIT4_D = M_CRIT5_D = 3.0                                                        */
    int32_t distance_measure;             /* -e (fparams.h distance_measure): 0 or 'e' euclid (clusterlib.c:933), 'b' city block = the MEAN of the three
                                             absolute differences (clusterlib.c:1022-1083); anything else: FPK_ERR_BAD_ARG                             */
    int32_t lig_interface_method;         /* dpocket -E (dparams.h:41-44): 0 or 1 = atoms contacted by spheres near the ligand (neighbor.c:313),
                                             2 = atoms within lig_interface_dist of a ligand atom (neighbor.c:67 get_mol_atm_neigh)              */
    int32_t clustering_method;            /* -C (fparams.h clustering_method): 0 or 's' single linkage (clusterlib.c:3514), 'm' complete (:3705), 'a' average
                                             (:3781), 'c' centroid (:3337), each cut by cuttree_distance (:3235) at clust_max_dist. m / a / c keep a
                                             dense f64 distance matrix per structure (4 S^2 bytes for S spheres) in a pool of half the free device
                                             memory, at most 16 GB per worker context (FPK_LINK_POOL_MB overrides): a structure whose matrix alone exceeds it, or one large enough for
                                             the whole-GPU route, gets status FPK_ERR_BAD_ARG; anything else: the call fails with FPK_ERR_BAD_ARG   */
} fpk_params;

/* flattened s_pdb pair (reference headers/rpdb.h:69-98, atom.h:29-77). Arrays are caller-owned, read-only. */
typedef struct fpk_structure {
    int32_t natoms;            /* pdb->natoms (all atoms of the ligand-stripped structure, hydrogens included) */
    const float *x, *y, *z;    /* s_atm.x,y,z                                           */
    const float *radius;       /* s_atm.radius (vdW, pertable.c:120-140)                */
    const float *electroneg;   /* s_atm.electroneg (pertable.c:80-92)                   */
    const float *bfactor;      /* s_atm.bfactor                                         */
    const int32_t *atom_id;    /* s_atm.id      (uniqueness key of contacted atoms, pocket.c:1298)  */
    const int32_t *res_id;     /* s_atm.res_id  (uniqueness key of residues, descriptors.c:408)     */
    const int8_t *aa_idx;      /* index 0..19 chosen by set_aa_desc()'s switch on res_name (descriptors.c:453-507), -1 if none */
    const uint8_t *flags;      /* FPK_ATOM_* */
    float avg_bfactor, min_bfactor, max_bfactor; /* s_pdb stats (rpdb.c:1429-1454) */

    int32_t natoms_w;          /* pdb_w_lig->natoms; 0 with same_pdb=1 means "pdb_w_lig is pdb" (mdpocket.c:844) */
    const float *wx, *wy, *wz, *wradius, *welectroneg;
    const uint8_t *wflags;
    int32_t same_pdb;          /* 1: pdb_w_lig == pdb, so asa.c:315 "a != cura" excludes the atom itself */

    int32_t n_xlig;            /* pdb->n_xlig_atoms (explicit ligand, -r / -a), 0 if none */
    const float *xlig;         /* xyz triplets */

    /* dpocket (reference src/dpocket.c:171-300): the known ligand, as indices into the pdb_w_lig atom arrays (pdb_cplx_l->latm_lig);
     * n_lig = 0 for plain fpocket. lig_mol = 0,1,.. per ligand atom: consecutive runs of equal (chain, res_id), as
     * count_atm_prop_vert_neigh groups them (neighbor.c:619-637). atom_name_key = the atom name of every pdb atom packed into an int
     * (atm_corsp compares names, atom.c:281); NULL: atoms match by id only. */
    int32_t n_lig;
    const int32_t *lig_atoms;
    const int32_t *lig_mol;
    const int32_t *atom_name_key;
} fpk_structure;

/* numeric part of s_desc (reference headers/descriptors.h:34-88) + s_pocket scalars (pocket.h:37-52) */
typedef struct fpk_desc {
    float score, drug_score;
    float hydrophobicity_score, volume_score, volume, convex_hull_volume, prop_polar_atm,
          mean_asph_ray, masph_sacc, apolar_asphere_prop, mean_loc_hyd_dens, as_density, as_max_dst,
          flex, nas_norm, polarity_score_norm, mean_loc_hyd_dens_norm, prop_asapol_norm,
          as_density_norm, as_max_dst_norm,
          surf_vdw14, surf_vdw22, surf_pol_vdw14, surf_pol_vdw22, surf_apol_vdw14, surf_apol_vdw22,
          as_max_r;
    float bary[3];             /* s_pocket.bary as left by reIndexPockets (refine.c:192-240) */
    int32_t nb_asph, polarity_score, charge_score, n_abpa;
    int32_t nAlphaApol, nAlphaPol;
    int32_t n_atoms;           /* unique contacted atoms */
    int32_t aa_compo[20];
    int32_t first_sphere;      /* lowest sphere index of the pocket */
    int32_t final_rank;        /* 1-based rank after drop+sort, 0 if the pocket was dropped */
} fpk_desc;

/*
 * Result of one detection. All arrays are owned by the library and released by fpk_result_free().
 * Sphere order is canonical: ascending lexicographic order of the sorted 4-tuple of Delaunay site
 * indices (the reference's order is qhull's facet order, which cannot be reproduced).
 * "Clusters" are the pockets before dropSmallNpolarPockets (every one gets full descriptors,
 * fpocket.c:183); cluster order = order of lowest sphere index, as apply_clustering leaves it.
 */
typedef struct fpk_result {
    int32_t status;            /* FPK_OK or FPK_ERR_* for this structure */
    int32_t n_sites;           /* heavy atoms fed to the Delaunay triangulation */
    int32_t n_tets;            /* finite Delaunay tetrahedra */
    int32_t *tets;             /* [n_tets*4] site indices, each tet ascending, tets in lexicographic order; NULL unless want_tets */
    int32_t n_spheres;         /* kept alpha spheres = pockets->vertices->nvert */
    float *sph_xyzr;           /* [n_spheres*4] s_vvertice.x,y,z,ray */
    float *sph_bary;           /* [n_spheres*3] s_vvertice.bary */
    int32_t *sph_neigh;        /* [n_spheres*4] indices into the pdb atom arrays (s_vvertice.neigh[]) */
    uint8_t *sph_type;         /* [n_spheres] 0 = M_APOLAR_AS, 1 = M_POLAR_AS */
    int32_t *sph_cluster;      /* [n_spheres] 0-based cluster (pre-drop pocket) */
    int32_t n_clusters;
    int32_t *cl_off;           /* [n_clusters+1] offsets into cl_sph */
    int32_t *cl_sph;           /* [n_spheres] sphere indices grouped by cluster, ascending inside a cluster */
    int32_t *cl_atom_off;      /* [n_clusters+1] offsets into cl_atoms */
    int32_t *cl_atoms;         /* unique contacted atoms per cluster, first-seen order (pocket.c:1280-1320) */
    fpk_desc *desc;            /* [n_clusters] */
    int32_t n_pockets;         /* surviving pockets */
    int32_t *rank_to_cluster;  /* [n_pockets] cluster index of the pocket ranked k+1 (psorting.c:64-165) */
    float *atom_fx;            /* [natoms*4] per-atom ASA side effects: abpa, a0, dA, abpa_sourrounding_prob (asa.c:337,392-396) */
    void *_owner;              /* library bookkeeping */
    /* ---- dpocket ligand queries (neighbor.c), filled when the structure carried ligand atoms; sets in ascending index order ---- */
    int32_t n_xspheres;        /* get_mol_vert_neigh (neighbor.c:192): spheres with dist(centre, some ligand atom) < lig_interface_dist */
    int32_t *xspheres;         /* [n_xspheres] sphere indices: the explicit pocket of dpocket.c:341-349 */
    fpk_desc *xdesc;           /* its descriptor record: set_descriptors(interface atoms, these spheres) (dpocket.c:350); NULL if no sphere qualifies.
                                  Not normalised, not scored, exactly like the reference's edesc */
    int32_t n_iface;           /* get_mol_ctd_atm_neigh(.., M_INTERFACE_SEARCH) (neighbor.c:313): atoms contacted by such a sphere and */
    int32_t *iface_atoms;      /* within 8.0 A (neighbor.h:30) of the ligand atom that reaches the sphere; indices into the pdb atoms   */
    float *pk_ovlp;            /* [n_pockets] rank order: atm_corsp(interface, pocket atoms) in % (atom.c:264, dpocket.c:249)           */
    float *pk_dst;             /* [n_pockets] min distance ligand atom - pocket barycentre (dpocket.c:253-261)                           */
    float *pk_c4;              /* [n_pockets] count_atm_prop_vert_neigh(.., lig_crit_dist) (neighbor.c:598, dpocket.c:266)               */
    float *pk_c5;              /* [n_pockets] count_pocket_lig_vert_ovlp(.., lig_crit_dist) (neighbor.c:490, dpocket.c:268)              */
    float lig_volume;          /* get_mol_volume_ptr(ligand atoms, nb_mcv_iter) (atom.c:156, dpocket.c:245): Monte-Carlo volume of the ligand's vdW
                                  spheres; the reference draws from time-seeded rand(), here Philox keyed by mc_seed                      */
} fpk_result;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* devices = the GPUs the library may use: fpk_detect_batch hands its chunks to all of them, fpk_md_push_frames gives every one a
 * contiguous share of the frames and sums the grids (reference loops fpmain.c:61-76, mdpocket.c:235-271).
 * devices == NULL: $FPK_DEVICES ("all" or "0,2,3") if set, else the current device. Fails loudly without a GPU. */
int fpk_init(const int *devices, int ndev);
int fpk_device_count(void);                   /* devices in use (0 before fpk_init) */
void fpk_shutdown(void);
const char *fpk_last_error(void);
const char *fpk_version(void);
void fpk_default_params(fpk_params *p);        /* init_def_fparams (fparams.c:63-95) */

/* ---- fpocket: replaces search_pocket() (fpocket.c:58) --------------------------------------- */
int fpk_detect(const fpk_structure *in, const fpk_params *p, fpk_result *out);
/* the "-F list" loop (fpmain.c:61-76) as one call; per-structure status in out[k].status, never aborts the batch */
int fpk_detect_batch(const fpk_structure *in, int n, const fpk_params *p, fpk_result *out);
void fpk_result_free(fpk_result *r);

/* staged form of fpk_detect_batch, for callers that keep inputs resident in HBM (bench.py "value") */
typedef struct fpk_batch fpk_batch;
fpk_batch *fpk_batch_create(const fpk_structure *in, int n, const fpk_params *p);  /* pack + H2D */
int fpk_batch_run(fpk_batch *b);                                                   /* kernels only, returns after sync */
int fpk_batch_fetch(fpk_batch *b, fpk_result *out);                                /* D2H + unpack */
float fpk_batch_last_kernel_ms(const fpk_batch *b, int which);                     /* CUDA-event time of stage `which` in the last run */
long long fpk_batch_counter(const fpk_batch *b, int which);                        /* 0 sites,1 tets,2 spheres,3 clusters,4 pockets,5 launches,6 h2d bytes,7 d2h bytes; b == NULL: 6/7 = bytes moved by fpk_detect_batch so far */
void fpk_batch_destroy(fpk_batch *b);

/* ---- mdpocket: replaces the per-frame body of mdpocket_detect (mdpocket.c:142-175,235-271) --- */
typedef struct fpk_md fpk_md;
/* topology = frame-independent atom properties (coordinates in `topo` are ignored) */
fpk_md *fpk_md_begin(const fpk_structure *topo, const fpk_params *p, int use_drug_score /* -S, mdpbase.c:127 */);
/* init_md_grid from one frame's kept spheres (mdpbase.c:444-502): fills origin[3], dims[3] */
int fpk_md_grid_from_frame(fpk_md *m, const float *xyz, float origin[3], int dims[3]);
int fpk_md_set_grid(fpk_md *m, const float origin[3], const int dims[3]);
/* detect + calculate_md_dens_grid + update_md_grid for nframes frames; xyz = [nframes][natoms][3] host floats */
int fpk_md_push_frames(fpk_md *m, const float *xyz, int nframes);
/* device pointers of the two accumulators (float[nx*ny*nz], on the first device, already summed over the library's devices) so that
 * a caller running one process per GPU can sum them across ranks in place (NCCL all-reduce) and then read them with fpk_md_fetch_grids.
 * With -S the increments are floats (drug scores): they are summed as 32.32 fixed-point integers so that the result does not depend on
 * the order in which frames reach a cell (same input, same bytes); the float grids are refreshed from those sums by this call and by
 * fpk_md_fetch_grids when frames were pushed since the last refresh (never otherwise: what the caller summed in place stays). */
int fpk_md_device_grids(fpk_md *m, void **dens, void **freq, size_t *ncell);
int fpk_md_fetch_grids(fpk_md *m, float *dens, float *freq);    /* D2H, un-normalised sums (normalize_grid stays host, mdpbase.c:354) */
/* mdpocket's characterize mode (mdpocket.c:360-660), the per-snapshot body of its loop (:428-460): full detection of the frame (ASA and
 * volumes on, as that mode keeps them), then the spheres of the surviving pockets that lie within 1.0 A of a point of the wanted zone, once
 * per such point (extract_wanted_vertices, :695-745), the atoms they contact and the descriptor record set_descriptors makes of that list
 * (never normalised or scored), with the side effects of its ASA pass on the frame's atoms. out[f] is filled for every frame. */
typedef struct fpk_zone {
    int32_t status;            /* FPK_OK, FPK_ERR_NO_POCKETS (no sphere of a surviving pocket near the zone) or the frame's detection status */
    int32_t n_spheres_frame;   /* kept alpha spheres of the frame */
    int32_t n_pockets_frame;   /* surviving pockets of the frame */
    int32_t n_entries;         /* length of the zone's sphere list (repeats included) */
    int32_t *spheres;          /* [n_entries] index of the sphere in the frame's canonical sphere list */
    float *sph_xyzr;           /* [n_entries*4] what write_pqr_vert prints (mdpocket.c:444) */
    uint8_t *sph_type;         /* [n_entries] */
    int32_t *sph_pocket;       /* [n_entries] 1-based rank of the pocket the sphere belongs to (vertice->resid, refine.c:223: the residue number write_pqr_vert prints) */
    int32_t n_atoms;
    int32_t *atoms;            /* [n_atoms] get_pocket_contacted_atms of the list, first-seen order */
    fpk_desc desc;             /* set_descriptors of the list: volume and hull volume always computed */
    float *atom_fx;            /* [natoms*4] per-atom ASA side effects after the zone's pass (last writer: the zone) */
    void *_owner;
} fpk_zone;
int fpk_md_characterize(fpk_md *m, const float *xyz, int nframes, const float *zone_xyz, int nzone, fpk_zone *out);
void fpk_zone_free(fpk_zone *z);
long long fpk_md_counter(const fpk_md *m, int which);            /* 0 frames, 1 spheres scattered, 2 launches, 3 devices that took frames */
void fpk_md_end(fpk_md *m);

#ifdef __cplusplus
}
#endif
#endif /* FPOCKET_CUDA_H */
