/*
 * TEST INFRASTRUCTURE — the CPU oracle. Not part of the product; nothing under fpocket_b200/
 * may include, link or call this. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Plain-C restatement of the reference's alpha-sphere hot path (Discngine/fpocket,
 * search_pocket(), src/fpocket.c:58-225, and the mdpocket grid scatter, src/mdpbase.c).
 * PARITY PINNED: validated bit-for-bit against the unmodified reference built from source
 * (oracle/_ref/ref_dump, see tests/test_oracle_vs_reference.py and tests/golden/).
 *
 * The POD types are the ones of the product's C ABI (include/fpocket_cuda.h) so the same
 * test harness drives both.
 */
#ifndef FPK_ORACLE_H
#define FPK_ORACLE_H
#include "../include/fpocket_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Delaunay sites: heavy atoms in file order (voronoi.c:94-107). Returns n_sites; site2atom[n_sites];
 * ki[3*n_sites] = coordinates in units of 1e-6 A exactly as "%.6f" prints them (voronoi.c:130);
 * xd[3*n_sites] = what qhull parses from that text (strtod). */
int orc_sites(const fpk_structure *in, int *site2atom, long long *ki, double *xd);

/* Exact 3-D Delaunay triangulation (incremental Bowyer-Watson, exact integer predicates, symbolic
 * perturbation). tets_out is malloc'ed: 4*ntets site indices, canonical form (each tet ascending,
 * tets sorted). Also returns the hull volume (sum of tet volumes, f64) if vol != NULL.
 * Returns 0, or FPK_ERR_DEGENERATE. */
int orc_delaunay(int n, const long long *ki, int **tets_out, int *ntets, double *vol);

/* canonicalise a tet list in place: sort the 4 ids of each tet, then sort tets lexicographically */
void orc_canonicalize_tets(int *tets, int ntets);

/* rng_mode for the Monte-Carlo volume (voronoi.c:1150-1233):
 *   0 = libc rand() after srand(mc_seed): reproduces ref_dump --seed S bit for bit
 *   1 = Philox4x32-10 keyed by mc_seed, counter (sample, replicate, cluster): what the CUDA path uses */
#define ORC_RNG_LIBC   0
#define ORC_RNG_PHILOX 1

/* The whole of search_pocket() after qhull: filter/typing (voronoi.c:369-521,969-1087), clustering
 * (clusterlib.c:933,3235), pockets (pocket.c:77, refine.c:58,192), descriptors (descriptors.c:158-513,
 * asa.c:240-454, voronoi.c:1150-1291), normalisation + scores (pocket.c:664, pscoring.c:241,325),
 * drop (refine.c:114) and sort (psorting.c:64).
 * tets: 4*ntets site indices IN THE ORDER AND VERTEX ORDER TO BE USED (qhull's, or canonical).
 * If tets == NULL the oracle triangulates itself (canonical order).
 * out is filled with malloc'ed arrays; free with orc_result_free. */
int orc_search_pocket(const fpk_structure *in, const fpk_params *p, const int *tets, int ntets,
                      int rng_mode, fpk_result *out);
void orc_result_free(fpk_result *r);
/* dpocket ligand queries on a finished result (called by orc_search_pocket when in->n_lig > 0): fills xspheres, iface_atoms, pk_* */
void orc_ligand_queries(const fpk_structure *in, const fpk_params *p, fpk_result *out);
/* mdpocket characterize mode (mdpocket.c:436-460, :695-745): hand the wanted zone in before orc_search_pocket, read the zone's record back after it.
 * The zone pointer is borrowed; pass NULL to switch the mode off again. */
void orc_set_wanted_zone(const float *xyz, int n);
/* -e: 'e' (default) or 'b' city block for the clustering distance; a process-wide switch, reset it to 'e' after use */
void orc_set_distance_measure(int c);
/* the clustering step alone (tests): n centres, cut, -e measure ('e' | 'b'), -C method ('s' | 'm' | 'a' | 'c'); returns the number of clusters */
int orc_cluster_points(const float *xyz, int n, float clust_max_dist, int measure, int method, int *label);
/* dpocket: 1 (default) interface = atoms contacted by the spheres near the ligand, 2 = atoms within the distance of a ligand atom (dparams.h:41-44) */
void orc_set_interface_method(int m);
int orc_get_wanted(const int **idx, int *nidx, const int **atoms, int *natoms, fpk_desc *d);

/* mdpocket grids (mdpbase.c). Spheres of surviving pockets in final pocket order. */
void orc_md_grid_geometry(const fpk_result *r, float origin[3], int dims[3]);           /* init_md_grid :444-502 */
void orc_md_scatter(const fpk_result *r, int use_drug_score, const float origin[3], const int dims[3],
                    float *dens, float *freq, float *refg);                              /* :108-262, refg is cleared first (:275) */

/* single pieces exported for unit tests */
void orc_circumcentre(const double *p0, const double *p1, const double *p2, const double *p3, double c[3]);
int orc_orient3d(const long long *a, const long long *b, const long long *c, const long long *d);
int orc_insphere(const long long *a, const long long *b, const long long *c, const long long *d, const long long *e);
void orc_spiral_points(float *pts /* 300 */);                                            /* asa.c:438-454 */
void orc_philox4x32(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned out[4]);

#ifdef __cplusplus
}
#endif
#endif
