// build/dpocket_fast: the reference's `dpocket -f list.txt` (src/dpmain.c, src/dpocket.c:83-300) with NO reference object in it: native PDB
// reader in ligand mode (rpdb_open / rpdb_read with the ligand's residue name), ONE fpk_detect_batch call per group of complexes (pockets,
// explicit-pocket spheres and record, interface atoms, overlap / distance / consensus criteria per pocket: kernels K1-K6), native writers of
// every <code>_out/ directory and of the three tables dpout_explicitp.txt / dpout_fpocketp.txt / dpout_fpocketnp.txt (rows in list order).
//
//   dpocket_fast -f list.txt [-o prefix] [-d dist] [-e | -E] [-m f] [-M f] [-D f] [-i n] [-A n] [-p f] [-v n] [--threads T] [--group G] [--quiet] [--stats]
//
// list.txt: "<complex.pdb> <ligand residue name>" per line (dparams.c:206-240). Both interface definitions (-e: atoms contacted by spheres near the
// ligand, -E: atoms within 4 A of a ligand atom, dparams.h:41-44) are computed by the library (K6), as is the ligand's own Monte-Carlo volume (column
// lig_vol), which has no reference stream to copy (libc rand() seeded with the time): it is drawn from a counter-based stream keyed by the seed.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "fpocket_hostio.h"

int main(int argc, char **argv)
{
    fpk_params P;
    fpk_default_params(&P);
    std::string listf, prefix;
    int threads = 0, group = 0, quiet = 0, want_stats = 0;
    for (int a = 1; a < argc; a++) {
        const std::string o = argv[a];
        auto val = [&]() -> const char * { if (a + 1 >= argc) { fprintf(stderr, "! option %s needs a value\n", o.c_str()); exit(2); } return argv[++a]; };
        if (o == "-f") listf = val();
        else if (o == "-o") { prefix = val(); const size_t dot = prefix.rfind('.'); if (dot != std::string::npos) prefix.erase(dot); }      // remove_ext (dparams.c:127)
        else if (o == "-d") P.lig_interface_dist = (float)atof(val());
        else if (o == "-e") { P.lig_interface_method = 1; P.lig_interface_dist = 4.0f; }      // M_INTERFACE_METHOD1, M_VERT_LIG_NEIG_DIST
        else if (o == "-E") { P.lig_interface_method = 2; P.lig_interface_dist = 4.0f; }      // M_INTERFACE_METHOD2, M_LIG_NEIG_DIST (dparams.c:116-119)
        else if (o == "-m") P.asph_min_size = (float)atof(val());
        else if (o == "-M") P.asph_max_size = (float)atof(val());
        else if (o == "-D") P.clust_max_dist = (float)atof(val());
        else if (o == "-i") P.min_pock_nb_asph = atoi(val());
        else if (o == "-A") P.min_apol_neigh = atoi(val());
        else if (o == "-p") P.refine_min_apolar_asphere_prop = (float)atof(val());
        else if (o == "-v") P.nb_mcv_iter = atoi(val());
        else if (o == "--threads") threads = atoi(val());
        else if (o == "--group") group = atoi(val());
        else if (o == "--quiet") quiet = 1;
        else if (o == "--stats") want_stats = 1;
        else { fprintf(stderr, "> Unknown option '%s'. Ignoring it.\n", o.c_str()); }
    }
    if (listf.empty()) { fprintf(stderr, "usage: %s -f list.txt [-o prefix] [-d dist] [fpocket detection options] [--threads T] [--group G] [--quiet] [--stats]\n", argv[0]); return 2; }
    FILE *fl = fopen(listf.c_str(), "r");
    if (!fl) { fprintf(stderr, "! File %s doesn't exist\n", listf.c_str()); return 2; }
    std::vector<std::string> files, ligs;
    char buf[1024], cb[512], lb[64];
    while (fgets(buf, 210, fl)) if (sscanf(buf, "%511s\t%63s", cb, lb) == 2) { files.push_back(cb); ligs.push_back(lb); }      // dparams.c:222-226
    fclose(fl);
    if (files.empty()) { fprintf(stderr, "! No complex in %s\n", listf.c_str()); return 2; }
    for (const auto &f : files) if (strstr(f.c_str(), ".cif")) { fprintf(stderr, "! dpocket_fast reads PDB files only: use build/dpocket_cuda\n"); return 3; }
    P.mc_seed = (uint64_t)time(nullptr);
    if (getenv("FPOCKET_MC_SEED")) P.mc_seed = strtoull(getenv("FPOCKET_MC_SEED"), nullptr, 10);
    if (fpk_init(nullptr, 0) != FPK_OK) { fprintf(stderr, "! %s\n", fpk_last_error()); return 4; }
    std::vector<const char *> pp, ll;
    for (size_t k = 0; k < files.size(); k++) { pp.push_back(files[k].c_str()); ll.push_back(ligs[k].c_str()); }
    const std::string o0 = prefix.empty() ? "dpout_explicitp.txt" : prefix + "_exp.txt", o1 = prefix.empty() ? "dpout_fpocketp.txt" : prefix + "_fp.txt",
                      o2 = prefix.empty() ? "dpout_fpocketnp.txt" : prefix + "_fpn.txt";
    const char *outs[3] = {o0.c_str(), o1.c_str(), o2.c_str()};
    fpkio_stats S;
    // the reference decides its write mode from the name of the "input file", which for dpocket is the list (dparams.c:98, fparams.c:397-407)
    const char wmode = strstr(listf.c_str(), ".pdb") ? 'p' : 'd';
    if (strstr(listf.c_str(), ".cif")) { fprintf(stderr, "! a list name containing .cif selects the mmCIF writers: use build/dpocket_cuda\n"); return 3; }
    const int rc = fpkio_run_dpocket(pp.data(), ll.data(), (int)pp.size(), &P, fpk_detect_batch, fpk_result_free, outs, wmode, group, threads, quiet, &S);
    if (rc != FPKIO_OK) fprintf(stderr, "! dpocket_fast: %d (%s)\n", rc, fpk_last_error());
    if (want_stats)
        printf("dpocket_fast: %d complexes, %d written (%lld pockets), %d unreadable / without ligand, %d failed | %.3f s wall, %.1f complexes/s | "
               "read %.3f cpu-s, detect %.3f s, write %.3f cpu-s\n", S.n_files, S.n_written, S.n_pockets, S.n_read_failed, S.n_detect_failed, S.wall_s,
               S.wall_s > 0 ? S.n_files / S.wall_s : 0.0, S.read_cpu_s, S.detect_s, S.write_cpu_s);
    fpk_shutdown();
    return (rc != FPKIO_OK || S.n_read_failed || S.n_detect_failed) ? 1 : 0;
}
