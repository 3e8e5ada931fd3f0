// build/fpocket_fast: the reference's `fpocket -f X.pdb` / `fpocket -F list.txt` (src/fpmain.c:52-103) with NO reference object in it:
// native PDB reader and writers (libfpocket_hostio) around libfpocket_cuda, as a three-stage pipeline
//     reader pool  ->  one fpk_detect_batch per group of files  ->  writer pool.
// It leaves, next to every input, the <code>_out/ directory `fpocket -f` writes (PDB mode, fparams.c:403-407).
//
//   fpocket_fast -f X.pdb | -F list.txt  [-m f] [-M f] [-D f] [-i n] [-A n] [-p f] [-v n] [-c A,B | -k A,B]
//                [--threads T] [--group G] [--quiet] [--stats] [--devices 0,1,..]
//
// --devices: one worker process per listed GPU (independent structures: no collective), file k of the list goes to worker k mod N; each
// worker is this same program restricted to its device through CUDA_VISIBLE_DEVICES, with the I/O pool divided among them.
//
// -r resnumber:resname:chain (explicit ligand) and -P res:ins:chain.res:ins:chain (explicit pocket, -u) are handled as the reference does.
// Options outside the restated route (mmCIF input, -d, -x, -C != s, -e != e) are refused with a pointer to
// build/fpocket_cuda_batch (reference reader / writers around the same library). There is no CPU detection path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <thread>
#include <vector>

#include <sys/wait.h>
#include <unistd.h>

#include "fpocket_hostio.h"

static int refuse(const char *what)
{
    fprintf(stderr, "! fpocket_fast does not handle %s: use build/fpocket_cuda_batch (the reference's reader and writers on the same library)\n", what);
    return 3;
}

int main(int argc, char **argv)
{
    fpk_params P;
    fpk_default_params(&P);
    fpkio_read_opts ro;
    memset(&ro, 0, sizeof ro);
    std::vector<std::string> files;
    int threads = 0, group = 0, quiet = 0, want_stats = 0, slice_k = 0, slice_n = 1;
    std::vector<std::string> devices;
    auto chains = [&](const char *arg, int kept) {                 // fparams.c:203-240: names separated by ',' or ':'
        char tmp[512];
        snprintf(tmp, sizeof tmp, "%s", arg);
        ro.n_chains = 0;
        for (char *t = strtok(tmp, ",:"); t && ro.n_chains < 20; t = strtok(nullptr, ",:")) snprintf(ro.chains[ro.n_chains++], 21, "%s", t);
        ro.chains_are_kept = kept;
    };
    for (int a = 1; a < argc; a++) {
        const std::string o = argv[a];
        auto val = [&]() -> const char * { if (a + 1 >= argc) { fprintf(stderr, "! option %s needs a value\n", o.c_str()); exit(2); } return argv[++a]; };
        if (o == "-f" || o == "--file") files.push_back(val());
        else if (o == "-F" || o == "--fileList") {
            const char *lp = val();
            FILE *fl = fopen(lp, "r");
            if (!fl) { fprintf(stderr, "! File %s doesn't exist\n", lp); return 2; }
            char line[4096];
            while (fgets(line, sizeof line, fl)) {
                size_t L = strlen(line);
                while (L && (line[L - 1] == '\n' || line[L - 1] == '\r' || line[L - 1] == ' ')) line[--L] = 0;
                if (L) files.push_back(line);
            }
            fclose(fl);
        }
        else if (o == "-m" || o == "--min_alpha_size") P.asph_min_size = (float)atof(val());
        else if (o == "-M" || o == "--max_alpha_size") P.asph_max_size = (float)atof(val());
        else if (o == "-D" || o == "--clustering_distance") P.clust_max_dist = (float)atof(val());
        else if (o == "-i" || o == "--min_spheres_per_pocket") P.min_pock_nb_asph = atoi(val());
        else if (o == "-A" || o == "--number_apol_asph_pocket") P.min_apol_neigh = atoi(val());
        else if (o == "-p" || o == "--ratio_apol_spheres_pocket") P.refine_min_apolar_asphere_prop = (float)atof(val());
        else if (o == "-v" || o == "--iterations_volume_mc") P.nb_mcv_iter = atoi(val());
        else if (o == "-r" || o == "--custom_ligand") { fpkio_parse_custom_ligand(val(), &ro); }
        else if (o == "-P" || o == "--custom_pocket") { fpkio_parse_custom_pocket(val(), &ro); }
        else if (o == "-a" || o == "--chain_as_ligand") {        // fparams.c:186-201
            char tmp[512];
            snprintf(tmp, sizeof tmp, "%s", val());
            ro.n_lig_chains = 0;
            for (char *t = strtok(tmp, ",:"); t && ro.n_lig_chains < 20; t = strtok(nullptr, ",:")) snprintf(ro.lig_chains[ro.n_lig_chains++], 21, "%s", t);
        }
        else if (o == "-l" || o == "--model_number") ro.model_number = atoi(val());
        else if (o == "-u" || o == "--min_n_explicit_pocket") P.min_n_explicit_pocket_atoms = atoi(val());
        else if (o == "-c" || o == "--drop_chains") chains(val(), 0);
        else if (o == "-k" || o == "--keep_chains") chains(val(), 1);
        else if (o == "-C" || o == "--clustering_method") { const char m = val()[0]; if (m != 's' && m != 'm' && m != 'a' && m != 'c') return refuse("-C other than s, m, a, c (clusterlib's linkages)"); P.clustering_method = m; }
        else if (o == "-e" || o == "--clustering_measure") { const char m = val()[0]; if (m != 'e' && m != 'b') return refuse("-e other than e (euclidean) or b (city block)"); P.distance_measure = m; }
        else if (o == "-w" || o == "--write_mode") {            // fparams.c:167-183: d | b(oth) | p(db) | m(mcif) | cif
            const char *w = val();
            if (w[0] != 'd' && w[0] != 'b' && w[0] != 'p' && w[0] != 'm' && strcmp(w, "cif")) printf("Writing mode invalid, set to default\n");
            else ro.write_mode = !strcmp(w, "cif") ? 'm' : (w[0] == 'd' ? 'p' : w[0]);
        }
        else if (o == "--threads") threads = atoi(val());
        else if (o == "--group") group = atoi(val());
        else if (o == "--devices") { char tmp[256]; snprintf(tmp, sizeof tmp, "%s", val()); for (char *t = strtok(tmp, ","); t; t = strtok(nullptr, ",")) devices.push_back(t); }
        else if (o == "--slice") { if (sscanf(val(), "%d/%d", &slice_k, &slice_n) != 2 || slice_n < 1 || slice_k < 0 || slice_k >= slice_n) { fprintf(stderr, "! bad --slice\n"); return 2; } }
        else if (o == "--quiet") quiet = 1;
        else if (o == "--stats") want_stats = 1;
        else if (o == "-d" || o == "-x" || o == "-y" || o == "-b") return refuse(o.c_str());
        else { fprintf(stderr, "! unknown option %s\n", o.c_str()); return 2; }
    }
    if (files.empty()) { fprintf(stderr, "usage: %s -f X.pdb | -F list.txt [fpocket detection options] [--threads T] [--group G] [--quiet] [--stats] [--devices 0,1,..]\n", argv[0]); return 2; }
    for (const auto &f : files) if (strstr(f.c_str(), ".cif")) return refuse("mmCIF input");
    if (ro.n_xpocket > 0 && ro.has_xlig) {                   // fparams.c:409-413
        fprintf(stderr, "\nERROR: you specified an explicit ligand (-r) AND an explicit pocke (-P) in the same fpocket run. This is currently not allowed, please use either the one or the other.\n\n");
        return 2;
    }
    P.skip_clustering = (ro.has_xlig || ro.n_xpocket > 0 || ro.n_lig_chains > 0) ? 1 : 0;     // fpocket.c:114 (-a sets xlig_resnumber = 0)
    P.xpocket_active = ro.n_xpocket > 0 ? 1 : 0;
    if (devices.size() > 1) {
        // one worker per GPU: re-run this command line with --slice k/N under CUDA_VISIBLE_DEVICES=<device k>
        const int nd = (int)devices.size();
        int hw = (int)std::thread::hardware_concurrency();
        if (threads <= 0) threads = hw > 0 ? std::max(2, hw / nd) : 4;
        std::vector<pid_t> kids;
        for (int k = 0; k < nd; k++) {
            pid_t pid = fork();
            if (pid < 0) { perror("fork"); return 4; }
            if (pid == 0) {
                setenv("CUDA_VISIBLE_DEVICES", devices[k].c_str(), 1);
                std::vector<std::string> av;
                for (int a = 0; a < argc; a++) {
                    if (!strcmp(argv[a], "--devices") || !strcmp(argv[a], "--threads")) { a++; continue; }
                    av.push_back(argv[a]);
                }
                char sl[32], th[16];
                snprintf(sl, sizeof sl, "%d/%d", k, nd);
                snprintf(th, sizeof th, "%d", threads);
                av.push_back("--slice"); av.push_back(sl); av.push_back("--threads"); av.push_back(th);
                std::vector<char *> cav;
                for (auto &x : av) cav.push_back(const_cast<char *>(x.c_str()));
                cav.push_back(nullptr);
                execv("/proc/self/exe", cav.data());
                perror("execv");
                _exit(127);
            }
            kids.push_back(pid);
        }
        int worst = 0;
        for (pid_t pid : kids) { int st = 0; waitpid(pid, &st, 0); const int rc = WIFEXITED(st) ? WEXITSTATUS(st) : 128; if (rc > worst) worst = rc; }
        return worst;
    }
    if (devices.size() == 1) setenv("CUDA_VISIBLE_DEVICES", devices[0].c_str(), 1);
    if (slice_n > 1) {
        std::vector<std::string> mine;
        for (size_t k = 0; k < files.size(); k++) if ((int)(k % (size_t)slice_n) == slice_k) mine.push_back(files[k]);
        files.swap(mine);
        if (files.empty()) return 0;
    }
    P.mc_seed = (uint64_t)time(nullptr);                                       // voronoi.c:87 srand(time(NULL))
    if (getenv("FPOCKET_MC_SEED")) P.mc_seed = strtoull(getenv("FPOCKET_MC_SEED"), nullptr, 10);
    if (fpk_init(nullptr, 0) != FPK_OK) { fprintf(stderr, "! %s\n", fpk_last_error()); return 4; }
    std::vector<const char *> paths;
    for (const auto &f : files) paths.push_back(f.c_str());
    fpkio_stats S;
    const int rc = fpkio_run_list(paths.data(), (int)paths.size(), &ro, &P, fpk_detect_batch, fpk_result_free, group, threads, quiet, &S);
    if (rc != FPKIO_OK) fprintf(stderr, "! fpk_detect_batch: %s\n", fpk_last_error());
    if (want_stats)
        printf("fpocket_fast: %d files, %d written (%lld pockets), %d unreadable, %d without pockets, %d failed | %.3f s wall, %.1f structures/s | "
               "read %.3f cpu-s, detect %.3f s, write %.3f cpu-s | %.1f MB in, %.1f MB out\n",
               S.n_files, S.n_written, S.n_pockets, S.n_read_failed, S.n_no_pockets, S.n_detect_failed, S.wall_s,
               S.wall_s > 0 ? S.n_files / S.wall_s : 0.0, S.read_cpu_s, S.detect_s, S.write_cpu_s, S.bytes_in / 1e6, S.bytes_out / 1e6);
    fpk_shutdown();
    return (rc != FPKIO_OK || S.n_read_failed || S.n_detect_failed) ? 1 : 0;
}
