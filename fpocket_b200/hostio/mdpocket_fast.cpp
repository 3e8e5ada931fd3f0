// build/mdpocket_fast: the reference's mdpocket in detection mode (src/mdpmain.c, src/mdpocket.c:71-340) with NO reference object in it:
// native topology reader, native DCD reader (or a list of PDB snapshots read by a pool of threads), fpk_md_* for every frame (detection +
// density / frequency grid scatter on the GPU), native writers of the five output files. Frames are read into one buffer while the library
// works on the other.
//
//   mdpocket_fast --trajectory_file traj.dcd --trajectory_format dcd -f topology.pdb [-o prefix] [-S] [-m f] [-M f] [-D f] [-i n] [-A n] [-p f]
//   mdpocket_fast --pdb_list snapshots.txt [...]                       (short options of the reference: -t -O -f -L -o -S)
//
// Other trajectory formats (the reference reads them through the binary molfile plugins), characterize mode (--selected_pocket), user grids and
// energy grids are refused with a pointer to build/mdpocket_cuda (the reference's program on the same library). No CPU detection path.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "fpocket_hostio.h"

static int refuse(const char *what)
{
    fprintf(stderr, "! mdpocket_fast does not handle %s: use build/mdpocket_cuda (the reference's program on the same library)\n", what);
    return 3;
}

int main(int argc, char **argv)
{
    fpk_params P;
    fpk_default_params(&P);
    P.do_asa_and_volume = 0;                                    // mdpocket.c:89
    std::string traj, fmt, top, listf, prefix;
    int score = 0, want_stats = 0;
    for (int a = 1; a < argc; a++) {
        const std::string o = argv[a];
        auto val = [&]() -> const char * { if (a + 1 >= argc) { fprintf(stderr, "! option %s needs a value\n", o.c_str()); exit(2); } return argv[++a]; };
        if (o == "-t" || o == "--trajectory_file") traj = val();
        else if (o == "-O" || o == "--trajectory_format") fmt = val();
        else if (o == "-f" || o == "--file") top = val();
        else if (o == "-L" || o == "--pdb_list") listf = val();
        else if (o == "-o" || o == "--output_prefix") prefix = val();
        else if (o == "-S" || o == "--druggability_score") score = 1;
        else if (o == "-m") P.asph_min_size = (float)atof(val());
        else if (o == "-M") P.asph_max_size = (float)atof(val());
        else if (o == "-D") P.clust_max_dist = (float)atof(val());
        else if (o == "-i") P.min_pock_nb_asph = atoi(val());
        else if (o == "-A") P.min_apol_neigh = atoi(val());
        else if (o == "-p") P.refine_min_apolar_asphere_prop = (float)atof(val());
        else if (o == "--stats") want_stats = 1;
        else if (o == "--devices") setenv("FPK_DEVICES", val(), 1);       // "all" or "0,1,..": the library shards every block of frames over them and sums the grids
        else if (o == "-w" || o == "--selected_pocket") return refuse("characterize mode (--selected_pocket)");
        else if (o == "-P" || o == "-T" || o == "-x" || o == "-a") return refuse(o.c_str());
        else { fprintf(stderr, "! unknown option %s\n", o.c_str()); return 2; }
    }
    const bool use_traj = !traj.empty();
    if (use_traj == !listf.empty()) { fprintf(stderr, "usage: %s --trajectory_file T.dcd --trajectory_format dcd -f top.pdb | --pdb_list list.txt  [-o prefix] [-S] [--devices all|0,1,..] [fpocket detection options]\n", argv[0]); return 2; }
    if (use_traj && strncmp(fmt.c_str(), "dcd", 3)) return refuse("trajectory formats other than dcd");
    std::vector<std::string> snaps;
    if (!use_traj) {
        FILE *fl = fopen(listf.c_str(), "r");
        if (!fl) { fprintf(stderr, "! File %s doesn't exist\n", listf.c_str()); return 2; }
        char line[4096];
        while (fgets(line, sizeof line, fl)) {
            size_t L = strlen(line);
            while (L && (line[L - 1] == '\n' || line[L - 1] == '\r' || line[L - 1] == ' ')) line[--L] = 0;
            if (L) snaps.push_back(line);
        }
        fclose(fl);
        if (snaps.empty()) { fprintf(stderr, "! no snapshot in %s\n", listf.c_str()); return 2; }
        top = snaps[0];
    }
    if (top.empty()) { fprintf(stderr, "! the topology PDB (-f) is missing\n"); return 2; }
    // ---- topology: frame-independent atom properties (mdpocket.c:190-193 / :839) ----
    fpkio_file *T = nullptr;
    int rc = fpkio_read_pdb(top.c_str(), nullptr, &T);
    if (rc) { fprintf(stderr, "! PDB reading failed on %s (%d)\n", top.c_str(), rc); return 1; }
    fpk_structure st;
    if (fpkio_md_topology(T, &st)) return refuse("a topology with HETATM groups outside the kept list");
    const int na = st.natoms;
    if (fpk_init(nullptr, 0) != FPK_OK) { fprintf(stderr, "! %s\n", fpk_last_error()); return 4; }
    fpk_md *md = fpk_md_begin(&st, &P, score);
    if (!md) { fprintf(stderr, "! libfpocket_cuda: %s\n", fpk_last_error()); return 4; }
    const int BLOCK = 2368 * (fpk_device_count() > 0 ? fpk_device_count() : 1);      // frames handed to the library at a time (a share of 2,368 per device)
    std::vector<float> buf[2];
    buf[0].resize((size_t)3 * na * BLOCK); buf[1].resize((size_t)3 * na * BLOCK);
    float origin[3];
    int dims[3], nframes = 0, cur = 0, failed = 0, have_grid = 0;
    std::thread pusher;
    int push_rc = 0;
    auto wait = [&]() { if (pusher.joinable()) pusher.join(); if (push_rc) { fprintf(stderr, "! libfpocket_cuda (%d): %s\n", push_rc, fpk_last_error()); failed = 1; } };
    auto flush = [&](int n) {
        if (n <= 0) return;
        wait();
        if (failed) return;
        float *b = buf[cur].data();
        if (!have_grid) {                                        // frame 1 defines the grid (mdpocket.c:151-154)
            int r = fpk_md_grid_from_frame(md, b, origin, dims);
            if (!r) r = fpk_md_set_grid(md, origin, dims);
            if (r) { fprintf(stderr, "! libfpocket_cuda (%d): %s\n", r, fpk_last_error()); failed = 1; return; }
            have_grid = 1;
        }
        push_rc = 0;
        pusher = std::thread([&, b, n] { push_rc = fpk_md_push_frames(md, b, n); });
        nframes += n;
        cur ^= 1;
    };
    if (use_traj) {
        fpkio_dcd *D = nullptr;
        int dn = 0, df = 0;
        rc = fpkio_dcd_open(traj.c_str(), &D, &dn, &df);
        if (rc) { fprintf(stderr, "Error: could not open file '%s' for reading.\n", traj.c_str()); return 1; }
        if (dn != na) { fprintf(stderr, "Number of atoms in topology and MD trajectory do not coincide : %d vs %d\n", dn, na); return 1; }
        // the first frame alone fixes the grid, as in the reference; then full blocks
        int n = fpkio_dcd_read(D, buf[cur].data(), 1);
        while (n > 0 && !failed) { flush(n); n = fpkio_dcd_read(D, buf[cur].data(), BLOCK); }
        fpkio_dcd_close(D);
    } else {
        // snapshots: parsed by a few threads at a time, coordinates copied in list order (atom properties are snapshot 1's for every frame)
        const int nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        size_t done = 0;
        int nb = 0;
        while (done < snaps.size() && !failed) {
            const size_t chunk = std::min(snaps.size() - done, (size_t)std::min(BLOCK - nb, 256));
            std::vector<int> ok(chunk, 0);
            std::vector<std::thread> th;
            float *dst0 = buf[cur].data() + (size_t)3 * na * nb;
            for (int t = 0; t < nt; t++)
                th.emplace_back([&, t] {
                    for (size_t k = (size_t)t; k < chunk; k += (size_t)nt) {
                        fpkio_file *S = nullptr;
                        if (fpkio_read_pdb(snaps[done + k].c_str(), nullptr, &S) || fpkio_natoms(S, 0) != na) { fpkio_file_free(S); continue; }
                        const fpk_structure *ss = fpkio_structure(S);
                        float *d = dst0 + (size_t)3 * na * k;
                        for (int j = 0; j < na; j++) { d[3 * j] = ss->x[j]; d[3 * j + 1] = ss->y[j]; d[3 * j + 2] = ss->z[j]; }
                        fpkio_file_free(S);
                        ok[k] = 1;
                    }
                });
            for (auto &x : th) x.join();
            for (size_t k = 0; k < chunk; k++) if (!ok[k]) { fprintf(stderr, "! PDB reading failed on snapshot %zu (or its atom count differs from snapshot 1)\n", done + k + 1); failed = 1; }
            if (failed) break;
            nb += (int)chunk;
            done += chunk;
            if (nframes == 0 && nb >= 1) {                       // the first frame alone, then the rest of what was read
                flush(1);
                if (nb > 1) { memcpy(buf[cur].data(), buf[cur ^ 1].data() + (size_t)3 * na, sizeof(float) * 3 * (size_t)na * (nb - 1)); }
                nb -= 1;
            }
            if (nb == BLOCK || done == snaps.size()) { flush(nb); nb = 0; }
        }
    }
    wait();
    printf("<mdpocket> %d snapshots analysed on the GPU\n", nframes);
    if (failed || !have_grid || nframes == 0) { fpk_md_end(md); return 1; }
    const size_t ncell = (size_t)dims[0] * dims[1] * dims[2];
    std::vector<float> dens(ncell), freq(ncell);
    if (fpk_md_fetch_grids(md, dens.data(), freq.data())) { fprintf(stderr, "! libfpocket_cuda: %s\n", fpk_last_error()); return 4; }
    fpk_md_end(md);
    const std::string pf = prefix.empty() ? "mdpout" : prefix;
    const std::string o0 = prefix.empty() ? "mdpout_freq_grid.dx" : pf + "_freq.dx", o1 = prefix.empty() ? "mdpout_dens_grid.dx" : pf + "_dens.dx",
                      o2 = pf + "_freq_iso_0_5.pdb", o3 = pf + "_dens_iso_8.pdb", o4 = pf + "_all_atom_pdensities.pdb";
    const char *paths[5] = {o0.c_str(), o1.c_str(), o2.c_str(), o3.c_str(), o4.c_str()};
    long long bytes = 0;
    rc = fpkio_md_write_outputs(T, dens.data(), freq.data(), origin, dims, nframes, &P, score, paths, &bytes);
    if (rc) { fprintf(stderr, "! could not write the output files (%d)\n", rc); return 1; }
    if (want_stats) printf("mdpocket_fast: %d frames, grid %d x %d x %d, %.1f MB written\n", nframes, dims[0], dims[1], dims[2], bytes / 1e6);
    fpkio_file_free(T);
    fpk_shutdown();
    return 0;
}
