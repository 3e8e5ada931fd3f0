// libfpocket_cuda: host side of the C ABI declared in include/fpocket_cuda.h.
//
// Host work here is packing (flatten the caller's arrays into pinned staging buffers, one H2D copy per
// array), launching the kernels and unpacking (slices of a few large D2H buffers). Every stage of the
// hot path runs on the GPU; there is no CPU implementation behind these entry points.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <numeric>
#include <string>
#include <vector>

#include "k_hull_md.cuh"
#include "k_linkage.cuh"

using namespace fpk;

// ------------------------------------------------------------------------------------------------------------------
// Last error: every thread keeps its own message; the most recent one of ANY thread is also kept process-wide, because the library
// itself works on helper threads (fpk_detect_batch, fpk_md_push_frames) and callers report from another thread than the one that failed.
static thread_local std::string g_err;
static std::mutex g_err_mu;
static std::string g_err_last;
static thread_local std::string g_err_ret;
static std::atomic<long long> g_tot_h2d(0), g_tot_d2h(0);      // bytes moved by fpk_detect_batch since load (fpk_batch_counter(NULL, 6 / 7))
static std::vector<int> g_devs;                                // the devices fpk_init was given (fpk_detect_batch and fpk_md_* shard over them)
static int g_sms = 0;
static bool g_init = false;

static void set_err(const std::string &m)
{
    g_err = m;
    std::lock_guard<std::mutex> g(g_err_mu);
    g_err_last = m;
}
static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    set_err(buf);
    return code;
}
#define CK(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) return fail(FPK_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" const char *fpk_last_error(void)
{
    if (!g_err.empty()) return g_err.c_str();
    std::lock_guard<std::mutex> g(g_err_mu);
    g_err_ret = g_err_last;
    return g_err_ret.c_str();
}
extern "C" const char *fpk_version(void) { return "fpocket_b200 0.1 (sm_100a)"; }

extern "C" void fpk_default_params(fpk_params *p)
{
    memset(p, 0, sizeof *p);
    p->asph_min_size = (float)3.4; p->asph_max_size = (float)6.2; p->clust_max_dist = (float)2.4;
    p->refine_min_apolar_asphere_prop = (float)0.0; p->min_as_density = (float)0.7;
    p->min_apol_neigh = 3; p->nb_mcv_iter = 300; p->min_pock_nb_asph = 15;
    p->do_asa_and_volume = 1; p->min_n_explicit_pocket_atoms = 4; p->mc_seed = 20260101ull;
    p->lig_interface_dist = (float)4.0; p->lig_crit_dist = (float)3.0;
    p->distance_measure = 'e'; p->lig_interface_method = 1; p->clustering_method = 's';
}

// asa.c:438-454 get_points_on_sphere(100): computed with the host libm exactly as the reference does
static void spiral_points(float *pts)
{
    const int nop = 100;
    float inc = (float)(3.1415926535897932384626433832795 * (3.0 - sqrt(5.0)));
    float off = (float)(2.0 / (double)(float)nop);
    for (int k = 0; k < nop; k++) {
        float y = (float)((double)(((float)k) * off) - 1.0 + ((double)off / 2.0));
        float r = (float)sqrt(1.0 - (double)(y * y));
        float phi = (float)k * inc;
        pts[3 * k] = (float)(cos((double)phi) * (double)r);
        pts[3 * k + 1] = y;
        pts[3 * k + 2] = (float)(sin((double)phi) * (double)r);
    }
}

extern "C" int fpk_init(const int *devices, int ndev)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(FPK_ERR_NO_DEVICE, "no CUDA device (%s); libfpocket_cuda has no CPU fallback", e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    std::vector<int> devs;
    if (devices && ndev > 0) devs.assign(devices, devices + ndev);
    else if (const char *ev = getenv("FPK_DEVICES")) {          // "all" or a comma-separated list: lets any program on the library use several GPUs
        if (!strcmp(ev, "all")) for (int d = 0; d < count; d++) devs.push_back(d);
        else for (const char *q = ev; *q;) { devs.push_back(atoi(q)); while (*q && *q != ',') q++; if (*q == ',') q++; }
    }
    if (devs.empty()) { int dev = 0; cudaGetDevice(&dev); devs.push_back(dev); }
    for (size_t i = 0; i < devs.size(); i++) {
        if (devs[i] < 0 || devs[i] >= count) return fail(FPK_ERR_BAD_ARG, "fpk_init: device %d does not exist (%d devices)", devs[i], count);
        for (size_t j = 0; j < i; j++) if (devs[j] == devs[i]) return fail(FPK_ERR_BAD_ARG, "fpk_init: device %d listed twice", devs[i]);
    }
    float pts[300];
    spiral_points(pts);
    int sms = 0;
    for (int dev : devs) {
        cudaDeviceProp pr;
        CK(cudaGetDeviceProperties(&pr, dev));
        if (pr.major < 10) return fail(FPK_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", dev, pr.major, pr.minor);
        CK(cudaSetDevice(dev));
        sms = sms ? std::min(sms, pr.multiProcessorCount) : pr.multiProcessorCount;
        CK(cudaMemcpyToSymbol(c_spiral, pts, sizeof pts));
        // dynamic shared memory of the Delaunay kernels: opted in ONCE per device to the device maximum (a per-batch setting would race
        // between the host threads of fpk_detect_batch); each launch then asks for what its batch needs
        cudaFuncAttributes fa;
        CK(cudaFuncGetAttributes(&fa, k_delaunay_filter));
        CK(cudaFuncSetAttribute(k_delaunay_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(pr.sharedMemPerBlockOptin - fa.sharedSizeBytes)));
        CK(cudaFuncGetAttributes(&fa, k_delaunay_grid));
        CK(cudaFuncSetAttribute(k_delaunay_grid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(pr.sharedMemPerBlockOptin - fa.sharedSizeBytes)));
        CK(cudaFuncGetAttributes(&fa, k_hull_volume));
        CK(cudaFuncSetAttribute(k_hull_volume, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(pr.sharedMemPerBlockOptin - fa.sharedSizeBytes)));
        CK(cudaFuncSetAttribute(k_hull_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(HW_WARPS * sizeof(HullWarp))));
    }
    CK(cudaSetDevice(devs[0]));
    g_devs = devs; g_sms = sms;
    g_init = true;
    return FPK_OK;
}
extern "C" int fpk_device_count(void) { return g_init ? (int)g_devs.size() : 0; }
static void pool_clear();
static void pin_clear();
extern "C" void fpk_shutdown(void) { pool_clear(); pin_clear(); g_init = false; }

static int ensure_init()
{
    if (g_init) { cudaSetDevice(g_devs[0]); return FPK_OK; }
    return fpk_init(nullptr, 0);
}

// ---- how a host thread waits for its stream ---------------------------------------------------------------------------------------
// cudaStreamSynchronize spins on a core until the device is done; FPK_SYNC=block waits on a cudaEventBlockingSync event instead (the
// thread sleeps). Measured on one B200 with the process pinned to 2, 4 and 16 cores (4 library threads; profiles/README.md): e2e
// 7,102 / 7,191 / 7,218 structures/s spinning, 7,191 / 7,233 / 7,239 blocking - the host threads are not what bounds the call, so
// spinning stays the default and blocking is a knob for boxes where the cores are needed elsewhere.
// Round 2, 8 GPUs on a 32-core box (4 cores per rank; profiles/r02_8gpu_trace.txt): fpk_detect_batch is the same either way, but mdpocket's push
// loop (a staging helper thread beside the spinning caller) gains 10 % when the waits sleep. Default therefore: block when this process has fewer
// than 8 cores per device to itself (torchrun's LOCAL_WORLD_SIZE tells how many ranks share the box), spin otherwise; FPK_SYNC=spin|block overrides.
static bool block_sync()
{
    static const int mode = []() {
        if (const char *e = getenv("FPK_SYNC")) return !strcmp(e, "block") ? 1 : 0;
        const int cores = (int)std::thread::hardware_concurrency();
        const char *lw = getenv("LOCAL_WORLD_SIZE");
        const int ranks = lw ? std::max(1, atoi(lw)) : 1;
        return (cores > 0 && cores / ranks < 8) ? 1 : 0;
    }();
    return mode != 0;
}
static cudaError_t wait_stream()
{
    if (!block_sync()) return cudaStreamSynchronize(cudaStreamPerThread);
    static thread_local cudaEvent_t ev = nullptr;
    if (!ev) { cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming); if (e != cudaSuccess) return e; }
    cudaError_t e = cudaEventRecord(ev, cudaStreamPerThread);
    return e != cudaSuccess ? e : cudaEventSynchronize(ev);
}

// ------------------------------------------------------------------------------------------------------------------
struct Arena {
    char *base = nullptr;
    size_t cap = 0, used = 0;
    template <class T> T *take(size_t n)
    {
        used = (used + 255) & ~(size_t)255;
        T *p = reinterpret_cast<T *>(base + used);
        used += n * sizeof(T);
        return p;
    }
    void release() { if (base) cudaFree(base); base = nullptr; cap = used = 0; }
};
// first pass with base == nullptr just measures
template <class F> static int arena_build(Arena &A, F layout)
{
    Arena m; m.base = nullptr; m.used = 0;
    layout(m);
    size_t need = m.used + 4096;
    if (need > A.cap) {
        need += need / 4;                  // headroom: pooled contexts see chunks of slightly different sizes
        A.release();
        cudaError_t e = cudaMalloc(&A.base, need);
        if (e != cudaSuccess) return fail(FPK_ERR_CUDA, "cudaMalloc(%zu bytes): %s", need, cudaGetErrorString(e));
        A.cap = need;
    }
    A.used = 0;
    layout(A);
    return FPK_OK;
}

// std::vector that does not zero-fill on resize (the arrays are overwritten by a D2H copy right away)
template <class T> struct NoInit : std::allocator<T> {
    template <class U> struct rebind { using other = NoInit<U>; };
    template <class U, class... A> void construct(U *p, A &&...a)
    {
        if constexpr (sizeof...(A) == 0) ::new ((void *)p) U; else ::new ((void *)p) U(std::forward<A>(a)...);
    }
};
template <class T> using rawvec = std::vector<T, NoInit<T>>;

// Pinned host blocks for results, recycled across fetches: a D2H copy into pinned memory is a plain DMA (no driver staging, no
// page faults on fresh pages), and allocating pinned memory is far too slow to do per chunk.
static std::mutex g_pin_mu;
static std::vector<std::pair<char *, size_t>> g_pin_pool;
static char *pin_get(size_t need, size_t *cap)
{
    {
        std::lock_guard<std::mutex> g(g_pin_mu);
        int best = -1;
        for (int i = 0; i < (int)g_pin_pool.size(); i++)
            if (g_pin_pool[i].second >= need && (best < 0 || g_pin_pool[i].second < g_pin_pool[best].second)) best = i;
        if (best >= 0) {
            char *p = g_pin_pool[best].first; *cap = g_pin_pool[best].second;
            g_pin_pool.erase(g_pin_pool.begin() + best);
            return p;
        }
    }
    char *p = nullptr;
    size_t want = need + need / 4 + 4096;
    if (cudaHostAlloc((void **)&p, want, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *cap = want;
    return p;
}
static void pin_put(char *p, size_t cap)
{
    if (!p) return;
    {
        std::lock_guard<std::mutex> g(g_pin_mu);
        if (g_pin_pool.size() < 24) { g_pin_pool.emplace_back(p, cap); return; }
    }
    cudaFreeHost(p);
}
static void pin_clear()
{
    std::lock_guard<std::mutex> g(g_pin_mu);
    for (auto &x : g_pin_pool) cudaFreeHost(x.first);
    g_pin_pool.clear();
}

struct HostResults {            // owner of the host arrays of one fetch: one pinned block, carved up
    std::atomic<int> refs{0};
    char *blk = nullptr; size_t cap = 0, used = 0;
    std::vector<int> tets;      // only with want_tets
    template <class T> T *take(size_t n)
    {
        used = (used + 255) & ~(size_t)255;
        T *p = reinterpret_cast<T *>(blk + used);
        used += n * sizeof(T);
        return p;
    }
    ~HostResults() { pin_put(blk, cap); }
};

struct fpk_batch {
    int n = 0;
    int n_cap = 0;                     // mdpocket: frames the skeleton was built for (n may be set lower, md_prepare)
    int dev = -1;                      // the device this context's arenas live on
    fpk_params prm;
    std::vector<fpk_structure> sane;   // the caller's structures, unusable ones replaced by an empty structure
    std::vector<int> bad;              // FPK_ERR_BAD_ARG for those
    std::vector<StructDev> S;
    std::vector<int> natoms;
    long long tot_atoms = 0, tot_w = 0, tot_sites = 0, tot_xlig = 0, tot_sphcap = 0, tot_tetcap = 0, tot_lig = 0;
    bool has_names = false;
    std::vector<int> lig_n;
    LigDev Lg;                         // dpocket ligand queries (tot_lig > 0)
    std::vector<int> h_ni;
    std::vector<float> h_ligvol;
    int max_sites = 0, max_w = 0, max_sphcap = 0;
    // structures above k1_max_sites are triangulated one at a time by the WHOLE GPU (k_delaunay_grid), the others by one block each (K1)
    int k1_max_sites = 0x7fffffff, max_sites_big = 0, max_sphcap_big = 0, max_sphcap_small = 0;
    std::vector<int> big;              // indices of the former
    GridDel *dGrid = nullptr; int nGridBlocks = 0; size_t grid_dsm = 0;
    DelScratch hGridX;
    ClusterGridScratch hCluG; int nCluGridBlocks = 0;
    AtomGrid AG;                       // cell grid over the heavy atoms of every structure (K4's surrounding atoms)
    int md_shared_props = 0;           // mdpocket: all frames share the property arrays
    // stage 1 device
    Arena A1, A2;
    Arena AL;                         // -C m | a | c: pool of distance matrices (k_linkage), sized from the sphere counts of the run
    BatchDev B;
    StructDev *dS = nullptr;
    DelScratch *dDel = nullptr; int nDel = 0;
    int k1_picap = 0, hull_picap = 0; size_t k1_dsm = 0, hull_dsm = 0;   // shared-memory plan of the two Delaunay kernels
    ClusterScratch *dClu = nullptr; int nClu = 0;
    int *d_counters = nullptr;         // [8] work counters
    char *hin = nullptr; size_t hin_cap = 0, in_bytes = 0;     // pinned host mirror of the arena's input prefix (one H2D copy)
    int *d_order = nullptr;
    // stage 2
    PocketDev Pk;
    int *d_clbase = nullptr, *d_sbase = nullptr;
    DelScratch *dHull = nullptr; int nHull = 0;
    int *d_hull_fb = nullptr; int nHullWarp = 0;     // pockets the warp-per-pocket hull kernel hands to the Delaunay route
    unsigned char *d_small_done = nullptr;           // clusters k_descriptors_small (one warp each) has described
    DescScratch *dDesc = nullptr; int nDesc = 0;
    // compact outputs (device)
    float4 *c_xyzr = nullptr; float *c_bary = nullptr; int4 *c_neigh = nullptr; uint8_t *c_type = nullptr; int *c_cluster = nullptr;
    int *c_cl_sph = nullptr, *c_cl_off = nullptr;
    std::vector<int> h_counts, h_clbase, h_sbase;
    int *pin_counts = nullptr; size_t pin_counts_cap = 0;      // pinned staging of the per-structure counts (read back without spinning)
    long long tot_spheres = 0, tot_clusters = 0, tot_tets = 0;
    // timing / accounting
    cudaEvent_t ev[8];
    float ms[8] = {0};
    long long launches = 0, h2d = 0, d2h = 0;
    bool events = false;
    ~fpk_batch()
    {
        int cur = -1;
        cudaGetDevice(&cur);
        if (dev >= 0 && dev != cur) cudaSetDevice(dev);
        A1.release(); A2.release(); AL.release();
        if (hin) cudaFreeHost(hin);
        if (pin_counts) cudaFreeHost(pin_counts);
        if (events) for (auto &e : ev) cudaEventDestroy(e);
        if (dev >= 0 && dev != cur && cur >= 0) cudaSetDevice(cur);
    }
};

static inline size_t sph_cap_for(const fpk_params &p, int nsites, double factor_override)
{
    // typical proteins keep ~0.6 spheres per heavy atom at default radii; everything (7 tets/atom) otherwise
    double f = factor_override > 0 ? factor_override : ((p.asph_max_size <= 6.3f && p.asph_min_size >= 3.3f) ? 1.6 : 7.2);
    size_t slack = 64;
    if (factor_override <= 0)
        if (const char *e = getenv("FPK_SPH_CAP_FACTOR")) { f = atof(e); slack = 0; }      // test knob: forces the FPK_ERR_CAPACITY re-run (tests/test_gpu_parity.py)
    return (size_t)(f * nsites) + slack;
}

// what makes ONE structure unusable; the batch goes on without it (the reference's -F loop moves on to the next file, fpmain.c:61-76)
static const char *structure_defect(const fpk_structure &s)
{
    if (s.natoms < 0 || s.natoms_w < 0 || s.n_xlig < 0 || s.n_lig < 0) return "negative count";
    if (s.natoms > 0 && (!s.x || !s.y || !s.z || !s.radius || !s.electroneg || !s.bfactor || !s.atom_id || !s.res_id || !s.aa_idx || !s.flags)) return "NULL atom array";
    if (s.natoms_w > 0 && (!s.wx || !s.wy || !s.wz || !s.wradius || !s.welectroneg || !s.wflags)) return "NULL pdb_w_lig array";
    if (s.natoms_w == 0 && !s.same_pdb) return "natoms_w == 0 requires same_pdb = 1";
    if (s.n_xlig > 0 && !s.xlig) return "NULL explicit-ligand array";
    if (s.n_lig > 0 && (!s.lig_atoms || !s.lig_mol)) return "NULL ligand array";
    for (int i = 0; i < s.n_lig; i++)
        if (s.lig_atoms[i] < 0 || s.lig_atoms[i] >= (s.natoms_w ? s.natoms_w : s.natoms)) return "ligand atom index out of range";
    // contacted atoms are made unique by id (pocket.c:1298) but ASA works on pointers (asa.c:51): both coincide iff ids are unique
    bool inc = true;
    for (int i = 1; i < s.natoms && inc; i++) inc = s.atom_id[i] > s.atom_id[i - 1];
    if (!inc) {
        std::vector<int> ids(s.atom_id, s.atom_id + s.natoms);
        std::sort(ids.begin(), ids.end());
        if (std::adjacent_find(ids.begin(), ids.end()) != ids.end()) return "duplicate atom ids are not supported";
    }
    return nullptr;
}

static int batch_pack(fpk_batch *b, const fpk_structure *in, int n, const fpk_params *p, double cap_factor)
{
    b->n = n; b->prm = *p;
    b->S.resize(n); b->natoms.resize(n); b->lig_n.assign(n, 0);
    b->sane.assign(in, in + n); b->bad.assign(n, 0);
    b->max_sites = b->max_w = b->max_sphcap = 0; b->h2d = b->d2h = 0;
    b->max_sites_big = b->max_sphcap_big = b->max_sphcap_small = 0; b->big.clear();
    // One block per structure is the right grain for a batch; a large complex (BASELINE configs[3]) or a lone structure would leave the
    // GPU idle: from 12,000 sites on - or 2,500 when the call holds at most four structures - the whole GPU works on ONE triangulation.
    {
        int thr = 12000;
        if (const char *e = getenv("FPK_GRID_MIN_SITES")) thr = std::max(4, atoi(e));
        else if (n <= 4) thr = 2500;
        b->k1_max_sites = thr - 1;
    }
    long long ao = 0, wo = 0, so = 0, xo = 0, sc = 0, tc = 0, lo = 0;
    b->has_names = false;
    for (int k = 0; k < n; k++) {
        if (const char *why = structure_defect(in[k])) {     // reported in out[k].status; the rest of the batch is processed
            fail(FPK_ERR_BAD_ARG, "structure %d: %s", k, why);
            b->bad[k] = FPK_ERR_BAD_ARG;
            fpk_structure &e = b->sane[k];
            memset(&e, 0, sizeof e);
            e.same_pdb = 1;
        }
        const fpk_structure &s = b->sane[k];
        lo += s.n_lig;
        b->lig_n[k] = s.n_lig;
        if (s.n_lig > 0 && s.atom_name_key) b->has_names = true;
        StructDev &d = b->S[k];
        memset(&d, 0, sizeof d);
        int ns = 0;
        for (int i = 0; i < s.natoms; i++) ns += !(s.flags[i] & FPK_ATOM_H);
        d.atom_off = (int)ao; d.natoms = s.natoms; d.prop_off = (int)ao;
        d.w_off = (int)wo; d.natoms_w = s.natoms_w;
        d.site_off = (int)so; d.nsites = ns;
        d.xlig_off = (int)xo; d.n_xlig = s.n_xlig;
        d.same_pdb = s.same_pdb;
        d.avg_b = s.avg_bfactor; d.min_b = s.min_bfactor; d.max_b = s.max_bfactor;
        d.sph_off = (int)sc; d.sph_cap = (int)sph_cap_for(*p, ns, cap_factor);
        d.tet_off = (int)tc; d.tet_cap = p->want_tets ? 8 * ns + 64 : 0;
        b->natoms[k] = s.natoms;
        ao += s.natoms; wo += s.natoms_w; so += ns; xo += s.n_xlig; sc += d.sph_cap; tc += d.tet_cap;
        if (ns > b->k1_max_sites) { b->big.push_back(k); b->max_sites_big = std::max(b->max_sites_big, ns); b->max_sphcap_big = std::max(b->max_sphcap_big, d.sph_cap); }
        else { b->max_sites = std::max(b->max_sites, ns); b->max_sphcap_small = std::max(b->max_sphcap_small, d.sph_cap); }
        b->max_w = std::max(b->max_w, s.natoms_w ? s.natoms_w : s.natoms);
        b->max_sphcap = std::max(b->max_sphcap, d.sph_cap);
        if (ao > 2000000000ll || sc + n > 2000000000ll || tc > 500000000ll) return fail(FPK_ERR_BAD_ARG, "batch too large for 32-bit offsets; split it");
    }
    b->tot_atoms = ao; b->tot_w = wo; b->tot_sites = so; b->tot_xlig = xo; b->tot_sphcap = sc; b->tot_tetcap = tc; b->tot_lig = lo;
    return FPK_OK;
}

static size_t del_scratch_layout(Arena &a, DelScratch &x, int nmax, int scap, int nthreads, bool with_filter)
{
    int npad = 1;
    while (npad < nmax) npad <<= 1;
    int cap = 10 * nmax + 256;
    x.cap = cap;
    x.P = a.take<double>(3 * (size_t)nmax);
    x.keys = a.take<unsigned long long>(npad);
    x.order = a.take<int>(nmax);
    x.tet = a.take<int>(8 * (size_t)cap); x.claim = a.take<int>(2 * (size_t)cap); x.mark = a.take<int>(2 * (size_t)cap);
    x.free0 = a.take<int>(cap); x.free1 = a.take<int>(cap);
    x.hint = a.take<int>(32 * 32 * 32);
    const int nwarps = (nthreads + 31) / 32;
    x.cav = a.take<int>((size_t)nwarps * CAVMAX); x.newid = a.take<int>((size_t)nwarps * CAVMAX * 4);
    x.tinfo = a.take<int>(4 * (size_t)nwarps); x.cursor = a.take<int>(2 * (size_t)nwarps);
    x.bigcav = a.take<int>(cap); x.bignewid = a.take<int>(4 * (size_t)cap);
    x.cnt = a.take<int>(C_COUNT); x.rstart = a.take<int>(40); x.rep = a.take<int>(nmax); x.Pi = a.take<int>(3 * (size_t)nmax);
    x.scap = with_filter ? scap : 0;
    x.rawkey = a.take<int4>(with_filter ? scap : 1); x.rawxyzr = a.take<float4>(with_filter ? scap : 1);
    x.sitecnt = a.take<int>(with_filter ? nmax + 1 : 1); x.bucket = a.take<int>(with_filter ? scap : 1);
    return a.used;
}

static int batch_upload(fpk_batch *b, const fpk_structure *in)
{
    const int n = b->n;
    if (in) in = b->sane.data();          // unusable structures were replaced by empty ones in batch_pack
    cudaGetDevice(&b->dev);
    int occ1 = 2, occ3 = 4;
    // K1 shared-memory plan. The kernel is bound by the latency of dependent tet / claim loads, so what counts is the number of
    // resident warps: three blocks of eight warps per SM at 80 registers (launch bounds) beat two blocks at 128 by 11 % even though
    // the sites then come through L1 / L2 instead of shared memory (profiles/README.md: 192 -> 171.5 ms per 3,000 structures; at two
    // blocks per SM the shared-memory tile made no difference: 193.6 vs 192.1). The tile (int32 lattice offsets, brought in by TMA
    // bulk copies) is kept when it fits beside the WarpCav records at three blocks per SM, i.e. for structures up to ~2,000 sites.
    {
        const size_t fixed = FPK_BD_BYTES + (FPK_K1_THREADS / 32) * sizeof(WarpCav), stat = 2048, sm_total = 227 * 1024, per_block_reserved = 1024;
        const size_t need = fixed + 12 * (size_t)b->max_sites + 128;          // + slack for the 16-byte aligned TMA window
        const size_t share = (sm_total - FPK_K1_MINBLOCKS * per_block_reserved) / FPK_K1_MINBLOCKS - stat;
        if (need <= share) { b->k1_picap = b->max_sites; b->k1_dsm = need; }
        else { b->k1_picap = 0; b->k1_dsm = fixed; }
        if (getenv("FPK_K1_SITES_GLOBAL")) { b->k1_picap = 0; b->k1_dsm = fixed; }     // tuning knob (profiles/): never stage the sites
    }
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, k_delaunay_filter, FPK_K1_THREADS, b->k1_dsm);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3, k_cluster, 256, 0);
    occ1 = std::max(1, std::min(occ1, 16)); occ3 = std::max(1, std::min(occ3, 4));
    if (const char *e = getenv("FPK_K1_BLOCKS_PER_SM")) occ1 = std::max(1, std::min(occ1, atoi(e)));     // tuning knob (profiles/)
    b->nDel = std::max(1, std::min(n, g_sms * occ1));
    b->nClu = std::max(1, std::min(n, g_sms * occ3));
    std::vector<DelScratch> hDel(b->nDel);
    std::vector<ClusterScratch> hClu(b->nClu);
    BatchDev &B = b->B;
    memset(&B, 0, sizeof B);
    int rc = arena_build(b->A1, [&](Arena &a) {
        b->dS = a.take<StructDev>(n);
        float *ax = a.take<float>(b->tot_atoms), *ay = a.take<float>(b->tot_atoms), *az = a.take<float>(b->tot_atoms);
        float *ar = a.take<float>(b->tot_atoms), *ae = a.take<float>(b->tot_atoms), *ab = a.take<float>(b->tot_atoms);
        int *aid = a.take<int>(b->tot_atoms), *ares = a.take<int>(b->tot_atoms);
        int8_t *aaa = a.take<int8_t>(b->tot_atoms);
        uint8_t *afl = a.take<uint8_t>(b->tot_atoms);
        float *wx = a.take<float>(b->tot_w), *wy = a.take<float>(b->tot_w), *wz = a.take<float>(b->tot_w), *wr = a.take<float>(b->tot_w),
              *we = a.take<float>(b->tot_w);
        uint8_t *wfl = a.take<uint8_t>(b->tot_w);
        int *s2a = a.take<int>(b->tot_sites);
        float *xl = a.take<float>(3 * b->tot_xlig + 1);
        B.ax = ax; B.ay = ay; B.az = az; B.arad = ar; B.aen = ae; B.abf = ab; B.aid = aid; B.ares = ares; B.aaa = aaa; B.afl = afl;
        B.wx = wx; B.wy = wy; B.wz = wz; B.wrad = wr; B.wen = we; B.wfl = wfl; B.site2atom = s2a; B.xlig = xl;
        b->d_order = a.take<int>(n);
        b->dDel = a.take<DelScratch>(b->nDel);
        b->dClu = a.take<ClusterScratch>(b->nClu);
        b->Lg.lig_off = a.take<int>(n + 1); b->Lg.lig_idx = a.take<int>(b->tot_lig + 1); b->Lg.lig_mol = a.take<int>(b->tot_lig + 1);
        b->Lg.name_key = b->has_names ? a.take<int>(b->tot_atoms + 1) : nullptr;
        int *cbase = a.take<int>(n + 1);
        b->AG.cbase = cbase;
        a.take<char>(0);
        b->in_bytes = a.used;                      // everything above is input: mirrored in pinned host memory, one copy
        if (b->tot_lig > 0) {
            b->Lg.xpos = a.take<int>(b->tot_sphcap + n + 1); b->Lg.ipos = a.take<int>(b->tot_atoms + n + 1);
            b->Lg.ilist = a.take<int>(b->tot_atoms + n + 1); b->Lg.ni = a.take<int>(n);
        }
        {
            long long cells = 0;
            for (int k = 0; k < n; k++) cells += (b->S[k].natoms_w ? b->S[k].natoms_w : b->S[k].natoms) / 4 + 66;
            b->AG.cellend = a.take<int>((size_t)cells + 2);
            b->AG.cellatom = a.take<int>((size_t)b->tot_w + (size_t)b->tot_atoms + 2);      // addressed at w_off, or tot_w + atom_off
            b->AG.same_base = (int)b->tot_w; b->AG.min_atoms = K4_GRID_MIN_ATOMS;
            b->AG.geom = a.take<float4>(n); b->AG.dims = a.take<int4>(n);
            b->AG.cbox = a.take<float4>(2 * (((size_t)b->tot_w + (size_t)b->tot_atoms) / 32 + (size_t)n + 2));
        }
        B.s_xyzr = a.take<float4>(b->tot_sphcap); B.s_bary = a.take<float>(3 * b->tot_sphcap);
        B.s_neigh = a.take<int4>(b->tot_sphcap); B.s_key = a.take<int4>(b->tot_sphcap);
        B.s_type = a.take<uint8_t>(b->tot_sphcap); B.s_cluster = a.take<int>(b->tot_sphcap);
        B.cl_off = a.take<int>(b->tot_sphcap + 2 * (size_t)n + 2); B.cl_sph = a.take<int>(b->tot_sphcap);
        B.counts = a.take<int>((size_t)n * CNT_STRIDE);
        b->d_counters = a.take<int>(8);
        B.tets_out = b->tot_tetcap ? a.take<int>(4 * b->tot_tetcap) : nullptr;
        for (int i = 0; i < b->nDel; i++) del_scratch_layout(a, hDel[i], std::max(4, b->max_sites), std::max(64, b->max_sphcap_small), FPK_K1_THREADS, true);
        if (!b->big.empty()) {
            b->dGrid = a.take<GridDel>(1);
            del_scratch_layout(a, b->hGridX, b->max_sites_big, b->max_sphcap_big, FPK_K1_THREADS, true);
            b->hCluG.cellcnt = a.take<int>(CELLMAX + 2); b->hCluG.cellsph = a.take<int>(b->max_sphcap_big + 1);
            b->hCluG.tmp = a.take<int>(b->max_sphcap_big + 2); b->hCluG.bb = a.take<float>(6 * (size_t)g_sms * 8);
            b->hCluG.glob = a.take<int>(4); b->hCluG.scap = b->max_sphcap_big;
        }
        for (int i = 0; i < b->nClu; i++) {
            const int sc = std::max(64, b->max_sphcap_small);
            hClu[i].cellcnt = a.take<int>(CELLMAX + 2); hClu[i].cellsph = a.take<int>(sc + 1);
            hClu[i].tmp = a.take<int>(sc + 2); hClu[i].scap = sc;
        }
    });
    if (rc) return rc;
    B.nstruct = n; B.S = b->dS; B.prm = b->prm; B.work_order = b->d_order; B.work_counter = b->d_counters; B.k1_max_sites = b->k1_max_sites;
    if (!b->big.empty()) {
        int occg = 1;
        b->grid_dsm = FPK_BD_BYTES + (FPK_K1_THREADS / 32) * sizeof(WarpCav);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occg, k_delaunay_grid, FPK_K1_THREADS, b->grid_dsm);
        // ONE block per SM: the kernel is bound by its grid barriers (82 % of warp samples), whose cost grows with the number of blocks, not by
        // the number of points in flight - measured on the 150,000-atom complex: 25.2 ms with 148 blocks, 28.7 with 296, 27.8 with 444
        // (profiles/r02_k1g_blocks_per_sm.log)
        occg = std::max(1, std::min(occg, 1));
        if (const char *e = getenv("FPK_GRID_BLOCKS_PER_SM")) occg = std::max(1, std::min(FPK_K1_MINBLOCKS, atoi(e)));      // tuning knob (profiles/)
        b->nGridBlocks = g_sms * occg;
        int occc = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occc, k_cluster_grid, 256, 0);
        b->nCluGridBlocks = g_sms * std::max(1, std::min(occc, 4));
    }
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return b->S[x].nsites > b->S[y].nsites; });
    std::vector<int> hcb(n + 1, 0);
    for (int k = 0; k < n; k++) hcb[k + 1] = hcb[k] + (b->S[k].natoms_w ? b->S[k].natoms_w : b->S[k].natoms) / 4 + 66;
    if (!in) {          // mdpocket skeleton: descriptors only; properties and coordinates are uploaded by the caller
        CK(cudaMemcpy(const_cast<int *>(b->AG.cbase), hcb.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(b->dS, b->S.data(), sizeof(StructDev) * n, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(b->d_order, order.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(b->dDel, hDel.data(), sizeof(DelScratch) * b->nDel, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(b->dClu, hClu.data(), sizeof(ClusterScratch) * b->nClu, cudaMemcpyHostToDevice));
        b->h2d += (long long)(sizeof(StructDev) + sizeof(int)) * n;
    } else {
        // fill the pinned mirror at the arena's own offsets, then ONE asynchronous copy of the whole input prefix
        if (b->in_bytes > b->hin_cap) {
            if (b->hin) cudaFreeHost(b->hin);
            b->hin = nullptr; b->hin_cap = 0;
            const size_t want = b->in_bytes + b->in_bytes / 8 + 4096;
            cudaError_t e = cudaHostAlloc((void **)&b->hin, want, cudaHostAllocPortable);
            if (e != cudaSuccess) return fail(FPK_ERR_CUDA, "cudaHostAlloc(%zu bytes): %s", want, cudaGetErrorString(e));
            b->hin_cap = want;
        }
        char *H = b->hin;
        const char *D0 = b->A1.base;
        auto mir = [&](const void *dev) { return H + (reinterpret_cast<const char *>(dev) - D0); };
        memcpy(mir(b->dS), b->S.data(), sizeof(StructDev) * n);
        memcpy(mir(b->d_order), order.data(), sizeof(int) * n);
        memcpy(mir(b->dDel), hDel.data(), sizeof(DelScratch) * b->nDel);
        memcpy(mir(b->dClu), hClu.data(), sizeof(ClusterScratch) * b->nClu);
        memcpy(mir(b->AG.cbase), hcb.data(), sizeof(int) * (n + 1));
        float *hx = (float *)mir(B.ax), *hy = (float *)mir(B.ay), *hz = (float *)mir(B.az), *hr = (float *)mir(B.arad), *he = (float *)mir(B.aen),
              *hb = (float *)mir(B.abf);
        int *hid = (int *)mir(B.aid), *hres = (int *)mir(B.ares), *hs2a = (int *)mir(B.site2atom);
        int8_t *haa = (int8_t *)mir(B.aaa);
        uint8_t *hfl = (uint8_t *)mir(B.afl), *hwf = (uint8_t *)mir(B.wfl);
        float *hwx = (float *)mir(B.wx), *hwy = (float *)mir(B.wy), *hwz = (float *)mir(B.wz), *hwr = (float *)mir(B.wrad), *hwe = (float *)mir(B.wen);
        float *hxl = (float *)mir(B.xlig);
        for (int k = 0; k < n; k++) {
            const fpk_structure &q = in[k];
            const StructDev &d = b->S[k];
            const size_t na = (size_t)q.natoms, nw = (size_t)q.natoms_w;
            if (na) {
                memcpy(hx + d.atom_off, q.x, 4 * na); memcpy(hy + d.atom_off, q.y, 4 * na); memcpy(hz + d.atom_off, q.z, 4 * na);
                memcpy(hr + d.atom_off, q.radius, 4 * na); memcpy(he + d.atom_off, q.electroneg, 4 * na); memcpy(hb + d.atom_off, q.bfactor, 4 * na);
                memcpy(hid + d.atom_off, q.atom_id, 4 * na); memcpy(hres + d.atom_off, q.res_id, 4 * na);
                memcpy(haa + d.atom_off, q.aa_idx, na); memcpy(hfl + d.atom_off, q.flags, na);
            }
            if (nw) {
                memcpy(hwx + d.w_off, q.wx, 4 * nw); memcpy(hwy + d.w_off, q.wy, 4 * nw); memcpy(hwz + d.w_off, q.wz, 4 * nw);
                memcpy(hwr + d.w_off, q.wradius, 4 * nw); memcpy(hwe + d.w_off, q.welectroneg, 4 * nw); memcpy(hwf + d.w_off, q.wflags, nw);
            }
            int *s2 = hs2a + d.site_off;                                      // site map (voronoi.c:94-107 h_tr)
            int so = 0;
            for (int i = 0; i < q.natoms; i++) if (!(q.flags[i] & FPK_ATOM_H)) s2[so++] = i;
            if (q.n_xlig) memcpy(hxl + 3 * (size_t)d.xlig_off, q.xlig, sizeof(float) * 3 * q.n_xlig);
        }
        {
            int *hlo = (int *)mir(b->Lg.lig_off), *hli = (int *)mir(b->Lg.lig_idx), *hlm = (int *)mir(b->Lg.lig_mol);
            int *hnk = b->has_names ? (int *)mir(b->Lg.name_key) : nullptr;
            int lo = 0;
            for (int k = 0; k < n; k++) {
                hlo[k] = lo;
                if (in[k].n_lig > 0) { memcpy(hli + lo, in[k].lig_atoms, 4 * (size_t)in[k].n_lig); memcpy(hlm + lo, in[k].lig_mol, 4 * (size_t)in[k].n_lig); lo += in[k].n_lig; }
                if (hnk) {
                    if (in[k].atom_name_key) memcpy(hnk + b->S[k].atom_off, in[k].atom_name_key, 4 * (size_t)in[k].natoms);
                    else for (int i = 0; i < in[k].natoms; i++) hnk[b->S[k].atom_off + i] = -1 - i;      // no names: never equal
                }
            }
            hlo[n] = lo;
        }
        CK(cudaMemcpyAsync(b->A1.base, H, b->in_bytes, cudaMemcpyHostToDevice, cudaStreamPerThread));
        b->h2d += (long long)b->in_bytes;
    }
    if (!b->events) { for (auto &e : b->ev) CK(cudaEventCreate(&e)); b->events = true; }
    return FPK_OK;
}

__global__ void k_compact(BatchDev B, const int *sbase, float4 *xyzr, float *bary, int4 *neigh, uint8_t *type, int *cluster, int *cl_sph,
                          int *cl_off, const int *clbase)
{
    const int k = blockIdx.x;
    const StructDev S = B.S[k];
    const int *counts = B.counts + (size_t)k * CNT_STRIDE;
    const int ns = (counts[CNT_STATUS] == FPK_OK || counts[CNT_STATUS] == FPK_ERR_NO_POCKETS || counts[CNT_STATUS] == FPK_ERR_NO_SPHERES) ? counts[CNT_SPHERES] : 0;
    const int nc = counts[CNT_CLUSTERS];
    const int d0 = sbase[k];
    for (int i = threadIdx.x; i < ns; i += blockDim.x) {
        const int s = S.sph_off + i;
        xyzr[d0 + i] = B.s_xyzr[s]; neigh[d0 + i] = B.s_neigh[s]; type[d0 + i] = B.s_type[s]; cluster[d0 + i] = nc ? B.s_cluster[s] : 0;
        bary[3 * (size_t)(d0 + i)] = B.s_bary[3 * (size_t)s]; bary[3 * (size_t)(d0 + i) + 1] = B.s_bary[3 * (size_t)s + 1];
        bary[3 * (size_t)(d0 + i) + 2] = B.s_bary[3 * (size_t)s + 2];
        cl_sph[d0 + i] = nc ? B.cl_sph[s] : 0;
    }
    if (counts[CNT_XP] == 1)        // dpocket's explicit pocket: its sphere list sits after the partition
        for (int i = threadIdx.x; i < counts[CNT_NX]; i += blockDim.x) cl_sph[d0 + ns + i] = B.cl_sph[S.sph_off + ns + i];
    for (int i = threadIdx.x; i <= nc; i += blockDim.x) cl_off[clbase[k] + k + i] = B.cl_off[S.sph_off + k + i];
}

static int fetch_counts(fpk_batch *b);

// -C m | a | c (k_linkage.cuh): the sphere counts come back first (one more host round trip, only on this route), the pool of distance
// matrices is sized from them (at most 16 GB per worker context), and the structures are clustered in as many waves as the pool needs. A structure on the whole-GPU route, or
// whose matrix alone exceeds the pool, is refused with FPK_ERR_BAD_ARG (the reference would need the same n^2 / 2 doubles and O(n^3) time).
static int run_linkage(fpk_batch *b)
{
    const int n = b->n;
    { int rc = fetch_counts(b); if (rc) return rc; }
    size_t freeb = 0, totb = 0;
    CK(cudaMemGetInfo(&freeb, &totb));
    // half of what is free, at most 16 GB: up to eight worker contexts of fpk_detect_batch run side by side on a device, each with its own pool
    size_t budget = std::min((freeb + b->AL.cap) / 2, (size_t)16 << 30);
    if (const char *e = getenv("FPK_LINK_POOL_MB")) budget = (size_t)atof(e) << 20;
    std::vector<LinkItem> items;
    std::vector<int> wave_start, refused;
    size_t used = 0, pool = 0;
    for (int k = 0; k < n; k++) {
        const int *c = &b->h_counts[(size_t)k * CNT_STRIDE];
        if (c[CNT_STATUS] != FPK_OK || c[CNT_SPHERES] < 2) continue;
        const size_t need = (linkage_bytes((size_t)c[CNT_SPHERES]) + 255) & ~(size_t)255;
        if (b->S[k].nsites > b->k1_max_sites || need > budget) { refused.push_back(k); continue; }
        if (wave_start.empty() || used + need > budget) { wave_start.push_back((int)items.size()); used = 0; }
        items.push_back(LinkItem{k, c[CNT_SPHERES], (unsigned long long)used});
        used += need;
        pool = std::max(pool, used);
    }
    for (int k : refused) {
        const int v = FPK_ERR_BAD_ARG;
        CK(cudaMemcpyAsync(b->B.counts + (size_t)k * CNT_STRIDE + CNT_STATUS, &v, sizeof v, cudaMemcpyHostToDevice, cudaStreamPerThread));
        CK(wait_stream());
    }
    if (items.empty()) return FPK_OK;
    LinkItem *d_items = nullptr;
    unsigned char *d_pool = nullptr;
    int rc = arena_build(b->AL, [&](Arena &a) { d_items = a.take<LinkItem>(items.size()); d_pool = a.take<unsigned char>(pool + 256); });
    if (rc) return rc;
    CK(cudaMemcpyAsync(d_items, items.data(), sizeof(LinkItem) * items.size(), cudaMemcpyHostToDevice, cudaStreamPerThread));
    CK(wait_stream());
    wave_start.push_back((int)items.size());
    for (size_t w = 0; w + 1 < wave_start.size(); w++) {
        k_linkage<<<wave_start[w + 1] - wave_start[w], 256, 0, cudaStreamPerThread>>>(b->B, d_items + wave_start[w], d_pool);
        b->launches++;
    }
    CK(cudaGetLastError());
    return FPK_OK;
}

static int batch_run(fpk_batch *b)
{
    const int n = b->n;
    BatchDev &B = b->B;
    b->launches = 0;
    CK(cudaMemsetAsync(b->d_counters, 0, sizeof(int) * 8));
    CK(cudaMemsetAsync(B.counts, 0, sizeof(int) * (size_t)n * CNT_STRIDE));
    CK(cudaEventRecord(b->ev[0]));
    for (int kb : b->big) {            // large structures: one cooperative launch each, every resident block on the same triangulation
        if (b->S[kb].nsites < 4 || kb >= n) continue;
        void *args[] = {(void *)&B, (void *)&b->hGridX, (void *)&b->dGrid, (void *)&kb};
        CK(cudaLaunchCooperativeKernel((void *)k_delaunay_grid, dim3(b->nGridBlocks), dim3(FPK_K1_THREADS), args, b->grid_dsm, cudaStreamPerThread));
        b->launches++;
    }
    k_delaunay_filter<<<b->nDel, FPK_K1_THREADS, b->k1_dsm>>>(B, b->dDel, b->k1_picap);
    CK(cudaEventRecord(b->ev[1]));
    {
        const int cm = b->prm.clustering_method;
        if ((cm == 'm' || cm == 'a' || cm == 'c') && !b->prm.skip_clustering) { int rcl = run_linkage(b); if (rcl) return rcl; }
    }
    for (int kb : b->big) {
        if (kb >= n) continue;
        void *args[] = {(void *)&B, (void *)&b->hCluG, (void *)&kb};
        CK(cudaLaunchCooperativeKernel((void *)k_cluster_grid, dim3(b->nCluGridBlocks), dim3(256), args, 0, cudaStreamPerThread));
        b->launches++;
    }
    k_cluster<<<b->nClu, 256>>>(B, b->dClu, b->d_counters + 1);
    const bool lig = b->tot_lig > 0 && b->prm.skip_clustering == 0;
    if (lig) { k_ligand_select<<<n, 256>>>(B, b->Lg); b->launches++; }
    if (b->prm.do_asa_and_volume) { k_atom_grid<<<n, 256>>>(B, b->AG); b->launches++; }
    CK(cudaEventRecord(b->ev[2]));
    b->launches += 2;
    // ---- the one host round trip: per-structure counts -> exact sizes of the pocket-level arrays -----------------------
    { int rcq = fetch_counts(b); if (rcq) return rcq; }
    b->h_clbase.assign(n + 1, 0); b->h_sbase.assign(n + 1, 0);
    int max_ns = 1;
    long long tt = 0;
    for (int k = 0; k < n; k++) {
        const int *c = &b->h_counts[(size_t)k * CNT_STRIDE];
        bool have = c[CNT_STATUS] == FPK_OK || c[CNT_STATUS] == FPK_ERR_NO_SPHERES || c[CNT_STATUS] == FPK_ERR_NO_POCKETS;
        int ns = have ? c[CNT_SPHERES] : 0;
        b->h_clbase[k + 1] = b->h_clbase[k] + c[CNT_CLUSTERS];
        b->h_sbase[k + 1] = b->h_sbase[k] + ns + (have && c[CNT_XP] == 1 ? c[CNT_NX] : 0);      // room for the explicit pocket's pocket-level arrays
        max_ns = std::max(max_ns, ns);
        tt += c[CNT_TETS];
    }
    b->tot_clusters = b->h_clbase[n]; b->tot_spheres = b->h_sbase[n]; b->tot_tets = tt;
    const long long NC = b->tot_clusters, NS = b->tot_spheres;
    int occH = 8, occD = 8;
    // the pocket's centres stay in the block's global scratch (L1 / L2): measured 43.1 vs 45.1 ms per 3,000 structures with the shared tile
    b->hull_picap = getenv("FPK_HULL_SITES_SHARED") ? std::min(max_ns, 1024) : 0;
    b->hull_dsm = FPK_BD_BYTES + (HULL_THREADS / 32) * sizeof(WarpCav) + 12 * (size_t)b->hull_picap + 16;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occH, k_hull_volume, HULL_THREADS, b->hull_dsm);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occD, k_descriptors, K4_THREADS, 0);
    occH = std::max(1, std::min(occH, FPK_HULL_MINBLOCKS)); occD = std::max(1, std::min(occD, 8));
    b->nHull = (int)std::max(1ll, std::min<long long>(NC, (long long)g_sms * occH));
    int occW = FPK_HW_MINBLOCKS;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occW, k_hull_warp, HW_THREADS, HW_WARPS * sizeof(HullWarp));
    b->nHullWarp = (int)std::max(1ll, std::min<long long>((NC + 32 * HW_WARPS - 1) / (32 * HW_WARPS), (long long)g_sms * std::max(1, occW)));
    b->nDesc = (int)std::max(1ll, std::min<long long>(NC, (long long)g_sms * occD));
    std::vector<DelScratch> hHull(b->nHull);
    std::vector<DescScratch> hDesc(b->nDesc);
    PocketDev &Pk = b->Pk;
    memset(&Pk, 0, sizeof Pk);
    const bool full = b->prm.do_asa_and_volume != 0;
    int rc = arena_build(b->A2, [&](Arena &a) {
        b->d_clbase = a.take<int>(n + 1); b->d_sbase = a.take<int>(n + 1);
        Pk.desc = a.take<fpk_desc>(NC + 1);
        Pk.cl_atoms = a.take<int>(4 * NS + 4); Pk.ent_u = a.take<int>(4 * NS + 4); Pk.touch_off = a.take<int>(5 * NS + 5);
        Pk.touch_list = a.take<int>(4 * NS + 4); Pk.acc = a.take<int>(16 * NS + 16); Pk.fx_val = a.take<float>(16 * NS + 16);
        Pk.fx_who = a.take<int>(2 * b->tot_atoms + 2); Pk.atom_fx = a.take<float>(4 * b->tot_atoms + 4);
        Pk.rank_to_cluster = a.take<int>(NS + n + 1);
        b->c_xyzr = a.take<float4>(NS + 1); b->c_bary = a.take<float>(3 * NS + 3); b->c_neigh = a.take<int4>(NS + 1);
        b->c_type = a.take<uint8_t>(NS + 1); b->c_cluster = a.take<int>(NS + 1); b->c_cl_sph = a.take<int>(NS + 1);
        b->c_cl_off = a.take<int>(NC + n + 1);
        b->dHull = a.take<DelScratch>(b->nHull); b->dDesc = a.take<DescScratch>(b->nDesc);
        b->d_hull_fb = a.take<int>(NC + 1);
        b->d_small_done = a.take<unsigned char>(NC + 1);
        if (b->tot_lig > 0) { b->Lg.crit = a.take<float>(4 * NC + 4); b->Lg.lig_vol = a.take<float>(n + 1); }
        if (full) for (int i = 0; i < b->nHull; i++) del_scratch_layout(a, hHull[i], max_ns, 0, HULL_THREADS, false);
        size_t sacap = 64;
        while (sacap < (size_t)b->max_w + 1) sacap <<= 1;          // the list is sorted by a bitonic network: padded to a power of two
        const size_t nbits = b->max_w > K4_GRID_MIN_ATOMS ? (size_t)b->max_w / 32 + 2 : 1;       // the bit map is only used from K4_GRID_MIN_ATOMS atoms on
        for (int i = 0; i < b->nDesc; i++) {
            hDesc[i].sa = a.take<int>(sacap); hDesc[i].bits = a.take<unsigned>(nbits);
            hDesc[i].sph = a.take<float4>(K4_STAGE_MAX); hDesc[i].sphk = a.take<int>(K4_STAGE_MAX);
            hDesc[i].sphmc = a.take<float4>(K4_STAGE_MAX); hDesc[i].mcent = a.take<unsigned short>(K4_MC_GENT);
        }
    });
    if (rc) return rc;
    Pk.clbase = b->d_clbase; Pk.sbase = b->d_sbase; Pk.total_clusters = (int)NC;
    Pk.xp_ilist = lig ? b->Lg.ilist : nullptr; Pk.xp_ni = lig ? b->Lg.ni : nullptr;
    CK(cudaMemcpyAsync(b->d_clbase, b->h_clbase.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpyAsync(b->d_sbase, b->h_sbase.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice));
    if (full) CK(cudaMemcpyAsync(b->dHull, hHull.data(), sizeof(DelScratch) * b->nHull, cudaMemcpyHostToDevice));
    CK(cudaMemcpyAsync(b->dDesc, hDesc.data(), sizeof(DescScratch) * b->nDesc, cudaMemcpyHostToDevice));
    CK(cudaMemsetAsync(Pk.desc, 0, sizeof(fpk_desc) * (size_t)(NC + 1)));
    CK(cudaMemsetAsync(Pk.fx_who, 0xff, sizeof(int) * (2 * (size_t)b->tot_atoms + 2)));
    CK(cudaMemsetAsync(Pk.atom_fx, 0, sizeof(float) * (4 * (size_t)b->tot_atoms + 4)));
    // stage-2 arrays are addressed through the compact sphere base: rebase the per-structure offsets used by the pocket kernels
    // (the kernels compute R = sph_off + cl_off[c]; with sph_off replaced by sbase for these arrays). See PocketDev.
    CK(cudaEventRecord(b->ev[3]));
    if (NC > 0) {
        if (full) {
            // K4b: one warp per pocket (incremental exact hull in shared memory); what it declines goes through the block-per-pocket Delaunay
            static const int hw_enable = getenv("FPK_HULL_WARP") ? atoi(getenv("FPK_HULL_WARP")) : 1;
            k_hull_warp<<<b->nHullWarp, HW_THREADS, HW_WARPS * sizeof(HullWarp)>>>(B, Pk, b->d_counters + 4, b->d_hull_fb, b->d_counters + 5, hw_enable);
            k_hull_volume<<<b->nHull, HULL_THREADS, b->hull_dsm>>>(B, Pk, b->dHull, b->d_counters + 2, b->hull_picap, b->d_hull_fb, b->d_counters + 5);
            b->launches += 2;
        }
        CK(cudaEventRecord(b->ev[4]));
        // small clusters (fewer than min_pock_nb_asph spheres: nine in ten) by one warp each, the others by one block each
        static const int small_enable = getenv("FPK_K4_SMALL") ? atoi(getenv("FPK_K4_SMALL")) : 1;
        CK(cudaMemsetAsync(b->d_small_done, 0, (size_t)NC + 1, cudaStreamPerThread));
        if (small_enable) {
            const int nsb = (int)std::max(1ll, std::min<long long>((NC + 32 * K4S_WARPS - 1) / (32 * K4S_WARPS), (long long)g_sms * 8));
            k_descriptors_small<<<nsb, K4S_THREADS>>>(B, Pk, b->d_counters + 6, b->AG, b->d_small_done);
            b->launches++;
        }
        k_descriptors<<<b->nDesc, K4_THREADS>>>(B, Pk, b->dDesc, b->d_counters + 3, b->AG, b->d_small_done);
        CK(cudaEventRecord(b->ev[5]));
        k_finalize<<<n, 128>>>(B, Pk);
        if (full) { k_atom_fx<<<(unsigned)NC, 64>>>(B, Pk); b->launches++; }
        b->launches += 2;
    } else {
        CK(cudaEventRecord(b->ev[4])); CK(cudaEventRecord(b->ev[5]));
    }
    if (lig) CK(cudaMemsetAsync(b->Lg.lig_vol, 0, sizeof(float) * (size_t)n, cudaStreamPerThread));
    if (lig && NC > 0) { k_ligand<<<n, 256>>>(B, Pk, b->Lg); b->launches++; }
    k_compact<<<n, 128>>>(B, b->d_sbase, b->c_xyzr, b->c_bary, b->c_neigh, b->c_type, b->c_cluster, b->c_cl_sph, b->c_cl_off, b->d_clbase);
    b->launches++;
    CK(cudaEventRecord(b->ev[6]));
    // refresh counts (pockets / final status written by k_finalize); the wait inside is the end of the run
    { int rcq = fetch_counts(b); if (rcq) return rcq; }
    CK(cudaGetLastError());
    for (int i = 0; i < 6; i++) cudaEventElapsedTime(&b->ms[i], b->ev[i], b->ev[i + 1]);
#ifdef FPK_K4_PROFILE
    {
        unsigned long long h[24] = {0}, z[24] = {0};
        cudaMemcpyFromSymbol(h, g_k4prof, sizeof h);
        cudaMemcpyToSymbol(g_k4prof, z, sizeof z);
        fprintf(stderr, "k_descriptors split: %llu clusters below min_pock_nb_asph, %.1f M block cycles (%.0f each); %llu others, %.1f M block cycles (%.0f each)\n",
                h[1], h[0] / 1e6, h[1] ? (double)h[0] / h[1] : 0.0, h[3], h[2] / 1e6, h[3] ? (double)h[2] / h[3] : 0.0);
        if (h[1]) fprintf(stderr, "k_descriptors split, the first kind by phase (cycles per cluster): fetch + contacted atoms %.0f, CSR %.0f, sums / surroundings / SASA %.0f, per-atom epilogue %.0f, record %.0f\n",
                          (double)h[4] / h[1], (double)h[5] / h[1], (double)h[6] / h[1], (double)h[7] / h[1], (double)h[8] / h[1]);
        if (h[12]) fprintf(stderr, "k_descriptors split, %llu clusters of >= 256 spheres (%.0f on average), k-cycles per cluster: contacted atoms %.0f, CSR %.0f, then from there: warp 0's sums done at %.0f, "
                           "surroundings at %.0f, last SASA atom at %.0f, Monte-Carlo (all) at %.0f; per-atom epilogue %.0f, record %.0f\n", h[12], (double)h[13] / h[12],
                           h[14] / 1e3 / h[12], h[15] / 1e3 / h[12], h[16] / 1e3 / h[12], h[17] / 1e3 / h[12], h[18] / 1e3 / h[12], h[19] / 1e3 / h[12], h[20] / 1e3 / h[12], h[21] / 1e3 / h[12]);
    }
#endif
    cudaEventElapsedTime(&b->ms[7], b->ev[0], b->ev[6]);
    return FPK_OK;
}

static int fetch_counts(fpk_batch *b)
{
    const size_t cnt = (size_t)b->n * CNT_STRIDE;
    b->h_counts.resize(cnt);
    if (cnt > b->pin_counts_cap) {
        if (b->pin_counts) cudaFreeHost(b->pin_counts);
        b->pin_counts = nullptr; b->pin_counts_cap = 0;
        CK(cudaHostAlloc((void **)&b->pin_counts, sizeof(int) * (cnt + cnt / 4 + 64), cudaHostAllocPortable));
        b->pin_counts_cap = cnt + cnt / 4 + 64;
    }
    CK(cudaMemcpyAsync(b->pin_counts, b->B.counts, sizeof(int) * cnt, cudaMemcpyDeviceToHost, cudaStreamPerThread));
    CK(wait_stream());
    memcpy(b->h_counts.data(), b->pin_counts, sizeof(int) * cnt);
    b->d2h += (long long)(sizeof(int) * cnt);
    return FPK_OK;
}

template <class T> static int d2h(fpk_batch *b, T *dst, const void *src, size_t count)
{
    if (count) { CK(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyDeviceToHost, cudaStreamPerThread)); b->d2h += (long long)(count * sizeof(T)); }
    return FPK_OK;
}

static inline bool in_has_lig(const fpk_batch *b, int k) { return b->lig_n[k] > 0; }

static int batch_fetch(fpk_batch *b, fpk_result *out)
{
    const int n = b->n;
    HostResults *H = new HostResults();
    const size_t NS = (size_t)b->tot_spheres, NC = (size_t)b->tot_clusters, NA = (size_t)b->tot_atoms;
    const size_t need = 16 * NS + 12 * NS + 16 * NS + NS + 4 * NS + 4 * NS + 4 * (NC + n) + sizeof(fpk_desc) * NC + 16 * NS + 4 * (NS + n) + 16 * NA +
                        4 * (NC + n) + 16 * NS + 256 * 24 + (b->tot_lig > 0 ? 4 * (NS + n + 1) + 4 * (NA + n + 1) + 32 * (NC + 1) : 0);
    H->blk = pin_get(need, &H->cap);
    if (!H->blk) { delete H; return fail(FPK_ERR_CUDA, "cudaHostAlloc(%zu bytes) for the results failed", need); }
    float *xyzr = H->take<float>(4 * NS), *bary = H->take<float>(3 * NS), *atom_fx = H->take<float>(4 * NA);
    int *neigh = H->take<int>(4 * NS), *cluster = H->take<int>(NS), *cl_sph = H->take<int>(NS), *cl_off = H->take<int>(NC + n);
    int *sparse = H->take<int>(4 * NS), *rank = H->take<int>(NS + n), *cl_atom_off = H->take<int>(NC + n), *cl_atoms = H->take<int>(4 * NS);
    uint8_t *type = H->take<uint8_t>(NS);
    fpk_desc *desc = H->take<fpk_desc>(NC);
    int rc;
    if ((rc = d2h(b, xyzr, b->c_xyzr, 4 * NS)) || (rc = d2h(b, bary, b->c_bary, 3 * NS)) || (rc = d2h(b, neigh, b->c_neigh, 4 * NS)) ||
        (rc = d2h(b, type, b->c_type, NS)) || (rc = d2h(b, cluster, b->c_cluster, NS)) || (rc = d2h(b, cl_sph, b->c_cl_sph, NS)) ||
        (rc = d2h(b, cl_off, b->c_cl_off, NC + n)) || (rc = d2h(b, desc, b->Pk.desc, NC)) || (rc = d2h(b, sparse, b->Pk.cl_atoms, 4 * NS)) ||
        (rc = d2h(b, rank, b->Pk.rank_to_cluster, NS + n)) || (rc = d2h(b, atom_fx, b->Pk.atom_fx, 4 * NA))) { delete H; return rc; }
    const bool lig = b->tot_lig > 0 && b->prm.skip_clustering == 0;
    int *ilist = nullptr;
    float *crit = nullptr, *pk4 = nullptr;
    if (lig) {
        ilist = H->take<int>(NA + n + 1); crit = H->take<float>(4 * NC + 4); pk4 = H->take<float>(4 * NC + 4);
        b->h_ni.resize(n); b->h_ligvol.resize(n);
        if ((rc = d2h(b, ilist, b->Lg.ilist, NA + n)) || (rc = d2h(b, crit, b->Lg.crit, 4 * NC)) ||
            (rc = d2h(b, b->h_ni.data(), b->Lg.ni, (size_t)n)) || (rc = d2h(b, b->h_ligvol.data(), b->Lg.lig_vol, (size_t)n))) { delete H; return rc; }
    }
    CK(wait_stream());
    if (b->B.tets_out) {
        H->tets.resize(4 * (size_t)b->tot_tets + 4);
        size_t o = 0;
        for (int k = 0; k < n; k++) {
            int nt = b->h_counts[(size_t)k * CNT_STRIDE + CNT_TETS];
            if (nt > b->S[k].tet_cap) nt = b->S[k].tet_cap;
            if (nt) CK(cudaMemcpy(&H->tets[o], b->B.tets_out + 4 * (size_t)b->S[k].tet_off, sizeof(int) * 4 * nt, cudaMemcpyDeviceToHost));
            // canonical order: lexicographic (each tet is already ascending)
            struct T4 { int v[4]; };
            T4 *t = reinterpret_cast<T4 *>(&H->tets[o]);
            std::sort(t, t + nt, [](const T4 &x, const T4 &y) { return std::lexicographical_compare(x.v, x.v + 4, y.v, y.v + 4); });
            o += 4 * (size_t)nt;
            b->d2h += (long long)sizeof(int) * 4 * nt;
        }
    }
    // compact the per-cluster atom lists (device layout: 4 slots per sphere of the cluster, first n_atoms used)
    size_t tetpos = 0, apos = 0;
    for (int k = 0; k < n; k++) {
        const int *c = &b->h_counts[(size_t)k * CNT_STRIDE];
        fpk_result &r = out[k];
        memset(&r, 0, sizeof r);
        r.status = b->bad[k] ? b->bad[k] : c[CNT_STATUS];
        r.n_sites = b->S[k].nsites;
        r.n_tets = c[CNT_TETS];
        const bool have = c[CNT_STATUS] == FPK_OK || c[CNT_STATUS] == FPK_ERR_NO_SPHERES || c[CNT_STATUS] == FPK_ERR_NO_POCKETS;
        const int xp = (have && c[CNT_XP] == 1) ? 1 : 0;
        const int s0 = b->h_sbase[k], ns = have ? c[CNT_SPHERES] : 0, c0 = b->h_clbase[k], nc = b->h_clbase[k + 1] - c0;
        r.n_spheres = ns; r.n_clusters = nc - xp; r.n_pockets = c[CNT_POCKETS];
        if (b->B.tets_out) { r.tets = &H->tets[tetpos]; tetpos += 4 * (size_t)std::min(r.n_tets, b->S[k].tet_cap); }
        r.sph_xyzr = xyzr + 4 * (size_t)s0; r.sph_bary = bary + 3 * (size_t)s0;
        r.sph_neigh = neigh + 4 * (size_t)s0; r.sph_type = type + s0; r.sph_cluster = cluster + s0;
        r.cl_off = cl_off + c0 + k; r.cl_sph = cl_sph + s0;
        r.desc = desc + c0;
        r.rank_to_cluster = rank + s0 + k;
        r.atom_fx = atom_fx + 4 * (size_t)b->S[k].atom_off;
        int *ao = cl_atom_off + c0 + k;
        r.cl_atoms = cl_atoms + apos;
        ao[0] = 0;
        for (int q = 0; q < nc; q++) {
            const int *src = sparse + 4 * ((size_t)s0 + r.cl_off[q]);
            const int na = r.desc[q].n_atoms;
            memcpy(cl_atoms + apos, src, sizeof(int) * (size_t)na);
            apos += (size_t)na;
            ao[q + 1] = ao[q] + na;
        }
        r.cl_atom_off = ao;
        if (lig && in_has_lig(b, k)) {
            if (xp) { r.n_xspheres = c[CNT_NX]; r.xspheres = r.cl_sph + r.cl_off[nc - 1]; r.xdesc = r.desc + (nc - 1); }
            r.n_iface = b->h_ni[k]; r.iface_atoms = ilist + b->S[k].atom_off + k;
            r.lig_volume = b->h_ligvol[k];
            float *o = pk4 + 4 * (size_t)c0;
            const int np = r.n_pockets;
            r.pk_ovlp = o; r.pk_dst = o + np; r.pk_c4 = o + 2 * np; r.pk_c5 = o + 3 * np;
            for (int q = 0; q < np; q++) for (int z = 0; z < 4; z++) o[z * np + q] = crit[4 * ((size_t)c0 + q) + z];
        }
        r._owner = H;
        H->refs++;
    }
    if (H->refs == 0) delete H;
    return FPK_OK;
}

extern "C" void fpk_result_free(fpk_result *r)
{
    if (!r || !r->_owner) return;
    HostResults *H = static_cast<HostResults *>(r->_owner);
    memset(r, 0, sizeof *r);
    if (H->refs.fetch_sub(1) == 1) delete H;
}

// parameter values the kernels have no path for (there is no CPU fallback to hand them to)
static const char *params_defect(const fpk_params *p)
{
    static thread_local char msg[160];
    if (p->distance_measure != 0 && p->distance_measure != 'e' && p->distance_measure != 'b') {
        snprintf(msg, sizeof msg, "distance measure '%c' is not supported (-e e or -e b; there is no CPU fallback)", p->distance_measure);
        return msg;
    }
    const int cm = p->clustering_method;
    if (cm != 0 && cm != 's' && cm != 'm' && cm != 'a' && cm != 'c') {
        snprintf(msg, sizeof msg, "clustering method '%c' is not one of clusterlib's s, m, a, c (clusterlib.c:3886)", cm);
        return msg;
    }
    if (p->lig_interface_method < 0 || p->lig_interface_method > 2) {
        snprintf(msg, sizeof msg, "ligand interface method %d (dparams.h:41-44 knows 1 and 2)", p->lig_interface_method);
        return msg;
    }
    return nullptr;
}

extern "C" fpk_batch *fpk_batch_create(const fpk_structure *in, int n, const fpk_params *p)
{
    if (ensure_init()) return nullptr;
    if (!in || n <= 0 || !p) { fail(FPK_ERR_BAD_ARG, "fpk_batch_create: bad arguments"); return nullptr; }
    if (const char *why = params_defect(p)) { fail(FPK_ERR_BAD_ARG, "%s", why); return nullptr; }
    fpk_batch *b = new fpk_batch();
    if (batch_pack(b, in, n, p, 0.0) || batch_upload(b, in)) { delete b; return nullptr; }
    // the upload was queued on THIS thread's stream: it must have landed before the handle can be run from any other thread (stream)
    if (wait_stream() != cudaSuccess) { fail(FPK_ERR_CUDA, "fpk_batch_create: upload failed"); delete b; return nullptr; }
    return b;
}
extern "C" int fpk_batch_run(fpk_batch *b) { return b ? batch_run(b) : fail(FPK_ERR_BAD_ARG, "null batch"); }
extern "C" int fpk_batch_fetch(fpk_batch *b, fpk_result *out) { return (b && out) ? batch_fetch(b, out) : fail(FPK_ERR_BAD_ARG, "null argument"); }
extern "C" float fpk_batch_last_kernel_ms(const fpk_batch *b, int which) { return (b && which >= 0 && which < 8) ? b->ms[which] : -1.0f; }
extern "C" long long fpk_batch_counter(const fpk_batch *b, int which)
{
    if (!b) return which == 6 ? g_tot_h2d.load() : (which == 7 ? g_tot_d2h.load() : -1);
    switch (which) {
    case 0: return b->tot_sites;
    case 1: return b->tot_tets;
    case 2: return b->tot_spheres;
    case 3: return b->tot_clusters;
    case 4: { long long p = 0; for (int k = 0; k < b->n; k++) p += b->h_counts.empty() ? 0 : b->h_counts[(size_t)k * CNT_STRIDE + CNT_POCKETS]; return p; }
    case 5: return b->launches;
    case 6: return b->h2d;
    case 7: return b->d2h;
    case 8: { long long p = 0; for (int k = 0; k < b->n; k++) p += b->h_counts.empty() ? 0 : b->h_counts[(size_t)k * CNT_STRIDE + CNT_SUBROUNDS]; return p; }
    case 9: { long long p = 0; for (int k = 0; k < b->n; k++) p += b->h_counts.empty() ? 0 : b->h_counts[(size_t)k * CNT_STRIDE + CNT_RETRIES]; return p; }
    case 11: case 12: case 13: case 14: { long long p = 0; for (int k = 0; k < b->n; k++) p += b->h_counts.empty() ? 0 : b->h_counts[(size_t)k * CNT_STRIDE + CNT_WALK + (which - 11)]; return p; }
    case 10: { long long p = 0; for (int k = 0; k < b->n; k++) p += b->h_counts.empty() ? 0 : b->h_counts[(size_t)k * CNT_STRIDE + CNT_NBIG]; return p; }
    }
    return -1;
}
extern "C" void fpk_batch_destroy(fpk_batch *b) { delete b; }

// ---- fpk_detect_batch: the batch is cut into chunks that a few host threads push through pack -> H2D -> kernels -> D2H ->
// unpack, each on its own CUDA stream (the library is compiled with --default-stream per-thread), so that one chunk's copies and
// host-side packing overlap another chunk's kernels. Worker contexts (device arenas, pinned mirrors) are pooled across calls.
static std::mutex g_pool_mu;
static std::vector<fpk_batch *> g_pool;        // idle contexts of all devices (a context's arenas stay on the device that made them)
static fpk_batch *pool_get(int dev)
{
    {
        std::lock_guard<std::mutex> g(g_pool_mu);
        for (size_t i = 0; i < g_pool.size(); i++)
            if (g_pool[i]->dev == dev) { fpk_batch *b = g_pool[i]; g_pool.erase(g_pool.begin() + i); return b; }
    }
    fpk_batch *b = new fpk_batch();
    b->dev = dev;
    return b;
}
static void pool_put(fpk_batch *b)
{
    {
        std::lock_guard<std::mutex> g(g_pool_mu);
        // two concurrent callers x four threads per device (libfpocket_hostio) keep their contexts: a cudaFree here would serialise the device
        if (g_pool.size() < 8 * std::max<size_t>(1, g_devs.size())) { g_pool.push_back(b); return; }
    }
    delete b;
}
static void pool_clear()
{
    std::lock_guard<std::mutex> g(g_pool_mu);
    for (fpk_batch *b : g_pool) delete b;
    g_pool.clear();
}

static std::atomic<long long> g_t_pack(0), g_t_up(0), g_t_run(0), g_t_fetch(0);     // host wall time per stage, ns (FPK_TRACE=1 prints them)
static int detect_chunk(fpk_batch *b, const fpk_structure *in, int n, const fpk_params *p, fpk_result *out)
{
    typedef std::chrono::steady_clock clk;
    auto t0 = clk::now();
    int rc = batch_pack(b, in, n, p, 0.0);
    auto t1 = clk::now();
    if (!rc) rc = batch_upload(b, in);
    auto t2 = clk::now();
    if (!rc) rc = batch_run(b);
    auto t3 = clk::now();
    if (!rc) rc = batch_fetch(b, out);
    auto t4 = clk::now();
    g_t_pack += (t1 - t0).count(); g_t_up += (t2 - t1).count(); g_t_run += (t3 - t2).count(); g_t_fetch += (t4 - t3).count();
    g_tot_h2d += b->h2d; g_tot_d2h += b->d2h;
    return rc;
}

extern "C" int fpk_detect_batch(const fpk_structure *in, int n, const fpk_params *p, fpk_result *out)
{
    if (!out) return fail(FPK_ERR_BAD_ARG, "null result array");
    if (!in || n <= 0 || !p) return fail(FPK_ERR_BAD_ARG, "fpk_detect_batch: bad arguments");
    if (const char *why = params_defect(p)) return fail(FPK_ERR_BAD_ARG, "%s", why);
    int rc = ensure_init();
    if (rc) return rc;
    memset(out, 0, sizeof(fpk_result) * (size_t)n);      // whatever happens, every out[k] can be handed to fpk_result_free
    const int ndev = (int)g_devs.size();
    // Chunking: K1 keeps 3 x 148 = 444 blocks resident, so a launch wants well over that many structures, and the last round of chunks
    // should keep every host thread busy: ~1,200 structures per chunk, the number of chunks a multiple of the thread count when the batch
    // allows (measured on 6,144 structures, 4 threads: 512 -> 7,668, 1,024 -> 7,867, 1,536 -> 8,268 structures/s end to end).
    // With several devices (fpk_init's list) every device gets its own group of host threads; chunks are handed out from one counter,
    // so a faster or less loaded device simply takes more of them.
    // host threads per device: 4, or 8 when this process has at least 16 cores per device to itself (measured on the 16-core box, one GPU: e2e 11,100
    // structures/s with 4, 11,181 with 6, 11,322 with 8, profiles/r02f_e2e_threads.log; with 8 ranks on 32 cores more threads only fight for cores)
    int chunk = 0, per_dev = 4;
    {
        const int cores = (int)std::thread::hardware_concurrency();
        const char *lw = getenv("LOCAL_WORLD_SIZE");
        const int ranks = lw ? std::max(1, atoi(lw)) : 1;
        if (cores / ranks / std::max(1, ndev) >= 16) per_dev = 8;
    }
    if (const char *e = getenv("FPK_CHUNK")) chunk = std::max(1, atoi(e));
    if (const char *e = getenv("FPK_HOST_THREADS")) per_dev = std::max(1, std::min(8, atoi(e)));
    int nthreads = per_dev * ndev;
    if (chunk <= 0) {
        int nc = (n + 1199) / 1200;
        if (n >= 600 * nthreads) nc = std::max(nthreads, (nc / nthreads) * nthreads);
        else nc = std::max(1, std::min(nthreads, (n + 599) / 600));
        chunk = (n + nc - 1) / nc;
    }
    const int nchunks = (n + chunk - 1) / chunk;
    nthreads = std::min(nthreads, nchunks);
    if (nthreads <= 1) {
        fpk_batch *b = pool_get(g_devs[0]);
        for (int c = 0; c < nchunks && !rc; c++) rc = detect_chunk(b, in + (size_t)c * chunk, std::min(chunk, n - c * chunk), p, out + (size_t)c * chunk);
        pool_put(b);
    } else {
        std::atomic<int> next(0), err(0);
        std::vector<std::string> msgs(nthreads);
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++)
            th.emplace_back([&, t]() {
                const int dev = g_devs[t % ndev];         // thread t works for device t mod ndev: the first ndev threads cover every device
                cudaSetDevice(dev);
                fpk_batch *b = pool_get(dev);
                for (;;) {
                    const int c = next.fetch_add(1);
                    if (c >= nchunks || err.load()) break;
                    int r = detect_chunk(b, in + (size_t)c * chunk, std::min(chunk, n - c * chunk), p, out + (size_t)c * chunk);
                    if (r) { msgs[t] = g_err; err.store(r); break; }
                }
                pool_put(b);
            });
        for (auto &x : th) x.join();
        rc = err.load();
        if (rc) for (auto &m : msgs) if (!m.empty()) set_err(m);
    }
    if (getenv("FPK_TRACE"))
        fprintf(stderr, "fpk_detect_batch: %d structures, %d chunks, %d threads on %d device(s); cumulative host ms: pack %.1f upload %.1f run %.1f fetch %.1f\n", n, nchunks,
                nthreads, ndev, g_t_pack.load() / 1e6, g_t_up.load() / 1e6, g_t_run.load() / 1e6, g_t_fetch.load() / 1e6);
    if (rc) {
        // a CUDA or allocation failure (never a defect of one structure: those are reported per structure): nothing half-made is handed back
        const std::string keep = g_err;
        for (int k = 0; k < n; k++) { fpk_result_free(&out[k]); memset(&out[k], 0, sizeof out[k]); out[k].status = rc; }
        set_err(keep);
        return rc;
    }
    // a structure whose sphere buffer overflowed is re-run alone with the worst-case capacity (7 tets per atom)
    for (int k = 0; k < n; k++) {
        if (out[k].status != FPK_ERR_CAPACITY) continue;
        fpk_batch *b2 = new fpk_batch();
        fpk_result r2;
        if (!batch_pack(b2, in + k, 1, p, 7.5) && !batch_upload(b2, in + k) && !batch_run(b2) && !batch_fetch(b2, &r2)) {
            fpk_result_free(&out[k]);
            out[k] = r2;
        }
        delete b2;
    }
    return rc;
}

extern "C" int fpk_detect(const fpk_structure *in, const fpk_params *p, fpk_result *out)
{
    int rc = fpk_detect_batch(in, 1, p, out);
    return rc ? rc : out->status;
}

// =================================================== mdpocket =======================================================
// One MdDev per device of fpk_init's list: its own batch skeleton, frame staging and pair of grids. fpk_md_push_frames cuts a block of
// frames into one contiguous share per device (frames are independent once frame 1 has fixed the grid, mdpbase.c:444-502),
// a host thread per device pushes its share; the grids are summed into device 0 when they are read (md_reduce: peer copies + an add
// kernel, integer-valued floats or 32.32 fixed point with -S: exact and order-free either way). With one process per GPU (torchrun) the
// caller sums fpk_md_device_grids across ranks with NCCL instead.
struct MdDev {
    int dev = 0;
    fpk_batch *b = nullptr;            // re-used for every chunk of frames (geometry of the buffers is fixed)
    fpk_batch *bc = nullptr;           // characterize mode: the same with ASA and volumes on (mdpocket.c:360-660 keeps the default)
    Arena A3;                          // the zone batch of the current chunk
    float *d_zone = nullptr; int nzone_cap = 0;
    float *d_dens = nullptr, *d_freq = nullptr; unsigned *d_mask = nullptr; unsigned long long *d_nscat = nullptr;
    long long *d_dens_fx = nullptr, *d_freq_fx = nullptr;      // -S: 32.32 fixed-point accumulators (order-free sums), see MdGrid
    float *d_geom = nullptr;
    void *d_peer = nullptr; size_t peer_cap = 0;               // device 0 only: landing buffer for another device's grid
    int nwords = 0, nblocks = 0;
    long long frames = 0, launches = 0;
    bool used = false;                 // frames were scattered into these grids since the last reduction
    float *pin[2] = {nullptr, nullptr}, *d_xyz[2] = {nullptr, nullptr}; size_t stage_cap = 0;     // double-buffered frame staging
    void free_grids()
    {
        cudaFree(d_dens); cudaFree(d_freq); cudaFree(d_mask); cudaFree(d_nscat); cudaFree(d_dens_fx); cudaFree(d_freq_fx);
        d_dens = d_freq = nullptr; d_mask = nullptr; d_nscat = nullptr; d_dens_fx = d_freq_fx = nullptr;
    }
    ~MdDev()
    {
        cudaSetDevice(dev);
        for (int q = 0; q < 2; q++) { if (pin[q]) cudaFreeHost(pin[q]); if (d_xyz[q]) cudaFree(d_xyz[q]); }
        delete b; delete bc;
        A3.release();
        free_grids();
        cudaFree(d_geom); cudaFree(d_peer); cudaFree(d_zone);
    }
};
struct fpk_md {
    fpk_structure topo;
    std::vector<float> tx, ty, tz, trad, ten, tbf; std::vector<int> tid_, tres; std::vector<int8_t> taa; std::vector<uint8_t> tfl;
    fpk_params prm;
    int use_drug = 0;
    bool have_grid = false;
    bool float_stale = true;           // -S: the float grids have to be refreshed from the fixed-point sums before they are read
    float origin[3]; int dims[3];
    std::vector<MdDev *> D;            // D[0] holds the summed grids
    ~fpk_md() { for (MdDev *d : D) delete d; if (!g_devs.empty()) cudaSetDevice(g_devs[0]); }
};

static int md_prepare(fpk_md *m, MdDev *q, int nframes, bool characterize = false)
{
    // (re)build the batch skeleton for `nframes` frames: properties uploaded once, coordinates per push
    fpk_batch *&slot = characterize ? q->bc : q->b;
    // a skeleton built for more frames serves fewer (the last chunk of a push, a shorter push): all frames have the same shape, so the first
    // nframes entries of every array are exactly the skeleton of nframes frames; rebuilding it (pack, arenas, uploads) cost more than a chunk's kernels
    if (slot && slot->n_cap >= nframes) { slot->n = nframes; slot->B.nstruct = nframes; return FPK_OK; }
    delete slot;
    slot = new fpk_batch();
    fpk_batch *b = slot;
    const fpk_structure &t = m->topo;
    std::vector<fpk_structure> fr(nframes, t);
    fpk_params prm = m->prm;
    if (characterize) prm.do_asa_and_volume = 1;
    int rc = batch_pack(b, fr.data(), nframes, &prm, 0.0);
    if (rc) return rc;
    if (b->bad[0]) return FPK_ERR_BAD_ARG;
    // all frames share one copy of the properties and of the site map
    for (int k = 0; k < nframes; k++) { b->S[k].prop_off = 0; b->S[k].site_off = 0; b->S[k].w_off = 0; b->S[k].natoms_w = 0; b->S[k].same_pdb = 1; }
    b->tot_w = 0; b->tot_sites = b->S[0].nsites;
    rc = batch_upload(b, nullptr);
    if (rc) return rc;
    BatchDev &B = b->B;
    const size_t na = (size_t)t.natoms;
    CK(cudaMemcpy(const_cast<float *>(B.arad), t.radius, sizeof(float) * na, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<float *>(B.aen), t.electroneg, sizeof(float) * na, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<float *>(B.abf), t.bfactor, sizeof(float) * na, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<int *>(B.aid), t.atom_id, sizeof(int) * na, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<int *>(B.ares), t.res_id, sizeof(int) * na, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<int8_t *>(B.aaa), t.aa_idx, na, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<uint8_t *>(B.afl), t.flags, na, cudaMemcpyHostToDevice));
    std::vector<int> s2a;
    for (int i = 0; i < t.natoms; i++) if (!(t.flags[i] & FPK_ATOM_H)) s2a.push_back(i);
    CK(cudaMemcpy(const_cast<int *>(B.site2atom), s2a.data(), sizeof(int) * s2a.size(), cudaMemcpyHostToDevice));
    b->h2d += (long long)(na * 22 + s2a.size() * 4);
    b->n_cap = nframes;
    return FPK_OK;
}

// xyz is [frame][atom][3] (the molfile timestep layout, mdpocket.c:238-243): split into the SoA the kernels read, on the device
__global__ void k_md_deinterleave(const float *xyz, size_t tot, float *ax, float *ay, float *az)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
        ax[i] = xyz[3 * i]; ay[i] = xyz[3 * i + 1]; az[i] = xyz[3 * i + 2];
    }
}
__global__ void k_md_add_f32(float *dst, const float *src, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] += src[i];
}
__global__ void k_md_add_i64(long long *dst, const long long *src, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] += src[i];
}

// host -> pinned bounce buffer -> device staging buffer `slot`, on the calling thread's stream (returns after the copy has landed)
static int md_stage_chunk(fpk_md *m, MdDev *q, int slot, const float *xyz, int nframes)
{
    const size_t bytes = sizeof(float) * 3 * (size_t)m->topo.natoms * nframes;
    if (bytes > q->stage_cap) return fail(FPK_ERR_BAD_ARG, "md staging buffer too small");
    memcpy(q->pin[slot], xyz, bytes);
    CK(cudaMemcpyAsync(q->d_xyz[slot], q->pin[slot], bytes, cudaMemcpyHostToDevice, cudaStreamPerThread));
    CK(wait_stream());
    return FPK_OK;
}

static int md_ensure_staging(fpk_md *m, MdDev *q, int chunk_frames)
{
    const size_t need = sizeof(float) * 3 * (size_t)m->topo.natoms * chunk_frames;
    if (need <= q->stage_cap) return FPK_OK;
    for (int z = 0; z < 2; z++) {
        if (q->pin[z]) cudaFreeHost(q->pin[z]);
        if (q->d_xyz[z]) cudaFree(q->d_xyz[z]);
        q->pin[z] = nullptr; q->d_xyz[z] = nullptr;
        CK(cudaHostAlloc((void **)&q->pin[z], need, cudaHostAllocPortable));
        CK(cudaMalloc((void **)&q->d_xyz[z], need));
    }
    q->stage_cap = need;
    return FPK_OK;
}

extern "C" fpk_md *fpk_md_begin(const fpk_structure *topo, const fpk_params *p, int use_drug_score)
{
    if (ensure_init()) return nullptr;
    if (!topo || !p || topo->natoms <= 0) { fail(FPK_ERR_BAD_ARG, "fpk_md_begin: bad arguments"); return nullptr; }
    if (const char *why = params_defect(p)) { fail(FPK_ERR_BAD_ARG, "fpk_md_begin: %s", why); return nullptr; }
    if (!topo->radius || !topo->electroneg || !topo->bfactor || !topo->atom_id || !topo->res_id || !topo->aa_idx || !topo->flags) {
        fail(FPK_ERR_BAD_ARG, "fpk_md_begin: NULL atom property array");
        return nullptr;
    }
    fpk_md *m = new fpk_md();
    const int na = topo->natoms;
    // own copies of the topology (the caller's arrays need not outlive this call)
    m->tx.assign(na, 0.f); m->ty.assign(na, 0.f); m->tz.assign(na, 0.f);
    m->trad.assign(topo->radius, topo->radius + na); m->ten.assign(topo->electroneg, topo->electroneg + na);
    m->tbf.assign(topo->bfactor, topo->bfactor + na); m->tid_.assign(topo->atom_id, topo->atom_id + na);
    m->tres.assign(topo->res_id, topo->res_id + na); m->taa.assign(topo->aa_idx, topo->aa_idx + na);
    m->tfl.assign(topo->flags, topo->flags + na);
    m->topo = *topo;
    m->topo.x = m->tx.data(); m->topo.y = m->ty.data(); m->topo.z = m->tz.data(); m->topo.radius = m->trad.data();
    m->topo.electroneg = m->ten.data(); m->topo.bfactor = m->tbf.data(); m->topo.atom_id = m->tid_.data();
    m->topo.res_id = m->tres.data(); m->topo.aa_idx = m->taa.data(); m->topo.flags = m->tfl.data();
    m->topo.natoms_w = 0; m->topo.same_pdb = 1; m->topo.n_xlig = 0; m->topo.xlig = nullptr; m->topo.n_lig = 0; m->topo.lig_atoms = nullptr; m->topo.lig_mol = nullptr; m->topo.atom_name_key = nullptr;     // mdpocket.c:844 search_pocket(pdb, params, pdb)
    if (const char *why = structure_defect(m->topo)) { fail(FPK_ERR_BAD_ARG, "fpk_md_begin: %s", why); delete m; return nullptr; }
    m->prm = *p;
    m->prm.do_asa_and_volume = 0;          // mdpocket.c:89
    m->prm.want_tets = 0;
    m->use_drug = use_drug_score;
    for (int dev : g_devs) { MdDev *q = new MdDev(); q->dev = dev; m->D.push_back(q); }
    return m;
}

extern "C" int fpk_md_set_grid(fpk_md *m, const float origin[3], const int dims[3])
{
    if (!m || !origin || !dims || dims[0] <= 0 || dims[1] <= 0 || dims[2] <= 0) return fail(FPK_ERR_BAD_ARG, "fpk_md_set_grid: bad arguments");
    for (int q = 0; q < 3; q++) { m->origin[q] = origin[q]; m->dims[q] = dims[q]; }
    const size_t ncell = (size_t)dims[0] * dims[1] * dims[2];
    for (MdDev *q : m->D) {
        CK(cudaSetDevice(q->dev));
        q->free_grids();
        q->nwords = (int)((ncell + 31) / 32);
        q->nblocks = g_sms * 2;
        CK(cudaMalloc(&q->d_dens, sizeof(float) * ncell)); CK(cudaMalloc(&q->d_freq, sizeof(float) * ncell));
        CK(cudaMalloc(&q->d_mask, sizeof(unsigned) * (size_t)q->nwords * q->nblocks)); CK(cudaMalloc(&q->d_nscat, 8));
        CK(cudaMemset(q->d_dens, 0, sizeof(float) * ncell)); CK(cudaMemset(q->d_freq, 0, sizeof(float) * ncell)); CK(cudaMemset(q->d_nscat, 0, 8));
        if (m->use_drug) {
            CK(cudaMalloc(&q->d_dens_fx, sizeof(long long) * ncell)); CK(cudaMalloc(&q->d_freq_fx, sizeof(long long) * ncell));
            CK(cudaMemset(q->d_dens_fx, 0, sizeof(long long) * ncell)); CK(cudaMemset(q->d_freq_fx, 0, sizeof(long long) * ncell));
        }
        q->used = false;
    }
    CK(cudaSetDevice(m->D[0]->dev));
    m->have_grid = true; m->float_stale = true;
    return FPK_OK;
}

extern "C" int fpk_md_grid_from_frame(fpk_md *m, const float *xyz, float origin[3], int dims[3])
{
    if (!m || !xyz) return fail(FPK_ERR_BAD_ARG, "fpk_md_grid_from_frame: bad arguments");
    MdDev *q = m->D[0];
    CK(cudaSetDevice(q->dev));
    int rc = md_prepare(m, q, 1);
    if (!rc) rc = md_ensure_staging(m, q, 1);
    if (!rc) rc = md_stage_chunk(m, q, 0, xyz, 1);
    if (!rc) {
        const size_t tot = (size_t)m->topo.natoms;
        k_md_deinterleave<<<64, 256>>>(q->d_xyz[0], tot, const_cast<float *>(q->b->B.ax), const_cast<float *>(q->b->B.ay), const_cast<float *>(q->b->B.az));
    }
    if (!rc) rc = batch_run(q->b);
    if (rc) return rc;
    q->launches += q->b->launches;
    if (!q->d_geom) CK(cudaMalloc(&q->d_geom, sizeof(float) * 8));
    k_md_geometry<<<1, 32>>>(q->b->B, 0, q->d_geom);
    q->launches++;
    float g[6];
    CK(cudaMemcpy(g, q->d_geom, sizeof g, cudaMemcpyDeviceToHost));
    for (int z = 0; z < 3; z++) { origin[z] = g[z]; dims[z] = (int)g[3 + z]; }
    return FPK_OK;
}

// detect + scatter `nframes` frames on ONE device (the calling thread has made it current)
static int md_push_on(fpk_md *m, MdDev *q, const float *xyz, int nframes)
{
    const size_t na = (size_t)m->topo.natoms;
    // frames per launch of the pipeline; the next chunk is staged (host copy into pinned memory + H2D) meanwhile. Measured on the 20,000-frame
    // trajectory (profiles/r02_md_chunk_sweep2.log): 1,184 frames per chunk 10,333 frames/s, 2,368 9,068 (a 2,048-frame push: 10,264 / 10,324)
    int chunk_max = std::max(256, g_sms * 8);
    if (const char *e = getenv("FPK_MD_CHUNK")) chunk_max = std::max(64, atoi(e));      // tuning knob (profiles/)
    int rc = md_ensure_staging(m, q, std::min(chunk_max, nframes));
    typedef std::chrono::steady_clock clk;
    double t_run = 0, t_scatter = 0, t_prep = 0;
    if (rc) return rc;
    // chunk k+1 is staged (host copy into pinned memory + H2D) by a helper thread on its own stream while chunk k is detected
    const int nchunks = (nframes + chunk_max - 1) / chunk_max;
    rc = md_stage_chunk(m, q, 0, xyz, std::min(chunk_max, nframes));
    if (rc) return rc;
    for (int c = 0; c < nchunks; c++) {
        const int f0 = c * chunk_max, nf = std::min(chunk_max, nframes - f0);
        std::thread pre;
        int pre_rc = FPK_OK;
        std::string pre_err;
        if (c + 1 < nchunks) {
            const int f1 = f0 + chunk_max, nf1 = std::min(chunk_max, nframes - f1);
            pre = std::thread([&, f1, nf1, c]() {
                cudaSetDevice(q->dev);
                pre_rc = md_stage_chunk(m, q, (c + 1) & 1, xyz + 3 * na * (size_t)f1, nf1);
                if (pre_rc) pre_err = g_err;
            });
        }
        auto t0 = clk::now();
        rc = md_prepare(m, q, nf);
        auto t1 = clk::now();
        t_prep += std::chrono::duration<double, std::milli>(t1 - t0).count();
        if (!rc) {
            const size_t tot = na * (size_t)nf;
            k_md_deinterleave<<<g_sms * 4, 256>>>(q->d_xyz[c & 1], tot, const_cast<float *>(q->b->B.ax), const_cast<float *>(q->b->B.ay), const_cast<float *>(q->b->B.az));
            q->b->h2d += (long long)(sizeof(float) * 3 * tot);
            rc = batch_run(q->b);
            t_run += std::chrono::duration<double, std::milli>(clk::now() - t1).count();
        }
        if (!rc) {
            auto t2 = clk::now();
            MdGrid G;
            for (int z = 0; z < 3; z++) G.origin[z] = m->origin[z];
            G.nx = m->dims[0]; G.ny = m->dims[1]; G.nz = m->dims[2];
            G.dens = q->d_dens; G.freq = q->d_freq; G.mask = q->d_mask; G.nwords = q->nwords; G.use_drug_score = m->use_drug;
            G.dens_fx = q->d_dens_fx; G.freq_fx = q->d_freq_fx;
            k_md_scatter<<<std::min(q->nblocks, nf), 256>>>(q->b->B, q->b->Pk, G, q->d_nscat);
            cudaError_t e = cudaStreamSynchronize(cudaStreamPerThread);
            if (e != cudaSuccess) rc = fail(FPK_ERR_CUDA, "k_md_scatter: %s", cudaGetErrorString(e));
            q->launches += q->b->launches + 2;
            q->frames += nf;
            q->used = true;
            t_scatter += std::chrono::duration<double, std::milli>(clk::now() - t2).count();
        }
        if (pre.joinable()) pre.join();
        if (rc) return rc;
        if (pre_rc) { set_err(pre_err); return pre_rc; }
    }
    CK(cudaDeviceSynchronize());
    if (getenv("FPK_TRACE")) fprintf(stderr, "fpk_md_push_frames: device %d: %d frames, %d chunks; host ms: prepare %.1f detect %.1f scatter %.1f\n", q->dev, nframes, nchunks, t_prep, t_run, t_scatter);
    return FPK_OK;
}

extern "C" int fpk_md_push_frames(fpk_md *m, const float *xyz, int nframes)
{
    if (!m || !xyz || nframes <= 0) return fail(FPK_ERR_BAD_ARG, "fpk_md_push_frames: bad arguments");
    if (!m->have_grid) return fail(FPK_ERR_BAD_ARG, "fpk_md_push_frames: set the grid first (fpk_md_set_grid)");
    m->float_stale = true;
    const int ndev = (int)m->D.size();
    const size_t na = (size_t)m->topo.natoms;
    // contiguous shares of at least 64 frames per device; a short push (the reference's frame-by-frame loop) stays on device 0
    const int nuse = std::max(1, std::min(ndev, nframes / 64));
    if (nuse == 1) {
        CK(cudaSetDevice(m->D[0]->dev));
        return md_push_on(m, m->D[0], xyz, nframes);
    }
    std::vector<int> rcs(nuse, FPK_OK);
    std::vector<std::string> msgs(nuse);
    std::vector<std::thread> th;
    for (int d = 0; d < nuse; d++) {
        const int f0 = (int)((long long)nframes * d / nuse), f1 = (int)((long long)nframes * (d + 1) / nuse);
        th.emplace_back([&, d, f0, f1]() {
            cudaSetDevice(m->D[d]->dev);
            rcs[d] = md_push_on(m, m->D[d], xyz + 3 * na * (size_t)f0, f1 - f0);
            if (rcs[d]) msgs[d] = g_err;
        });
    }
    for (auto &t : th) t.join();
    cudaSetDevice(m->D[0]->dev);
    for (int d = 0; d < nuse; d++) if (rcs[d]) { set_err(msgs[d]); return rcs[d]; }
    return FPK_OK;
}

// sum the grids of devices 1.. into device 0 and clear them (so that a later reduction does not count them twice)
static int md_reduce(fpk_md *m)
{
    MdDev *z = m->D[0];
    const size_t ncell = (size_t)m->dims[0] * m->dims[1] * m->dims[2];
    const size_t esz = m->use_drug ? sizeof(long long) : sizeof(float);
    CK(cudaSetDevice(z->dev));
    for (size_t d = 1; d < m->D.size(); d++) {
        MdDev *q = m->D[d];
        if (!q->used) continue;
        if (z->peer_cap < ncell * esz) { cudaFree(z->d_peer); z->d_peer = nullptr; z->peer_cap = 0; CK(cudaMalloc(&z->d_peer, ncell * esz)); z->peer_cap = ncell * esz; }
        for (int g = 0; g < 2; g++) {
            void *src = m->use_drug ? (void *)(g ? q->d_freq_fx : q->d_dens_fx) : (void *)(g ? q->d_freq : q->d_dens);
            CK(cudaMemcpyPeer(z->d_peer, z->dev, src, q->dev, ncell * esz));
            if (m->use_drug) k_md_add_i64<<<g_sms * 4, 256>>>(g ? z->d_freq_fx : z->d_dens_fx, (const long long *)z->d_peer, ncell);
            else k_md_add_f32<<<g_sms * 4, 256>>>(g ? z->d_freq : z->d_dens, (const float *)z->d_peer, ncell);
            CK(cudaStreamSynchronize(cudaStreamPerThread));
            z->launches++;
        }
        CK(cudaSetDevice(q->dev));
        if (m->use_drug) { CK(cudaMemset(q->d_dens_fx, 0, ncell * esz)); CK(cudaMemset(q->d_freq_fx, 0, ncell * esz)); }
        else { CK(cudaMemset(q->d_dens, 0, ncell * esz)); CK(cudaMemset(q->d_freq, 0, ncell * esz)); }
        q->used = false;
        CK(cudaSetDevice(z->dev));
        m->float_stale = true;
    }
    return FPK_OK;
}

// -S: the float grids are views of the fixed-point accumulators, refreshed when frames were pushed since they were last handed out.
// (Not on every call: a caller that summed the float grids across ranks in place - fpk_md_device_grids, NCCL all-reduce - must read
// back what it summed, not this rank's share again.)
static int md_refresh_float_grids(fpk_md *m)
{
    int rc = md_reduce(m);
    if (rc) return rc;
    MdDev *z = m->D[0];
    if (!z->d_dens_fx || !m->float_stale) { m->float_stale = false; return FPK_OK; }
    const size_t ncell = (size_t)m->dims[0] * m->dims[1] * m->dims[2];
    k_md_fixed_to_float<<<g_sms * 4, 256>>>(z->d_dens_fx, z->d_dens, ncell);
    k_md_fixed_to_float<<<g_sms * 4, 256>>>(z->d_freq_fx, z->d_freq, ncell);
    z->launches += 2;
    CK(cudaStreamSynchronize(cudaStreamPerThread));
    m->float_stale = false;
    return FPK_OK;
}

extern "C" int fpk_md_device_grids(fpk_md *m, void **dens, void **freq, size_t *ncell)
{
    if (!m || !m->have_grid) return fail(FPK_ERR_BAD_ARG, "no grid");
    { int rc = md_refresh_float_grids(m); if (rc) return rc; }
    if (dens) *dens = m->D[0]->d_dens;
    if (freq) *freq = m->D[0]->d_freq;
    if (ncell) *ncell = (size_t)m->dims[0] * m->dims[1] * m->dims[2];
    return FPK_OK;
}
extern "C" int fpk_md_fetch_grids(fpk_md *m, float *dens, float *freq)
{
    if (!m || !m->have_grid) return fail(FPK_ERR_BAD_ARG, "no grid");
    size_t ncell = (size_t)m->dims[0] * m->dims[1] * m->dims[2];
    { int rc = md_refresh_float_grids(m); if (rc) return rc; }
    if (dens) CK(cudaMemcpy(dens, m->D[0]->d_dens, sizeof(float) * ncell, cudaMemcpyDeviceToHost));
    if (freq) CK(cudaMemcpy(freq, m->D[0]->d_freq, sizeof(float) * ncell, cudaMemcpyDeviceToHost));
    return FPK_OK;
}
// ---- characterize mode (mdpocket.c:360-660): full detection of every frame, then the zone's sphere list described as one pocket --------
struct ZoneOwner { std::atomic<int> refs{0}; std::vector<char> blk; };
extern "C" void fpk_zone_free(fpk_zone *z)
{
    if (!z || !z->_owner) return;
    ZoneOwner *o = static_cast<ZoneOwner *>(z->_owner);
    memset(z, 0, sizeof *z);
    if (o->refs.fetch_sub(1) == 1) delete o;
}

static int md_characterize_chunk(fpk_md *m, MdDev *q, const float *xyz, int nf, int nzone, fpk_zone *out)
{
    int rc = md_prepare(m, q, nf, true);
    if (!rc) rc = md_ensure_staging(m, q, nf);
    if (!rc) rc = md_stage_chunk(m, q, 0, xyz, nf);
    if (rc) return rc;
    fpk_batch *b = q->bc;
    const size_t na = (size_t)m->topo.natoms, tot = na * (size_t)nf;
    k_md_deinterleave<<<g_sms * 4, 256>>>(q->d_xyz[0], tot, const_cast<float *>(b->B.ax), const_cast<float *>(b->B.ay), const_cast<float *>(b->B.az));
    rc = batch_run(b);
    if (rc) return rc;
    q->launches += b->launches + 1;
    q->frames += nf;
    for (int f = 0; f < nf; f++) {
        memset(&out[f], 0, sizeof out[f]);
        const int *c = &b->h_counts[(size_t)f * CNT_STRIDE];
        out[f].status = c[CNT_STATUS] == FPK_OK ? FPK_ERR_NO_POCKETS : c[CNT_STATUS];
        out[f].n_spheres_frame = c[CNT_SPHERES]; out[f].n_pockets_frame = c[CNT_POCKETS];
    }
    if (b->tot_clusters == 0 || nzone <= 0) return FPK_OK;
    // ---- pass 0: how many (sphere, zone point) entries does every frame have? -----------------------------------------------------
    ZoneDev Z;
    memset(&Z, 0, sizeof Z);
    Z.zone = q->d_zone; Z.nzone = nzone;
    int *d_ecount = nullptr;
    CK(cudaMalloc(&d_ecount, sizeof(int) * (size_t)nf));
    Z.ecount = d_ecount;
    k_zone_select<<<nf, 256>>>(b->B, b->Pk, Z, 0);
    std::vector<int> ecount(nf), ebase(nf + 1, 0), clbase2(nf + 1, 0);
    cudaError_t e = cudaMemcpy(ecount.data(), d_ecount, sizeof(int) * (size_t)nf, cudaMemcpyDeviceToHost);
    cudaFree(d_ecount);
    if (e != cudaSuccess) return fail(FPK_ERR_CUDA, "k_zone_select: %s", cudaGetErrorString(e));
    int maxE = 1;
    for (int f = 0; f < nf; f++) { ebase[f + 1] = ebase[f] + ecount[f]; clbase2[f + 1] = clbase2[f] + (ecount[f] > 0 ? 1 : 0); maxE = std::max(maxE, ecount[f]); }
    const long long E = ebase[nf], NC2 = clbase2[nf];
    q->launches++;
    if (E == 0) return FPK_OK;
    // ---- the zone batch: one structure per frame whose spheres are the entries, one cluster each -----------------------------------
    const int nHull2 = (int)std::min<long long>(NC2, 16);
    std::vector<DelScratch> hHull2(nHull2);
    StructDev *dS2 = nullptr;
    int *d_ebase = nullptr, *d_clbase2 = nullptr, *d_fb = nullptr, *d_cnt = nullptr;
    unsigned char *d_done = nullptr;
    DelScratch *dHull2 = nullptr;
    PocketDev P2;
    memset(&P2, 0, sizeof P2);
    rc = arena_build(q->A3, [&](Arena &a) {
        d_ebase = a.take<int>(nf + 1); d_clbase2 = a.take<int>(nf + 1); dS2 = a.take<StructDev>(nf);
        Z.entry_sphere = a.take<int>(E + 1); Z.entry_pocket = a.take<int>(E + 1);
        Z.z_xyzr = a.take<float4>(E + 1); Z.z_bary = a.take<float>(3 * E + 3); Z.z_neigh = a.take<int4>(E + 1); Z.z_type = a.take<uint8_t>(E + 1);
        Z.z_cl_sph = a.take<int>(E + 1); Z.z_cl_off = a.take<int>(E + 2 * (size_t)nf + 2);
        Z.z_counts = a.take<int>((size_t)nf * CNT_STRIDE); Z.z_mcid = a.take<int>(nf);
        P2.desc = a.take<fpk_desc>(NC2 + 1);
        P2.cl_atoms = a.take<int>(4 * E + 4); P2.ent_u = a.take<int>(4 * E + 4); P2.touch_off = a.take<int>(5 * E + 5);
        P2.touch_list = a.take<int>(4 * E + 4); P2.acc = a.take<int>(16 * E + 16); P2.fx_val = a.take<float>(16 * E + 16);
        P2.fx_who = a.take<int>(2 * b->tot_atoms + 2); P2.atom_fx = a.take<float>(4 * b->tot_atoms + 4);
        P2.rank_to_cluster = a.take<int>(E + nf + 1);
        d_fb = a.take<int>(NC2 + 1); d_done = a.take<unsigned char>(NC2 + 1); d_cnt = a.take<int>(8);
        dHull2 = a.take<DelScratch>(nHull2);
        for (int i = 0; i < nHull2; i++) del_scratch_layout(a, hHull2[i], maxE, 0, HULL_THREADS, false);
    });
    if (rc) return rc;
    Z.ebase = d_ebase;
    std::vector<StructDev> S2(b->S.begin(), b->S.begin() + nf);
    for (int f = 0; f < nf; f++) { S2[f].sph_off = ebase[f]; S2[f].sph_cap = ecount[f]; }
    CK(cudaMemcpyAsync(d_ebase, ebase.data(), sizeof(int) * (nf + 1), cudaMemcpyHostToDevice, cudaStreamPerThread));
    CK(cudaMemcpyAsync(d_clbase2, clbase2.data(), sizeof(int) * (nf + 1), cudaMemcpyHostToDevice, cudaStreamPerThread));
    CK(cudaMemcpyAsync(dS2, S2.data(), sizeof(StructDev) * nf, cudaMemcpyHostToDevice, cudaStreamPerThread));
    CK(cudaMemcpyAsync(dHull2, hHull2.data(), sizeof(DelScratch) * nHull2, cudaMemcpyHostToDevice, cudaStreamPerThread));
    CK(cudaMemsetAsync(P2.desc, 0, sizeof(fpk_desc) * (size_t)(NC2 + 1), cudaStreamPerThread));
    CK(cudaMemsetAsync(P2.fx_who, 0xff, sizeof(int) * (2 * (size_t)b->tot_atoms + 2), cudaStreamPerThread));
    CK(cudaMemcpyAsync(P2.atom_fx, b->Pk.atom_fx, sizeof(float) * 4 * (size_t)b->tot_atoms, cudaMemcpyDeviceToDevice, cudaStreamPerThread));   // the frame's own pass came first
    CK(cudaMemsetAsync(d_done, 0, (size_t)NC2 + 1, cudaStreamPerThread));
    CK(cudaMemsetAsync(d_cnt, 0, sizeof(int) * 8, cudaStreamPerThread));
    k_zone_select<<<nf, 256>>>(b->B, b->Pk, Z, 1);
    BatchDev B2 = b->B;
    B2.S = dS2; B2.s_xyzr = Z.z_xyzr; B2.s_bary = Z.z_bary; B2.s_neigh = Z.z_neigh; B2.s_type = Z.z_type;
    B2.cl_off = Z.z_cl_off; B2.cl_sph = Z.z_cl_sph; B2.counts = Z.z_counts;
    B2.prm.min_pock_nb_asph = 0;            // the zone's record is written for every snapshot: always with volume and hull (mdpocket.c:453)
    P2.clbase = d_clbase2; P2.sbase = d_ebase; P2.total_clusters = (int)NC2; P2.mc_id = Z.z_mcid; P2.sph_uid = Z.entry_sphere;
    const size_t hull_dsm = FPK_BD_BYTES + (HULL_THREADS / 32) * sizeof(WarpCav) + 16;
    const int nhw = (int)std::max(1ll, std::min<long long>((NC2 + 32 * HW_WARPS - 1) / (32 * HW_WARPS), (long long)g_sms * FPK_HW_MINBLOCKS));
    k_hull_warp<<<nhw, HW_THREADS, HW_WARPS * sizeof(HullWarp)>>>(B2, P2, d_cnt + 4, d_fb, d_cnt + 5, 1);
    k_hull_volume<<<nHull2, HULL_THREADS, hull_dsm>>>(B2, P2, dHull2, d_cnt + 2, 0, d_fb, d_cnt + 5);
    k_descriptors<<<(int)std::min<long long>(NC2, b->nDesc), K4_THREADS>>>(B2, P2, b->dDesc, d_cnt + 3, b->AG, d_done);
    k_atom_fx<<<(unsigned)NC2, 64>>>(B2, P2);
    q->launches += 5;
    // ---- results -----------------------------------------------------------------------------------------------------------------------
    ZoneOwner *own = new ZoneOwner();
    const size_t bytes = 8 * (size_t)E + 16 * (size_t)E + (size_t)E + 16 * (size_t)E + sizeof(fpk_desc) * (size_t)NC2 + 16 * (size_t)b->tot_atoms + 1024;
    own->blk.resize(bytes);
    char *w = own->blk.data();
    auto carve = [&](size_t n) { char *r = w; w += (n + 15) & ~(size_t)15; return r; };
    int *h_sph = (int *)carve(4 * (size_t)E);
    int *h_pk = (int *)carve(4 * (size_t)E);
    float *h_xyzr = (float *)carve(16 * (size_t)E);
    uint8_t *h_type = (uint8_t *)carve((size_t)E);
    int *h_atoms = (int *)carve(16 * (size_t)E);
    fpk_desc *h_desc = (fpk_desc *)carve(sizeof(fpk_desc) * (size_t)NC2);
    float *h_fx = (float *)carve(16 * (size_t)b->tot_atoms);
    e = cudaMemcpy(h_sph, Z.entry_sphere, 4 * (size_t)E, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(h_pk, Z.entry_pocket, 4 * (size_t)E, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(h_xyzr, Z.z_xyzr, 16 * (size_t)E, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(h_type, Z.z_type, (size_t)E, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(h_atoms, P2.cl_atoms, 16 * (size_t)E, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(h_desc, P2.desc, sizeof(fpk_desc) * (size_t)NC2, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(h_fx, P2.atom_fx, 16 * (size_t)b->tot_atoms, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { delete own; return fail(FPK_ERR_CUDA, "characterize: %s", cudaGetErrorString(e)); }
    for (int f = 0; f < nf; f++) {
        fpk_zone &z = out[f];
        z.atom_fx = h_fx + 4 * (size_t)b->S[f].atom_off;
        z._owner = own; own->refs++;
        if (ecount[f] == 0) continue;
        z.status = FPK_OK;
        z.n_entries = ecount[f];
        z.spheres = h_sph + ebase[f]; z.sph_xyzr = h_xyzr + 4 * (size_t)ebase[f]; z.sph_type = h_type + ebase[f]; z.sph_pocket = h_pk + ebase[f];
        z.desc = h_desc[clbase2[f]];
        z.desc.first_sphere = z.spheres[0];
        z.n_atoms = z.desc.n_atoms;
        z.atoms = h_atoms + 4 * (size_t)ebase[f];          // device layout: 4 slots per entry, the first n_atoms used
    }
    return FPK_OK;
}

extern "C" int fpk_md_characterize(fpk_md *m, const float *xyz, int nframes, const float *zone_xyz, int nzone, fpk_zone *out)
{
    if (!m || !xyz || nframes <= 0 || !out || nzone < 0 || (nzone > 0 && !zone_xyz)) return fail(FPK_ERR_BAD_ARG, "fpk_md_characterize: bad arguments");
    MdDev *q = m->D[0];
    CK(cudaSetDevice(q->dev));
    if (nzone > q->nzone_cap) { cudaFree(q->d_zone); q->d_zone = nullptr; q->nzone_cap = 0; CK(cudaMalloc(&q->d_zone, sizeof(float) * 3 * (size_t)nzone)); q->nzone_cap = nzone; }
    if (nzone > 0) CK(cudaMemcpy(q->d_zone, zone_xyz, sizeof(float) * 3 * (size_t)nzone, cudaMemcpyHostToDevice));
    const size_t na = (size_t)m->topo.natoms;
    const int chunk = 256;
    for (int f0 = 0; f0 < nframes; f0 += chunk) {
        const int nf = std::min(chunk, nframes - f0);
        int rc = md_characterize_chunk(m, q, xyz + 3 * na * (size_t)f0, nf, nzone, out + f0);
        if (rc) { for (int f = 0; f < f0 + nf; f++) fpk_zone_free(&out[f]); return rc; }
    }
    return FPK_OK;
}

extern "C" long long fpk_md_counter(const fpk_md *m, int which)
{
    if (!m) return -1;
    long long v = 0;
    for (const MdDev *q : m->D) {
        if (which == 0) v += q->frames;
        else if (which == 2) v += q->launches;
        else if (which == 1 && q->d_nscat) { unsigned long long u = 0; cudaSetDevice(q->dev); cudaMemcpy(&u, q->d_nscat, 8, cudaMemcpyDeviceToHost); v += (long long)u; }
        else if (which == 3) v += q->frames > 0 ? 1 : 0;          // devices that took a share of the frames
    }
    if (!m->D.empty()) cudaSetDevice(m->D[0]->dev);
    return v;
}
extern "C" void fpk_md_end(fpk_md *m) { delete m; }
