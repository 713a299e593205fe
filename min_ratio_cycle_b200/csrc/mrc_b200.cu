// mrc_b200.cu -- host side of libmrc_b200.so: device-resident graph, batched probes with early
// negative-cycle detection, the numeric multisection driver and the exact (int64) Stern-Brocot driver.
// C ABI declared in include/mrc_b200.h; each entry point cites the reference method it replaces.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "mrc_kernels.cuh"
#include "mrc_probe.cuh"
#define MRC_MAX_DEVICES 64
#define MRC_SPARSE_MIN_EDGES 1500000   /* below this a dense sweep is launch-bound (~10 us) and a worklist pass does not pay */
#include "mrc_search.h"
#include "mrc_batch.cuh"
#include "mrc_host.h"

typedef __int128 i128;

thread_local std::string g_mrc_err;

struct PatchRec {
    ll k;
    ll *d_idx;
    double *d_oldc, *d_oldt;
};

struct ProbeOut {
    int verdict = -1;
    int len = 0;
    int minv = -1;               // smallest vertex of the cycle
    int start = -1;              // vertex the device walk (and its edge list) starts from
    double sumc = 0, sumt = 0;   // predecessor-walk order (device)
    ll isumc = 0, isumt = 0;
    ll passes = 0;
};

struct mrc_graph {
    int device = 0, sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    ll n = 0, m = 0;
    int *d_src = nullptr, *d_dst = nullptr;
    double *d_cost = nullptr, *d_time = nullptr;
    ll *d_icost = nullptr, *d_itime = nullptr;
    // workspace
    void *d_dist = nullptr;      // [n][KT] f64 or i64
    ull *d_pred = nullptr;       // [n][KT] (src vertex << 32 | edge index)
    int *d_jump = nullptr;       // [n][KT]
    int *d_cyc_edges = nullptr;  // [KMAX][cap]
    int *d_best_edges = nullptr; // [cap] edge list of the best cycle of the running solve (stash_cycle)
    ll cyc_cap = 0;
    DevStatus *d_status = nullptr, *h_status = nullptr;
    int *d_g_src = nullptr; double *d_g_c = nullptr, *d_g_t = nullptr; ll *d_g_ic = nullptr, *d_g_it = nullptr;
    uint8_t *d_mask = nullptr;
    ull *d_keys = nullptr;
    int *d_verr = nullptr;
    // host mirrors needed by the exact path (tight-subgraph DFS and first-(u,v)-edge lookup)
    std::vector<int> h_src, h_dst;
    std::vector<ll> h_icost, h_itime, h_starts;
    ll max_abs_ic = 0, max_abs_it = 0, max_it = 1;
    double min_ratio_i = 0, max_ratio_i = 0;
    bool has_int = false;
    // dist cache: the last exact K=1 probe that converged left potentials in d_dist
    bool pot_valid = false; ll pot_a = 0, pot_b = 0;
    std::vector<PatchRec> patches;
    ProbeOut last_res[MRC_KMAX];   // last probe's per-candidate results (cycles stay on the device until the next probe)
    bool last_exact = false; int last_K = 0;
    bool scen_on = false; int scen_pe[MRC_KMAX]; double scen_dc[MRC_KMAX], scen_dt[MRC_KMAX];   // scenario batching of the current probe
    int relax_blocks[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
    int stop_on_neg = 0;         // option "stop_on_neg": minimum K from which mrc_solve_numeric ends rounds early (0 = never)
    int check_blocks[4] = {0, 0, 0, 0};
    int burst_init = 2, burst_max = 16;
    // sparse (worklist) passes: out-edge index by source, worklists + membership bitmaps, built on first use
    int sparse = 0;              // option "sparse" (env MRC_B200_SPARSE): bursts run as worklist passes once few vertices are still lowered
    int sparse_div = 32;         // option "sparse_div": ... i.e. fewer than n / sparse_div
    bool sparse_ready = false, bm1_dirty = false;
    int *d_out_off = nullptr, *d_out_edge = nullptr, *d_wl[2] = {nullptr, nullptr};
    unsigned *d_bm[2] = {nullptr, nullptr};
    size_t bm_bytes = 0;
    int sparse_blocks[2] = {0, 0};
    ll sparse_min_edges = MRC_SPARSE_MIN_EDGES;   // option "sparse_min_edges": smaller graphs sweep densely (launch-bound anyway)
    int sparse_look = 0;         // option "sparse_look": dense bursts are at most this long while the sparse passes wait to take over (0 = by graph size)
    int sparse_bps = 2;          // CTAs per SM of the cooperative sparse kernel (env MRC_B200_SPARSE_BPS, experiments)
    int prune_above = 1;         // option "prune_above": cycle checks stop examining candidates above the lowest one with a cycle
    int gate = 1;                // option "gate": launches of a burst return at once after a launch that moved nothing
    int patience = 0;            // option "patience": mrc_solve_numeric passes MRC_F_PATIENCE to its probes
    int narrow = 1;              // option "narrow": late passes run the narrowest kernel covering the undecided candidates
    int compact = 1;             // option "compact": a probe narrowed to ONE live slot continues on a compact [n][1] copy of its column
    ll compact_min_edges = 2000000;   // ... on graphs where a sweep is bandwidth-bound (smaller ones are launch-bound)
    int pdl = 1;                 // option "pdl": sweep i > 0 of a burst is launched with programmatic stream serialisation (queued while sweep i-1 drains)
    int cert_first = 1;          // option "cert_first": mrc_solve_numeric passes MRC_F_CERT_FIRST to the probes of rounds that hold a cycle
    int tick_max = 8;            // option "tick_max": longest tick (passes between two exchanges) of the sharded search
    int check_growth = 50;       // option "check_growth" (0 = the doubling schedule 2, 6, 14, 30, ...): see ProbeRun::enqueue_burst
    int check_every = 0;         // option "check_every" (experiments): cycle checks every so many passes after the first instead of 2, 6, 14, 30, ...
    int k_after_cycle = 1;       // option "k_after_cycle": candidates per round once the search holds a cycle (0 = as many as before).  Multisection
                                 // brackets lambda*; with a cycle in hand what is left is Newton steps on its ratio and the final certifying probe --
                                 // single-candidate probes that run 1-wide in the persistent kernel instead of dragging K - 1 companions along
    int dprobe = 1;              // option "dprobe": with several ranks, probes of a search that holds a cycle are DISTRIBUTED (every rank sweeps
                                 // a slice of the edge stream, mrc_comm.cuh) instead of run by one rank while the others wait
    int dprobe_group = 0;        // option "dprobe_group": sweeps of a distributed probe between two exchanges (0 = max(2, ranks / 2))
    int slice_W = 0, slice_me = -1;   // distributed probe: cached slice of this rank (graph_slice)
    ll slice_e[2] = {0, 0}, slice_v[2] = {0, 0};
    int qstart = 1;              // option "qstart": the first round of a search without a known cycle places its candidates at low quantiles of the
                                 // edge-ratio distribution (see plan_first_round) instead of uniformly over [min c/t - 1, max c/t + 1]
    ll qstart_min_edges = 20000; // ... on graphs of at least this many edges (smaller ones keep the reference's midpoints)
    std::vector<double> qsample; // sorted sample of cost/time over the edges (first use)
    int persist = 2;             // option "persist" (env MRC_B200_PERSIST): a probe is ONE cooperative launch (mrc_probe.cuh) instead of bursts + host
                                 // round trips.  0 never, 1 whenever the kernel exists (K <= 4), 2 (default) for single-candidate probes: measured at C3
                                 // (profiles/r02_persist.txt) the K = 1 searches gain 10-30 %, the K = 4 ones lose the register double buffer of the
                                 // crowded first sweeps (165 instead of 116 us) and come out even or behind
    int probe_blocks[MRC_PW][2] = {{0, 0}, {0, 0}, {0, 0}};
    ProbeDev *d_probe = nullptr, *h_probe = nullptr;
    void *d_cdist = nullptr; ull *d_cpred = nullptr; int *d_cjump = nullptr;   // compact column (compact survivor)
    void *d_base = nullptr;      // [n] potentials of the base probe of mrc_scenarios_certify
    int *d_order = nullptr;      // mrc_graph_build: index among the preprocessed insertion-order edges of sorted edge i
    int *d_starts = nullptr;     // [n+1] CSR by destination (built on first use by mrc_cycle_weights if the graph came sorted)
    unsigned char *d_keep = nullptr;   // mrc_graph_build: 1 = input edge survived _remove_dominated_edges
    ll m_input = 0;
    mrc_stats stats;
};

extern "C" int mrc_abi_version(void) { return MRC_ABI_VERSION; }
extern "C" const char *mrc_last_error(void) { return g_mrc_err.c_str(); }
extern "C" int mrc_device_count(int *count) {
    if (!count) return set_err(MRC_ERR_ARG, "count is NULL");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *count = 0; cudaGetLastError(); return set_err(MRC_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e)); }
    *count = c;
    return c > 0 ? MRC_OK : set_err(MRC_ERR_NO_DEVICE, "no CUDA device");
}

static int kt_of(int K) { return K <= 1 ? 1 : K <= 2 ? 2 : K <= 4 ? 4 : 8; }
static int kt_index(int KT) { return KT == 1 ? 0 : KT == 2 ? 1 : KT == 4 ? 2 : 3; }

template <int K> static const void *relax_fn(bool exact, bool scen) {
    if (exact) return (const void *)relax_kernel<K, true, false>;
    return scen ? (const void *)relax_kernel<K, false, true> : (const void *)relax_kernel<K, false, false>;
}
static const void *relax_fn_of(int KT, bool exact, bool scen = false) {
    switch (KT) {
        case 1: return relax_fn<1>(exact, scen);
        case 2: return relax_fn<2>(exact, scen);
        case 4: return relax_fn<4>(exact, scen);
        default: return relax_fn<8>(exact, scen);
    }
}
static const void *probe_fn_of(int KT, bool exact) {
    switch (KT) {
        case 1: return exact ? (const void *)probe_kernel<1, true> : (const void *)probe_kernel<1, false>;
        case 2: return exact ? (const void *)probe_kernel<2, true> : (const void *)probe_kernel<2, false>;
        default: return exact ? (const void *)probe_kernel<4, true> : (const void *)probe_kernel<4, false>;
    }
}
static const void *check_fn_of(int KT) {
    switch (KT) {
        case 1: return (const void *)cycle_check_kernel<1>;
        case 2: return (const void *)cycle_check_kernel<2>;
        case 4: return (const void *)cycle_check_kernel<4>;
        default: return (const void *)cycle_check_kernel<8>;
    }
}

// Device facts that do not depend on the graph are queried once per process and device (several graphs, devices and
// host threads may share the library: one record per device, filled under its own once-flag).
struct DeviceInfo {
    std::once_flag once;
    int rc = MRC_OK;
    std::string err;
    int sms = 0, occ_relax[4][2], occ_check[4], occ_probe[MRC_PW][2];
};
static DeviceInfo g_devinfo[MRC_MAX_DEVICES];
static std::mutex g_probe_mu[MRC_MAX_DEVICES];
static int device_info_fill(int device, DeviceInfo &d) {
    CUDA_TRY(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, device));
    for (int KT : {1, 2, 4, 8})
        for (int ex = 0; ex < 2; ex++) {
            const int i = kt_index(KT);
            CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_relax[i][ex], relax_fn_of(KT, ex), MRC_RELAX_THREADS, 0));
            if (ex == 0) CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_check[i], check_fn_of(KT), 256, 0));
            if (i < MRC_PW) CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.occ_probe[i][ex], probe_fn_of(KT, ex), MRC_RELAX_THREADS, 0));
        }
    return MRC_OK;
}
static int device_info(int device, const DeviceInfo **out) {
    if (device < 0 || device >= MRC_MAX_DEVICES) return set_err(MRC_ERR_ARG, "device %d out of range", device);
    DeviceInfo &d = g_devinfo[device];
    std::call_once(d.once, [&]() { d.rc = device_info_fill(device, d); if (d.rc) d.err = g_mrc_err; });
    if (d.rc) { g_mrc_err = d.err; return d.rc; }
    *out = &d;
    return MRC_OK;
}

// Pinned host blocks for the status / result read-backs of a graph handle ([DevStatus][ProbeDev]).  cudaMallocHost and
// cudaFreeHost cost 0.1-0.2 ms each; callers that create a handle per solve (bench.py's e2e leg, the facade) get a block
// back from this small process-wide cache instead.
static std::mutex g_pin_mu;
static std::vector<void *> g_pin_free;
static constexpr size_t kPinBytes = ((sizeof(DevStatus) + 255) & ~(size_t)255) + sizeof(ProbeDev);
static int pin_get(void **out) {
    {
        std::lock_guard<std::mutex> lock(g_pin_mu);
        if (!g_pin_free.empty()) { *out = g_pin_free.back(); g_pin_free.pop_back(); return MRC_OK; }
    }
    CUDA_TRY(cudaMallocHost(out, kPinBytes));
    return MRC_OK;
}
static void pin_put(void *p) {
    {
        std::lock_guard<std::mutex> lock(g_pin_mu);
        if (g_pin_free.size() < 16) { g_pin_free.push_back(p); return; }
    }
    cudaFreeHost(p);
}

extern "C" int mrc_graph_destroy(mrc_graph *g) {
    if (!g) return MRC_OK;
    cudaSetDevice(g->device);
    for (auto &p : g->patches) { cudaFreeAsync(p.d_idx, g->stream); cudaFreeAsync(p.d_oldc, g->stream); cudaFreeAsync(p.d_oldt, g->stream); }
    if (g->stream) {
        void *ptrs[] = {g->d_src, g->d_dst, g->d_cost, g->d_time, g->d_icost, g->d_itime, g->d_dist, g->d_pred, g->d_jump,
                        g->d_cyc_edges, g->d_status, g->d_g_src, g->d_g_c, g->d_g_t, g->d_g_ic, g->d_g_it, g->d_mask,
                        g->d_keys, g->d_verr, g->d_out_off, g->d_out_edge,
                        g->d_wl[0], g->d_wl[1], g->d_bm[0], g->d_bm[1], g->d_best_edges, g->d_probe, g->d_cdist, g->d_cpred, g->d_cjump, g->d_base, g->d_order, g->d_starts, g->d_keep};
        for (void *q : ptrs) if (q) cudaFreeAsync(q, g->stream);
        cudaStreamSynchronize(g->stream);
    }
    if (g->h_status) pin_put(g->h_status);   // (h_probe lives in the same block)
    for (auto &e : g->ev) if (e) cudaEventDestroy(e);
    if (g->stream) cudaStreamDestroy(g->stream);
    delete g;
    return MRC_OK;
}

// launch geometry: persistent grids sized to whole waves of the 148 SMs
static int graph_geometry(mrc_graph *g, const DeviceInfo *di) {
    const ll n = g->n, m = g->m;
    for (int KT : {1, 2, 4, 8})
        for (int ex = 0; ex < 2; ex++) {
            const int i = kt_index(KT);
            if (di->occ_relax[i][ex] < 1 || di->occ_check[i] < 1) return set_err(MRC_ERR_CUDA, "kernel K=%d does not fit on an SM", KT);
            const int rows = KT == 8 ? RelaxTune<8>::rows : RelaxTune<1>::rows;
            const ll chunks = (m + 32 * rows - 1) / (32 * rows);   // one chunk per warp step
            const ll need = std::max<ll>(1, (chunks + MRC_RELAX_THREADS / 32 - 1) / (MRC_RELAX_THREADS / 32));
            g->relax_blocks[i][ex] = (int)std::min<ll>(need, (ll)di->occ_relax[i][ex] * g->sms);
            const ll needc = std::max<ll>(MRC_KMAX, (n * KT + 255) / 256);
            g->check_blocks[i] = (int)std::min<ll>(needc, (ll)di->occ_check[i] * g->sms);
            if (i < MRC_PW) {
                if (di->occ_probe[i][ex] < 1) return set_err(MRC_ERR_CUDA, "probe kernel K=%d does not fit on an SM", KT);
                // one grid serves the sweeps (a chunk per warp step) and the check (n * KT entries, one CTA per candidate)
                const ll want = std::max<ll>(std::max<ll>(need, (n * KT + 255) / 256), MRC_KMAX);
                g->probe_blocks[i][ex] = (int)std::min<ll>(want, (ll)di->occ_probe[i][ex] * g->sms);
            }
        }
    return MRC_OK;
}

// Allocation of everything a graph handle owns on `device` (edge arrays uninitialised), launch geometry, options.
static int graph_new(int64_t n, int64_t m, int device, bool with_int, mrc_graph **out) {
    *out = nullptr;
    if (n <= 0 || m <= 0) return set_err(MRC_ERR_ARG, "n and m must be positive (n=%lld m=%lld)", (ll)n, (ll)m);
    if (n > 0x7ffffff0ll || m > 0x7ffffff0ll) return set_err(MRC_ERR_ARG, "n, m must fit int32");
    int ndev = 0;
    MRC_TRY(mrc_device_count(&ndev));
    if (device < 0 || device >= ndev) return set_err(MRC_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    {   // keep freed device memory cached in the stream-ordered pool: create/destroy cycles reuse it
        cudaMemPool_t pool;
        CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep = ~0ull;
        CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    mrc_graph *g = new mrc_graph();
    memset(&g->stats, 0, sizeof g->stats);
    g->device = device; g->n = n; g->m = m;
    int rc = [&]() -> int {
        const DeviceInfo *di = nullptr;
        MRC_TRY(device_info(device, &di));
        g->sms = di->sms;
        CUDA_TRY(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
        for (auto &e : g->ev) CUDA_TRY(cudaEventCreate(&e));
        const size_t m4 = ((size_t)m + 3) & ~(size_t)3;   // vector loads never read past the allocation
        DALLOC(g->d_src, m4 * 4); DALLOC(g->d_dst, m4 * 4);
        DALLOC(g->d_cost, m4 * 8); DALLOC(g->d_time, m4 * 8);
        if (with_int) {
            g->has_int = true;
            DALLOC(g->d_icost, m4 * 8); DALLOC(g->d_itime, m4 * 8);
            DALLOC(g->d_mask, (size_t)m);
        }
        DALLOC(g->d_dist, (size_t)n * MRC_KMAX * 8);
        DALLOC(g->d_pred, (size_t)n * MRC_KMAX * 8);
        DALLOC(g->d_jump, (size_t)n * MRC_KMAX * 4);
        g->cyc_cap = n;
        DALLOC(g->d_cyc_edges, (size_t)g->cyc_cap * MRC_KMAX * 4);
        DALLOC(g->d_status, sizeof(DevStatus));
        CUDA_TRY(cudaMemsetAsync(g->d_status, 0, sizeof(DevStatus), g->stream));
        {
            void *blk = nullptr;
            MRC_TRY(pin_get(&blk));
            g->h_status = (DevStatus *)blk;
            g->h_probe = (ProbeDev *)((char *)blk + ((sizeof(DevStatus) + 255) & ~(size_t)255));
        }
        DALLOC(g->d_probe, sizeof(ProbeDev));
        DALLOC(g->d_keys, 16);
        DALLOC(g->d_verr, 4);
        MRC_TRY(graph_geometry(g, di));
        if (const char *env = getenv("MRC_B200_SPARSE")) g->sparse = atoi(env) != 0;   // default of option "sparse"
        if (const char *env = getenv("MRC_B200_SPARSE_DIV")) g->sparse_div = std::max(1, atoi(env));
        if (const char *env = getenv("MRC_B200_SPARSE_BPS")) g->sparse_bps = std::max(1, atoi(env));
        if (const char *env = getenv("MRC_B200_SPARSE_LOOK")) g->sparse_look = std::max(0, atoi(env));
        if (const char *env = getenv("MRC_B200_SPARSE_MIN_EDGES")) g->sparse_min_edges = atoll(env);   // parity suite: 0
        if (const char *env = getenv("MRC_B200_COMPACT_MIN_EDGES")) g->compact_min_edges = atoll(env);   // parity suite: 0
        if (const char *env = getenv("MRC_B200_PERSIST")) g->persist = std::max(0, std::min(2, atoi(env)));
        if (const char *env = getenv("MRC_B200_DPROBE")) g->dprobe = atoi(env) != 0;
        if (const char *env = getenv("MRC_B200_DPROBE_GROUP")) g->dprobe_group = std::max(0, std::min(atoi(env), 64));
        return MRC_OK;
    }();
    if (rc) { mrc_graph_destroy(g); return rc; }
    *out = g;
    return MRC_OK;
}

// input validation on the device (the host never touches the 24 B/edge again); synchronises the stream
static int graph_validate(mrc_graph *g) {
    CUDA_TRY(cudaMemsetAsync(g->d_verr, 0, 4, g->stream));
    validate_edges_kernel<<<(int)std::min<ll>((g->m + 255) / 256, (ll)g->sms * 8), 256, 0, g->stream>>>(
        g->d_src, g->d_dst, g->d_time, g->n, g->m, g->d_verr);
    CUDA_TRY(cudaGetLastError());
    int verr = 0;
    CUDA_TRY(cudaMemcpyAsync(&verr, g->d_verr, 4, cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    { g->stats.other_launches++; g->stats.kernel_launches++; }
    if (verr & 1) return set_err(MRC_ERR_ARG, "an edge has an endpoint outside [0,n)");
    if (verr & 2) return set_err(MRC_ERR_ARG, "edges are not destination-sorted");
    if (verr & 4) return set_err(MRC_ERR_ARG, "edge time must be strictly positive");
    return MRC_OK;
}

extern "C" int mrc_graph_create(int64_t n, int64_t m, const int32_t *src, const int32_t *dst, const double *cost,
                                const double *time, const int64_t *icost, const int64_t *itime, int device,
                                mrc_graph **out) {
    if (!out) return set_err(MRC_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (!src || !dst || !cost || !time) return set_err(MRC_ERR_ARG, "edge arrays are NULL");
    if ((icost == nullptr) != (itime == nullptr)) return set_err(MRC_ERR_ARG, "icost/itime must both be set or both NULL");
    mrc_graph *g = nullptr;
    MRC_TRY(graph_new(n, m, device, icost != nullptr, &g));
    int rc = [&]() -> int {
        CUDA_TRY(cudaMemcpyAsync(g->d_src, src, (size_t)m * 4, cudaMemcpyHostToDevice, g->stream));
        CUDA_TRY(cudaMemcpyAsync(g->d_dst, dst, (size_t)m * 4, cudaMemcpyHostToDevice, g->stream));
        CUDA_TRY(cudaMemcpyAsync(g->d_cost, cost, (size_t)m * 8, cudaMemcpyHostToDevice, g->stream));
        CUDA_TRY(cudaMemcpyAsync(g->d_time, time, (size_t)m * 8, cudaMemcpyHostToDevice, g->stream));
        g->stats.h2d_bytes += (int64_t)m * 24;
        if (icost) {
            CUDA_TRY(cudaMemcpyAsync(g->d_icost, icost, (size_t)m * 8, cudaMemcpyHostToDevice, g->stream));
            CUDA_TRY(cudaMemcpyAsync(g->d_itime, itime, (size_t)m * 8, cudaMemcpyHostToDevice, g->stream));
            g->stats.h2d_bytes += (int64_t)m * 16;
            g->h_src.assign(src, src + m); g->h_dst.assign(dst, dst + m);
            g->h_icost.assign(icost, icost + m); g->h_itime.assign(itime, itime + m);
            g->h_starts.assign((size_t)n + 1, 0);
            for (ll e = 0; e < m; e++) {
                if ((unsigned)src[e] >= (unsigned)n || (unsigned)dst[e] >= (unsigned)n)
                    return set_err(MRC_ERR_ARG, "edge %lld has an endpoint outside [0,n)", e);
                g->h_starts[dst[e] + 1]++;
            }
            for (ll v = 0; v < n; v++) g->h_starts[v + 1] += g->h_starts[v];
            double mn = INFINITY, mx = -INFINITY;
            for (ll e = 0; e < m; e++) {
                if (itime[e] <= 0) return set_err(MRC_ERR_ARG, "edge %lld: integer time must be positive", e);
                g->max_abs_ic = std::max<ll>(g->max_abs_ic, icost[e] < 0 ? -icost[e] : icost[e]);
                g->max_abs_it = std::max<ll>(g->max_abs_it, itime[e]);
                const double r = (double)icost[e] / (double)itime[e];   // solver.py:754
                mn = std::min(mn, r); mx = std::max(mx, r);
            }
            g->max_it = g->max_abs_it;
            g->min_ratio_i = mn; g->max_ratio_i = mx;
        }
        return graph_validate(g);
    }();
    if (rc) { mrc_graph_destroy(g); return rc; }
    *out = g;
    return MRC_OK;
}

// ---- host -> device upload of pageable arrays ---------------------------------------------------------------------
// A plain cudaMemcpyAsync from pageable memory is staged by the driver through its own pinned buffer on ONE thread
// (~10 GB/s here).  Four host threads each narrow / copy their quarter of the array into pinned chunks of their own and
// queue the DMA of a chunk while filling the next: the copy runs at several memcpy streams' worth of bandwidth, and the
// int64 -> int32 narrowing of vertex ids costs no pass of its own.
struct UploadPool {
    std::mutex mu;
    static constexpr int T = 4, NBUF = 2;
    static constexpr size_t CHUNK = 4u << 20;
    void *buf[T][NBUF] = {};
    bool ready = false;
};
static UploadPool g_upload;
static int upload_array(int device, void *d_dst, const void *h_src, size_t count, int src_bytes, int dst_bytes) {
    std::lock_guard<std::mutex> lock(g_upload.mu);
    CUDA_TRY(cudaSetDevice(device));
    if (!g_upload.ready) {
        for (int t = 0; t < UploadPool::T; t++)
            for (int b = 0; b < UploadPool::NBUF; b++) CUDA_TRY(cudaMallocHost(&g_upload.buf[t][b], UploadPool::CHUNK));
        g_upload.ready = true;
    }
    const size_t per_chunk = UploadPool::CHUNK / (size_t)dst_bytes;
    std::vector<int> rcs(UploadPool::T, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < UploadPool::T; t++)
        th.emplace_back([&, t]() {
            if (cudaSetDevice(device) != cudaSuccess) { rcs[t] = 1; return; }
            cudaStream_t st;
            cudaEvent_t ev[UploadPool::NBUF];
            if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { rcs[t] = 1; return; }
            for (auto &e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            const size_t lo = count * t / UploadPool::T, hi = count * (t + 1) / UploadPool::T;
            int b = 0;
            bool used[UploadPool::NBUF] = {};
            for (size_t i = lo; i < hi; i += per_chunk, b ^= 1) {
                const size_t cnt = std::min(per_chunk, hi - i);
                if (used[b]) cudaEventSynchronize(ev[b]);
                if (src_bytes == dst_bytes) memcpy(g_upload.buf[t][b], (const char *)h_src + i * src_bytes, cnt * src_bytes);
                else {   // int64 -> int32 (values were range-checked by the caller or are checked on the device)
                    const int64_t *s = (const int64_t *)h_src + i;
                    int32_t *d = (int32_t *)g_upload.buf[t][b];
                    for (size_t j = 0; j < cnt; j++) d[j] = (int32_t)s[j];
                }
                if (cudaMemcpyAsync((char *)d_dst + i * dst_bytes, g_upload.buf[t][b], cnt * dst_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) rcs[t] = 1;
                cudaEventRecord(ev[b], st);
                used[b] = true;
            }
            if (cudaStreamSynchronize(st) != cudaSuccess) rcs[t] = 1;
            for (auto &e : ev) cudaEventDestroy(e);
            cudaStreamDestroy(st);
        });
    for (auto &x : th) x.join();
    for (int r : rcs) if (r) return set_err(MRC_ERR_CUDA, "host-to-device upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    return MRC_OK;
}

__global__ void range_check_kernel(const int *src, const int *dst, const double *time, ll n, ll m, int *err) {
    int bad = 0;
    for (ll e = (ll)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (ll)gridDim.x * blockDim.x) {
        if ((unsigned)src[e] >= (unsigned)n || (unsigned)dst[e] >= (unsigned)n) bad |= 1;
        if (!(time[e] > 0.0)) bad |= 4;
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicOr(err, bad);
}

/*
 * _preprocess_graph + _build_arrays_if_needed + upload in ONE step (solver.py:450-575): host arrays in INSERTION order
 * (vertex ids int32 or int64), everything else on the device -- range checks, removal of dominated parallel edges,
 * stable sort by destination, CSR offsets -- and the sorted arrays stay where the probes need them (no device -> host ->
 * device round trip of 240 MB at m = 10^7).  The reference-order host mirrors (_U, _V, _C, _T, order) can be fetched
 * later with mrc_graph_fetch_arrays.  Numeric graphs only.
 */
extern "C" int mrc_graph_build(int64_t n, int64_t m, const void *src, const void *dst, int index_bytes, const double *cost,
                               const double *time, int remove_dominated, int device, mrc_graph **out, int64_t *m_out,
                               int64_t *dup_pairs, int64_t *removed) {
    if (!out) return set_err(MRC_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (!src || !dst || !cost || !time) return set_err(MRC_ERR_ARG, "edge arrays are NULL");
    if (index_bytes != 4 && index_bytes != 8) return set_err(MRC_ERR_ARG, "index_bytes must be 4 or 8");
    mrc_graph *g = nullptr;
    MRC_TRY(graph_new(n, m, device, false, &g));
    int *d_s = nullptr, *d_d = nullptr;
    double *d_c = nullptr, *d_t = nullptr;
    int rc = [&]() -> int {
        const size_t M = (size_t)m;
        DALLOC(d_s, M * 4); DALLOC(d_d, M * 4); DALLOC(d_c, M * 8); DALLOC(d_t, M * 8);
        DALLOC(g->d_order, M * 4); DALLOC(g->d_starts, ((size_t)n + 1) * 4); DALLOC(g->d_keep, M);
        CUDA_TRY(cudaStreamSynchronize(g->stream));   // allocations are stream-ordered; the upload threads use their own streams
        MRC_TRY(upload_array(device, d_s, src, M, index_bytes, 4));
        MRC_TRY(upload_array(device, d_d, dst, M, index_bytes, 4));
        MRC_TRY(upload_array(device, d_c, cost, M, 8, 8));
        MRC_TRY(upload_array(device, d_t, time, M, 8, 8));
        g->stats.h2d_bytes += (int64_t)m * 24;
        CUDA_TRY(cudaMemsetAsync(g->d_verr, 0, 4, g->stream));
        range_check_kernel<<<(int)std::min<ll>((m + 255) / 256, (ll)g->sms * 8), 256, 0, g->stream>>>(d_s, d_d, d_t, n, m, g->d_verr);
        CUDA_TRY(cudaGetLastError());
        int verr = 0;
        CUDA_TRY(cudaMemcpyAsync(&verr, g->d_verr, 4, cudaMemcpyDeviceToHost, g->stream));
        CUDA_TRY(cudaStreamSynchronize(g->stream));
        if (verr & 1) return set_err(MRC_ERR_ARG, "an edge has an endpoint outside [0,n)");
        if (verr & 4) return set_err(MRC_ERR_ARG, "edge time must be strictly positive");
        int64_t mk = m, dups = 0, rem = 0;
        MRC_TRY(mrc_build_sorted_device(g->stream, n, m, d_s, d_d, d_c, d_t, remove_dominated, g->d_src, g->d_dst, g->d_cost,
                                        g->d_time, g->d_order, g->d_starts, g->d_keep, &mk, &dups, &rem));
        g->m_input = m;
        g->m = mk;
        const DeviceInfo *di = nullptr;
        MRC_TRY(device_info(device, &di));
        MRC_TRY(graph_geometry(g, di));
        { g->stats.other_launches += 12; g->stats.kernel_launches += 12; }
        if (m_out) *m_out = mk;
        if (dup_pairs) *dup_pairs = dups;
        if (removed) *removed = rem;
        return MRC_OK;
    }();
    for (void *q : {(void *)d_s, (void *)d_d, (void *)d_c, (void *)d_t}) if (q) cudaFreeAsync(q, g->stream);
    if (rc) { mrc_graph_destroy(g); return rc; }
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    *out = g;
    return MRC_OK;
}

/* Destination-sorted arrays of a graph back on the host (any pointer may be NULL): the reference's _U/_V/_C/_T
 * (solver.py:488-491) for callers that read them; order[i] (mrc_graph_build only) = insertion index, among the edges that
 * survived preprocessing, of sorted edge i; keep[j] (mrc_graph_build only, length = number of INPUT edges) = 1 where
 * input edge j survived _remove_dominated_edges. */
extern "C" int mrc_graph_fetch_arrays(mrc_graph *g, int32_t *src, int32_t *dst, double *cost, double *time, int32_t *order,
                                      uint8_t *keep) {
    if (!g) return set_err(MRC_ERR_ARG, "graph is NULL");
    if ((order && !g->d_order) || (keep && !g->d_keep)) return set_err(MRC_ERR_ARG, "order / keep exist only for graphs made by mrc_graph_build");
    CUDA_TRY(cudaSetDevice(g->device));
    const size_t M = (size_t)g->m;
    if (src) CUDA_TRY(cudaMemcpyAsync(src, g->d_src, M * 4, cudaMemcpyDeviceToHost, g->stream));
    if (dst) CUDA_TRY(cudaMemcpyAsync(dst, g->d_dst, M * 4, cudaMemcpyDeviceToHost, g->stream));
    if (cost) CUDA_TRY(cudaMemcpyAsync(cost, g->d_cost, M * 8, cudaMemcpyDeviceToHost, g->stream));
    if (time) CUDA_TRY(cudaMemcpyAsync(time, g->d_time, M * 8, cudaMemcpyDeviceToHost, g->stream));
    if (order) CUDA_TRY(cudaMemcpyAsync(order, g->d_order, M * 4, cudaMemcpyDeviceToHost, g->stream));
    if (keep) CUDA_TRY(cudaMemcpyAsync(keep, g->d_keep, (size_t)g->m_input, cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    g->stats.d2h_bytes += (int64_t)M * ((src ? 4 : 0) + (dst ? 4 : 0) + (cost ? 8 : 0) + (time ? 8 : 0) + (order ? 4 : 0));
    return MRC_OK;
}

// per hop u -> v of a cycle: among the edges u -> v (segment of v in the destination-sorted arrays) the first one with
// the smallest cost/time (solver.py:1309-1313); edge index or -1.  One warp per hop.
__global__ void hop_edges_kernel(const int *cyc, ll L, const int *starts, const int *src, const double *cost, const double *time,
                                 int *hop_edge, double *hop_c, double *hop_t) {
    const int lane = threadIdx.x & 31;
    const ll warp = ((ll)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((ll)gridDim.x * blockDim.x) >> 5;
    for (ll i = warp; i < L; i += nw) {
        const int u = cyc[i], v = cyc[i + 1];
        const int lo = starts[v], hi = starts[v + 1];
        double best = __longlong_as_double(0x7ff0000000000000ll);
        int be = 0x7fffffff;
        for (int e = lo + lane; e < hi; e += 32)
            if (src[e] == u) {
                const double r = __ddiv_rn(cost[e], time[e]);
                if (r < best || (r == best && e < be)) { best = r; be = e; }
            }
        for (int o = 16; o; o >>= 1) {
            const double rb = __shfl_xor_sync(0xffffffffu, best, o);
            const int eb = __shfl_xor_sync(0xffffffffu, be, o);
            if (eb != 0x7fffffff && (be == 0x7fffffff || rb < best || (rb == best && eb < be))) { best = rb; be = eb; }
        }
        if (lane == 0) {
            hop_edge[i] = be == 0x7fffffff ? -1 : be;
            hop_c[i] = be == 0x7fffffff ? 0.0 : cost[be];
            hop_t[i] = be == 0x7fffffff ? 0.0 : time[be];
        }
    }
}

/* _compute_cycle_weights_float + _verify_cycle_exists (solver.py:1296-1327, :1563-1582) on the resident arrays:
 * closed cycle of L hops (L + 1 vertices) in, per hop the edge the reference's rule picks; sums accumulated in cycle
 * order from 0.0 like the reference.  *missing_hop = index of the first hop without an edge, or -1. */
extern "C" int mrc_cycle_weights(mrc_graph *g, int64_t L, const int32_t *closed_cycle, double *sum_cost, double *sum_time,
                                 int32_t *missing_hop, int32_t *hop_edges) {
    if (!g || L < 0 || (L && !closed_cycle)) return set_err(MRC_ERR_ARG, "bad cycle arguments");
    if (missing_hop) *missing_hop = -1;
    if (sum_cost) *sum_cost = 0.0;
    if (sum_time) *sum_time = 0.0;
    if (L == 0) return MRC_OK;
    for (ll i = 0; i <= L; i++)
        if (closed_cycle[i] < 0 || closed_cycle[i] >= g->n) return set_err(MRC_ERR_ARG, "cycle vertex out of range");
    CUDA_TRY(cudaSetDevice(g->device));
    if (!g->d_starts) {
        DALLOC(g->d_starts, ((size_t)g->n + 1) * 4);
        in_offsets_kernel<<<(int)std::min<ll>((g->n + 256) / 256, (ll)g->sms * 8), 256, 0, g->stream>>>(g->d_dst, g->m, g->n, g->d_starts);
        CUDA_TRY(cudaGetLastError());
    }
    int *d_cyc = nullptr, *d_he = nullptr;
    double *d_hc = nullptr, *d_ht = nullptr;
    std::vector<int> he((size_t)L);
    std::vector<double> hc((size_t)L), ht((size_t)L);
    int rc = [&]() -> int {
        DALLOC(d_cyc, ((size_t)L + 1) * 4); DALLOC(d_he, (size_t)L * 4); DALLOC(d_hc, (size_t)L * 8); DALLOC(d_ht, (size_t)L * 8);
        CUDA_TRY(cudaMemcpyAsync(d_cyc, closed_cycle, ((size_t)L + 1) * 4, cudaMemcpyHostToDevice, g->stream));
        hop_edges_kernel<<<(int)std::min<ll>((L * 32 + 255) / 256, 1024), 256, 0, g->stream>>>(d_cyc, L, g->d_starts, g->d_src, g->d_cost,
                                                                                            g->d_time, d_he, d_hc, d_ht);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(he.data(), d_he, (size_t)L * 4, cudaMemcpyDeviceToHost, g->stream));
        CUDA_TRY(cudaMemcpyAsync(hc.data(), d_hc, (size_t)L * 8, cudaMemcpyDeviceToHost, g->stream));
        CUDA_TRY(cudaMemcpyAsync(ht.data(), d_ht, (size_t)L * 8, cudaMemcpyDeviceToHost, g->stream));
        CUDA_TRY(cudaStreamSynchronize(g->stream));
        return MRC_OK;
    }();
    for (void *q : {(void *)d_cyc, (void *)d_he, (void *)d_hc, (void *)d_ht}) if (q) cudaFreeAsync(q, g->stream);
    if (rc) return rc;
    { g->stats.other_launches++; g->stats.kernel_launches++; }
    double sc = 0.0, st = 0.0;
    for (ll i = 0; i < L; i++) {
        if (he[i] < 0) { if (missing_hop && *missing_hop < 0) *missing_hop = (int32_t)i; continue; }
        sc += hc[i]; st += ht[i];
        }
    if (hop_edges) for (ll i = 0; i < L; i++) hop_edges[i] = he[i];
    if (sum_cost) *sum_cost = sc;
    if (sum_time) *sum_time = st;
    return MRC_OK;
}

extern "C" int mrc_graph_info(const mrc_graph *g, int64_t *n, int64_t *m, int *device, int *has_int) {
    if (!g) return set_err(MRC_ERR_ARG, "graph is NULL");
    if (n) *n = g->n;
    if (m) *m = g->m;
    if (device) *device = g->device;
    if (has_int) *has_int = g->has_int;
    return MRC_OK;
}
extern "C" int mrc_get_stats(const mrc_graph *g, mrc_stats *out) {
    if (!g || !out) return set_err(MRC_ERR_ARG, "NULL argument");
    *out = g->stats;
    return MRC_OK;
}
extern "C" int mrc_reset_stats(mrc_graph *g) {
    if (!g) return set_err(MRC_ERR_ARG, "graph is NULL");
    memset(&g->stats, 0, sizeof g->stats);
    return MRC_OK;
}

static double key_to_f64(ull k) {
    ull b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double d;
    memcpy(&d, &b, 8);
    return d;
}
extern "C" int mrc_numeric_bounds(mrc_graph *g, double *lo, double *hi) {
    if (!g || !lo || !hi) return set_err(MRC_ERR_ARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(g->device));
    ull init[2] = {~0ull, 0ull}, keys[2];
    CUDA_TRY(cudaMemcpyAsync(g->d_keys, init, 16, cudaMemcpyHostToDevice, g->stream));
    int blocks = (int)std::min<ll>((g->m + 255) / 256, (ll)g->sms * 8);
    ratio_bounds_kernel<<<blocks, 256, 0, g->stream>>>(g->d_cost, g->d_time, g->m, g->d_keys, g->d_keys + 1);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(keys, g->d_keys, 16, cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    { g->stats.other_launches++; g->stats.kernel_launches++; }
    *lo = key_to_f64(keys[0]) - 1.0;   // solver.py:1022-1023
    *hi = key_to_f64(keys[1]) + 1.0;
    return MRC_OK;
}

/*
 * get_graph_properties (solver.py:339-387) computed on the resident arrays instead of eight NumPy sweeps over the
 * host copies (160 ms at m = 10^7; here two streaming passes + one pass over the vertices, < 0.5 ms).
 * out[0..12] = isolated, source, sink vertices, max in-degree, max out-degree, cost min, max, mean, std,
 * time min, max, mean, std (population std, as np.std).
 */
extern "C" int mrc_graph_properties(mrc_graph *g, double *out) {
    if (!g || !out) return set_err(MRC_ERR_ARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(g->device));
    const ll n = g->n, m = g->m;
    unsigned *d_deg = nullptr;   // indeg[n], outdeg[n], cnt[8]
    double *d_acc = nullptr;     // acc[4], keys[4] (as ull)
    DALLOC(d_deg, ((size_t)2 * n + 8) * 4);
    DALLOC(d_acc, 8 * 8);
    struct { double acc[4]; ull keys[4]; } h;
    for (int i = 0; i < 4; i++) h.acc[i] = 0.0;
    h.keys[0] = ~0ull; h.keys[1] = 0ull; h.keys[2] = ~0ull; h.keys[3] = 0ull;
    CUDA_TRY(cudaMemsetAsync(d_deg, 0, ((size_t)2 * n + 8) * 4, g->stream));
    CUDA_TRY(cudaMemcpyAsync(d_acc, &h, sizeof h, cudaMemcpyHostToDevice, g->stream));
    const int eb = (int)std::min<ll>((m + 255) / 256, (ll)g->sms * 8), vb = (int)std::min<ll>((n + 255) / 256, (ll)g->sms * 8);
    props_edges_kernel<<<eb, 256, 0, g->stream>>>(g->d_src, g->d_dst, g->d_cost, g->d_time, m, d_deg, d_deg + n, d_acc,
                                                  reinterpret_cast<ull *>(d_acc + 4));
    props_deviation_kernel<<<eb, 256, 0, g->stream>>>(g->d_cost, g->d_time, m, d_acc);
    props_vertices_kernel<<<vb, 256, 0, g->stream>>>(d_deg, d_deg + n, n, d_deg + 2 * n);
    CUDA_TRY(cudaGetLastError());
    unsigned cnt[8];
    CUDA_TRY(cudaMemcpyAsync(&h, d_acc, sizeof h, cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaMemcpyAsync(cnt, d_deg + 2 * n, sizeof cnt, cudaMemcpyDeviceToHost, g->stream));
    cudaFreeAsync(d_deg, g->stream); cudaFreeAsync(d_acc, g->stream);
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    { g->stats.other_launches += 3; g->stats.kernel_launches += 3; }
    for (int i = 0; i < 5; i++) out[i] = (double)cnt[i];
    out[5] = key_to_f64(h.keys[0]); out[6] = key_to_f64(h.keys[1]);
    out[7] = h.acc[0] / (double)m; out[8] = sqrt(h.acc[2] / (double)m);
    out[9] = key_to_f64(h.keys[2]); out[10] = key_to_f64(h.keys[3]);
    out[11] = h.acc[1] / (double)m; out[12] = sqrt(h.acc[3] / (double)m);
    return MRC_OK;
}

extern "C" int mrc_graph_patch_edges(mrc_graph *g, int64_t k, const int64_t *idx, const double *dcost,
                                     const double *dtime) {
    if (!g || k < 0 || (k && (!idx || !dcost || !dtime))) return set_err(MRC_ERR_ARG, "bad patch arguments");
    if (k == 0) return MRC_OK;
    for (ll i = 0; i < k; i++)
        if (idx[i] < 0 || idx[i] >= g->m) return set_err(MRC_ERR_ARG, "edge index out of range");
    {   // an edge listed twice would race on its saved weight: mrc_graph_restore could not bring the original back
        std::vector<int64_t> sorted(idx, idx + k);
        std::sort(sorted.begin(), sorted.end());
        if (std::adjacent_find(sorted.begin(), sorted.end()) != sorted.end()) return set_err(MRC_ERR_ARG, "edge index listed twice in one patch");
    }
    CUDA_TRY(cudaSetDevice(g->device));
    PatchRec p;
    p.k = k; p.d_idx = nullptr; p.d_oldc = p.d_oldt = nullptr;
    double *d_dc = nullptr, *d_dt = nullptr;
    int bad = 0;
    int rc = [&]() -> int {
        DALLOC(p.d_idx, (size_t)k * 8); DALLOC(p.d_oldc, (size_t)k * 8); DALLOC(p.d_oldt, (size_t)k * 8);
        DALLOC(d_dc, (size_t)k * 8); DALLOC(d_dt, (size_t)k * 8);
        CUDA_TRY(cudaMemcpyAsync(p.d_idx, idx, (size_t)k * 8, cudaMemcpyHostToDevice, g->stream));
        CUDA_TRY(cudaMemcpyAsync(d_dc, dcost, (size_t)k * 8, cudaMemcpyHostToDevice, g->stream));
        CUDA_TRY(cudaMemcpyAsync(d_dt, dtime, (size_t)k * 8, cudaMemcpyHostToDevice, g->stream));
        CUDA_TRY(cudaMemsetAsync(g->d_verr, 0, 4, g->stream));
        patch_edges_kernel<<<(int)((k + 127) / 128), 128, 0, g->stream>>>(g->d_cost, g->d_time, p.d_idx, d_dc, d_dt, k,
                                                                          p.d_oldc, p.d_oldt, g->d_verr);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(&bad, g->d_verr, 4, cudaMemcpyDeviceToHost, g->stream));
        CUDA_TRY(cudaStreamSynchronize(g->stream));   // the host delta arrays may be reused by the caller
        if (bad) {   // every kernel of the path relies on time > 0 (ratio bounds, the sign test of a cycle): undo and refuse
            unpatch_edges_kernel<<<(int)((k + 127) / 128), 128, 0, g->stream>>>(g->d_cost, g->d_time, p.d_idx, k, p.d_oldc, p.d_oldt);
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cudaStreamSynchronize(g->stream));
            return set_err(MRC_ERR_ARG, "Edge transit time must be strictly positive after the perturbation");
        }
        return MRC_OK;
    }();
    if (d_dc) cudaFreeAsync(d_dc, g->stream);
    if (d_dt) cudaFreeAsync(d_dt, g->stream);
    if (rc) {
        for (void *q : {(void *)p.d_idx, (void *)p.d_oldc, (void *)p.d_oldt}) if (q) cudaFreeAsync(q, g->stream);
        return rc;
    }
    g->patches.push_back(p);
    g->pot_valid = false;
    { g->stats.other_launches++; g->stats.kernel_launches++; }
    g->stats.h2d_bytes += k * 24;
    return MRC_OK;
}
extern "C" int mrc_graph_restore(mrc_graph *g) {
    if (!g) return set_err(MRC_ERR_ARG, "graph is NULL");
    CUDA_TRY(cudaSetDevice(g->device));
    while (!g->patches.empty()) {
        PatchRec p = g->patches.back();
        g->patches.pop_back();
        unpatch_edges_kernel<<<(int)((p.k + 127) / 128), 128, 0, g->stream>>>(g->d_cost, g->d_time, p.d_idx, p.k,
                                                                              p.d_oldc, p.d_oldt);
        CUDA_TRY(cudaGetLastError());
        cudaFreeAsync(p.d_idx, g->stream); cudaFreeAsync(p.d_oldc, g->stream); cudaFreeAsync(p.d_oldt, g->stream);
        { g->stats.other_launches++; g->stats.kernel_launches++; }
    }
    return MRC_OK;
}

// ------------------------------------------------------------------------------------------- probe engine
// Narrowest kernel that still covers every undecided candidate slot: an aligned sub-range of width 1, 2 or 4 of
// the KT slots (alignment keeps the 128/256-bit dist loads legal).  Most rounds spend their late passes on a
// single slow candidate; serving it with the K=1 kernel saves the arithmetic and the registers of the decided ones.
static void narrow_candidates(int KT, unsigned active, int *kc, int *k0) {
    *kc = KT; *k0 = 0;
    for (int w = 1; w < KT; w *= 2)
        for (int o = 0; o < KT; o += w)
            if (active && !(active & ~(((1u << w) - 1u) << o))) { *kc = w; *k0 = o; return; }
}
// `ra` := ra_full restricted to the narrowest aligned candidate sub-range; returns its width
static int narrow_args(const mrc_graph *g, int KT, const RelaxArgs &ra_full, RelaxArgs &ra) {
    ra = ra_full;
    ra.ks = KT; ra.k0 = 0;
    int kc = KT, k0 = 0;
    if (g->narrow) narrow_candidates(KT, ra.active, &kc, &k0);
    if (kc < KT) {
        const size_t esz = 8;   // double and int64 slots alike
        ra.dist = (char *)ra.dist + (size_t)k0 * esz;
        ra.pred += k0;
        ra.active >>= k0;
        ra.k0 = k0;
        for (int j = 0; j < kc; j++) {
            ra.lam[j] = ra_full.lam[k0 + j]; ra.a[j] = ra_full.a[k0 + j]; ra.b[j] = ra_full.b[k0 + j];
            ra.pe[j] = ra_full.pe[k0 + j]; ra.pdc[j] = ra_full.pdc[k0 + j]; ra.pdt[j] = ra_full.pdt[k0 + j];
        }
    }
    return kc;
}
// Compact survivor: candidate slot `k` of the [n][KT] layout continues on the [n][1] copy of its column.
static void compact_args(const mrc_graph *g, int k, const RelaxArgs &ra_full, RelaxArgs &ra) {
    ra = ra_full;
    ra.dist = g->d_cdist; ra.pred = g->d_cpred; ra.ks = 1; ra.k0 = k; ra.active = 1u;
    ra.lam[0] = ra_full.lam[k]; ra.a[0] = ra_full.a[k]; ra.b[0] = ra_full.b[k];
    ra.pe[0] = ra_full.pe[k]; ra.pdc[0] = ra_full.pdc[k]; ra.pdt[0] = ra_full.pdt[k];
}
static int launch_relax_args(mrc_graph *g, int kc, bool exact, const RelaxArgs &ra) {
    void *args[] = {(void *)&ra};
    g->stats.kernel_launches++;
    const int blocks = g->relax_blocks[kt_index(kc)][exact ? 1 : 0];
    if (g->pdl && ra.gate) {   // a sweep that follows another sweep of its burst: queue it while the previous one drains
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(MRC_RELAX_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = g->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CUDA_TRY(cudaLaunchKernelExC(&cfg, relax_fn_of(kc, exact, ra.scen != 0), args));
        return MRC_OK;
    }
    CUDA_TRY(cudaLaunchKernel(relax_fn_of(kc, exact, ra.scen != 0), dim3(blocks), dim3(MRC_RELAX_THREADS), args, 0, g->stream));
    return MRC_OK;
}
static int launch_relax(mrc_graph *g, int KT, bool exact, const RelaxArgs &ra_full, int compact_k = -1) {
    RelaxArgs ra;
    if (compact_k >= 0) { compact_args(g, compact_k, ra_full, ra); return launch_relax_args(g, 1, exact, ra); }
    const int kc = narrow_args(g, KT, ra_full, ra);
    return launch_relax_args(g, kc, exact, ra);
}
static int launch_check(mrc_graph *g, int KT, const CheckArgs &ca) {
    const int blocks = g->check_blocks[kt_index(KT)];
    void *args[] = {(void *)&ca};
    g->stats.kernel_launches++;
    CUDA_TRY(cudaLaunchCooperativeKernel(check_fn_of(KT), dim3(blocks), dim3(256), args, 0, g->stream));
    return MRC_OK;
}

// Out-edge index, worklists and bitmaps of the sparse passes (built once per graph, on first use).
static int ensure_sparse(mrc_graph *g) {
    if (g->sparse_ready) return MRC_OK;
    MRC_TRY(mrc_build_out_csr(g->stream, g->n, g->m, g->d_src, &g->d_out_off, &g->d_out_edge));
    g->bm_bytes = (size_t)((g->n + 31) / 32) * 4;
    for (int i = 0; i < 2; i++) {
        DALLOC(g->d_bm[i], g->bm_bytes);
        DALLOC(g->d_wl[i], (size_t)g->n * 4);
        CUDA_TRY(cudaMemsetAsync(g->d_bm[i], 0, g->bm_bytes, g->stream));
    }
    for (int ex = 0; ex < 2; ex++) {
        int occ = 0;
        const void *fn = ex ? (const void *)relax_sparse_kernel<true> : (const void *)relax_sparse_kernel<false>;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 256, 0));
        if (occ < 1) return set_err(MRC_ERR_CUDA, "sparse relaxation kernel does not fit on an SM");
        g->sparse_blocks[ex] = std::min(occ, g->sparse_bps) * g->sms;
    }
    { g->stats.other_launches += 3; g->stats.kernel_launches += 3; }
    g->sparse_ready = true;
    return MRC_OK;
}
static int launch_sparse(mrc_graph *g, int KT, bool exact, const RelaxArgs &ra_full, int first, int cidx, int npasses,
                         int compact_k = -1) {
    SparseArgs sa;
    memset(&sa, 0, sizeof sa);
    if (compact_k >= 0) { compact_args(g, compact_k, ra_full, sa.r); sa.kc = 1; }
    else sa.kc = narrow_args(g, KT, ra_full, sa.r);
    sa.out_off = g->d_out_off; sa.out_edge = g->d_out_edge;
    sa.bm[0] = g->d_bm[0]; sa.bm[1] = g->d_bm[1]; sa.wl[0] = g->d_wl[0]; sa.wl[1] = g->d_wl[1];
    sa.st = g->d_status; sa.first = first; sa.cidx = cidx; sa.npasses = npasses;
    void *args[] = {(void *)&sa};
    g->stats.kernel_launches++;
    const void *fn = exact ? (const void *)relax_sparse_kernel<true> : (const void *)relax_sparse_kernel<false>;
    CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(g->sparse_blocks[exact ? 1 : 0]), dim3(256), args, 0, g->stream));
    return MRC_OK;
}
static int ensure_compact(mrc_graph *g) {
    if (g->d_cdist) return MRC_OK;
    DALLOC(g->d_cdist, (size_t)g->n * 8);
    DALLOC(g->d_cpred, (size_t)g->n * 8);
    DALLOC(g->d_cjump, (size_t)g->n * 4);
    return MRC_OK;
}

/*
 * Batched Bellman-Ford for K candidates.  Numeric: lam[k]; exact: a[k]/b[k].
 * Per burst: <= burst relaxation launches back to back, then ONE cooperative cycle check, then one
 * small D2H of the status block.  A candidate is decided when
 *   - some pass of the burst changed nothing for it            -> NONNEG (true fixed point)
 *   - the predecessor graph holds a cycle that verifies negative-> NEG
 *   - (numeric) the cycle's ratio is >= lambda but within the rounding noise of the walk -> TIE
 *   - n passes done and still relaxing (reference rule :1252-1257)       -> NEG_NOCYCLE
 *
 * ProbeRun is that loop cut at the host round trip, so that a driver can put work of its own between the burst
 * and the read-back (the multi-GPU driver enqueues an all-gather of the per-candidate records there):
 *   begin()           arguments, dist = 0, pred = none
 *   enqueue_burst()   relaxation launches + cycle check of the next burst on the graph's stream; no host wait
 *   enqueue_readback() + (caller: cudaStreamSynchronize) + collect()   verdicts of that burst
 *   drop()            a candidate whose verdict is implied by something the caller knows
 */
struct ProbeRun {
    mrc_graph *g = nullptr;
    bool exact = false;
    int K = 0, KT = 1;
    unsigned flags = 0;
    double lam[MRC_KMAX];
    ll a[MRC_KMAX], b[MRC_KMAX];
    ProbeOut *res = nullptr;
    RelaxArgs ra;
    CheckArgs ca;
    unsigned active = 0;
    ll passes = 0, next_check = 0, max_passes = 0;
    int burst = 0;
    bool any_neg = false;
    int new_verdicts = 0;
    bool use_sparse = false, sparse_mode = false, may_compact = false;
    int sp_in = 0, sp_ci = 0, look0 = 0, look = 0;
    int ck = -1;   // slot served from the compact column, or -1
    const void *warm_from = nullptr;   // [n] potentials every candidate starts from instead of zero (set before begin())
    bool device_init = false;          // the persistent probe kernel initialises dist / pred itself (set before begin())
    // the burst in flight
    int nb = 0, dense_launches = 0;
    bool sparse_burst = false, do_check = false;

    int begin(mrc_graph *g_, bool exact_, int K_, const double *lam_, const ll *a_, const ll *b_, double slack, unsigned flags_,
              ProbeOut *res_);
    int plan_burst(int cap) const;   // passes the next burst would run (0: the pass limit is reached)
    int enqueue_burst(int cap = 0);
    int enqueue_readback();
    int collect();
    void drop(int k, int verdict) { if ((active >> k) & 1u) { res[k].verdict = verdict; res[k].passes = passes; active &= ~(1u << k); } }
    void finish();
};

int ProbeRun::begin(mrc_graph *g_, bool exact_, int K_, const double *lam_, const ll *a_, const ll *b_, double slack,
                    unsigned flags_, ProbeOut *res_) {
    g = g_; exact = exact_; K = K_; flags = flags_; res = res_;
    if (!g) return set_err(MRC_ERR_ARG, "graph is NULL");
    if (K < 1 || K > MRC_KMAX) return set_err(MRC_ERR_ARG, "K must be in [1,%d]", MRC_KMAX);
    if (exact && !g->has_int) return set_err(MRC_ERR_NOT_INTEGER, "Exact mode requires integer weights");
    if (exact && !g->patches.empty()) return set_err(MRC_ERR_ARG, "exact probes on a patched graph are not supported");
    if (!exact && !(slack >= 0.0)) return set_err(MRC_ERR_ARG, "cycle slack must be >= 0");   // dist <= 0 is what makes the bit-pattern atomic a min
    CUDA_TRY(cudaSetDevice(g->device));
    KT = kt_of(K);
    const ll n = g->n, m = g->m;
    g->pot_valid = false;
    for (int k = 0; k < K; k++) {
        if (exact) { a[k] = a_[k]; b[k] = b_[k]; lam[k] = 0; } else { lam[k] = lam_[k]; a[k] = 0; b[k] = 1; }
    }
    if (exact) {
        for (int k = 0; k < K; k++) {
            if (b[k] <= 0) return set_err(MRC_ERR_ARG, "denominator must be positive");
            const i128 wmax = (i128)b[k] * g->max_abs_ic + (i128)(a[k] < 0 ? -(i128)a[k] : (i128)a[k]) * g->max_abs_it;
            if (wmax * (i128)(n + 1) >= ((i128)1 << 62))
                return set_err(MRC_ERR_OVERFLOW, "int64 overflow: |b*cost - a*time| * n exceeds 2^62 for a=%lld b=%lld", a[k], b[k]);
        }
    }
    memset(&ra, 0, sizeof ra);
    ra.src = g->d_src; ra.dst = g->d_dst; ra.cost = g->d_cost; ra.time = g->d_time;
    ra.icost = g->d_icost; ra.itime = g->d_itime; ra.dist = g->d_dist; ra.pred = g->d_pred; ra.m = m;
    ra.slack = slack; ra.n = n; ra.examined = &g->d_status->examined; ra.decided = &g->d_status->decided;
    memset(&ca, 0, sizeof ca);
    for (int k = 0; k < K; k++) {
        if (exact) { ra.a[k] = ca.a[k] = a[k]; ra.b[k] = ca.b[k] = b[k]; } else ra.lam[k] = ca.lam[k] = lam[k];
        res[k] = ProbeOut();
    }
    ra.scen = ca.scen = (!exact && g->scen_on) ? 1 : 0;
    for (int k = 0; k < MRC_KMAX; k++) {
        ra.pe[k] = ca.pe[k] = (ra.scen && k < K) ? g->scen_pe[k] : -1;
        ra.pdc[k] = ca.pdc[k] = (ra.scen && k < K) ? g->scen_dc[k] : 0.0;
        ra.pdt[k] = ca.pdt[k] = (ra.scen && k < K) ? g->scen_dt[k] : 0.0;
    }
    ca.src = g->d_src; ca.cost = g->d_cost; ca.time = g->d_time; ca.icost = g->d_icost; ca.itime = g->d_itime;
    ca.pred = g->d_pred; ca.dist = g->d_dist; ca.jump = g->d_jump; ca.cyc_edges = g->d_cyc_edges; ca.st = g->d_status;
    ca.n = n; ca.cap = g->cyc_cap; ca.exact = exact;
    if ((flags & MRC_F_MONOTONE_PRUNE) && g->prune_above) {
        bool asc = true;
        for (int k = 1; k < K; k++)
            asc &= exact ? ((i128)a[k - 1] * b[k] < (i128)a[k] * b[k - 1]) : (lam[k - 1] < lam[k]);
        ca.prune_above = asc ? 1 : 0;
    }
    { int r = 1; while (((ll)1 << r) < n + 1) r++; ca.rounds = std::min(r + 1, MRC_MAX_ROUNDS); }

    if (device_init) {
    } else if (warm_from) {
        broadcast_column_kernel<<<std::min<int>((int)((n * KT + 255) / 256), 8 * g->sms), 256, 0, g->stream>>>(
            (const ull *)warm_from, (ull *)g->d_dist, KT, n);
        CUDA_TRY(cudaGetLastError());
        { g->stats.other_launches++; g->stats.kernel_launches++; }
    } else CUDA_TRY(cudaMemsetAsync(g->d_dist, 0, (size_t)n * KT * 8, g->stream));   // dist = 0 (:1148 / :829)
    if (!device_init) {
        CUDA_TRY(cudaMemsetAsync(g->d_pred, 0xff, (size_t)n * KT * 8, g->stream));   // pred = none (:1154)
        CUDA_TRY(cudaMemsetAsync(&g->d_status->decided, 0, sizeof(unsigned), g->stream));
    }
    active = (1u << K) - 1u;
    max_passes = n;   // n-1 passes + the detection pass
    passes = 0;
    burst = g->burst_init;
    next_check = g->burst_init;   // cycle checks after 2, 6, 10, 15, 22, 33, 49, ... passes ("check_growth"); convergence flags every burst
    g->stats.probes += K;
    any_neg = false;
    new_verdicts = 0;
    // sparse passes (see relax_sparse_kernel): decided burst by burst from the number of vertices the last pass lowered
    use_sparse = g->sparse && m >= g->sparse_min_edges;
    if (use_sparse) MRC_TRY(ensure_sparse(g));
    sparse_mode = false;
    sp_in = sp_ci = 0;
    // dense sweeps until the active count is looked at again: every look is a host round trip (~30 us) plus a
    // compaction, worth it when a sweep costs 60-100 us (m >= 5 M edges), not when it costs 10-25 us
    look0 = g->sparse_look > 0 ? g->sparse_look : (m >= 5000000 ? 2 : MRC_MAX_BURST);
    look = look0;
    // compact survivor: once ONE candidate is left of a K >= 2 probe its column moves to [n][1] buffers
    may_compact = g->compact && KT > 1 && m >= g->compact_min_edges;
    ck = -1;
    nb = 0;
    return MRC_OK;
}

int ProbeRun::plan_burst(int cap) const {
    int want = (int)std::min<ll>(std::min<ll>(burst, std::max<ll>(1, next_check - passes)), max_passes - passes);
    // while sweeping densely with the sparse passes available, look at the active count every few sweeps:
    // a worklist pass is 5-10x cheaper than a sweep as soon as it applies
    if (use_sparse && !sparse_mode && want > look) want = look;
    if (cap > 0 && want > cap) want = cap;
    return want;
}

int ProbeRun::enqueue_burst(int cap) {
    const ll n = g->n;
    nb = plan_burst(cap);
    if (nb <= 0) {
        for (int k = 0; k < K; k++) drop(k, MRC_V_NEG_NOCYCLE);
        nb = 0;
        return MRC_OK;
    }
    if (may_compact && ck < 0 && active && !(active & (active - 1))) {
        MRC_TRY(ensure_compact(g));
        ck = __builtin_ctz(active);
        column_copy_kernel<<<std::min<int>((int)((n + 255) / 256), 8 * g->sms), 256, 0, g->stream>>>(
            (const ull *)g->d_dist + ck, g->d_pred + ck, KT, (ull *)g->d_cdist, g->d_cpred, 1, n);
        CUDA_TRY(cudaGetLastError());
        { g->stats.other_launches++; g->stats.kernel_launches++; }
    }
    CUDA_TRY(cudaMemsetAsync(g->d_status, 0, offsetof(DevStatus, cnt), g->stream));   // changed[] + examined
    CUDA_TRY(cudaEventRecord(g->ev[0], g->stream));
    ra.active = active;
    dense_launches = 0;
    sparse_burst = sparse_mode;
    if (sparse_burst) {
        // one cooperative launch runs the nb worklist passes of this burst
        ra.changed = &g->d_status->changed[0];
        ra.act_out = nullptr; ra.gate = nullptr;
        MRC_TRY(launch_sparse(g, KT, exact, ra, sp_in, sp_ci, nb, ck));
        g->bm1_dirty = true;   // either bitmap may now hold bits of a worklist that is never consumed
    }
    for (int i = 0; i < nb && !sparse_burst; i++) {
        ra.changed = &g->d_status->changed[i];
        ra.act_out = nullptr;
        if (use_sparse && i == nb - 1) {
            // the last dense launch of a burst records which vertices it lowered: the seed of a worklist
            CUDA_TRY(cudaMemsetAsync(g->d_bm[0], 0, g->bm_bytes, g->stream));
            ra.act_out = g->d_bm[0];
        }
        dense_launches++;
        // launch i > 0 skips the candidates launch i-1 did not move (fixed point) and returns at once if that is all of them
        ra.gate = (i > 0 && g->gate) ? &g->d_status->changed[i - 1] : nullptr;
        MRC_TRY(launch_relax(g, KT, exact, ra, ck));
    }
    if (use_sparse && !sparse_burst) {
        // bm[1] may hold the membership bits of a worklist nobody consumed (a probe that ended on a verified cycle,
        // or a switch back to dense sweeps): the next worklist pass deduplicates through it, so it must be clean
        if (g->bm1_dirty) { CUDA_TRY(cudaMemsetAsync(g->d_bm[1], 0, g->bm_bytes, g->stream)); g->bm1_dirty = false; }
        CUDA_TRY(cudaMemsetAsync(&g->d_status->act_count, 0, 4 * sizeof(unsigned), g->stream));   // act_count + act_cnt[3]
        compact_active_kernel<<<std::min<int>((int)((g->bm_bytes / 4 + 255) / 256), 4 * g->sms), 256, 0, g->stream>>>(
            g->d_bm[0], (int)(g->bm_bytes / 4), g->d_wl[0], g->d_status, 0);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaEventRecord(g->ev[1], g->stream));
    do_check = passes + nb >= next_check || passes + nb >= max_passes;
    if (do_check) {
        ca.gate = g->gate ? &g->d_status->changed[nb - 1] : nullptr;
        if (ck >= 0) {   // the survivor's predecessor graph lives in the compact column: a 1-wide check, results in slot 0
            CheckArgs cc = ca;
            cc.pred = g->d_cpred; cc.dist = g->d_cdist; cc.jump = g->d_cjump; cc.cyc_edges = g->d_cyc_edges + (ll)ck * g->cyc_cap;
            cc.active = 1u; cc.k0 = ck; cc.prune_above = 0;
            cc.lam[0] = ca.lam[ck]; cc.a[0] = ca.a[ck]; cc.b[0] = ca.b[ck];
            cc.pe[0] = ca.pe[ck]; cc.pdc[0] = ca.pdc[ck]; cc.pdt[0] = ca.pdt[ck];
            MRC_TRY(launch_check(g, 1, cc));
        } else {
            ca.active = active; ca.k0 = 0;
            MRC_TRY(launch_check(g, KT, ca));
        }
        // cycle checks after 2, 6, 14, 30, ... passes, or ("check_growth" > 0) every max(4, growth % of the passes so far): a
        // negative probe is noticed at the first check after its cycle closed
        next_check = g->check_every > 0 ? next_check + g->check_every
                     : g->check_growth > 0 ? next_check + std::max<ll>(4, next_check * g->check_growth / 100) : 2 * next_check + 2;
    }
    CUDA_TRY(cudaEventRecord(g->ev[2], g->stream));
    return MRC_OK;
}

int ProbeRun::enqueue_readback() {
    CUDA_TRY(cudaMemcpyAsync(g->h_status, g->d_status, sizeof(DevStatus), cudaMemcpyDeviceToHost, g->stream));
    return MRC_OK;
}

// after the stream has been synchronised: verdicts of the burst enqueue_burst() launched
int ProbeRun::collect() {
    if (nb <= 0) return MRC_OK;
    const ll n = g->n, m = g->m;
    float ms = 0;
    const DevStatus &hs = *g->h_status;
    // launches / passes that did work: the first of the burst, and each one whose predecessor still moved something
    int ran = nb;
    if (g->gate || sparse_burst)
        for (ran = 1; ran < nb && (hs.changed[ran - 1] & active); ran++) {}
    // one event pair per burst (an event between every two launches was measured: it breaks the back-to-back issue of the
    // kernels and costs ~15 us per launch): the few gated-out launches of a burst (~4 us each) count as relaxation time
    CUDA_TRY(cudaEventElapsedTime(&ms, g->ev[0], g->ev[1]));
    if (sparse_burst) g->stats.sparse_ms += ms; else g->stats.relax_ms += ms;
    CUDA_TRY(cudaEventElapsedTime(&ms, g->ev[1], g->ev[2])); g->stats.check_ms += ms;
    if (sparse_burst) {
        g->stats.sparse_passes += ran;
        g->stats.other_launches += 1;
        sp_in ^= nb & 1; sp_ci = (sp_ci + nb) % 3;
        if ((ll)hs.act_count * g->sparse_div > 2 * n) sparse_mode = false;   // the worklist grew back: sweep densely again
    } else {
        g->stats.relax_launches += ran;
        g->stats.relax_stream_bytes += (int64_t)ran * m * 24;
        g->stats.other_launches += nb - ran + (use_sparse ? 1 : 0);
        g->stats.kernel_launches += use_sparse ? 1 : 0;   // the compaction (sweeps and check are counted where they are launched)
        if (use_sparse) {
            const ll scaled = (ll)hs.act_cnt[0] * g->sparse_div;   // size of the worklist compacted from the last sweep
            if (scaled < n) { sparse_mode = true; sp_in = 0; sp_ci = 0; }
            // far from the threshold: look again later (every look costs a host round trip)
            look = scaled > 8 * n ? 8 * look0 : scaled > 4 * n ? 4 * look0 : scaled > 2 * n ? 2 * look0 : look0;
        }
    }
    // per-candidate work of the burst: launch i relaxed candidate k iff no earlier launch of the burst had left it unmoved
    {
        int64_t slots = 0;
        for (int k = 0; k < K; k++) {
            if (!((active >> k) & 1u)) continue;
            int w = 1;
            while (w < nb && ((hs.changed[w - 1] >> k) & 1u)) w++;
            slots += w;
        }
        if (sparse_burst) g->stats.relax_edge_visits += (int64_t)hs.examined * __builtin_popcount(active);
        else g->stats.relax_edge_visits += slots * m;
        g->stats.relax_edge_slots += slots * m;
    }
    if (do_check) {
        g->stats.check_launches++; g->stats.false_cycles_cut += hs.cuts;
        g->stats.check_doubling_ms += 1e-6 * (double)hs.ns_check[0]; g->stats.check_extract_ms += 1e-6 * (double)hs.ns_check[1];
        g->stats.check_rounds += (int64_t)hs.check_rounds;
    }
    g->stats.d2h_bytes += sizeof(DevStatus);
    for (int k = 0; k < K; k++) {
        if (!((active >> k) & 1u)) continue;
        int conv_at = -1;
        for (int i = 0; i < nb && conv_at < 0; i++) if (!((hs.changed[i] >> k) & 1u)) conv_at = i;
        ProbeOut &o = res[k];
        const int ko = ck >= 0 ? 0 : k;   // slot of candidate k in the check's output
        if (conv_at >= 0) { o.verdict = MRC_V_NONNEG; o.passes = passes + conv_at + 1; }
        else if (do_check && ((hs.cyc_mask >> ko) & 1u) && hs.out[ko].found) {
            // the check kernel extracted a cycle and tested it against lambda on the device:
            // found == 1 verified negative, 2 rounding tie; cycles that are neither were cut there
            const CycleOut &c = hs.out[ko];
            o.verdict = (c.found == 1) ? MRC_V_NEG : MRC_V_TIE;
            o.len = c.len; o.minv = c.minv; o.start = c.pad; o.sumc = c.sumc; o.sumt = c.sumt; o.isumc = c.isumc; o.isumt = c.isumt;
            o.passes = passes + nb;
            any_neg |= c.found == 1;
        }
        if (o.verdict >= 0) { active &= ~(1u << k); new_verdicts++; }
    }
    passes += nb;
    if (flags & MRC_F_MONOTONE_PRUNE) {
        // verdicts are monotone in lambda: NEG at x implies NEG above x, NONNEG at x implies NONNEG below x
        for (int k = 0; k < K; k++) {
            if (res[k].verdict != MRC_V_NEG && res[k].verdict != MRC_V_NONNEG) continue;
            for (int j = 0; j < K; j++) {
                if (!((active >> j) & 1u)) continue;
                bool above = exact ? ((i128)a[j] * b[k] > (i128)a[k] * b[j]) : (lam[j] > lam[k]);
                bool below = exact ? ((i128)a[j] * b[k] < (i128)a[k] * b[j]) : (lam[j] < lam[k]);
                // a verified cycle of ratio r is negative for every lambda > r, not only above lambda_k
                if (res[k].verdict == MRC_V_NEG && !exact && res[k].sumt > 0 && lam[j] > res[k].sumc / res[k].sumt) above = true;
                if (res[k].verdict == MRC_V_NEG && exact && (i128)a[j] * res[k].isumt > (i128)res[k].isumc * b[j]) above = true;
                if (res[k].verdict == MRC_V_NEG && above) drop(j, MRC_V_SKIPPED_NEG);
                if (res[k].verdict == MRC_V_NONNEG && below) drop(j, MRC_V_SKIPPED_NONNEG);
            }
        }
    }
    // ... but not before the second check (pass 6): the candidates closest to lambda* close their cycle a few
    // passes later than the far ones and are the ones worth waiting for (measured: stopping at pass 2
    // needs 10 rounds instead of 4).
    if ((flags & MRC_F_STOP_ON_NEG) && active && do_check && passes >= 6) {
        bool neg_now = false;
        for (int k = 0; k < K; k++) neg_now |= res[k].verdict == MRC_V_NEG;
        if (neg_now) for (int k = 0; k < K; k++) drop(k, MRC_V_ABANDONED);
    }
    // MRC_F_CERT_FIRST: nothing turned negative by the second check and the certifier (last candidate) is still
    // relaxing -- its companions can only raise lo, which is worthless if the certifier converges and unnecessary if it
    // does not (its cycle then moves hi); it continues alone (1-wide kernel, compact column)
    if ((flags & MRC_F_CERT_FIRST) && do_check && passes >= 6 && K > 1 && ((active >> (K - 1)) & 1u) && (active & (active - 1))) {
        bool neg_seen = false;
        for (int k = 0; k < K; k++) neg_seen |= res[k].verdict == MRC_V_NEG || res[k].verdict == MRC_V_TIE || res[k].verdict == MRC_V_SKIPPED_NEG;
        if (!neg_seen) for (int k = 0; k + 1 < K; k++) drop(k, MRC_V_ABANDONED);
    }
    // MRC_F_PATIENCE: a negative cycle is in hand and a whole check interval (as long as everything before it)
    // brought no further verdict -- what is still open is a slow candidate next to lambda*; leave it undecided.
    if ((flags & MRC_F_PATIENCE) && active && do_check) {
        if (any_neg && new_verdicts == 0) for (int k = 0; k < K; k++) drop(k, MRC_V_ABANDONED);
        new_verdicts = 0;
    }
    burst = std::min(burst * 2, g->burst_max);
    nb = 0;
    return MRC_OK;
}

void ProbeRun::finish() {
    if (exact && K == 1 && res[0].verdict == MRC_V_NONNEG) { g->pot_valid = true; g->pot_a = a[0]; g->pot_b = b[0]; }
    for (int k = 0; k < K; k++) g->last_res[k] = res[k];
    g->last_exact = exact; g->last_K = K;
}

// arguments of a persistent probe from a ProbeRun whose begin() has run (with device_init set)
static int probe_args(mrc_graph *g, ProbeRun &pr, unsigned flags, const void *warm, ProbeArgs &pa) {
    const int K = pr.K, KT = pr.KT;
    if (pr.may_compact) MRC_TRY(ensure_compact(g));
    memset(&pa, 0, sizeof pa);
    pa.r = pr.ra;
    pa.r.ks = KT; pa.r.k0 = 0; pa.r.active = (1u << K) - 1u;
    pa.r.changed = nullptr; pa.r.gate = nullptr; pa.r.decided = nullptr; pa.r.act_out = nullptr;
    pa.c = pr.ca;
    pa.c.active = pa.r.active; pa.c.k0 = 0; pa.c.gate = nullptr;
    pa.K = K; pa.flags = flags; pa.asc = pr.ca.prune_above; pa.max_passes = pr.max_passes;
    pa.first_check = g->burst_init; pa.check_growth = g->check_growth; pa.check_every = g->check_every;
    pa.may_compact = pr.may_compact ? 1 : 0;
    pa.cdist = g->d_cdist; pa.cpred = g->d_cpred; pa.cjump = g->d_cjump;
    pa.use_sparse = pr.use_sparse ? 1 : 0; pa.sparse_div = g->sparse_div; pa.look0 = g->sparse_look > 0 ? g->sparse_look : 2;
    pa.out_off = g->d_out_off; pa.out_edge = g->d_out_edge;
    pa.bm[0] = g->d_bm[0]; pa.bm[1] = g->d_bm[1]; pa.wl[0] = g->d_wl[0]; pa.wl[1] = g->d_wl[1];
    pa.bm_words = (int)(g->bm_bytes / 4);
    pa.warm = warm;
    pa.out = g->d_probe;
    pa.dW = 1;
    return MRC_OK;
}
// launch, wait, read the result block; verdicts and statistics of the probe
static int probe_launch_collect(mrc_graph *g, ProbeRun &pr, const ProbeArgs &pa, ProbeOut *res) {
    const int K = pr.K, KT = pr.KT;
    const bool exact = pr.exact;
    const int blocks = g->probe_blocks[kt_index(KT)][exact ? 1 : 0];
    {   // one probe per device at a time: its arguments live in the device's constant memory (c_probe)
        std::lock_guard<std::mutex> lock(g_probe_mu[g->device]);
        CUDA_TRY(cudaMemcpyToSymbolAsync(c_probe, &pa, sizeof pa, 0, cudaMemcpyHostToDevice, g->stream));
        CUDA_TRY(cudaLaunchCooperativeKernel(probe_fn_of(KT, exact), dim3(blocks), dim3(MRC_RELAX_THREADS), nullptr, 0, g->stream));
        CUDA_TRY(cudaMemcpyAsync(g->h_probe, g->d_probe, sizeof(ProbeDev), cudaMemcpyDeviceToHost, g->stream));
        CUDA_TRY(cudaStreamSynchronize(g->stream));
    }
    g->bm1_dirty = g->bm1_dirty || pr.use_sparse;
    const ProbeDev &o = *g->h_probe;
    for (int k = 0; k < K; k++) {
        ProbeOut &r = res[k];
        r.verdict = o.verdict[k]; r.passes = o.passes[k];
        if (r.verdict == MRC_V_NEG || r.verdict == MRC_V_TIE) {
            const CycleOut &c = o.cyc[k];
            r.len = c.len; r.minv = c.minv; r.start = c.pad; r.sumc = c.sumc; r.sumt = c.sumt; r.isumc = c.isumc; r.isumt = c.isumt;
        }
    }
    mrc_stats &st = g->stats;
    for (int i = 0; i < MRC_PW; i++) {
        st.relax_launches += o.n_sweep[i]; st.relax_ms += 1e-6 * (double)o.ns_sweep[i];
        st.relax_stream_bytes += (int64_t)o.n_sweep[i] * pa.r.m * 24;
    }
    st.crowded_sweeps += o.n_crowded; st.crowded_ms += 1e-6 * (double)o.ns_crowded;
    st.sparse_passes += o.n_sparse; st.sparse_ms += 1e-6 * (double)o.ns_sparse;
    st.check_launches += o.n_check; st.check_ms += 1e-6 * (double)o.ns_check;
    st.false_cycles_cut += o.cuts;
    st.check_doubling_ms += 1e-6 * (double)o.ns_chk_phase[0]; st.check_extract_ms += 1e-6 * (double)o.ns_chk_phase[1]; st.check_rounds += (int64_t)o.check_rounds;
    st.relax_edge_slots += (int64_t)o.slots * pa.r.m;
    st.relax_edge_visits += (int64_t)o.slots * pa.r.m + (int64_t)o.sparse_visits;
    st.probe_launches += 1; st.probe_ms += 1e-6 * (double)o.ns_total;
    st.kernel_launches += 1;
    st.d2h_bytes += sizeof(ProbeDev);
    pr.finish();
    return MRC_OK;
}

/*
 * The same probe as ONE cooperative launch (mrc_probe.cuh): the pass loop, the check schedule, narrowing, the compact
 * survivor, the switch to worklist passes and pruning by monotonicity run on the device; the host launches, waits once
 * and reads one result block.
 */
static int run_probe_persist(mrc_graph *g, bool exact, int K, const double *lam, const ll *a, const ll *b, double slack,
                             unsigned flags, ProbeOut *res, const void *warm = nullptr) {
    ProbeRun pr;
    pr.device_init = true;
    MRC_TRY(pr.begin(g, exact, K, lam, a, b, slack, flags, res));
    ProbeArgs pa;
    MRC_TRY(probe_args(g, pr, flags, warm, pa));
    return probe_launch_collect(g, pr, pa, res);
}

static int run_probe(mrc_graph *g, bool exact, int K, const double *lam, const ll *a, const ll *b, double slack,
                     unsigned flags, ProbeOut *res) {
    if (g && (g->persist == 1 ? K <= 4 : g->persist == 2 ? K == 1 : false) && K >= 1 && !g->scen_on &&
        !(flags & (MRC_F_STOP_ON_NEG | MRC_F_PATIENCE)))
        return run_probe_persist(g, exact, K, lam, a, b, slack, flags, res);
    ProbeRun pr;
    MRC_TRY(pr.begin(g, exact, K, lam, a, b, slack, flags, res));
    while (pr.active) {
        MRC_TRY(pr.enqueue_burst());
        if (pr.nb <= 0) break;
        MRC_TRY(pr.enqueue_readback());
        CUDA_TRY(cudaStreamSynchronize(g->stream));
        MRC_TRY(pr.collect());
    }
    pr.finish();
    return MRC_OK;
}

/*
 * Bring the cycle of candidate k to the host: CLOSED vertex list in forward order starting at the
 * smallest vertex, and cost/time sums accumulated in that order over the edges Bellman-Ford used.
 */
static int fetch_cycle(mrc_graph *g, bool exact, int k, const ProbeOut &o, std::vector<int> &verts, double *sc,
                       double *st, ll *isc, ll *ist, const int *edges = nullptr) {
    const ll L = o.len;
    verts.clear();
    if (L <= 0) return MRC_OK;
    if (L > g->cyc_cap) return set_err(MRC_ERR_ARG, "cycle longer than buffer");
    if (!g->d_g_src) {
        DALLOC(g->d_g_src, (size_t)g->cyc_cap * 8);
        DALLOC(g->d_g_c, (size_t)g->cyc_cap * 8); DALLOC(g->d_g_t, (size_t)g->cyc_cap * 8);
        DALLOC(g->d_g_ic, (size_t)g->cyc_cap * 8); DALLOC(g->d_g_it, (size_t)g->cyc_cap * 8);
    }
    const int blocks = (int)std::min<ll>((L + 127) / 128, 1024);
    gather_cycle_kernel<<<blocks, 128, 0, g->stream>>>(edges ? edges : g->d_cyc_edges + (ll)k * g->cyc_cap, L, g->d_src, g->d_cost,
                                                       g->d_time, g->d_icost, g->d_itime, exact, g->d_g_src, g->d_g_c,
                                                       g->d_g_t, g->d_g_ic, g->d_g_it);
    CUDA_TRY(cudaGetLastError());
    std::vector<int> hsrc((size_t)L * 2);
    std::vector<double> hc, ht;
    std::vector<ll> hic, hit;
    CUDA_TRY(cudaMemcpyAsync(hsrc.data(), g->d_g_src, (size_t)L * 8, cudaMemcpyDeviceToHost, g->stream));
    if (exact) {
        hic.resize(L); hit.resize(L);
        CUDA_TRY(cudaMemcpyAsync(hic.data(), g->d_g_ic, (size_t)L * 8, cudaMemcpyDeviceToHost, g->stream));
        CUDA_TRY(cudaMemcpyAsync(hit.data(), g->d_g_it, (size_t)L * 8, cudaMemcpyDeviceToHost, g->stream));
    } else {
        hc.resize(L); ht.resize(L);
        CUDA_TRY(cudaMemcpyAsync(hc.data(), g->d_g_c, (size_t)L * 8, cudaMemcpyDeviceToHost, g->stream));
        CUDA_TRY(cudaMemcpyAsync(ht.data(), g->d_g_t, (size_t)L * 8, cudaMemcpyDeviceToHost, g->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    { g->stats.other_launches++; g->stats.kernel_launches++; }
    g->stats.d2h_bytes += L * 20;
    // The device recorded the BACKWARD walk from `start`: b_0 = start, b_{j+1} = src(e_j), b_L = start, edge e_j
    // runs b_{j+1} -> b_j.  Forward cycle: f_i = b_{L-i}, forward edge i = e_{L-1-i}.  Rotate it so that it starts
    // at the smallest vertex, and accumulate the sums in that order.
    auto fvert = [&](ll i) { return i == 0 ? o.start : hsrc[(size_t)(L - 1 - i)]; };   // f_i for 0 <= i < L
    ll r = 0;
    for (ll i = 0; i < L; i++) if (fvert(i) == o.minv) { r = i; break; }
    verts.reserve((size_t)L + 1);
    for (ll i = 0; i < L; i++) verts.push_back(fvert((r + i) % L));
    verts.push_back(o.minv);
    if (exact) {
        ll c = 0, t = 0;
        for (ll i = 0; i < L; i++) { const ll j = L - 1 - (r + i) % L; c += hic[j]; t += hit[j]; }
        if (isc) *isc = c;
        if (ist) *ist = t;
    } else {
        double c = 0.0, t = 0.0;
        for (ll i = 0; i < L; i++) {
            const ll j = L - 1 - (r + i) % L;
            double ce = hc[j], te = ht[j];
            if (g->scen_on && hsrc[L + j] == g->scen_pe[k]) { ce += g->scen_dc[k]; te += g->scen_dt[k]; }
            c += ce; t += te;
        }
        if (sc) *sc = c;
        if (st) *st = t;
    }
    return MRC_OK;
}

// The lambda search only needs the RATIO of a better cycle to go on (the device sums); its edge list is set aside on
// the device (no host round trip) and brought to the host once, at the end of the solve.
static int stash_cycle(mrc_graph *g, int k, const ProbeOut &o) {
    if (o.len > g->cyc_cap) return set_err(MRC_ERR_ARG, "cycle longer than buffer");
    if (!g->d_best_edges) DALLOC(g->d_best_edges, (size_t)g->cyc_cap * 4);
    CUDA_TRY(cudaMemcpyAsync(g->d_best_edges, g->d_cyc_edges + (ll)k * g->cyc_cap, (size_t)o.len * 4, cudaMemcpyDeviceToDevice, g->stream));
    return MRC_OK;
}

template <typename S>
static int probe_common(mrc_graph *g, bool exact, const double *lam, const ll *a, const ll *b, int K, double slack,
                        uint32_t flags, int32_t *verdict, int32_t *cyc_len, int32_t *cyc_buf, int64_t cyc_cap,
                        S *sum_cost, S *sum_time, int64_t *passes) {
    ProbeOut res[MRC_KMAX];
    MRC_TRY(run_probe(g, exact, K, lam, a, b, slack, flags, res));
    std::vector<int> verts;
    for (int k = 0; k < K; k++) {
        if (verdict) verdict[k] = res[k].verdict;
        if (passes) passes[k] = res[k].passes;
        if (cyc_len) cyc_len[k] = 0;
        if (sum_cost) sum_cost[k] = 0;
        if (sum_time) sum_time[k] = 0;
        if ((res[k].verdict == MRC_V_NEG || res[k].verdict == MRC_V_TIE) && (flags & MRC_F_NO_CYCLES) && !g->scen_on) {
            if (cyc_len) cyc_len[k] = (int32_t)res[k].len + 1;
            if (sum_cost) sum_cost[k] = exact ? (S)res[k].isumc : (S)res[k].sumc;
            if (sum_time) sum_time[k] = exact ? (S)res[k].isumt : (S)res[k].sumt;
        } else if (res[k].verdict == MRC_V_NEG || res[k].verdict == MRC_V_TIE) {
            double sc = 0, st = 0; ll isc = 0, ist = 0;
            MRC_TRY(fetch_cycle(g, exact, k, res[k], verts, &sc, &st, &isc, &ist));
            if (cyc_len) cyc_len[k] = (int32_t)verts.size();
            if (cyc_buf && cyc_cap > 0)
                for (size_t i = 0; i < verts.size() && (int64_t)i < cyc_cap; i++) cyc_buf[(size_t)k * cyc_cap + i] = verts[i];
            if (sum_cost) sum_cost[k] = exact ? (S)isc : (S)sc;
            if (sum_time) sum_time[k] = exact ? (S)ist : (S)st;
        }
    }
    return MRC_OK;
}

extern "C" int mrc_fetch_cycle(mrc_graph *g, int k, int32_t *cycle, int64_t cycle_cap, int64_t *cycle_len) {
    if (!g || !cycle_len) return set_err(MRC_ERR_ARG, "NULL argument");
    if (k < 0 || k >= g->last_K) return set_err(MRC_ERR_ARG, "candidate index out of range of the last probe");
    const ProbeOut &o = g->last_res[k];
    if (o.verdict != MRC_V_NEG && o.verdict != MRC_V_TIE) { *cycle_len = 0; return MRC_OK; }
    CUDA_TRY(cudaSetDevice(g->device));
    std::vector<int> verts;
    MRC_TRY(fetch_cycle(g, g->last_exact, k, o, verts, nullptr, nullptr, nullptr, nullptr));
    *cycle_len = (int64_t)verts.size();
    if (cycle) for (size_t i = 0; i < verts.size() && (int64_t)i < cycle_cap; i++) cycle[i] = verts[i];
    return MRC_OK;
}

extern "C" int mrc_fetch_cycle_sums(mrc_graph *g, int k, int32_t *cycle, int64_t cycle_cap, int64_t *cycle_len,
                                    double *sum_cost, double *sum_time) {
    if (!g || !cycle_len) return set_err(MRC_ERR_ARG, "NULL argument");
    if (k < 0 || k >= g->last_K) return set_err(MRC_ERR_ARG, "candidate index out of range of the last probe");
    const ProbeOut &o = g->last_res[k];
    if (o.verdict != MRC_V_NEG && o.verdict != MRC_V_TIE) { *cycle_len = 0; return MRC_OK; }
    CUDA_TRY(cudaSetDevice(g->device));
    std::vector<int> verts;
    double sc = 0, st = 0; ll isc = 0, ist = 0;
    MRC_TRY(fetch_cycle(g, g->last_exact, k, o, verts, &sc, &st, &isc, &ist));
    *cycle_len = (int64_t)verts.size();
    if (cycle) for (size_t i = 0; i < verts.size() && (int64_t)i < cycle_cap; i++) cycle[i] = verts[i];
    if (sum_cost) *sum_cost = g->last_exact ? (double)isc : sc;
    if (sum_time) *sum_time = g->last_exact ? (double)ist : st;
    return MRC_OK;
}

extern "C" int mrc_probe_numeric(mrc_graph *g, const double *lam, int K, double slack, uint32_t flags,
                                 int32_t *verdict, int32_t *cyc_len, int32_t *cyc_buf, int64_t cyc_cap,
                                 double *sum_cost, double *sum_time, int64_t *passes) {
    if (!lam) return set_err(MRC_ERR_ARG, "lam is NULL");
    return probe_common<double>(g, false, lam, nullptr, nullptr, K, slack, flags, verdict, cyc_len, cyc_buf, cyc_cap,
                                sum_cost, sum_time, passes);
}
extern "C" int mrc_probe_numeric_scenarios(mrc_graph *g, const double *lam, int K, double slack, uint32_t flags,
                                           const int64_t *patch_edge, const double *dcost, const double *dtime,
                                           int32_t *verdict, int32_t *cyc_len, int32_t *cyc_buf, int64_t cyc_cap,
                                           double *sum_cost, double *sum_time, int64_t *passes) {
    if (!g || !lam || !patch_edge || !dcost || !dtime) return set_err(MRC_ERR_ARG, "NULL argument");
    if (K < 1 || K > MRC_KMAX) return set_err(MRC_ERR_ARG, "K must be in [1,%d]", MRC_KMAX);
    for (int k = 0; k < K; k++) {
        if (patch_edge[k] >= g->m) return set_err(MRC_ERR_ARG, "edge index out of range");
        g->scen_pe[k] = patch_edge[k] < 0 ? -1 : (int)patch_edge[k];
        g->scen_dc[k] = dcost[k]; g->scen_dt[k] = dtime[k];
    }
    g->scen_on = true;
    // pruning by monotonicity is meaningless here: the candidates belong to K different graphs
    const int rc = probe_common<double>(g, false, lam, nullptr, nullptr, K, slack, flags & ~MRC_F_MONOTONE_PRUNE, verdict,
                                        cyc_len, cyc_buf, cyc_cap, sum_cost, sum_time, passes);
    g->scen_on = false;
    return rc;
}

/*
 * sensitivity_analysis (solver.py:1699-1729), the part that dominates it: S single-edge scenarios, for each the question
 * "is there a negative cycle at lam[s] on the graph with edge[s] perturbed?" -- lam[s] being the ratio of the cycle the
 * caller holds under that scenario's weights minus a sliver, so NONNEG means that cycle is still optimal there.
 * One base probe at lam_base on the UNPERTURBED graph leaves converged potentials; every group of K scenarios then
 * starts from them instead of zero.  An edge whose cost went up leaves those potentials feasible (one sweep that
 * lowers nothing); one whose cost went down lowers a few vertices around it.  Either way a scenario costs a sweep or
 * two instead of the 20-50 a cold probe needs, and no host language loop runs between the groups.
 */
extern "C" int mrc_scenarios_certify(mrc_graph *g, int64_t S, const int64_t *edge, const double *dcost, const double *dtime,
                                     const double *lam, double lam_base, double slack, int K, int32_t *verdict,
                                     int64_t *passes) {
    if (!g || S < 0 || (S && (!edge || !dcost || !dtime || !lam || !verdict))) return set_err(MRC_ERR_ARG, "NULL argument");
    if (K < 1 || K > MRC_KMAX) return set_err(MRC_ERR_ARG, "K must be in [1,%d]", MRC_KMAX);
    for (ll s = 0; s < S; s++)
        if (edge[s] >= g->m) return set_err(MRC_ERR_ARG, "edge index out of range");
    if (S == 0) return MRC_OK;
    CUDA_TRY(cudaSetDevice(g->device));
    // base potentials: the fixed point at lam_base on the unperturbed graph (if it does not exist the groups start cold)
    ProbeOut base;
    MRC_TRY(run_probe(g, false, 1, &lam_base, nullptr, nullptr, slack, 0, &base));
    const bool warm = base.verdict == MRC_V_NONNEG;
    if (warm) {
        if (!g->d_base) DALLOC(g->d_base, (size_t)g->n * 8);
        CUDA_TRY(cudaMemcpyAsync(g->d_base, g->d_dist, (size_t)g->n * 8, cudaMemcpyDeviceToDevice, g->stream));   // K = 1 layout: [n][1]
    }
    ProbeOut res[MRC_KMAX];
    for (ll s0 = 0; s0 < S; s0 += K) {
        const int kk = (int)std::min<ll>(K, S - s0);
        for (int k = 0; k < kk; k++) {
            g->scen_pe[k] = edge[s0 + k] < 0 ? -1 : (int)edge[s0 + k];
            g->scen_dc[k] = dcost[s0 + k]; g->scen_dt[k] = dtime[s0 + k];
        }
        g->scen_on = true;
        ProbeRun pr;
        pr.warm_from = warm ? g->d_base : nullptr;
        int rc = pr.begin(g, false, kk, lam + s0, nullptr, nullptr, slack, 0, res);
        while (!rc && pr.active) {
            rc = pr.enqueue_burst();
            if (rc || pr.nb <= 0) break;
            rc = pr.enqueue_readback();
            if (!rc && cudaStreamSynchronize(g->stream) != cudaSuccess) rc = set_err(MRC_ERR_CUDA, "stream synchronize failed");
            if (!rc) rc = pr.collect();
        }
        g->scen_on = false;
        if (rc) return rc;
        pr.finish();
        for (int k = 0; k < kk; k++) { verdict[s0 + k] = res[k].verdict; if (passes) passes[s0 + k] = res[k].passes; }
    }
    return MRC_OK;
}

extern "C" int mrc_probe_exact(mrc_graph *g, const int64_t *a, const int64_t *b, int K, uint32_t flags,
                               int32_t *verdict, int32_t *cyc_len, int32_t *cyc_buf, int64_t cyc_cap,
                               int64_t *sum_cost, int64_t *sum_time, int64_t *passes) {
    if (!a || !b) return set_err(MRC_ERR_ARG, "a/b is NULL");
    return probe_common<int64_t>(g, true, nullptr, (const ll *)a, (const ll *)b, K, 0.0, flags, verdict, cyc_len,
                                 cyc_buf, cyc_cap, sum_cost, sum_time, passes);
}

/*
 * Measurement aid: `passes` relaxation launches for K lambdas from dist == 0, each timed with its own pair
 * of CUDA events on the library stream (no cycle checks, no early exit).  ms_per_pass[passes].
 */
extern "C" int mrc_bench_relax(mrc_graph *g, const double *lam, int K, int passes, float *ms_per_pass) {
    if (!g || !lam || !ms_per_pass || K < 1 || K > MRC_KMAX || passes < 1) return set_err(MRC_ERR_ARG, "bad arguments");
    CUDA_TRY(cudaSetDevice(g->device));
    const int KT = kt_of(K);
    RelaxArgs ra;
    memset(&ra, 0, sizeof ra);
    ra.src = g->d_src; ra.dst = g->d_dst; ra.cost = g->d_cost; ra.time = g->d_time;
    ra.dist = g->d_dist; ra.pred = g->d_pred; ra.m = g->m; ra.slack = 1e-15;
    ra.active = (1u << K) - 1u;
    ra.changed = &g->d_status->changed[0];
    for (int k = 0; k < MRC_KMAX; k++) { ra.lam[k] = k < K ? lam[k] : 0.0; ra.pe[k] = -1; }
    g->pot_valid = false;
    CUDA_TRY(cudaMemsetAsync(g->d_dist, 0, (size_t)g->n * KT * 8, g->stream));
    CUDA_TRY(cudaMemsetAsync(g->d_pred, 0xff, (size_t)g->n * KT * 8, g->stream));
    std::vector<cudaEvent_t> ev((size_t)passes + 1);
    for (auto &e : ev) CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaEventRecord(ev[0], g->stream));
    ra.n = g->n; ra.examined = &g->d_status->examined;
    for (int i = 0; i < passes; i++) {
        MRC_TRY(launch_relax(g, KT, false, ra));
        CUDA_TRY(cudaEventRecord(ev[i + 1], g->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    for (int i = 0; i < passes; i++) CUDA_TRY(cudaEventElapsedTime(&ms_per_pass[i], ev[i], ev[i + 1]));
    for (auto &e : ev) cudaEventDestroy(e);
    return MRC_OK;
}
extern "C" int mrc_set_option(mrc_graph *g, const char *name, int value) {
    if (!g || !name) return set_err(MRC_ERR_ARG, "NULL argument");
    std::string s(name);
    if (s == "stop_on_neg") g->stop_on_neg = std::max(0, value);
    else if (s == "narrow") g->narrow = value != 0;
    else if (s == "compact") g->compact = value != 0;
    else if (s == "compact_min_edges") g->compact_min_edges = std::max(0, value);
    else if (s == "persist") g->persist = std::max(0, std::min(2, value));
    else if (s == "qstart") g->qstart = value != 0;
    else if (s == "dprobe") g->dprobe = value != 0;
    else if (s == "k_after_cycle") g->k_after_cycle = std::max(0, std::min(value, MRC_KMAX));
    else if (s == "dprobe_group") g->dprobe_group = std::max(0, std::min(value, 64));
    else if (s == "qstart_min_edges") g->qstart_min_edges = std::max(0, value);
    else if (s == "check_every") g->check_every = std::max(0, value);
    else if (s == "check_growth") g->check_growth = std::max(0, value);
    else if (s == "pdl") g->pdl = value != 0;
    else if (s == "cert_first") g->cert_first = value != 0;
    else if (s == "tick_max") g->tick_max = std::max(1, std::min(value, MRC_MAX_BURST));
    else if (s == "gate") g->gate = value != 0;
    else if (s == "sparse") g->sparse = value != 0;
    else if (s == "sparse_div") g->sparse_div = std::max(1, value);
    else if (s == "sparse_look") g->sparse_look = std::max(0, value);
    else if (s == "sparse_min_edges") g->sparse_min_edges = std::max(0, value);
    else if (s == "sparse_poison") {   // test hook: garbage in both membership bitmaps, as an interrupted probe may leave it
        CUDA_TRY(cudaSetDevice(g->device));
        MRC_TRY(ensure_sparse(g));
        CUDA_TRY(cudaMemsetAsync(g->d_bm[0], 0xff, g->bm_bytes, g->stream));
        CUDA_TRY(cudaMemsetAsync(g->d_bm[1], 0xff, g->bm_bytes, g->stream));
        g->bm1_dirty = true;   // ... and the library knows no more than that
    }
    else if (s == "prune_above") g->prune_above = value != 0;
    else if (s == "patience") g->patience = value != 0;
    else if (s == "burst_init") g->burst_init = std::max(1, std::min(value, MRC_MAX_BURST));
    else if (s == "burst_max") g->burst_max = std::max(1, std::min(value, MRC_MAX_BURST));
    else return set_err(MRC_ERR_ARG, "unknown option %s", name);
    return MRC_OK;
}

// ------------------------------------------------------------------------------- numeric lambda search
// Where the first round looks.  The reference starts from [min c/t - 1, max c/t + 1] (solver.py:1019-1023) and halves it;
// uniform candidates over that bracket spend the first rounds far from lambda*: the minimum cycle ratio is a weighted mean of
// edge ratios and, with many cycles to choose from, sits in the low tail of their distribution (C3: the 2.5 % quantile of
// c/t, while the bracket's midpoint is near the 60 % quantile).  So the first round's candidates are low QUANTILES of a
// sample of the edge ratios -- uniform in probability mass, not in value.  Any lambda in (lo, hi) is a legitimate candidate;
// the rounds after the first are the usual ones (mrc_plan), so a graph whose optimum is not in the tail loses one round.
__global__ void ratio_sample_kernel(const double *cost, const double *time, ll m, int S, double *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S) {
        const ll e = (ll)(((double)i + 0.5) * (double)m / (double)S);
        out[i] = __ddiv_rn(cost[e], time[e]);
    }
}
static int ensure_qsample(mrc_graph *g) {
    if (!g->qsample.empty()) return MRC_OK;
    const int S = (int)std::min<ll>(g->m, 2048);
    double *d = nullptr;
    DALLOC(d, (size_t)S * 8);
    std::vector<double> h((size_t)S);
    ratio_sample_kernel<<<(S + 255) / 256, 256, 0, g->stream>>>(g->d_cost, g->d_time, g->m, S, d);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h.data(), d, (size_t)S * 8, cudaMemcpyDeviceToHost, g->stream));
    cudaFreeAsync(d, g->stream);
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    { g->stats.other_launches++; g->stats.kernel_launches++; }
    g->stats.d2h_bytes += (int64_t)S * 8;
    std::sort(h.begin(), h.end());
    g->qsample.swap(h);
    return MRC_OK;
}
// quantile levels of a first round of K candidates: a geometric ladder from q_lo to q_hi (K = 4: 1 %, 2.9 %, 8.5 %, 25 %)
static void first_round_levels(int K, double *q) {
    if (K == 1) { q[0] = 0.05; return; }
    const double q_lo = K == 2 ? 0.02 : K <= 4 ? 0.01 : K <= 8 ? 0.005 : 0.002, q_hi = K == 2 ? 0.10 : K <= 4 ? 0.25 : 0.5;
    for (int j = 0; j < K; j++) q[j] = q_lo * std::pow(q_hi / q_lo, (double)j / (double)(K - 1));
}
extern "C" int mrc_first_round_levels(int K, double *q) {
    if (K < 1 || K > 4096 || !q) return set_err(MRC_ERR_ARG, "bad arguments");
    first_round_levels(K, q);
    return MRC_OK;
}
// K candidates for the first round of a search over (lo, hi) without a known cycle; false = use the uniform plan
static bool plan_first_round(mrc_graph *g, double lo, double hi, int K, double *lam) {
    if (!g->qstart || g->m < g->qstart_min_edges || K < 1) return false;
    if (ensure_qsample(g) != MRC_OK) return false;
    std::vector<double> q((size_t)K);
    first_round_levels(K, q.data());
    const size_t S = g->qsample.size();
    size_t prev = 0;
    for (int j = 0; j < K; j++) {
        size_t idx = std::min(S - 1, (size_t)(q[j] * (double)(S - 1)));
        if (j > 0 && idx <= prev) idx = prev + 1;   // a long ladder (many ranks) is denser than the sample at its low end
        if (idx >= S) return false;
        prev = idx;
        lam[j] = g->qsample[idx];
        if (!(lam[j] > lo && lam[j] < hi) || (j > 0 && !(lam[j] > lam[j - 1]))) return false;   // ties / outside the bracket
    }
    return true;
}

extern "C" int mrc_plan_candidates(double lo, double hi, int have_cycle, double tol, int K_total, double *lam_out) {
    if (!lam_out || K_total < 1) return set_err(MRC_ERR_ARG, "bad plan arguments");
    mrc_plan(lo, hi, have_cycle, tol, K_total, lam_out);
    return MRC_OK;
}

extern "C" int mrc_reduce_round(int K_total, const double *lam, const int32_t *verdict, const double *cyc_ratio,
                                double *lo, double *hi, int *have_cycle, int *best_idx) {
    if (!lam || !verdict || !cyc_ratio || !lo || !hi || !have_cycle || !best_idx)
        return set_err(MRC_ERR_ARG, "NULL argument");
    *best_idx = mrc_reduce(K_total, lam, verdict, cyc_ratio, lo, hi, have_cycle);
    return MRC_OK;
}

extern "C" int mrc_solve_numeric(mrc_graph *g, double lo, double hi, int hi_is_cycle, int max_iter, double tol,
                                 double slack, int K, double max_seconds, double *ratio, double *sum_cost, double *sum_time,
                                 int32_t *cycle, int64_t cycle_cap, int64_t *cycle_len, int32_t *iterations,
                                 double *final_lo, double *final_hi) {
    if (!g) return set_err(MRC_ERR_ARG, "graph is NULL");
    if (K < 1 || K > MRC_KMAX) return set_err(MRC_ERR_ARG, "K must be in [1,%d]", MRC_KMAX);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<int> best_cycle, verts;
    double best_sc = 0, best_st = 0;
    bool have_best = false, best_pending = false;
    ProbeOut best_o;
    int have_cycle = hi_is_cycle ? 1 : 0, iters = 0, it = -1;
    double lam[MRC_KMAX], cr[MRC_KMAX];
    int32_t vd[MRC_KMAX];
    ProbeOut res[MRC_KMAX];
    const bool trace = getenv("MRC_B200_TRACE") != nullptr;
    for (int i = 0; i < max_iter; i++) {
        it = i;
        if (max_seconds > 0) {   // solver.py:1032-1041
            const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (el > max_seconds) return set_err(MRC_ERR_TIMEOUT, "Numeric solve exceeded time limit");
        }
        iters++;
        g->stats.rounds++;
        const int Kr = (have_cycle && g->k_after_cycle > 0) ? std::min(K, g->k_after_cycle) : K;   // candidates of this round
        if (!(i == 0 && !have_cycle && plan_first_round(g, lo, hi, Kr, lam))) MRC_TRY(mrc_plan_candidates(lo, hi, have_cycle, tol, Kr, lam));
        // Option "stop_on_neg": intermediate rounds only need a better cycle, so a round may end at the first
        // check (from pass 6 on) that verified one, leaving the slow non-negative candidates undecided; the round in
        // which nothing is negative still runs to convergence and proves the lower end.
        MRC_TRY(run_probe(g, false, Kr, lam, nullptr, nullptr, slack,
                          MRC_F_MONOTONE_PRUNE | ((g->stop_on_neg && Kr >= g->stop_on_neg) ? MRC_F_STOP_ON_NEG : 0u) |
                              (g->patience ? MRC_F_PATIENCE : 0u) | ((have_cycle && g->cert_first) ? MRC_F_CERT_FIRST : 0u), res));
        for (int k = 0; k < Kr; k++) {
            vd[k] = res[k].verdict;
            cr[k] = (vd[k] == MRC_V_NEG || vd[k] == MRC_V_TIE) ? res[k].sumc / res[k].sumt : NAN;
        }
        if (trace) {
            fprintf(stderr, "  round %d:", i);
            for (int k = 0; k < Kr; k++) fprintf(stderr, " [%.12g v%d p%lld len %d r=%.12g]", lam[k], vd[k], (long long)res[k].passes, res[k].len, cr[k]);
            fprintf(stderr, "\n");
        }
        int best = -1;
        MRC_TRY(mrc_reduce_round(Kr, lam, vd, cr, &lo, &hi, &have_cycle, &best));
        if (best >= 0) {
            MRC_TRY(stash_cycle(g, best, res[best]));
            best_o = res[best];
            have_best = best_pending = true;
        }
        if (hi - lo <= tol) break;   // solver.py:1075-1076
    }
    if (best_pending) {
        MRC_TRY(fetch_cycle(g, false, 0, best_o, verts, &best_sc, &best_st, nullptr, nullptr, g->d_best_edges));
        best_cycle = verts;
    }
    if (!have_best && hi_is_cycle) {   // the caller's cycle (ratio == the initial hi) was not beaten
        if (iterations) *iterations = iters;
        if (final_lo) *final_lo = lo;
        if (final_hi) *final_hi = hi;
        if (cycle_len) *cycle_len = 0;
        if (ratio) *ratio = hi;
        if (max_iter > 0 && it == max_iter - 1 && hi - lo > tol)
            return set_err(MRC_ERR_CONVERGENCE, "Numeric solver failed to converge");
        return MRC_OK;
    }
    if (!have_best) {   // solver.py:1079-1085: one more try at the upper bound
        MRC_TRY(run_probe(g, false, 1, &hi, nullptr, nullptr, slack, 0, res));
        if (res[0].verdict == MRC_V_NEG) {
            MRC_TRY(fetch_cycle(g, false, 0, res[0], verts, &best_sc, &best_st, nullptr, nullptr));
            best_cycle = verts;
            have_best = true;
        }
    }
    if (iterations) *iterations = iters;
    if (final_lo) *final_lo = lo;
    if (final_hi) *final_hi = hi;
    if (!have_best) return set_err(MRC_ERR_NO_CYCLE, "No negative cycle found");   // solver.py:1087-1090
    if (cycle_len) *cycle_len = (int64_t)best_cycle.size();
    if (cycle)
        for (size_t i = 0; i < best_cycle.size() && (int64_t)i < cycle_cap; i++) cycle[i] = best_cycle[i];
    if (sum_cost) *sum_cost = best_sc;
    if (sum_time) *sum_time = best_st;
    if (ratio) *ratio = best_sc / best_st;   // solver.py:1293
    if (max_iter > 0 && it == max_iter - 1 && hi - lo > tol)   // solver.py:1102-1109
        return set_err(MRC_ERR_CONVERGENCE, "Numeric solver failed to converge");
    return MRC_OK;
}

// ------------------------------------------------------------------------------------------ exact mode
static int ensure_potentials(mrc_graph *g, ll a, ll b, bool *ok) {
    if (g->pot_valid && g->pot_a == a && g->pot_b == b) { *ok = true; return MRC_OK; }
    ProbeOut r;
    MRC_TRY(run_probe(g, true, 1, nullptr, &a, &b, 0.0, 0, &r));
    *ok = (r.verdict == MRC_V_NONNEG);
    return MRC_OK;
}

extern "C" int mrc_potentials_exact(mrc_graph *g, int64_t a, int64_t b, uint8_t *ok, int64_t *dist) {
    if (!g || !ok) return set_err(MRC_ERR_ARG, "NULL argument");
    bool good = false;
    MRC_TRY(ensure_potentials(g, a, b, &good));
    *ok = good;
    if (good && dist) {
        CUDA_TRY(cudaMemcpyAsync(dist, g->d_dist, (size_t)g->n * 8, cudaMemcpyDeviceToHost, g->stream));
        CUDA_TRY(cudaStreamSynchronize(g->stream));
        g->stats.d2h_bytes += g->n * 8;
    }
    return MRC_OK;
}

static int tight_mask_device(mrc_graph *g, ll a, ll b, bool *ok) {
    MRC_TRY(ensure_potentials(g, a, b, ok));
    if (!*ok) return MRC_OK;
    const int blocks = (int)std::min<ll>((g->m + 255) / 256, (ll)g->sms * 8);
    tight_mask_kernel<<<blocks, 256, 0, g->stream>>>(g->d_src, g->d_dst, g->d_icost, g->d_itime, (const ll *)g->d_dist,
                                                     a, b, g->m, g->d_mask);
    CUDA_TRY(cudaGetLastError());
    { g->stats.other_launches++; g->stats.kernel_launches++; }
    return MRC_OK;
}

extern "C" int mrc_tight_mask_exact(mrc_graph *g, int64_t a, int64_t b, uint8_t *ok, uint8_t *mask) {
    if (!g || !ok || !mask) return set_err(MRC_ERR_ARG, "NULL argument");
    bool good = false;
    MRC_TRY(tight_mask_device(g, a, b, &good));
    *ok = good;
    if (good) {
        CUDA_TRY(cudaMemcpyAsync(mask, g->d_mask, (size_t)g->m, cudaMemcpyDeviceToHost, g->stream));
        CUDA_TRY(cudaStreamSynchronize(g->stream));
        g->stats.d2h_bytes += g->m;
    }
    return MRC_OK;
}

/*
 * Host part of _find_zero_cycle_exact (solver.py:884-920): the reference's recursive three-colour DFS
 * over the tight subgraph, replayed with an explicit stack.  Adjacency of u lists tight edges in
 * ascending dst-sorted edge index (:886-888); roots 0..n-1 in order (:914); the first GRAY neighbour
 * closes the cycle, reported as [child-of-v, ..., u, v] (:903-910).
 */
static bool zero_cycle_dfs(const mrc_graph *g, const std::vector<uint8_t> &mask, std::vector<int> &cyc) {
    const ll n = g->n, m = g->m;
    std::vector<ll> adj_start((size_t)n + 1, 0);
    ll tight = 0;
    for (ll e = 0; e < m; e++) if (mask[e]) { adj_start[g->h_src[e] + 1]++; tight++; }
    if (!tight) return false;   // :881-882
    for (ll v = 0; v < n; v++) adj_start[v + 1] += adj_start[v];
    std::vector<int> adj((size_t)tight);
    std::vector<ll> cursor(adj_start.begin(), adj_start.end() - 1);
    for (ll e = 0; e < m; e++) if (mask[e]) adj[cursor[g->h_src[e]]++] = g->h_dst[e];
    std::copy(adj_start.begin(), adj_start.end() - 1, cursor.begin());
    std::vector<int8_t> color((size_t)n, 0);
    std::vector<int> parent((size_t)n, -1), stack;
    stack.reserve(1024);
    for (ll s = 0; s < n; s++) {
        if (color[s]) continue;
        stack.clear();
        stack.push_back((int)s); color[s] = 1;
        while (!stack.empty()) {
            const int u = stack.back();
            if (cursor[u] < adj_start[u + 1]) {
                const int v = adj[cursor[u]++];
                if (color[v] == 0) { parent[v] = u; color[v] = 1; stack.push_back(v); }
                else if (color[v] == 1) {
                    cyc.clear();
                    cyc.push_back(v);
                    for (int x = u; x != v && x != -1; x = parent[x]) cyc.push_back(x);
                    std::reverse(cyc.begin(), cyc.end());
                    return true;
                }
            } else { color[u] = 2; stack.pop_back(); }
        }
    }
    return false;
}

// _compute_cycle_weights_exact (solver.py:971-995): first inserted (u,v) edge == first hit in v's segment
static int cycle_weights_exact(const mrc_graph *g, const std::vector<int> &cyc, ll *sc, ll *st) {
    ll c = 0, t = 0;
    const size_t L = cyc.size();
    for (size_t i = 0; i < L; i++) {
        const int u = cyc[i], v = cyc[(i + 1) % L];
        ll k = g->h_starts[v];
        const ll e = g->h_starts[v + 1];
        for (; k < e; k++) if (g->h_src[k] == u) break;
        if (k == e) return set_err(MRC_ERR_EDGE_MISSING, "Edge %d -> %d not found in cycle extraction", u, v);
        c += g->h_icost[k]; t += g->h_itime[k];
    }
    *sc = c; *st = t;
    return MRC_OK;
}

static int zero_cycle(mrc_graph *g, ll a, ll b, bool *found, std::vector<int> &cyc, ll *sc, ll *st) {
    *found = false;
    bool ok = false;
    MRC_TRY(tight_mask_device(g, a, b, &ok));
    if (!ok) return MRC_OK;   // :866-867
    std::vector<uint8_t> mask((size_t)g->m);
    CUDA_TRY(cudaMemcpyAsync(mask.data(), g->d_mask, (size_t)g->m, cudaMemcpyDeviceToHost, g->stream));
    CUDA_TRY(cudaStreamSynchronize(g->stream));
    g->stats.d2h_bytes += g->m;
    if (!zero_cycle_dfs(g, mask, cyc)) return MRC_OK;
    MRC_TRY(cycle_weights_exact(g, cyc, sc, st));
    *found = true;
    return MRC_OK;
}

extern "C" int mrc_zero_cycle_exact(mrc_graph *g, int64_t a, int64_t b, uint8_t *found, int32_t *cycle,
                                    int64_t cycle_cap, int64_t *cycle_len, int64_t *sum_cost, int64_t *sum_time) {
    if (!g || !found) return set_err(MRC_ERR_ARG, "NULL argument");
    if (!g->has_int) return set_err(MRC_ERR_NOT_INTEGER, "Exact mode requires integer weights");
    std::vector<int> cyc;
    bool f = false;
    ll sc = 0, st = 0;
    MRC_TRY(zero_cycle(g, a, b, &f, cyc, &sc, &st));
    *found = f;
    if (cycle_len) *cycle_len = f ? (int64_t)cyc.size() : 0;
    if (f && cycle) for (size_t i = 0; i < cyc.size() && (int64_t)i < cycle_cap; i++) cycle[i] = cyc[i];
    if (sum_cost) *sum_cost = sc;
    if (sum_time) *sum_time = st;
    return MRC_OK;
}

static ll gcd_ll(ll a, ll b) {
    if (a < 0) a = -a;
    if (b < 0) b = -b;
    while (b) { ll t = a % b; a = b; b = t; }
    return a;
}
static ll floordiv_ll(ll a, ll b) {
    ll q = a / b, r = a % b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

/*
 * _stern_brocot_search (solver.py:734-797).  `neg(a,b)` answers _has_negative_cycle_exact and `zero(a,b)`
 * answers _find_zero_cycle_exact's "is not None"; strategy 0 backs them with GPU probes, strategy 1 with
 * exact comparisons against the lambda* found by Newton steps.
 */
extern "C" int mrc_solve_exact(mrc_graph *g, int64_t max_den, int64_t max_steps, int strategy, int32_t *cycle,
                               int64_t cycle_cap, int64_t *cycle_len, int64_t *sum_cost, int64_t *sum_time,
                               int64_t *a_out, int64_t *b_out, int64_t *iterations, int64_t *probes_ab,
                               int64_t probes_cap, int64_t *nprobes) {
    if (!g) return set_err(MRC_ERR_ARG, "graph is NULL");
    if (!g->has_int) return set_err(MRC_ERR_NOT_INTEGER, "Exact mode requires integer weights");
    if (max_den < 0) max_den = g->n * std::max<ll>(1, g->max_it);   // :746-748
    if (max_steps < 0) max_steps = 2 * max_den + 10;                 // :750-751
    ll aL = (ll)std::floor(g->min_ratio_i) - 1, bL = 1;              // :754-756
    ll aR = (ll)std::ceil(g->max_ratio_i) + 1, bR = 1;
    ll np = 0;
    if (nprobes) *nprobes = 0;

    // lambda* = ps/qs, known in advance only with strategy 1
    ll ps = 0, qs = 1;
    if (strategy == 1) {
        ProbeOut r;
        MRC_TRY(run_probe(g, true, 1, nullptr, &aL, &bL, 0.0, 0, &r));
        if (r.verdict != MRC_V_NONNEG) {
            if (probes_ab && probes_cap > 0) { probes_ab[0] = aL; probes_ab[1] = bL; }
            if (nprobes) *nprobes = 1;
            return set_err(MRC_ERR_LOWER_NEG, "Lower bound has negative cycle");
        }
        ll a = aR, b = bR;
        bool first = true;
        for (;;) {
            g->stats.rounds++;
            MRC_TRY(run_probe(g, true, 1, nullptr, &a, &b, 0.0, 0, &r));
            if (r.verdict == MRC_V_NONNEG) {
                if (first) {
                    if (probes_ab && probes_cap > 1) { probes_ab[0] = aL; probes_ab[1] = bL; probes_ab[2] = aR; probes_ab[3] = bR; }
                    if (nprobes) *nprobes = 2;
                    return set_err(MRC_ERR_UPPER_NONNEG, "Upper bound has no negative cycle");
                }
                ps = a; qs = b;
                break;
            }
            first = false;
            if (r.verdict != MRC_V_NEG) { strategy = 0; break; }   // pass limit without an extracted cycle: probe every step instead
            const ll gg = gcd_ll(r.isumc, r.isumt);
            a = r.isumc / gg; b = r.isumt / gg;   // Newton step: lambda <- ratio of the extracted cycle
        }
    }
    auto record = [&](ll a, ll b) {
        if (probes_ab && np < probes_cap) { probes_ab[2 * np] = a; probes_ab[2 * np + 1] = b; }
        np++;
        if (nprobes) *nprobes = np;
    };
    auto neg = [&](ll a, ll b, bool *out) -> int {
        record(a, b);
        if (strategy == 1) { *out = (i128)a * qs > (i128)ps * b; return MRC_OK; }
        ProbeOut r;
        MRC_TRY(run_probe(g, true, 1, nullptr, &a, &b, 0.0, 0, &r));
        *out = (r.verdict == MRC_V_NEG || r.verdict == MRC_V_NEG_NOCYCLE);
        return MRC_OK;
    };
    std::vector<int> cyc;
    ll sc = 0, st = 0;
    auto zero = [&](ll a, ll b, bool *found) -> int {
        if (strategy == 1 && (i128)a * qs != (i128)ps * b) { *found = false; return MRC_OK; }
        if (strategy == 1) { a = ps; b = qs; }   // same tight subgraph at any positive scale of (a,b)
        return zero_cycle(g, a, b, found, cyc, &sc, &st);
    };
    auto finish = [&](ll a, ll b, ll iters) -> int {
        const ll gg = gcd_ll(a, b);
        if (a_out) *a_out = floordiv_ll(a, gg);
        if (b_out) *b_out = floordiv_ll(b, gg);
        if (iterations) *iterations = iters;
        if (cycle_len) *cycle_len = (int64_t)cyc.size();
        if (cycle) for (size_t i = 0; i < cyc.size() && (int64_t)i < cycle_cap; i++) cycle[i] = cyc[i];
        if (sum_cost) *sum_cost = sc;
        if (sum_time) *sum_time = st;
        return MRC_OK;
    };
    bool hn = false, found = false;
    MRC_TRY(neg(aL, bL, &hn));
    if (hn) return set_err(MRC_ERR_LOWER_NEG, "Lower bound has negative cycle");          // :759-760
    MRC_TRY(neg(aR, bR, &hn));
    if (!hn) return set_err(MRC_ERR_UPPER_NONNEG, "Upper bound has no negative cycle");   // :761-762
    ll iters = 0;
    for (ll step = 0; step < max_steps; step++) {   // :766
        iters++;
        if (strategy == 0) g->stats.rounds++;
        ll aM = aL + aR, bM = bL + bR;              // :768
        if (bM > max_den) {                         // :770-780
            MRC_TRY(zero(aL, bL, &found));
            if (found) return finish(aL, bL, -1);
            ll k = std::max<ll>(1, floordiv_ll(max_den - bL, bR));
            if ((i128)k * (aR < 0 ? -(i128)aR : (i128)aR) > ((i128)1 << 62) || (i128)k * bR > ((i128)1 << 62))
                return set_err(MRC_ERR_OVERFLOW, "Stern-Brocot jump leaves int64");
            aM = aL + k * aR; bM = bL + k * bR;
        }
        MRC_TRY(neg(aM, bM, &hn));                  // :782
        if (hn) { aR = aM; bR = bM; continue; }     // :783-784
        MRC_TRY(zero(aM, bM, &found));              // :787
        if (found) return finish(aM, bM, iters);    // :788-793
        aL = aM; bL = bM;                           // :795
    }
    return set_err(MRC_ERR_SB_STEPS, "Stern-Brocot search did not converge");   // :797
}

// ------------------------------------------------------------------------------------ small-graph batches
static const void *batch_fn_of(int KT) {
    switch (KT) {
        case 1: return (const void *)batch_solve_kernel<1>;
        case 2: return (const void *)batch_solve_kernel<2>;
        case 4: return (const void *)batch_solve_kernel<4>;
        default: return (const void *)batch_solve_kernel<8>;
    }
}

/*
 * B independent small graphs, one CTA each, the whole lambda search on the device (mrc_batch.cuh).
 * Replaces a Python loop of B solve() calls (BASELINE config 4).  Every graph is destination-sorted with
 * LOCAL vertex ids; edge_off[B+1] delimits the graphs in the concatenated arrays.
 */
extern "C" int mrc_batch_solve_numeric(int64_t B, const int32_t *nverts, const int64_t *edge_off, const int32_t *src,
                                       const int32_t *dst, const double *cost, const double *time, int max_iter,
                                       double tol, double slack, int K, int device, double *ratio, double *sum_cost,
                                       double *sum_time, int32_t *cycle_len, int32_t *cycle_buf, int64_t cycle_cap,
                                       int32_t *iterations, int32_t *status, mrc_stats *stats_out) {
    if (B <= 0 || !nverts || !edge_off || !src || !dst || !cost || !time || !ratio || !sum_cost || !sum_time ||
        !cycle_len || !cycle_buf || !iterations || !status)
        return set_err(MRC_ERR_ARG, "NULL or empty batch argument");
    if (K < 1 || K > MRC_KMAX) return set_err(MRC_ERR_ARG, "K must be in [1,%d]", MRC_KMAX);
    int ndev = 0;
    MRC_TRY(mrc_device_count(&ndev));
    if (device < 0 || device >= ndev) return set_err(MRC_ERR_ARG, "device %d out of range", device);
    const int KT = kt_of(K);
    int n_max = 1, m_max = 1;
    const ll M = edge_off[B];
    for (ll g = 0; g < B; g++) {
        const ll mg = edge_off[g + 1] - edge_off[g];
        if (nverts[g] <= 0 || nverts[g] > 65535 || mg <= 0 || mg > 0x7fffffff)
            return set_err(MRC_ERR_ARG, "graph %lld: batch solver needs 1 <= n <= 65535 and m >= 1", g);
        n_max = std::max(n_max, (int)nverts[g]); m_max = std::max<ll>(m_max, mg);
        for (ll e = edge_off[g]; e < edge_off[g + 1]; e++) {
            if ((unsigned)src[e] >= (unsigned)nverts[g] || (unsigned)dst[e] >= (unsigned)nverts[g])
                return set_err(MRC_ERR_ARG, "graph %lld: endpoint outside [0,n)", g);
            if (!(time[e] > 0)) return set_err(MRC_ERR_ARG, "graph %lld: time must be strictly positive", g);
        }
    }
    n_max = (n_max + 3) & ~3; m_max = (m_max + 3) & ~3;
    const size_t smem = mrc_batch_smem(n_max, m_max, KT);
    CUDA_TRY(cudaSetDevice(device));
    int smem_cap = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    if (smem > (size_t)smem_cap - 4096)
        return set_err(MRC_ERR_ARG, "largest graph of the batch (n=%d, m=%d) needs %zu B of shared memory (> %d): use mrc_graph_create",
                       n_max, m_max, smem, smem_cap);
    cudaStream_t st;
    CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    BatchArgs a;
    memset(&a, 0, sizeof a);
    ll *d_off = nullptr;
    int *d_nv = nullptr, *d_src = nullptr, *d_dst = nullptr, *d_len = nullptr, *d_buf = nullptr, *d_it = nullptr, *d_status = nullptr;
    double *d_c = nullptr, *d_t = nullptr, *d_r = nullptr, *d_sc = nullptr, *d_st = nullptr;
    ull *d_cnt = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = [&]() -> int {
        CUDA_TRY(cudaMallocAsync((void **)&d_off, (B + 1) * 8, st)); CUDA_TRY(cudaMallocAsync((void **)&d_nv, B * 4, st));
        CUDA_TRY(cudaMallocAsync((void **)&d_src, M * 4, st)); CUDA_TRY(cudaMallocAsync((void **)&d_dst, M * 4, st));
        CUDA_TRY(cudaMallocAsync((void **)&d_c, M * 8, st)); CUDA_TRY(cudaMallocAsync((void **)&d_t, M * 8, st));
        CUDA_TRY(cudaMallocAsync((void **)&d_r, B * 8, st)); CUDA_TRY(cudaMallocAsync((void **)&d_sc, B * 8, st));
        CUDA_TRY(cudaMallocAsync((void **)&d_st, B * 8, st)); CUDA_TRY(cudaMallocAsync((void **)&d_len, B * 4, st));
        CUDA_TRY(cudaMallocAsync((void **)&d_buf, B * cycle_cap * 4, st)); CUDA_TRY(cudaMallocAsync((void **)&d_it, B * 4, st));
        CUDA_TRY(cudaMallocAsync((void **)&d_status, B * 4, st)); CUDA_TRY(cudaMallocAsync((void **)&d_cnt, 8, st));
        CUDA_TRY(cudaMemcpyAsync(d_off, edge_off, (B + 1) * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_nv, nverts, B * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_src, src, M * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_dst, dst, M * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_c, cost, M * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_t, time, M * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemsetAsync(d_cnt, 0, 8, st));
        CUDA_TRY(cudaMemsetAsync(d_buf, 0, B * cycle_cap * 4, st));
        a.B = (int)B; a.edge_off = d_off; a.nverts = d_nv; a.src = d_src; a.dst = d_dst; a.cost = d_c; a.time = d_t;
        a.tol = tol; a.slack = slack; a.max_iter = max_iter; a.K = K;
        a.ratio = d_r; a.sum_cost = d_sc; a.sum_time = d_st; a.cyc_len = d_len; a.cyc_buf = d_buf; a.cyc_cap = cycle_cap;
        a.iters = d_it; a.status = d_status; a.relax_count = d_cnt; a.n_max = n_max; a.m_max = m_max;
        CUDA_TRY(cudaFuncSetAttribute(batch_fn_of(KT), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
        void *args[] = {(void *)&a};
        CUDA_TRY(cudaEventRecord(e0, st));
        CUDA_TRY(cudaLaunchKernel(batch_fn_of(KT), dim3((unsigned)B), dim3(MRC_BATCH_THREADS), args, smem, st));
        CUDA_TRY(cudaEventRecord(e1, st));
        CUDA_TRY(cudaMemcpyAsync(ratio, d_r, B * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(sum_cost, d_sc, B * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(sum_time, d_st, B * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(cycle_len, d_len, B * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(cycle_buf, d_buf, B * cycle_cap * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(iterations, d_it, B * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(status, d_status, B * 4, cudaMemcpyDeviceToHost, st));
        ull cnt = 0;
        CUDA_TRY(cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (stats_out) {
            memset(stats_out, 0, sizeof *stats_out);
            stats_out->relax_launches = 1; stats_out->relax_edge_visits = (int64_t)cnt; stats_out->relax_ms = ms;
            stats_out->probes = B * K; stats_out->h2d_bytes = M * 24 + B * 12; stats_out->d2h_bytes = B * (36 + cycle_cap * 4);
        }
        return MRC_OK;
    }();
    void *ptrs[] = {d_off, d_nv, d_src, d_dst, d_c, d_t, d_r, d_sc, d_st, d_len, d_buf, d_it, d_status, d_cnt};
    for (void *q : ptrs) if (q) cudaFreeAsync(q, st);
    cudaStreamSynchronize(st);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    return rc;
}

#include "mrc_comm.cuh"
