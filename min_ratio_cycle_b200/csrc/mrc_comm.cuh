// mrc_comm.cuh -- several GPUs of one box behind the C ABI (SURVEY 8b "mrc_comm_init / mrc_comm_destroy", 8e):
// NCCL communicators (loaded at run time), graph replication by broadcast, and the lambda search with the
// candidates sharded over the ranks.  Included at the end of mrc_b200.cu (it drives ProbeRun).
//
// The reference has no counterpart (single process, single thread; its search loop is solver.py:1031-1076).
// What crosses NVLink: per tick 32 B per candidate (all-gather), at the end one cycle (broadcast), once per graph
// the edge arrays (broadcast).  The data path of a probe -- 24 B/edge/pass -- never leaves its GPU.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <thread>
#include <unistd.h>

// ------------------------------------------------------------------------------------------------ NCCL at run time
// libmrc_b200.so does not link NCCL: a process that already carries one (PyTorch's bundled copy has the same soname)
// must not get a second copy, and single-GPU users need none at all.
struct NcclApi {
    std::once_flag once;
    void *h = nullptr;
    std::string err;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static int nccl_api(NcclApi **out) {
    NcclApi &n = g_nccl;
    std::call_once(n.once, [&]() {
        const char *env = getenv("MRC_B200_NCCL");
        if (env && *env) n.h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        if (!n.h) n.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy this process already uses, if any
        if (!n.h) n.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!n.h) n.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!n.h) { n.err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return; }
        auto sym = [&](const char *name) { void *p = dlsym(n.h, name); if (!p && n.err.empty()) n.err = std::string("NCCL symbol missing: ") + name; return p; };
        n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
        n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
        n.CommInitAll = (decltype(n.CommInitAll))sym("ncclCommInitAll");
        n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
        n.AllGather = (decltype(n.AllGather))sym("ncclAllGather");
        n.Broadcast = (decltype(n.Broadcast))sym("ncclBroadcast");
        n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
        n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
        n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
    });
    if (!n.err.empty()) return set_err(MRC_ERR_COMM, "%s", n.err.c_str());
    *out = &n;
    return MRC_OK;
}
#define NCCL_TRY(api, x)                                                                                   \
    do {                                                                                                   \
        ncclResult_t r_ = (x);                                                                             \
        if (r_ != ncclSuccess)                                                                             \
            return set_err(MRC_ERR_COMM, "%s failed: %s (%s:%d)", #x, (api)->GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------------ communicators
#define MRC_REC 4   /* doubles per candidate record: lambda, state (-1 running, else MRC_V_*), cycle ratio, cycle length */
#define MRC_HDR 4   /* doubles per rank header: passes done, next check, burst, flags (bit0: no probe left to start) */

struct mrc_comm {
    int nranks = 1, rank = 0, device = 0;
    ncclComm_t comm = nullptr;
    std::vector<mrc_comm *> group;   // single-process form: one rank communicator per device (this object is only the handle)
    // per-tick exchange buffers (sized for K_local = MRC_KMAX)
    double *d_send = nullptr, *d_recv = nullptr, *h_send = nullptr, *h_recv = nullptr;
    int *d_cyc = nullptr;            // broadcast buffer of the winning cycle
    size_t cyc_cap = 0;
    // peer window of the distributed probe: replicas of dist / pred and a mailbox in plain cudaMalloc memory, mapped into
    // every other rank (CUDA IPC between processes, peer access inside one process)
    ll win_n = 0;                    // vertices the window was made for (0 = none yet)
    bool win_ok = false;             // every rank mapped every peer
    double *win_dist = nullptr;
    ull *win_pred = nullptr, *win_mbox = nullptr;
    void *win_peer[MRC_KMAX][3] = {};   // [rank] -> dist, pred, mbox of that rank (mine: the local pointers)
    bool win_ipc[MRC_KMAX] = {};        // opened with cudaIpcOpenMemHandle (to be closed)
    ull xseq = 1ull << 32;           // step number of the next distributed probe's first message
};
static size_t comm_rank_doubles() { return MRC_HDR + (size_t)MRC_KMAX * MRC_REC; }

static void comm_window_free(mrc_comm *c) {
    for (int r = 0; r < MRC_KMAX; r++) {
        if (c->win_ipc[r]) for (int j = 0; j < 3; j++) if (c->win_peer[r][j]) cudaIpcCloseMemHandle(c->win_peer[r][j]);
        c->win_ipc[r] = false;
        for (int j = 0; j < 3; j++) c->win_peer[r][j] = nullptr;
    }
    if (c->win_dist) cudaFree(c->win_dist);
    if (c->win_pred) cudaFree(c->win_pred);
    if (c->win_mbox) cudaFree(c->win_mbox);
    c->win_dist = nullptr; c->win_pred = nullptr; c->win_mbox = nullptr;
    c->win_n = 0; c->win_ok = false;
    cudaGetLastError();
}

static int comm_buffers(mrc_comm *c) {
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t per = comm_rank_doubles() * 8;
    CUDA_TRY(cudaMalloc((void **)&c->d_send, per));
    CUDA_TRY(cudaMalloc((void **)&c->d_recv, per * c->nranks));
    CUDA_TRY(cudaMallocHost((void **)&c->h_send, per));
    CUDA_TRY(cudaMallocHost((void **)&c->h_recv, per * c->nranks));
    return MRC_OK;
}

extern "C" int mrc_comm_unique_id(uint8_t *id) {
    if (!id) return set_err(MRC_ERR_ARG, "id is NULL");
    NcclApi *api;
    MRC_TRY(nccl_api(&api));
    static_assert(sizeof(ncclUniqueId) == MRC_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    NCCL_TRY(api, api->GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return MRC_OK;
}

extern "C" int mrc_comm_destroy(mrc_comm *c) {
    if (!c) return MRC_OK;
    for (mrc_comm *r : c->group) mrc_comm_destroy(r);
    if (c->comm || c->d_send) cudaSetDevice(c->device);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    if (c->d_send) cudaFree(c->d_send);
    if (c->d_recv) cudaFree(c->d_recv);
    if (c->d_cyc) cudaFree(c->d_cyc);
    comm_window_free(c);
    if (c->h_send) cudaFreeHost(c->h_send);
    if (c->h_recv) cudaFreeHost(c->h_recv);
    delete c;
    return MRC_OK;
}

extern "C" int mrc_comm_init_rank(int nranks, int rank, const uint8_t *id, int device, mrc_comm **out) {
    if (!out || !id) return set_err(MRC_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (nranks < 1 || rank < 0 || rank >= nranks) return set_err(MRC_ERR_ARG, "bad rank %d of %d", rank, nranks);
    int ndev = 0;
    MRC_TRY(mrc_device_count(&ndev));
    if (device < 0 || device >= ndev) return set_err(MRC_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    NcclApi *api;
    MRC_TRY(nccl_api(&api));
    CUDA_TRY(cudaSetDevice(device));
    mrc_comm *c = new mrc_comm();
    c->nranks = nranks; c->rank = rank; c->device = device;
    int rc = [&]() -> int {
        ncclUniqueId u;
        memcpy(&u, id, sizeof u);
        NCCL_TRY(api, api->CommInitRank(&c->comm, nranks, u, rank));
        return comm_buffers(c);
    }();
    if (rc) { mrc_comm_destroy(c); return rc; }
    *out = c;
    return MRC_OK;
}

extern "C" int mrc_comm_init(int ndev, const int *devs, mrc_comm **out) {
    if (!out || !devs) return set_err(MRC_ERR_ARG, "NULL argument");
    *out = nullptr;
    int have = 0;
    MRC_TRY(mrc_device_count(&have));
    if (ndev < 1 || ndev > have) return set_err(MRC_ERR_ARG, "%d devices requested, %d present", ndev, have);
    for (int i = 0; i < ndev; i++) {
        if (devs[i] < 0 || devs[i] >= have) return set_err(MRC_ERR_ARG, "device %d out of range", devs[i]);
        for (int j = 0; j < i; j++) if (devs[j] == devs[i]) return set_err(MRC_ERR_ARG, "device %d listed twice", devs[i]);
    }
    NcclApi *api;
    MRC_TRY(nccl_api(&api));
    mrc_comm *c = new mrc_comm();
    c->nranks = ndev; c->rank = -1; c->device = devs[0];
    int rc = [&]() -> int {
        std::vector<ncclComm_t> comms((size_t)ndev, nullptr);
        NCCL_TRY(api, api->CommInitAll(comms.data(), ndev, devs));
        for (int i = 0; i < ndev; i++) {
            mrc_comm *r = new mrc_comm();
            r->nranks = ndev; r->rank = i; r->device = devs[i]; r->comm = comms[i];
            c->group.push_back(r);
        }
        for (mrc_comm *r : c->group) MRC_TRY(comm_buffers(r));
        return MRC_OK;
    }();
    if (rc) { mrc_comm_destroy(c); return rc; }
    *out = c;
    return MRC_OK;
}

extern "C" int mrc_comm_info(const mrc_comm *c, int *nranks, int *rank, int *device) {
    if (!c) return set_err(MRC_ERR_ARG, "comm is NULL");
    if (nranks) *nranks = c->nranks;
    if (rank) *rank = c->rank;
    if (device) *device = c->device;
    return MRC_OK;
}

// ------------------------------------------------------------------------------------------------ replicas
static int replicate_rank(mrc_comm *c, int root, int64_t n, int64_t m, const int32_t *src, const int32_t *dst,
                          const double *cost, const double *time, mrc_graph **out) {
    *out = nullptr;
    if (!c || c->rank < 0) return set_err(MRC_ERR_ARG, "a rank communicator is required");
    if (root < 0 || root >= c->nranks) return set_err(MRC_ERR_ARG, "root %d out of range", root);
    const bool is_root = c->rank == root;
    if (is_root && (!src || !dst || !cost || !time)) return set_err(MRC_ERR_ARG, "the root rank must pass the edge arrays");
    NcclApi *api;
    MRC_TRY(nccl_api(&api));
    mrc_graph *g = nullptr;
    MRC_TRY(graph_new(n, m, c->device, false, &g));
    int rc = [&]() -> int {
        // root: the four H2D copies run on a second stream, each broadcast starts when its array has landed -- the
        // NVLink broadcast of one array overlaps the PCIe copy of the next
        cudaStream_t up = nullptr;
        cudaEvent_t landed[4] = {nullptr, nullptr, nullptr, nullptr};
        void *dptr[4] = {g->d_src, g->d_dst, g->d_cost, g->d_time};
        const void *hptr[4] = {src, dst, cost, time};
        const size_t bytes[4] = {(size_t)m * 4, (size_t)m * 4, (size_t)m * 8, (size_t)m * 8};
        int rc2 = [&]() -> int {
            if (is_root) {
                CUDA_TRY(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
                for (int i = 0; i < 4; i++) {
                    CUDA_TRY(cudaEventCreateWithFlags(&landed[i], cudaEventDisableTiming));
                    CUDA_TRY(cudaMemcpyAsync(dptr[i], hptr[i], bytes[i], cudaMemcpyHostToDevice, up));
                    CUDA_TRY(cudaEventRecord(landed[i], up));
                }
                g->stats.h2d_bytes += (int64_t)m * 24;
            }
            for (int i = 0; i < 4; i++) {
                if (is_root) CUDA_TRY(cudaStreamWaitEvent(g->stream, landed[i], 0));
                if (c->nranks > 1) NCCL_TRY(api, api->Broadcast(dptr[i], dptr[i], bytes[i], ncclChar, root, c->comm, g->stream));
            }
            return graph_validate(g);   // synchronises g->stream
        }();
        for (auto &e : landed) if (e) cudaEventDestroy(e);
        if (up) { cudaStreamSynchronize(up); cudaStreamDestroy(up); }
        return rc2;
    }();
    if (rc) { mrc_graph_destroy(g); return rc; }
    *out = g;
    return MRC_OK;
}

extern "C" int mrc_graph_create_replicated(mrc_comm *c, int root, int64_t n, int64_t m, const int32_t *src,
                                           const int32_t *dst, const double *cost, const double *time, mrc_graph **out) {
    if (!out) return set_err(MRC_ERR_ARG, "out is NULL");
    return replicate_rank(c, root, n, m, src, dst, cost, time, out);
}

// ------------------------------------------------------------------------------------------------ distributed probe
// Once the search holds a cycle, what is left is a chain of single-candidate probes (the certifier of the best cycle, again
// after every better cycle): depth-bound, ~21 sweeps of the whole edge stream at C3, and more candidates in parallel do not
// shorten it.  The ranks therefore run such a probe TOGETHER: rank r sweeps the in-edges of its slice of the vertices
// (1/W of the stream), every value it lowers goes straight into the peers' replicas of dist over NVLink (an 8-byte atomic
// min per peer; after the first passes a pass lowers a few thousand vertices), and one word per rank and pass through
// mailboxes in peer memory tells everybody whether anybody moved.  No host, no NCCL inside the probe: it is one persistent
// kernel per rank (probe_kernel<1>, mrc_probe.cuh).  NCCL carries only the IPC handles, once.
struct WinHello {              // what a rank tells the others about its window (one all-gather)
    long long pid;
    int device, ok;
    char bus[16];          // PCI bus id of the rank's GPU: the ranks of a distributed probe must sit on DIFFERENT GPUs (their kernels wait
                           // for one another; two of them time-sliced on one GPU would never meet)
    void *raw[3];
    cudaIpcMemHandle_t h[3];
};
static_assert(sizeof(WinHello) <= (MRC_HDR + MRC_KMAX * MRC_REC) * 8, "WinHello must fit the per-rank exchange buffer");

static int comm_allgather_bytes(mrc_comm *c, cudaStream_t st, const void *mine, void *all, size_t bytes) {
    NcclApi *api;
    MRC_TRY(nccl_api(&api));
    const size_t per = comm_rank_doubles();
    CUDA_TRY(cudaMemcpyAsync(c->d_send, mine, bytes, cudaMemcpyHostToDevice, st));
    NCCL_TRY(api, api->AllGather(c->d_send, c->d_recv, per, ncclDouble, c->comm, st));
    std::vector<char> tmp(per * 8 * (size_t)c->nranks);
    CUDA_TRY(cudaMemcpyAsync(tmp.data(), c->d_recv, tmp.size(), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int r = 0; r < c->nranks; r++) memcpy((char *)all + (size_t)r * bytes, tmp.data() + (size_t)r * per * 8, bytes);
    return MRC_OK;
}

// Collective over the ranks of c: (re)build the peer window for graphs of n vertices.  Failure to map a peer is not an
// error of the search: the window is simply marked unusable on EVERY rank and the probes stay on one GPU.
static int comm_window(mrc_comm *c, mrc_graph *g) {
    const ll n = g->n;
    if (c->win_n == n) return MRC_OK;
    const int W = c->nranks, me = c->rank;
    comm_window_free(c);
    c->win_n = n;
    if (W < 2 || W > MRC_KMAX) return MRC_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    WinHello mine;
    memset(&mine, 0, sizeof mine);
    mine.pid = (long long)getpid(); mine.device = c->device; mine.ok = 1;
    if (cudaDeviceGetPCIBusId(mine.bus, sizeof mine.bus, c->device) != cudaSuccess) mine.ok = 0;
    const size_t mbox_bytes = (size_t)W * MRC_MBOX_SLOTS * 8;
    if (cudaMalloc((void **)&c->win_dist, (size_t)n * 8) != cudaSuccess || cudaMalloc((void **)&c->win_pred, (size_t)n * 8) != cudaSuccess ||
        cudaMalloc((void **)&c->win_mbox, mbox_bytes) != cudaSuccess) mine.ok = 0;
    if (mine.ok) {
        if (cudaMemset(c->win_mbox, 0, mbox_bytes) != cudaSuccess) mine.ok = 0;
        mine.raw[0] = c->win_dist; mine.raw[1] = c->win_pred; mine.raw[2] = c->win_mbox;
        for (int j = 0; j < 3 && mine.ok; j++) if (cudaIpcGetMemHandle(&mine.h[j], mine.raw[j]) != cudaSuccess) mine.ok = 0;
    }
    cudaGetLastError();
    std::vector<WinHello> all((size_t)W);
    MRC_TRY(comm_allgather_bytes(c, g->stream, &mine, all.data(), sizeof(WinHello)));
    int ok = mine.ok;
    for (int r = 0; r < W && ok; r++) ok &= all[r].ok;
    for (int r = 0; r < W && ok; r++)
        for (int q = 0; q < r; q++) if (!strncmp(all[r].bus, all[q].bus, sizeof mine.bus)) ok = 0;   // two ranks on one GPU
    for (int r = 0; r < W && ok; r++) {
        if (r == me) { for (int j = 0; j < 3; j++) c->win_peer[r][j] = mine.raw[j]; continue; }
        if (all[r].pid == mine.pid) {   // same process (mrc_comm_init): plain peer access
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, c->device, all[r].device) != cudaSuccess || !can) { ok = 0; break; }
            const cudaError_t e = cudaDeviceEnablePeerAccess(all[r].device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { ok = 0; break; }
            cudaGetLastError();
            for (int j = 0; j < 3; j++) c->win_peer[r][j] = all[r].raw[j];
        } else {
            for (int j = 0; j < 3; j++)
                if (cudaIpcOpenMemHandle(&c->win_peer[r][j], all[r].h[j], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { c->win_peer[r][j] = nullptr; ok = 0; }
            c->win_ipc[r] = true;
            if (!ok) break;
        }
    }
    cudaGetLastError();
    // every rank must come to the same conclusion
    WinHello verdict;
    memset(&verdict, 0, sizeof verdict);
    verdict.ok = ok;
    MRC_TRY(comm_allgather_bytes(c, g->stream, &verdict, all.data(), sizeof(WinHello)));
    for (int r = 0; r < W; r++) ok &= all[r].ok;
    c->win_ok = ok != 0;
    if (getenv("MRC_B200_TRACE") && me == 0) fprintf(stderr, "  peer window for n=%lld over %d ranks: %s\n", (long long)n, W, c->win_ok ? "mapped" : "unavailable");
    return MRC_OK;
}

// this rank's slice of a destination-sorted edge stream: vertices [v_lo, v_hi) and their in-edges [e_lo, e_hi), cut at vertex
// boundaries near m * r / W (every rank derives the same cuts from the same replica)
static int graph_slice(mrc_graph *g, int W, int me) {
    if (g->slice_W == W && g->slice_me == me) return MRC_OK;
    if (!g->d_starts) {
        DALLOC(g->d_starts, ((size_t)g->n + 1) * 4);
        in_offsets_kernel<<<(int)std::min<ll>((g->n + 256) / 256, (ll)g->sms * 8), 256, 0, g->stream>>>(g->d_dst, g->m, g->n, g->d_starts);
        CUDA_TRY(cudaGetLastError());
        { g->stats.other_launches++; g->stats.kernel_launches++; }
    }
    int v[2] = {0, (int)g->n}, e[2] = {0, (int)g->m};
    for (int j = 0; j < 2; j++) {
        const int r = me + j;
        if (r == 0 || r == W) continue;
        CUDA_TRY(cudaMemcpyAsync(&v[j], g->d_dst + g->m * r / W, 4, cudaMemcpyDeviceToHost, g->stream));
        CUDA_TRY(cudaStreamSynchronize(g->stream));
        CUDA_TRY(cudaMemcpyAsync(&e[j], g->d_starts + v[j], 4, cudaMemcpyDeviceToHost, g->stream));
        CUDA_TRY(cudaStreamSynchronize(g->stream));
    }
    g->slice_v[0] = v[0]; g->slice_v[1] = v[1]; g->slice_e[0] = e[0]; g->slice_e[1] = e[1];
    g->slice_W = W; g->slice_me = me;
    return MRC_OK;
}

// One single-candidate numeric probe run by all ranks of c together (collective; every rank gets the same verdict and, for a
// negative one, the same cycle in its own cycle buffer).
static int run_probe_dist(mrc_comm *c, mrc_graph *g, double lam, double slack, ProbeOut *res) {
    const int W = c->nranks, me = c->rank;
    MRC_TRY(graph_slice(g, W, me));
    ProbeRun pr;
    pr.device_init = true;
    MRC_TRY(pr.begin(g, false, 1, &lam, nullptr, nullptr, slack, 0, res));
    pr.use_sparse = false; pr.may_compact = false;
    ProbeArgs pa;
    MRC_TRY(probe_args(g, pr, 0, nullptr, pa));
    const ll e_lo = g->slice_e[0], e_hi = g->slice_e[1];
    pa.r.src += e_lo; pa.r.dst += e_lo; pa.r.cost += e_lo; pa.r.time += e_lo; pa.r.m = e_hi - e_lo;
    pa.r.dist = c->win_dist; pa.r.pred = c->win_pred;
    pa.c.dist = c->win_dist; pa.c.pred = c->win_pred;
    pa.dG = g->dprobe_group > 0 ? g->dprobe_group : std::max(2, W / 2);   // an exchange costs ~16 us, a sweep of the slice 26 / 13 / 7 us at 2 / 4 / 8 ranks
    pa.dW = W; pa.dme = me; pa.e_base = (unsigned)e_lo; pa.v_lo = g->slice_v[0]; pa.v_hi = g->slice_v[1];
    int np = 0;
    for (int r = 0; r < W; r++) {
        pa.peer_mbox[r] = (ull *)c->win_peer[r][2];
        if (r == me) continue;
        pa.pp.peer[np] = (ull *)c->win_peer[r][0];
        pa.peer_pred[np] = (ull *)c->win_peer[r][1];
        np++;
    }
    pa.pp.n_peers = np;
    pa.mbox = c->win_mbox;
    // A check costs a gather of pred over NVLink on top of the pointer doubling (~85 us) while a group of sweeps of the slices costs
    // 50 us at 8 ranks: first check after two groups, then at doubling distances (8, 16, 32, ... sweeps at 8 ranks; 2, 6, 12, ... at 2)
    pa.first_check = std::max(g->burst_init, 2 * (g->dprobe_group > 0 ? g->dprobe_group : std::max(2, W / 2)));
    pa.check_growth = 100;
    pa.seq0 = c->xseq;
    c->xseq += 1ull << 32;
    MRC_TRY(probe_launch_collect(g, pr, pa, res));
    if (getenv("MRC_B200_TRACE")) {
        double mn = 0; ll deep = 0;
        if (atoi(getenv("MRC_B200_TRACE")) >= 2) {   // (a 8 MB copy per probe: distorts the timeline)
            std::vector<double> h((size_t)g->n);
            CUDA_TRY(cudaMemcpy(h.data(), c->win_dist, (size_t)g->n * 8, cudaMemcpyDeviceToHost));
            for (double x : h) { mn = std::min(mn, x); deep += x < -1e4; }
        }
        const ProbeDev &o = *g->h_probe;
        fprintf(stderr, "    rank %d: slice e[%lld,%lld) v[%lld,%lld) G %d sweeps %u checks %u cuts %u verdict %d passes %lld min dist %.6g (%lld below -1e4) "
                        "sweep %.1f us/step; kernel %.3f ms = sweeps %.3f + checks %.3f + other %.3f (exchanges, waits)\n",
                me, (long long)e_lo, (long long)e_hi, (long long)pa.v_lo, (long long)pa.v_hi, pa.dG, o.n_sweep[0], o.n_check, o.cuts, res[0].verdict,
                (long long)res[0].passes, mn, (long long)deep, o.n_sweep[0] ? 1e-3 * (double)o.ns_sweep[0] / ((double)o.n_sweep[0] / pa.dG) : 0.0,
                1e-6 * (double)o.ns_total, 1e-6 * (double)o.ns_sweep[0], 1e-6 * (double)o.ns_check,
                1e-6 * ((double)o.ns_total - (double)o.ns_sweep[0] - (double)o.ns_check));
    }
    if (res[0].verdict < 0) return set_err(MRC_ERR_COMM, "distributed probe: a peer did not answer (rank %d of %d)", me, W);
    return MRC_OK;
}

// ------------------------------------------------------------------------------------------------ sharded search
// Bracket logic of a tick, shared by every rank (and testable without a device): see mrc_shard_fold in the header.
static int shard_fold(int nrec, const double *rec, double *lo, double *hi, int *have_cycle, uint8_t *implied) {
    double nlo = *lo, nhi = *hi;
    int hc = *have_cycle, best = -1;
    for (int i = 0; i < nrec; i++) {
        const double lam = rec[MRC_REC * i], ratio = rec[MRC_REC * i + 2];
        const int state = (int)rec[MRC_REC * i + 1];
        switch (state) {
            case MRC_V_NEG:
            case MRC_V_TIE:
                if (hc ? (ratio < nhi) : (ratio <= nhi)) { nhi = ratio; hc = 1; best = i; }
                if (state == MRC_V_TIE && lam > nlo) nlo = lam;
                break;
            case MRC_V_NEG_NOCYCLE:   // reference: hi = mid without a cycle (solver.py:1057-1060)
                if (lam < nhi) { nhi = lam; hc = 0; best = -1; }
                break;
            case MRC_V_NONNEG:
                if (lam > nlo) nlo = lam;
                break;
            default: break;
        }
    }
    if (nlo > nhi) nlo = nhi;   // rounding noise near lambda*
    if (implied)
        for (int i = 0; i < nrec; i++) {
            const double lam = rec[MRC_REC * i];
            const int state = (int)rec[MRC_REC * i + 1];
            // a cycle of ratio hi is negative at every lambda > hi; at lambda <= lo relaxation converges
            implied[i] = state < 0 && ((hc ? lam > nhi : lam >= nhi) || lam <= nlo);
        }
    *lo = nlo; *hi = nhi; *have_cycle = hc;
    return best;
}

extern "C" int mrc_shard_fold(int nrec, const double *rec, double *lo, double *hi, int *have_cycle, int *best_idx,
                              uint8_t *implied) {
    if (nrec < 0 || (nrec && !rec) || !lo || !hi || !have_cycle || !best_idx) return set_err(MRC_ERR_ARG, "NULL argument");
    *best_idx = shard_fold(nrec, rec, lo, hi, have_cycle, implied);
    return MRC_OK;
}

// Records of the burst that just ran, written on the device from the status block the relaxation launches and the
// cycle check filled in (the same decisions ProbeRun::collect takes on the host one stream-sync later): no host hop
// between the check and the all-gather.  One thread per candidate of the probe.
__global__ void pack_records_kernel(const DevStatus *st, int K, unsigned active, int nb, int do_check, int ck, int limit_hit,
                                    double *send /* header + K records */) {
    const int k = threadIdx.x;
    if (k >= K || !((active >> k) & 1u)) return;
    double *r = send + MRC_HDR + MRC_REC * k;
    int conv = 0;
    for (int i = 0; i < nb; i++) conv |= !((st->changed[i] >> k) & 1u);
    const int ko = ck >= 0 ? 0 : k;
    if (conv) { r[1] = MRC_V_NONNEG; return; }
    if (do_check && ((st->cyc_mask >> ko) & 1u) && st->out[ko].found) {
        const CycleOut &c = st->out[ko];
        r[1] = c.found == 1 ? MRC_V_NEG : MRC_V_TIE;
        r[2] = c.sumc / c.sumt;
        r[3] = (double)c.len;
        return;
    }
    if (limit_hit) r[1] = MRC_V_NEG_NOCYCLE;
}

struct ShardOut {
    int rc = MRC_OK;
    std::string err;
    std::vector<int> cycle;
    double sc = 0, st = 0, lo = 0, hi = 0;
    int iterations = 0, ticks = 0;
    bool have_best = false, converged = false;
};

// One rank of the sharded search (one process per GPU calls it directly; the single-process form runs it on one
// host thread per device).
static int shard_solve_rank(mrc_comm *c, mrc_graph *g, double lo, double hi, int max_iter, double tol, double slack,
                            int K, double max_seconds, ShardOut *out) {
    if (!c || c->rank < 0 || !g) return set_err(MRC_ERR_ARG, "a rank communicator and a graph are required");
    if (K < 1 || K > MRC_KMAX) return set_err(MRC_ERR_ARG, "K_local must be in [1,%d]", MRC_KMAX);
    if (g->device != c->device) return set_err(MRC_ERR_ARG, "graph lives on device %d, communicator on %d", g->device, c->device);
    NcclApi *api = nullptr;
    if (c->nranks > 1) MRC_TRY(nccl_api(&api));
    CUDA_TRY(cudaSetDevice(g->device));
    const auto t0 = std::chrono::steady_clock::now();
    const bool trace = getenv("MRC_B200_TRACE") != nullptr;
    const int W = c->nranks, me = c->rank, KT_all = W * K;
    const size_t per = comm_rank_doubles();
    std::vector<double> plan((size_t)KT_all), rec((size_t)KT_all * MRC_REC);
    std::vector<uint8_t> implied((size_t)KT_all);
    ProbeRun pr;
    ProbeOut res[MRC_KMAX];
    bool running = false;        // a probe of mine is in flight
    int have_cycle = 0, rounds = 0, ticks = 0;
    int best_owner = -1;
    ProbeOut best_o;             // valid on the owner
    bool exhausted = false;      // max_iter probes started: this rank only takes part in the exchange from now on
    int cap = g->burst_init;     // passes of the next tick, agreed by all ranks
    double plan_hi = hi;         // upper end of the bracket when my running probe was planned
    bool go_dist = false;        // a cycle is in hand and the peer window is mapped: the rest of the search is distributed probes
    if (W > 1 && g->dprobe) MRC_TRY(comm_window(c, g));   // collective the first time a graph of this size is seen
    for (int k = 0; k < K; k++) res[k] = ProbeOut();
    for (size_t i = 0; i < per; i++) c->h_send[i] = 0.0;
    for (int k = 0; k < K; k++) c->h_send[MRC_HDR + MRC_REC * k + 1] = (double)MRC_V_ABANDONED;
    for (;;) {
        if (max_seconds > 0) {   // solver.py:1032-1041; every rank runs the same clock test on (nearly) the same tick,
            const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();   // so it goes into the header
            if (el > max_seconds) c->h_send[3] = 2.0;
        }
        // ---- start a probe of K fresh candidates from the current bracket
        if (!running && !exhausted) {
            if (rounds >= max_iter) exhausted = true;
            else {
                if (!(rounds == 0 && !have_cycle && plan_first_round(g, lo, hi, KT_all, plan.data())))   // same sample, same plan on every replica
                    mrc_plan(lo, hi, have_cycle, tol, KT_all, plan.data());
                const double *mine = plan.data() + (size_t)me * K;
                int Kp = K;
                // With a cycle in hand the plan's last candidate is its certifier -- the one probe that can end the
                // search.  The last rank runs it ALONE (1-wide kernel on a compact column: 52 instead of 60 us per sweep,
                // cheaper first sweeps) instead of waiting for three companions that can only move lo.
                if (W > 1 && me == W - 1 && have_cycle) { mine = plan.data() + (size_t)KT_all - 1; Kp = 1; }
                MRC_TRY(pr.begin(g, false, Kp, mine, nullptr, nullptr, slack, MRC_F_MONOTONE_PRUNE, res));
                for (int k = 0; k < K; k++) {
                    double *r = c->h_send + MRC_HDR + MRC_REC * k;
                    r[0] = k < Kp ? mine[k] : 0.0; r[1] = k < Kp ? -1.0 : (double)MRC_V_ABANDONED; r[2] = NAN; r[3] = 0.0;
                }
                plan_hi = hi;
                running = true;
                rounds++;
                g->stats.rounds++;
            }
        }
        if (!running)   // nothing of mine is in flight: no record of this rank may read "running"
            for (int k = 0; k < K; k++) {
                double &state = c->h_send[MRC_HDR + MRC_REC * k + 1];
                if (state < 0) state = (double)MRC_V_ABANDONED;
            }
        // ---- one tick: a burst of at most `cap` passes (+ cycle check when due), records, all-gather, read-back
        if (running) {
            MRC_TRY(pr.enqueue_burst(cap));
            if (pr.nb <= 0) running = false;   // pass limit: the candidates were marked NEG_NOCYCLE by the probe
        }
        c->h_send[0] = running ? (double)(pr.passes + pr.nb) : 0.0;
        c->h_send[1] = running ? (double)pr.next_check : 0.0;
        c->h_send[2] = running ? (double)std::min(pr.burst * 2, g->burst_max) : 0.0;
        if (c->h_send[3] != 2.0) c->h_send[3] = exhausted ? 1.0 : 0.0;
        // records of candidates the host already knows the fate of travel in h_send; the kernel overwrites the ones
        // this burst decides
        CUDA_TRY(cudaMemcpyAsync(c->d_send, c->h_send, per * 8, cudaMemcpyHostToDevice, g->stream));
        if (running) {
            const int limit_hit = pr.passes + pr.nb >= pr.max_passes;
            pack_records_kernel<<<1, 32, 0, g->stream>>>(g->d_status, pr.K, pr.active, pr.nb, pr.do_check ? 1 : 0, pr.ck, limit_hit, c->d_send);
            CUDA_TRY(cudaGetLastError());
        }
        if (W > 1) NCCL_TRY(api, api->AllGather(c->d_send, c->d_recv, per, ncclDouble, c->comm, g->stream));
        else CUDA_TRY(cudaMemcpyAsync(c->d_recv, c->d_send, per * 8, cudaMemcpyDeviceToDevice, g->stream));
        CUDA_TRY(cudaMemcpyAsync(c->h_recv, c->d_recv, per * 8 * W, cudaMemcpyDeviceToHost, g->stream));
        if (running) MRC_TRY(pr.enqueue_readback());
        CUDA_TRY(cudaStreamSynchronize(g->stream));
        ticks++;
        g->stats.other_launches += running ? 1 : 0;
        g->stats.kernel_launches += running ? 1 : 0;
        g->stats.d2h_bytes += (int64_t)per * 8 * W;
        if (running) MRC_TRY(pr.collect());
        // keep my host copy of the records in step with what the device sent (decided candidates stay decided)
        for (int k = 0; k < K; k++)
            for (int j = 1; j < MRC_REC; j++) c->h_send[MRC_HDR + MRC_REC * k + j] = c->h_recv[(size_t)me * per + MRC_HDR + MRC_REC * k + j];
        // ---- fold every rank's records into the bracket (identical arithmetic on identical data on every rank)
        bool timeout = false, all_exhausted = true;
        for (int r = 0; r < W; r++) {
            const double *hdr = c->h_recv + (size_t)r * per;
            timeout |= hdr[3] == 2.0;
            all_exhausted &= hdr[3] == 1.0;
            for (int k = 0; k < K; k++)
                for (int j = 0; j < MRC_REC; j++) rec[((size_t)r * K + k) * MRC_REC + j] = hdr[MRC_HDR + MRC_REC * k + j];
        }
        const int best = shard_fold(KT_all, rec.data(), &lo, &hi, &have_cycle, implied.data());
        if (best >= 0) {
            best_owner = best / K;
            if (best_owner == me) {   // set the edge list aside before my next probe reuses the buffer
                const int k = best % K;
                MRC_TRY(stash_cycle(g, k, res[k]));
                best_o = res[k];
            }
        } else if (!have_cycle) best_owner = -1;
        if (timeout) return set_err(MRC_ERR_TIMEOUT, "Numeric solve exceeded time limit");
        // ---- my candidates whose verdict the bracket now implies
        if (running) {
            // a better cycle moved the upper end: the certifier the last rank is running (or the probe that kept it
            // from starting one) is obsolete -- it starts over from the new bracket on the next tick
            if (W > 1 && me == W - 1 && have_cycle && hi < plan_hi)
                for (int k = 0; k < pr.K; k++)
                    if ((pr.active >> k) & 1u) { pr.drop(k, MRC_V_ABANDONED); c->h_send[MRC_HDR + MRC_REC * k + 1] = (double)MRC_V_ABANDONED; }
            for (int k = 0; k < pr.K; k++)
                if (implied[(size_t)me * K + k]) {
                    const double lam = rec[((size_t)me * K + k) * MRC_REC];
                    pr.drop(k, lam <= lo ? MRC_V_SKIPPED_NONNEG : MRC_V_SKIPPED_NEG);
                    c->h_send[MRC_HDR + MRC_REC * k + 1] = (double)(lam <= lo ? MRC_V_SKIPPED_NONNEG : MRC_V_SKIPPED_NEG);
                }
            if (!pr.active) { pr.finish(); running = false; }
        }
        if (trace && me == 0) {
            const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            fprintf(stderr, "  tick %d t=%.3f ms cap %d lo %.9g hi %.12g hc %d best_owner %d |", ticks, el * 1e3, cap, lo, hi, have_cycle, best_owner);
            for (int i = 0; i < KT_all; i++)
                fprintf(stderr, " %s%.6g:%d%s", i % K == 0 ? "[" : "", rec[(size_t)i * MRC_REC], (int)rec[(size_t)i * MRC_REC + 1], implied[i] ? "i" : "");
            fprintf(stderr, "\n");
        }
        if (hi - lo <= tol) { out->converged = true; break; }   // solver.py:1075-1076, on the tick every rank sees it
        // ... not before the second tick: the candidates next to lambda* close their cycle a few sweeps after the far ones, and
        // starting the Newton chain from a far cycle costs two or three extra probes (measured on 2 GPUs: 2.8 -> 3.4 ms)
        // With 16 or more candidates on the quantile ladder the ones just above lambda* sit within a few per cent of it: their
        // cycle is almost always the optimal one, and if it is not a distributed Newton step costs less than a tick.
        if (have_cycle && W > 1 && c->win_ok && g->dprobe && ticks >= (KT_all >= 16 ? 1 : 2)) {   // same records, same fold: every rank leaves the tick loop on this tick
            if (running) { for (int k = 0; k < pr.K; k++) pr.drop(k, MRC_V_ABANDONED); pr.finish(); running = false; }
            go_dist = true;
            break;
        }
        // ---- length of the next tick.  Every rank runs at the pace of the slowest one (the all-gather waits), so the
        // tick follows the probe that is on the critical path: the LAST rank holds the top of the plan -- with a cycle
        // in hand its last candidate is the certifier of that cycle, the one probe that can end the search -- and its
        // burst schedule (2, 4, 8, .. passes, restarting whenever a better cycle moves the certifier) sets the length.
        // Measured on 8 GPUs (C3): with the shortest wish of ANY rank every tick carried somebody's two crowded first
        // sweeps + cycle check (~210 us) and the certifier's 21 sweeps took 2.4 ms instead of 1.2.
        int next_cap = g->burst_max;
        bool anyone = false;
        int lead_wish = 0;
        for (int r = 0; r < W; r++) {
            const double *hdr = c->h_recv + (size_t)r * per;
            bool r_running = false;
            for (int k = 0; k < K; k++) {
                const size_t i = (size_t)r * K + k;
                r_running |= (int)rec[i * MRC_REC + 1] < 0 && !implied[i];
            }
            int wish;
            if (r_running) wish = (int)std::max(1.0, std::min(hdr[2], hdr[1] - hdr[0]));
            else if (hdr[3] == 1.0) continue;   // nothing left to start on that rank
            else wish = g->burst_init;
            anyone = true;
            next_cap = std::min(next_cap, wish);
            if (r == W - 1) lead_wish = wish;
        }
        if (have_cycle && lead_wish > 0) next_cap = lead_wish;
        cap = std::max(1, std::min(next_cap, g->tick_max));
        if (!anyone && all_exhausted) break;   // max_iter probes everywhere and the bracket is still open
    }
    // ---- Newton steps on the best cycle's ratio, every probe run by all ranks together (run_probe_dist): the certifier of the
    // current cycle either converges (the cycle is optimal to within tol) or returns a better cycle, whose certifier is next
    while (go_dist && rounds < max_iter + 64) {
        double cert;
        mrc_plan(lo, hi, 1, tol, 1, &cert);
        ProbeOut r;
        MRC_TRY(run_probe_dist(c, g, cert, slack, &r));
        rounds++; ticks++;
        g->stats.rounds++;
        const double ratio = (r.verdict == MRC_V_NEG || r.verdict == MRC_V_TIE) ? r.sumc / r.sumt : NAN;
        if (trace && me == 0) {
            const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            fprintf(stderr, "  distributed probe %d t=%.3f ms lam %.12g verdict %d passes %lld ratio %.12g\n", rounds, el * 1e3, cert, r.verdict, (long long)r.passes, ratio);
        }
        if (r.verdict == MRC_V_NONNEG || r.verdict == MRC_V_TIE) lo = std::max(lo, cert);
        if ((r.verdict == MRC_V_NEG || r.verdict == MRC_V_TIE) && ratio < hi) {
            hi = ratio;
            best_owner = 0;   // every rank extracted the same cycle; rank 0 keeps it
            if (me == 0) { MRC_TRY(stash_cycle(g, 0, r)); best_o = r; }
        } else if (r.verdict == MRC_V_NEG) {
            lo = std::max(lo, std::min(hi, cert));   // a negative cycle that is not better than the one in hand: rounding noise at the optimum
        } else if (r.verdict == MRC_V_NEG_NOCYCLE) {
            hi = cert; have_cycle = 0; best_owner = -1;   // pass limit without a cycle to show (reference: hi = mid, solver.py:1057-1060)
            break;
        }
        if (lo > hi) lo = hi;
        if (hi - lo <= tol) { out->converged = true; break; }
    }
    out->iterations = rounds; out->ticks = ticks; out->lo = lo; out->hi = hi;
    // ---- the winning cycle: its owner brings it to the host and broadcasts it
    if (best_owner < 0) {
        // no cycle seen at all: the reference tries once more at the upper bound
        // (solver.py:1079-1085); every rank runs the probe, rank 0's cycle is the one
        ProbeOut r1;
        MRC_TRY(run_probe(g, false, 1, &hi, nullptr, nullptr, slack, 0, &r1));
        if (r1.verdict == MRC_V_NEG) { best_owner = 0; if (me == 0) { MRC_TRY(stash_cycle(g, 0, r1)); best_o = r1; } }
        // verdicts can differ between ranks only through racing predecessor stores, never in kind; still, rank 0 decides
        double flag = best_owner == 0 ? 1.0 : 0.0;
        if (W > 1) {
            CUDA_TRY(cudaMemcpyAsync(c->d_send, &flag, 8, cudaMemcpyHostToDevice, g->stream));
            NCCL_TRY(api, api->Broadcast(c->d_send, c->d_send, 1, ncclDouble, 0, c->comm, g->stream));
            CUDA_TRY(cudaMemcpyAsync(&flag, c->d_send, 8, cudaMemcpyDeviceToHost, g->stream));
            CUDA_TRY(cudaStreamSynchronize(g->stream));
        }
        best_owner = flag == 1.0 ? 0 : -1;
    }
    if (best_owner < 0) { out->have_best = false; return MRC_OK; }
    std::vector<int> verts;
    double msg[3] = {0, 0, 0};
    if (me == best_owner) {
        MRC_TRY(fetch_cycle(g, false, 0, best_o, verts, &msg[0], &msg[1], nullptr, nullptr, g->d_best_edges));
        msg[2] = (double)verts.size();
    }
    if (W > 1) {
        // ONE broadcast carries the sums, the length and (up to HEAD vertices of) the cycle; longer cycles need a second one
        const size_t HEAD = 2048, head_ints = 6 + HEAD;   // 3 doubles = 6 ints
        if (c->cyc_cap < head_ints) {
            if (c->d_cyc) CUDA_TRY(cudaFree(c->d_cyc));
            c->cyc_cap = head_ints;
            CUDA_TRY(cudaMalloc((void **)&c->d_cyc, c->cyc_cap * 4));
        }
        std::vector<int> pack(head_ints, 0);
        if (me == best_owner) {
            memcpy(pack.data(), msg, 24);
            memcpy(pack.data() + 6, verts.data(), std::min(verts.size(), HEAD) * 4);
            CUDA_TRY(cudaMemcpyAsync(c->d_cyc, pack.data(), head_ints * 4, cudaMemcpyHostToDevice, g->stream));
        }
        NCCL_TRY(api, api->Broadcast(c->d_cyc, c->d_cyc, head_ints, ncclInt, best_owner, c->comm, g->stream));
        if (me != best_owner) CUDA_TRY(cudaMemcpyAsync(pack.data(), c->d_cyc, head_ints * 4, cudaMemcpyDeviceToHost, g->stream));
        CUDA_TRY(cudaStreamSynchronize(g->stream));
        if (me != best_owner) memcpy(msg, pack.data(), 24);
        const size_t L = (size_t)msg[2];
        if (me != best_owner) verts.assign(pack.begin() + 6, pack.begin() + 6 + std::min(L, HEAD));
        if (L > HEAD) {
            if (c->cyc_cap < L) {
                CUDA_TRY(cudaFree(c->d_cyc));
                c->cyc_cap = L;
                CUDA_TRY(cudaMalloc((void **)&c->d_cyc, c->cyc_cap * 4));
            }
            if (me == best_owner) CUDA_TRY(cudaMemcpyAsync(c->d_cyc, verts.data(), L * 4, cudaMemcpyHostToDevice, g->stream));
            NCCL_TRY(api, api->Broadcast(c->d_cyc, c->d_cyc, L, ncclInt, best_owner, c->comm, g->stream));
            verts.resize(L);
            CUDA_TRY(cudaMemcpyAsync(verts.data(), c->d_cyc, L * 4, cudaMemcpyDeviceToHost, g->stream));
            CUDA_TRY(cudaStreamSynchronize(g->stream));
        }
    }
    if (trace && me == 0) fprintf(stderr, "  cycle handed over t=%.3f ms\n", 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    out->cycle = verts; out->sc = msg[0]; out->st = msg[1]; out->have_best = true;
    return MRC_OK;
}

static int shard_finish(const ShardOut &o, int max_iter, double tol, double *ratio, double *sum_cost, double *sum_time,
                        int32_t *cycle, int64_t cycle_cap, int64_t *cycle_len, int32_t *iterations, double *final_lo,
                        double *final_hi, int32_t *ticks) {
    if (iterations) *iterations = o.iterations;
    if (ticks) *ticks = o.ticks;
    if (final_lo) *final_lo = o.lo;
    if (final_hi) *final_hi = o.hi;
    if (cycle_len) *cycle_len = 0;
    if (!o.have_best) return set_err(MRC_ERR_NO_CYCLE, "No negative cycle found");   // solver.py:1087-1090
    if (cycle_len) *cycle_len = (int64_t)o.cycle.size();
    if (cycle) for (size_t i = 0; i < o.cycle.size() && (int64_t)i < cycle_cap; i++) cycle[i] = o.cycle[i];
    if (sum_cost) *sum_cost = o.sc;
    if (sum_time) *sum_time = o.st;
    if (ratio) *ratio = o.sc / o.st;
    if (!o.converged && o.hi - o.lo > tol) return set_err(MRC_ERR_CONVERGENCE, "Numeric solver failed to converge");   // solver.py:1102-1109
    (void)max_iter;
    return MRC_OK;
}

extern "C" int mrc_solve_numeric_sharded(mrc_comm *c, mrc_graph *g, double lo, double hi, int max_iter, double tol, double slack,
                                         int K_local, double max_seconds, double *ratio, double *sum_cost, double *sum_time,
                                         int32_t *cycle, int64_t cycle_cap, int64_t *cycle_len, int32_t *iterations,
                                         double *final_lo, double *final_hi, int32_t *ticks) {
    ShardOut o;
    MRC_TRY(shard_solve_rank(c, g, lo, hi, max_iter, tol, slack, K_local, max_seconds, &o));
    return shard_finish(o, max_iter, tol, ratio, sum_cost, sum_time, cycle, cycle_cap, cycle_len, iterations, final_lo, final_hi, ticks);
}

// ------------------------------------------------------------------------------------------------ single-process form
struct mrc_mgraph {
    mrc_comm *comm = nullptr;            // group handle (not owned)
    std::vector<mrc_graph *> replica;    // one per device of the group
};

// run fn(i) on one host thread per device of the group; first failure wins
template <typename F> static int group_run(mrc_comm *c, F fn) {
    const int W = (int)c->group.size();
    std::vector<int> rcs((size_t)W, MRC_OK);
    std::vector<std::string> errs((size_t)W);
    std::vector<std::thread> th;
    for (int i = 0; i < W; i++)
        th.emplace_back([&, i]() { rcs[i] = fn(i); if (rcs[i]) errs[i] = g_mrc_err; });   // g_mrc_err is thread-local
    for (auto &t : th) t.join();
    for (int i = 0; i < W; i++) if (rcs[i]) { g_mrc_err = errs[i]; return rcs[i]; }
    return MRC_OK;
}

extern "C" int mrc_mgraph_destroy(mrc_mgraph *mg) {
    if (!mg) return MRC_OK;
    for (mrc_graph *g : mg->replica) mrc_graph_destroy(g);
    delete mg;
    return MRC_OK;
}

extern "C" int mrc_mgraph_create(mrc_comm *c, int64_t n, int64_t m, const int32_t *src, const int32_t *dst,
                                 const double *cost, const double *time, mrc_mgraph **out) {
    if (!out) return set_err(MRC_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (!c || c->group.empty()) return set_err(MRC_ERR_ARG, "a group communicator (mrc_comm_init) is required");
    if (!src || !dst || !cost || !time) return set_err(MRC_ERR_ARG, "edge arrays are NULL");
    mrc_mgraph *mg = new mrc_mgraph();
    mg->comm = c;
    mg->replica.assign(c->group.size(), nullptr);
    const int rc = group_run(c, [&](int i) { return replicate_rank(c->group[i], 0, n, m, src, dst, cost, time, &mg->replica[i]); });
    if (rc) { mrc_mgraph_destroy(mg); return rc; }
    *out = mg;
    return MRC_OK;
}

extern "C" mrc_graph *mrc_mgraph_graph(mrc_mgraph *mg, int i) {
    return (mg && i >= 0 && i < (int)mg->replica.size()) ? mg->replica[i] : nullptr;
}

extern "C" int mrc_mgraph_solve_numeric(mrc_mgraph *mg, double lo, double hi, int max_iter, double tol, double slack, int K_local,
                                        double max_seconds, double *ratio, double *sum_cost, double *sum_time, int32_t *cycle,
                                        int64_t cycle_cap, int64_t *cycle_len, int32_t *iterations, double *final_lo,
                                        double *final_hi, int32_t *ticks) {
    if (!mg) return set_err(MRC_ERR_ARG, "mgraph is NULL");
    const int W = (int)mg->replica.size();
    std::vector<ShardOut> outs((size_t)W);
    MRC_TRY(group_run(mg->comm, [&](int i) {
        return shard_solve_rank(mg->comm->group[i], mg->replica[i], lo, hi, max_iter, tol, slack, K_local, max_seconds, &outs[i]);
    }));
    return shard_finish(outs[0], max_iter, tol, ratio, sum_cost, sum_time, cycle, cycle_cap, cycle_len, iterations, final_lo, final_hi, ticks);
}
