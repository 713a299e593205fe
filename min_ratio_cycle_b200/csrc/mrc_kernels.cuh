// mrc_kernels.cuh -- sm_100a device code of the parametric negative-cycle hot path.
//
// Replaces the NumPy inner body of the reference (min_ratio_cycle/solver.py):
//   relax_kernel      <- relax_once_kahan / relax_once_standard (:1156-1244) and the int64
//                        relax_once of _has_negative_cycle_exact (:831-849), K lambdas per edge read
//   cycle_check_kernel<- the predecessor walk of _detect_negative_cycle_numeric (:1262-1286), run as
//                        pointer doubling + a walk that sums cost/time over the edges actually used
//   tight_mask_kernel <- equal_mask of _find_zero_cycle_exact (:878-879)
//
// Data layout in HBM (all structure-of-arrays, 256-byte aligned):
//   src,dst  int32[m]   destination-sorted (reference order, solver.py:476-478)
//   cost,time f64[m]    /  icost,itime int64[m] (exact mode)            -> 24 B per edge per pass
//   dist     T[n][KT]   candidate-minor: the KT lambdas of one vertex share a 32/64/128-B sector group,
//                       so ONE random gather per edge serves every candidate
//   pred     u64[n][KT]   (src vertex << 32 | edge index) of the relaxation that last lowered dist, one
//                       8-byte store so vertex and edge always belong together (cross-warp order is still
//                       racy by design; cycles are verified on extraction)
//
// What bounds the sweep (tools/membench.cu, B200, C3 shape): the bare access pattern -- 24 B/edge stream + one
// random 8..32-byte gather per edge + the dist[dst] read, no arithmetic -- runs in 53-57 us (stream alone: 32 us).
// Every scattered lane of a gather is its own L1TEX wavefront (10 M per sweep, ~1 per clock per SM), so an
// edge-parallel sweep of a random digraph cannot go much below that however few instructions it issues.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mrc_b200.h"

namespace cg = cooperative_groups;
typedef long long ll;
typedef unsigned long long ull;

#define MRC_MAX_BURST 32
#define MRC_MAX_ROUNDS 40
#ifndef MRC_RELAX_THREADS
#define MRC_RELAX_THREADS 256
#endif
#define MRC_FULL 0xffffffffu

struct RelaxArgs {
    const int *src, *dst;
    const double *cost, *time;
    const ll *icost, *itime;
    void *dist;
    ull *pred;           // [n][KT]: (src vertex << 32) | edge index of the relaxation that last lowered dist; ~0 = none
    unsigned *changed;   // one word: bit k set when candidate k lowered some dist in this pass
    ll m;
    double lam[MRC_KMAX];
    ll a[MRC_KMAX], b[MRC_KMAX];
    double slack;
    unsigned active;     // bit k: candidate k still undecided
    // A launch may serve a SUB-RANGE of the KT candidate slots of the dist/pred layout (the slots still undecided):
    // the kernel template then is the sub-range width, `ks` stays the slot count of the layout (row stride of
    // dist and pred), the host pre-offsets dist/pred/lam/a/b/active/pe.. by the first slot `k0`, and `k0` only
    // shifts the bits of the changed word back.
    int ks, k0;
    // Convergence gate: the changed word of the PREVIOUS launch of the same burst (NULL for the first).  A candidate
    // that launch did not move sits at its fixed point and is skipped; if that holds for every candidate this launch
    // serves it returns at once -- the host enqueues whole bursts without looking, the device stops on the pass.
    const unsigned *gate;
    // Candidates a cycle check has already decided on the device (DevStatus::decided, bit per slot of the layout):
    // bursts enqueued before the host has read that check's result skip them.
    const unsigned *decided;
    unsigned *act_out;   // one bit per vertex, "dist[v] was lowered for some candidate" (seed of a worklist); NULL = not tracked
    ull *examined;       // worklist passes: edges actually relaxed
    ll n;
    // scenario batching (numeric only): candidate k sees edge pe[k] with cost + pdc[k], time + pdt[k]
    int scen;
    int pe[MRC_KMAX];
    double pdc[MRC_KMAX], pdt[MRC_KMAX];
};

struct CycleOut {
    int found, len, minv, pad;   // found: 0 none, 1 verified negative, 2 tie (rounding artefact)
    double sumc, sumt;   // numeric: sums along the predecessor walk
    double maxabs;       // numeric: max |dist| over the cycle's vertices (bounds the rounding noise of the walk)
    ll isumc, isumt;     // exact
};

struct DevStatus {
    unsigned changed[MRC_MAX_BURST];
    ull examined;                              // worklist passes: edges relaxed (cleared with `changed`)
    ull ns_check[2];                           // cycle check, globaltimer of thread 0: jump init + pointer doubling, extraction (cleared with `changed`)
    ull check_rounds;                          // ... doubling rounds run
    unsigned cnt[MRC_MAX_ROUNDS * MRC_KMAX];   // live vertices per doubling round and candidate
    int minlive[MRC_MAX_ROUNDS * MRC_KMAX];    // smallest live vertex per doubling round and candidate
    unsigned cyc_mask, rounds_run, pend, cuts;
    int rep[MRC_KMAX * 6];   // [MRC_MAX_EXTRACT][MRC_KMAX]
    CycleOut out[MRC_KMAX];
    ull sample_best[MRC_KMAX];   // cycle sampling: min over samples of (order key of the cycle ratio, 32 bit) << 32 | start vertex
    unsigned decided;        // candidates for which this probe's checks reported a cycle (bit per slot); cleared by the host per probe
    unsigned act_count;      // sparse mode: vertices lowered by the last relaxation pass (size of the next worklist)
    unsigned act_cnt[3];     // worklist sizes, rotating: pass p reads [p % 3], appends to [(p+1) % 3], zeroes [(p+2) % 3]
    unsigned xpad;
    ull xword[4];            // distributed probe: (step << 8 | OR of every rank's payload) of the last exchanges, for the CTAs of this GPU
};

// Sparse (worklist) relaxation, see relax_sparse_kernel
struct SparseArgs {
    RelaxArgs r;             // arrays, candidates, ks/k0 (pre-offset like a narrowed launch); r.changed = first word of the burst
    const int *out_off;      // [n+1] out-edge index by source
    const int *out_edge;     // [m]   edge indices grouped by source
    unsigned *bm[2];         // "already on the worklist" bitmaps, one bit per vertex
    int *wl[2];              // worklists (vertex ids)
    DevStatus *st;
    int first;               // buffer (0/1) holding the input worklist of the first pass
    int cidx;                // index into st->act_cnt of that worklist's size
    int npasses;
    int kc;                  // candidate slots served (1, 2, 4 or 8)
};

struct CheckArgs {
    const int *src;
    const double *cost, *time;
    const ll *icost, *itime;
    ull *pred;           // written only to cut a false cycle
    const void *dist;
    int *jump;
    int *cyc_edges;      // [KT][cap]
    DevStatus *st;
    ll n, cap;
    int rounds;          // ceil(log2(n+1)) + 1
    unsigned active;
    const unsigned *gate;   // changed word of the last relaxation launch: candidates it did not move are at their
                            // fixed point (no cycle to look for); all of them there -> the launch returns at once
    int exact;
    int prune_above;        // candidates are in increasing lambda order and the caller prunes by monotonicity: once the
                            // lowest candidate with a cycle is known, the ones above it are not examined further
    int k0;                 // first slot of the layout this (possibly narrowed) launch serves: shifts the bits of `decided`
    int samples;            // cycle sampling: warps per candidate that each walk one cycle of the predecessor graph (0 = off)
    double lam[MRC_KMAX];
    ll a[MRC_KMAX], b[MRC_KMAX];
    int scen;
    int pe[MRC_KMAX];
    double pdc[MRC_KMAX], pdt[MRC_KMAX];
};

// ---- cache-hinted accessors --------------------------------------------------------------------
// Edge stream: read once per pass -> evict-first, never allocate in L1.
// dist/pred/jump: shared between CTAs inside one launch -> L2 only (.cg), never a stale L1 line.
template <int K> struct DistVec;
template <> struct DistVec<1> {
    template <typename T> static __device__ __forceinline__ void load(T *r, const T *p) { r[0] = __ldcg(p); }
};
template <> struct DistVec<2> {
    static __device__ __forceinline__ void load(double *r, const double *p) {
        double2 v = __ldcg(reinterpret_cast<const double2 *>(p)); r[0] = v.x; r[1] = v.y;
    }
    static __device__ __forceinline__ void load(ll *r, const ll *p) {
        longlong2 v = __ldcg(reinterpret_cast<const longlong2 *>(p)); r[0] = v.x; r[1] = v.y;
    }
};
// 256-bit global loads (sm_100+, SASS LDG.E.256): the four candidates of a vertex arrive in ONE request,
// so a K=4 gather costs the same L1TEX wavefront as a K=1 gather.
template <> struct DistVec<4> {
    static __device__ __forceinline__ void load(double *r, const double *p) {
        asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];"
                     : "=d"(r[0]), "=d"(r[1]), "=d"(r[2]), "=d"(r[3]) : "l"(p));
    }
    static __device__ __forceinline__ void load(ll *r, const ll *p) {
        asm volatile("ld.global.cg.v4.b64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(r[0]), "=l"(r[1]), "=l"(r[2]), "=l"(r[3]) : "l"(p));
    }
};

// ---- arithmetic of one relaxation ----------------------------------------------------------------------
// numeric: w = cost - lam*time with a rounded product and a rounded difference (no FMA), exactly as
// NumPy evaluates `self._C - lam * self._T` (solver.py:1144); improve test `cand < dist[v] - slack`
// (:1161).  dist <= 0 always (starts at +0.0, only decreases), so for negative doubles
// "smaller value" == "larger bit pattern": atomicMax on the raw bits is a 64-bit atomic min.
// exact: w = b*icost - a*itime in int64 (solver.py:799-800), strict < without slack (:833).
// Where a sweep finds its candidates and its dist / pred columns.  ParamView: straight from the kernel parameters
// (constant-bank operands, no registers) -- every launch whose host picked the columns.  RegView<K>: a narrower set of
// columns picked ON THE DEVICE by the persistent probe kernel (mrc_probe.cuh); its few lambdas live in registers.
struct ParamView {
    const RelaxArgs &p;
    __device__ __forceinline__ double lam(int k) const { return p.lam[k]; }
    __device__ __forceinline__ ll a(int k) const { return p.a[k]; }
    __device__ __forceinline__ ll b(int k) const { return p.b[k]; }
    __device__ __forceinline__ void *dist() const { return p.dist; }
    __device__ __forceinline__ ull *pred() const { return p.pred; }
    __device__ __forceinline__ int ks() const { return p.ks; }
    __device__ __forceinline__ int k0() const { return p.k0; }
    __device__ __forceinline__ unsigned e_base() const { return 0u; }
    template <typename T> __device__ __forceinline__ void push(int, T) const {}
};
template <int K> struct RegView {
    double lam_[K];
    ll a_[K], b_[K];
    void *dist_;
    ull *pred_;
    int ks_, k0_;
    __device__ __forceinline__ double lam(int k) const { return lam_[k]; }
    __device__ __forceinline__ ll a(int k) const { return a_[k]; }
    __device__ __forceinline__ ll b(int k) const { return b_[k]; }
    __device__ __forceinline__ void *dist() const { return dist_; }
    __device__ __forceinline__ ull *pred() const { return pred_; }
    __device__ __forceinline__ int ks() const { return ks_; }
    __device__ __forceinline__ int k0() const { return k0_; }
    __device__ __forceinline__ unsigned e_base() const { return 0u; }
    template <typename T> __device__ __forceinline__ void push(int, T) const {}
};
// One candidate, one SLICE of the edge stream, the other slices on peer GPUs (distributed probe, mrc_probe.cuh): edge ids
// are offset by the slice start, and every value this GPU lowers is pushed into the peers' replicas of dist -- each vertex
// has a single owner (the GPU holding its in-edges), the push is an atomic min so that stores of different warps cannot
// land out of order, and only the few vertices a pass still lowers cross NVLink (8 bytes each per peer).
struct PeerPush {
    ull *peer[MRC_KMAX];   // dist replicas of the other ranks ([n][1]); n_peers entries
    int n_peers;
};
struct DistView {
    double lam_;
    ll a_, b_;
    void *dist_;
    ull *pred_;
    unsigned e_base_;
    const PeerPush *pp_;
    __device__ __forceinline__ double lam(int) const { return lam_; }
    __device__ __forceinline__ ll a(int) const { return a_; }
    __device__ __forceinline__ ll b(int) const { return b_; }
    __device__ __forceinline__ void *dist() const { return dist_; }
    __device__ __forceinline__ ull *pred() const { return pred_; }
    __device__ __forceinline__ int ks() const { return 1; }
    __device__ __forceinline__ int k0() const { return 0; }
    __device__ __forceinline__ unsigned e_base() const { return e_base_; }
    __device__ __forceinline__ void push(int v, double cand) const {
        const ull bits = (ull)__double_as_longlong(cand);
        for (int r = 0; r < pp_->n_peers; r++) atomicMax_system(pp_->peer[r] + v, bits);
    }
    __device__ __forceinline__ void push(int v, ll cand) const {
        for (int r = 0; r < pp_->n_peers; r++) atomicMin_system(reinterpret_cast<ll *>(pp_->peer[r]) + v, cand);
    }
};
template <typename V> __device__ __forceinline__ double mrc_cand(const V &vw, int k, double du, double c, double t) {
    return __dadd_rn(du, __dsub_rn(c, __dmul_rn(vw.lam(k), t)));
}
template <typename V> __device__ __forceinline__ ll mrc_cand(const V &vw, int k, ll du, ll c, ll t) {
    return du + (vw.b(k) * c - vw.a(k) * t);
}
// scenario batching: candidate k re-solves the graph with ONE edge perturbed (sensitivity_analysis, solver.py:1716-1721)
__device__ __forceinline__ void mrc_scenario_patch(const RelaxArgs &p, int k, int e, double &c, double &t) {
    if (e == p.pe[k]) { c = __dadd_rn(c, p.pdc[k]); t = __dadd_rn(t, p.pdt[k]); }
}
__device__ __forceinline__ void mrc_scenario_patch(const RelaxArgs &, int, int, ll &, ll &) {}
__device__ __forceinline__ bool mrc_improves(const RelaxArgs &p, double cand, double dv) {
    return cand < __dsub_rn(dv, p.slack);
}
__device__ __forceinline__ bool mrc_improves(const RelaxArgs &, ll cand, ll dv) { return cand < dv; }
__device__ __forceinline__ double mrc_identity(double) { return __longlong_as_double(0x7ff0000000000000ll); }
__device__ __forceinline__ ll mrc_identity(ll) { return 0x7fffffffffffffffll; }
// atomic min in two halves: the request (returns the value it found) and the test "did it lower dist?"
__device__ __forceinline__ double mrc_atomic_min(double *addr, double cand) {
    return __longlong_as_double((ll)atomicMax(reinterpret_cast<ull *>(addr), (ull)__double_as_longlong(cand)));
}
__device__ __forceinline__ ll mrc_atomic_min(ll *addr, ll cand) { return atomicMin(addr, cand); }
__device__ __forceinline__ bool mrc_lowered(double old, double cand) {
    return (ull)__double_as_longlong(old) < (ull)__double_as_longlong(cand);
}
__device__ __forceinline__ bool mrc_lowered(ll old, ll cand) { return cand < old; }
__device__ __forceinline__ bool mrc_atomic_lower(double *addr, double cand) { return mrc_lowered(mrc_atomic_min(addr, cand), cand); }
__device__ __forceinline__ bool mrc_atomic_lower(ll *addr, ll cand) { return mrc_lowered(mrc_atomic_min(addr, cand), cand); }

// ---- one row of 32 consecutive edges, K candidates ------------------------------------------------------
// Edges are destination-sorted, so the lanes of a row that share a destination are contiguous: a row is a
// sequence of RUNS.  Computed once per row (and only when something improves), shared by all K candidates.
struct RowRuns { int start, end; unsigned mask; };
__device__ __forceinline__ RowRuns mrc_row_runs(int v, int lane) {
    const int vprev = __shfl_up_sync(MRC_FULL, v, 1);
    const unsigned heads = __ballot_sync(MRC_FULL, lane == 0 || v != vprev);
    RowRuns r;
    r.start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
    const unsigned above = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));
    r.end = above ? (__ffs(above) - 2) : 31;
    r.mask = ((r.end == 31) ? 0xffffffffu : ((2u << r.end) - 1u)) & ~((1u << r.start) - 1u);
    return r;
}
// Several in-edges of one destination sit in adjacent lanes, and in the first passes of a probe several of them
// would lower dist[v] in the same instruction; if they all issued their atomic, the pred[v] store of a worse
// candidate could land last, pred would name an edge that did NOT produce dist[v], and the predecessor graph could
// close cycles that are not negative.  So per destination only ONE lane of the row issues the atomic: the one
// holding the smallest candidate (which also removes most same-address atomics of the early passes).
// numeric: an improving candidate is negative (cand < dist[v] - slack <= 0), more negative <=> larger raw bits, so
// the run minimum is a segmented MAX-scan over the upper 32 bits (one shuffle per step instead of two); lanes that
// tie there (candidates within 2^-20 of each other) are broken towards the lowest lane -- if that is not the exact
// minimum, dist[v] receives a slightly larger, still valid value and the true minimum is applied by the next pass.
__device__ __forceinline__ bool mrc_pick_winner(bool imp, double cand, const RowRuns &r, int lane) {
    const unsigned hi = (unsigned)((ull)__double_as_longlong(cand) >> 32);
    unsigned x = imp ? hi : 0u;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned a = __shfl_up_sync(MRC_FULL, x, off);
        if (lane - off >= r.start && a > x) x = a;
    }
    const unsigned runmax = __shfl_sync(MRC_FULL, x, r.end);
    const unsigned eq = __ballot_sync(MRC_FULL, imp && hi == runmax);
    return imp && (__ffs(eq & r.mask) - 1 == lane);
}
// exact: full 64-bit segmented min, lowest lane on ties (the reference's first-occurrence rule inside a row)
__device__ __forceinline__ bool mrc_pick_winner(bool imp, ll cand, const RowRuns &r, int lane) {
    ll x = imp ? cand : mrc_identity(ll(0));
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const ll a = __shfl_up_sync(MRC_FULL, x, off);
        if (lane - off >= r.start && a < x) x = a;
    }
    const ll runmin = __shfl_sync(MRC_FULL, x, r.end);
    const unsigned eq = __ballot_sync(MRC_FULL, imp && cand == runmin);
    return imp && (__ffs(eq & r.mask) - 1 == lane);
}

// The row is warp-synchronous: every lane calls it (valid == false for lanes past the end).  Steady state (nothing
// improves, almost every row after the first passes): two gathers, 4 flops + one compare per candidate, ONE vote.
// Control flow never waits for an atomic's return value before all atomics of the row are in flight.
template <int K, typename T, bool SCEN, typename V>
__device__ __forceinline__ void relax_row(const RelaxArgs &p, const V &vw, unsigned act, bool valid, int lane, unsigned e, int u, int v,
                                          T c, T t, unsigned &chg, unsigned *act_out) {
    // K = 8 is processed as two halves of 4 candidates (two 256-bit gathers from the same 64-B group, one after
    // the other), so its register footprint stays that of K = 4.
    constexpr int H = K > 4 ? 4 : K;
    T *dist = static_cast<T *>(vw.dist());
    const int ks = vw.ks();
#pragma unroll
    for (int h = 0; h < K; h += H) {
        const unsigned acth = (act >> h) & ((1u << H) - 1u);
        if (!acth) continue;   // warp-uniform: nothing active in this half
        T du[H], dv[H], cand[H];
        if (valid) {
            DistVec<H>::load(du, dist + (size_t)u * ks + h);
            DistVec<H>::load(dv, dist + (size_t)v * ks + h);
        }
        bool imp[H], any = false;
#pragma unroll
        for (int k = 0; k < H; k++) {
            T ck = c, tk = t;
            if (SCEN) mrc_scenario_patch(p, h + k, (int)e, ck, tk);
            cand[k] = mrc_cand(vw, h + k, du[k], ck, tk);
            imp[k] = valid && ((acth >> k) & 1u) && mrc_improves(p, cand[k], dv[k]);
            any |= imp[k];
        }
        if (!__any_sync(MRC_FULL, any)) continue;   // warp-uniform
        // ---- some lane improves some candidate: one winner per destination and candidate
        const RowRuns rr = mrc_row_runs(valid ? v : -1 - lane, lane);
        unsigned wins = 0;
#pragma unroll
        for (int k = 0; k < H; k++) {
            const unsigned im = __ballot_sync(MRC_FULL, imp[k]);
            if (im == 0) continue;               // warp-uniform
            bool w = imp[k];
            if (im & (im - 1)) w = mrc_pick_winner(imp[k], cand[k], rr, lane);
            if (w) wins |= 1u << k;
        }
        T old[H];
#pragma unroll
        for (int k = 0; k < H; k++)
            if ((wins >> k) & 1u) old[k] = mrc_atomic_min(dist + (size_t)v * ks + h + k, cand[k]);
        bool lowered = false;
#pragma unroll
        for (int k = 0; k < H; k++)
            if (((wins >> k) & 1u) && mrc_lowered(old[k], cand[k])) {
                __stcg(vw.pred() + (size_t)v * ks + h + k, ((ull)(unsigned)u << 32) | (e + vw.e_base()));
                vw.push(v, cand[k]);
                chg |= 1u << (h + k);
                lowered = true;
            }
        if (lowered && act_out) atomicOr(act_out + (v >> 5), 1u << (v & 31));
    }
}

template <bool EXACT> struct WeightT { typedef double type; };
template <> struct WeightT<true> { typedef ll type; };

// Edge-parallel relaxation.  A warp owns ROWS rows of 32 CONSECUTIVE edges per step; lane i takes edge
// base + 32*j + i, so every edge-stream load is one fully coalesced 128/256-B request and -- because
// the edges are destination-sorted -- the 32 dist[dst] reads of a row collapse to one or two lines.
// What remains is exactly one random line per edge (dist[src]); the kernel is bound by that L1TEX
// gather rate and by the 24 B/edge HBM stream.  Persistent grid: occupancy x 148 SMs, warps walk
// consecutive chunks so DRAM pages are streamed in order.
#ifndef MRC_TUNE_ROWS
#define MRC_TUNE_ROWS 2
#endif
#ifndef MRC_TUNE_MB
#define MRC_TUNE_MB 4
#endif
#ifndef MRC_TUNE_MB12
#define MRC_TUNE_MB12 4
#endif
#ifndef MRC_TUNE_PF
#define MRC_TUNE_PF (K <= 4)
#endif
template <int K> struct RelaxTune {
    static constexpr int rows = MRC_TUNE_ROWS, minblocks = K <= 2 ? MRC_TUNE_MB12 : MRC_TUNE_MB;
    // K <= 4: the edge stream of the NEXT chunk is requested before this chunk's gathers are issued, so the HBM latency
    // of the stream overlaps the L2 latency of the dist gathers and, in the update-heavy first passes, the winner
    // resolution (B200, C3: K=1 -8 % per steady pass; K=4 pass 1 157 -> 116 us, pass 2 103 -> 80, steady 60 -> 58);
    // at K = 8 (two halves) the extra registers cost more than the overlap wins (101 -> 120 us)
    static constexpr bool prefetch = MRC_TUNE_PF;
    // MRC_TUNE_L2PF (experiments): without the register double buffer, one prefetch.global.L2 per warp step pulls the lines
    // of the warp's next chunk into L2 (12 lines of 128 B, one per lane) -- costs no registers
#ifndef MRC_TUNE_L2PF
#define MRC_TUNE_L2PF 0
#endif
    static constexpr int l2pf = MRC_TUNE_L2PF;   // chunks ahead (0 = off)
};
// One sweep over the edge stream for the candidates in `act`; ORs "candidate k lowered something" into *changed.
// PF: 1 = register double buffer of the edge stream (the stand-alone kernels, K <= 4), 0 = none, 2 = none + one
// prefetch.global.L2 per warp step for the lines of the warp's next chunk (the persistent probe kernel, whose sweeps are
// called through the ABI and cannot afford the 12 registers of the double buffer)
template <int K, bool EXACT, bool SCEN, typename V, int PF = (RelaxTune<K>::prefetch ? 1 : (RelaxTune<K>::l2pf > 0 ? 2 : 0))>
__device__ __forceinline__ void relax_sweep(const RelaxArgs &p, const V &vw, unsigned act, unsigned *changed, unsigned *act_out) {
    typedef typename WeightT<EXACT>::type W;
    const W *cost = EXACT ? reinterpret_cast<const W *>(p.icost) : reinterpret_cast<const W *>(p.cost);
    const W *time = EXACT ? reinterpret_cast<const W *>(p.itime) : reinterpret_cast<const W *>(p.time);
    unsigned chg = 0;
    const int lane = threadIdx.x & 31;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned nwarps = (gridDim.x * blockDim.x) >> 5;
    constexpr int ROWS = RelaxTune<K>::rows;
    constexpr unsigned CH = 32 * ROWS;
    const unsigned m = (unsigned)p.m;
    const unsigned nchunks = (m + CH - 1) / CH, nfull = m / CH;   // chunks [0, nfull) need no bounds checks
    auto tail = [&](unsigned c) {   // the one partial chunk
#pragma unroll
        for (int j = 0; j < ROWS; j++) {
            const unsigned e = c * CH + 32 * j + lane;
            const bool ok = e < m;
            relax_row<K, W, SCEN>(p, vw, act, ok, lane, e, ok ? p.src[e] : 0, ok ? p.dst[e] : 0, ok ? cost[e] : W(0),
                                  ok ? time[e] : W(0), chg, act_out);
        }
    };
    if constexpr (PF != 1) {
        for (unsigned c = warp; c < nchunks; c += nwarps) {
            if (c >= nfull) { tail(c); break; }
            if constexpr (PF == 2) {
                const unsigned nx = c + (RelaxTune<K>::l2pf > 0 ? RelaxTune<K>::l2pf : 1) * nwarps;
                constexpr int L4 = CH * 4 / 128, L8 = CH * 8 / 128;   // lines of a chunk per int / f64 array
                if (nx < nfull && lane < 2 * L4 + 2 * L8) {
                    const char *a = lane < L4            ? (const char *)(p.src + (size_t)nx * CH) + lane * 128
                                    : lane < 2 * L4      ? (const char *)(p.dst + (size_t)nx * CH) + (lane - L4) * 128
                                    : lane < 2 * L4 + L8 ? (const char *)(cost + (size_t)nx * CH) + (lane - 2 * L4) * 128
                                                         : (const char *)(time + (size_t)nx * CH) + (lane - 2 * L4 - L8) * 128;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
                }
            }
            const unsigned e0 = c * CH + lane;
            int s[ROWS], d[ROWS];
            W cs[ROWS], tm[ROWS];
#pragma unroll
            for (int j = 0; j < ROWS; j++) {
                const unsigned e = e0 + 32 * j;
                s[j] = __ldcs(p.src + e); d[j] = __ldcs(p.dst + e);
                cs[j] = __ldcs(cost + e); tm[j] = __ldcs(time + e);
            }
#pragma unroll
            for (int j = 0; j < ROWS; j++)
                relax_row<K, W, SCEN>(p, vw, act, true, lane, e0 + 32 * j, s[j], d[j], cs[j], tm[j], chg, act_out);
        }
    } else {
        int s[ROWS], d[ROWS], s2[ROWS], d2[ROWS];
        W cs[ROWS], tm[ROWS], cs2[ROWS], tm2[ROWS];
        unsigned c = warp;
        if (c < nfull) {
            const unsigned e0 = c * CH + lane;
#pragma unroll
            for (int j = 0; j < ROWS; j++) {
                const unsigned e = e0 + 32 * j;
                s[j] = __ldcs(p.src + e); d[j] = __ldcs(p.dst + e);
                cs[j] = __ldcs(cost + e); tm[j] = __ldcs(time + e);
            }
        }
        while (c < nchunks) {
            if (c >= nfull) { tail(c); break; }
            const unsigned nx = c + nwarps;
            if (nx < nfull) {
                const unsigned e0 = nx * CH + lane;
#pragma unroll
                for (int j = 0; j < ROWS; j++) {
                    const unsigned e = e0 + 32 * j;
                    s2[j] = __ldcs(p.src + e); d2[j] = __ldcs(p.dst + e);
                    cs2[j] = __ldcs(cost + e); tm2[j] = __ldcs(time + e);
                }
            }
            const unsigned e0 = c * CH + lane;
#pragma unroll
            for (int j = 0; j < ROWS; j++)
                relax_row<K, W, SCEN>(p, vw, act, true, lane, e0 + 32 * j, s[j], d[j], cs[j], tm[j], chg, act_out);
#pragma unroll
            for (int j = 0; j < ROWS; j++) { s[j] = s2[j]; d[j] = d2[j]; cs[j] = cs2[j]; tm[j] = tm2[j]; }
            c = nx;
        }
    }
    // changed flag: warp vote, one atomic per warp at most
    chg = __reduce_or_sync(MRC_FULL, chg);
    if (lane == 0 && chg) atomicOr(changed, chg << vw.k0());
}

template <int K, bool EXACT, bool SCEN>
__global__ void __launch_bounds__(MRC_RELAX_THREADS, RelaxTune<K>::minblocks) relax_kernel(const __grid_constant__ RelaxArgs p) {
    // Programmatic dependent launch (sm_90+): the sweeps of a burst are launched with programmatic stream serialisation, so
    // this grid is already queued on the device while the previous sweep drains; it must not read what that sweep wrote
    // before griddepcontrol.wait returns (full completion + flush of the prerequisite grid), and it lets ITS successor be
    // queued right away.  Without the launch attribute both instructions are no-ops.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    unsigned act = p.active;
    if (p.decided) act &= ~(__ldcg(p.decided) >> p.k0);
    if (p.gate) act &= __ldcg(p.gate) >> p.k0;
    if (!act) return;
    relax_sweep<K, EXACT, SCEN>(p, ParamView{p}, act, p.changed, p.act_out);
}

// dist[v][0..ks) = base[v] for every vertex: a probe that starts from the converged potentials of a neighbouring
// problem instead of zero (scenario batches of sensitivity_analysis).  Any start <= 0 is legitimate: a pass that
// lowers nothing leaves feasible potentials, which is the certificate "no negative cycle", whatever the start was.
__global__ void broadcast_column_kernel(const ull *__restrict__ base, ull *__restrict__ dist, int ks, ll n) {
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n * ks; i += (ll)gridDim.x * blockDim.x)
        dist[i] = __ldg(base + i / ks);
}

// Column copies between the [n][ks] candidate layout and a compact [n][1] one: once a round has narrowed to a single
// live slot, its dist/pred column moves to compact buffers (16 MB at n = 1 M, ~5 us) so that the remaining sweeps
// gather from 8 MB of contiguous doubles instead of every fourth double of 32 MB.
__global__ void column_copy_kernel(const ull *__restrict__ sd, const ull *__restrict__ sp, int sks, ull *__restrict__ dd,
                                   ull *__restrict__ dp, int dks, ll n) {
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (ll)gridDim.x * blockDim.x) {
        dd[i * dks] = __ldcg(sd + i * sks);
        dp[i * dks] = __ldcg(sp + i * sks);
    }
}

// ---- sparse relaxation ------------------------------------------------------------------------------------
// After a handful of dense sweeps fewer than 1 % of the vertices are still being lowered (C3, certifying probe:
// 239 k, 82 k, 34 k, 17 k, 9 k, 5.7 k, ... 8 vertices in passes 1..20), and an edge can only relax if its SOURCE was
// lowered since it was last examined.  From then on a pass only touches the out-edges of the vertices the previous
// pass lowered: a worklist of vertex ids (deduplicated through a bitmap), an out-edge index by source, 8 lanes per
// vertex.  One COOPERATIVE launch runs a whole burst of passes (grid sync between passes, 9-12 us each instead of a
// 60-90 us sweep of the 240 MB edge stream), keeps the per-pass changed words the host's verdict logic reads, and
// stops on the pass whose worklist is empty -- a true fixed point (every edge was examined after its source's last
// change), so verdicts, potentials and parity are unchanged.
__global__ void compact_active_kernel(const unsigned *bm, int words, int *wl, DevStatus *st, int cidx) {
    const int lane = threadIdx.x & 31;
    for (int w0 = (blockIdx.x * blockDim.x + threadIdx.x) - lane; w0 < words; w0 += gridDim.x * blockDim.x) {
        const int w = w0 + lane;
        const unsigned bits = w < words ? __ldcg(bm + w) : 0u;
        const int c = __popc(bits);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&st->act_cnt[cidx], (unsigned)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        int pos = (int)base + incl - c;
        unsigned b = bits;
        while (b) {
            const int i = __ffs(b) - 1;
            b &= b - 1;
            wl[pos++] = (w << 5) | i;
        }
    }
}

// One worklist pass: the out-edges of the `cin` vertices of wl_in; vertices it lowers are appended to wl_out (counted in
// st->act_cnt[co], deduplicated through bm_out).  Called by every thread of a cooperative grid; the caller grid-syncs.
template <bool EXACT, typename V>
__device__ __forceinline__ void sparse_pass_device(const RelaxArgs &p, const V &vw, int kc, unsigned active, const int *out_off,
                                                   const int *out_edge, const int *wl_in, int *wl_out, unsigned *bm_in,
                                                   unsigned *bm_out, DevStatus *st, unsigned cin, int co, unsigned *changed) {
    typedef typename WeightT<EXACT>::type W;
    const W *cost = EXACT ? reinterpret_cast<const W *>(p.icost) : reinterpret_cast<const W *>(p.cost);
    const W *time = EXACT ? reinterpret_cast<const W *>(p.itime) : reinterpret_cast<const W *>(p.time);
    W *dist = static_cast<W *>(vw.dist());
    ull *pred = vw.pred();
    const int ks = vw.ks();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int sub = tid & 7;                                   // 8 lanes share one worklist vertex
    const int grp = tid >> 3, ngrp = (gridDim.x * blockDim.x) >> 3;
    unsigned chg = 0, seen = 0;
    for (unsigned i = grp; i < cin; i += ngrp) {
        const int u = __ldcg(wl_in + i);
        if (sub == 0) atomicAnd(bm_in + (u >> 5), ~(1u << (u & 31)));   // leave the bitmap clean for its next use
        W du[MRC_KMAX];
#pragma unroll
        for (int k = 0; k < MRC_KMAX; k++)
            du[k] = (k < kc && ((active >> k) & 1u)) ? __ldcg(dist + (size_t)u * ks + k) : W(0);
        const int beg = __ldg(out_off + u), end = __ldg(out_off + u + 1);
        for (int j = beg + sub; j < end; j += 8) {
            const int e = __ldg(out_edge + j);
            const int v = __ldg(p.dst + e);
            const W c = __ldg(cost + e), t = __ldg(time + e);
            seen++;
            bool any = false;
#pragma unroll
            for (int k = 0; k < MRC_KMAX; k++) {
                if (k >= kc || !((active >> k) & 1u)) continue;
                W ck = c, tk = t;
                mrc_scenario_patch(p, k, e, ck, tk);
                const W cand = mrc_cand(vw, k, du[k], ck, tk);
                const W dv = __ldcg(dist + (size_t)v * ks + k);
                if (mrc_improves(p, cand, dv) && mrc_atomic_lower(dist + (size_t)v * ks + k, cand)) {
                    __stcg(pred + (size_t)v * ks + k, ((ull)(unsigned)u << 32) | (unsigned)e);
                    chg |= 1u << k;
                    any = true;
                }
            }
            if (any) {
                const unsigned bit = 1u << (v & 31);
                if (!(atomicOr(bm_out + (v >> 5), bit) & bit)) wl_out[atomicAdd(&st->act_cnt[co], 1u)] = v;
            }
        }
    }
    chg = __reduce_or_sync(0xffffffffu, chg);
    seen = __reduce_add_sync(0xffffffffu, seen);
    if ((threadIdx.x & 31) == 0) {
        if (chg) atomicOr(changed, chg << vw.k0());
        if (seen) atomicAdd(p.examined, (ull)seen);
    }
}

template <bool EXACT>
__global__ void __launch_bounds__(256) relax_sparse_kernel(const __grid_constant__ SparseArgs a) {
    cg::grid_group grid = cg::this_grid();
    const RelaxArgs &p = a.r;
    DevStatus *st = a.st;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    int in = a.first, ci = a.cidx;
    for (int pass = 0; pass < a.npasses; pass++) {
        const int out = in ^ 1, co = ci == 2 ? 0 : ci + 1, cz = co == 2 ? 0 : co + 1;
        const unsigned cin = __ldcg(&st->act_cnt[ci]);          // final since the previous grid sync
        if (cin == 0) break;                                    // grid-uniform: fixed point
        if (tid == 0) st->act_cnt[cz] = 0;                      // next pass appends there; last read one pass ago
        sparse_pass_device<EXACT>(p, ParamView{p}, a.kc, p.active, a.out_off, a.out_edge, a.wl[in], a.wl[out], a.bm[in], a.bm[out],
                                  st, cin, co, p.changed + pass);
        grid.sync();
        in = out; ci = co;
    }
    if (tid == 0) st->act_count = st->act_cnt[ci];
}

// ---- predecessor-graph cycle check -----------------------------------------------------------------
// One cooperative launch:
//   0. jump[v][k] = src[pred[v][k]]  (or -1 at roots)
//   1. pointer doubling in place, <= ceil(log2 n)+1 rounds, stops as soon as every active candidate's
//      predecessor graph has collapsed to roots (it is a forest -> no cycle yet)
//   2. any vertex still alive leads into a cycle; its jump target lies ON the cycle.  Per candidate
//      the smallest such vertex picks the representative (deterministic for a given pred array)
//   3. one thread per candidate walks that cycle from its smallest vertex, records the edge list, sums
//      cost/time over exactly those edges and tests sum_cost/sum_time < lambda (exact: b*C - a*T < 0).
//      pred writes of different warps can still race, so a cycle of the predecessor graph need not be
//      negative: such a cycle is CUT (pred of its smallest vertex reset to -1), its vertices are marked
//      dead for this launch, and steps 2-3 repeat (<= MRC_MAX_EXTRACT times) on the remaining cycles.
#define MRC_MAX_EXTRACT 6
#define MRC_CYC_NEG 1
#define MRC_CYC_TIE 2
// ---- pieces of the check
template <int K, typename CV> __device__ __forceinline__ int cycle_step(const CV &cv, int k, int y) {
    return y < 0 ? -1 : (int)(__ldcg(cv.pred() + (ll)y * K + k) >> 32);
}
// one lap over pred from y, recording the edge list; returns the vertex it stopped at
template <int K, typename CV> __device__ __forceinline__ int cycle_lap(const CheckArgs &p, const CV &cv, int k, int y, ll cap, ll &len) {
    int cur = y;
    ll i = 0;
    do {
        const ull pe = __ldcg(cv.pred() + (ll)cur * K + k);
        if (i < p.cap) cv.cyc_edges()[(ll)k * p.cap + i] = (int)(unsigned)pe;
        cur = (int)(pe >> 32);
        i++;
    } while (cur != y && cur >= 0 && i < cap);
    len = i;
    return cur;
}
// The whole CTA: sums, smallest vertex and max |dist| over the recorded lap of candidate k (x = vertex the lap started from,
// -1 = none), then thread 0 tests it against lambda: negative -> reported in st->out[k] (+ decided bit); rounding artefact ->
// TIE; otherwise the cycle is false: pred of its smallest vertex is reset and its vertices are marked dead
// in jump.  Returns (thread 0) whether candidate k is settled for this check.
template <int K, typename CV>
__device__ __forceinline__ bool cycle_verify_block(const CheckArgs &p, const CV &cv, int k, int x, int len, bool mark_jump) {
    __shared__ double s_d[3][8];
    __shared__ ll s_l[2][8];
    __shared__ int s_m[8];
    DevStatus *st = p.st;
    bool settled = true;   // nothing left to examine for k unless a false cycle is cut below
    if (x < 0) return settled;   // block-uniform: no representative left, or the chain broke while we walked
    double sc = 0.0, stt = 0.0, mabs = 0.0;
    ll isc = 0, ist = 0;
    int minv = 0x7fffffff;
    __threadfence_block();
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        const int e = __ldcg(cv.cyc_edges() + (ll)k * p.cap + i);
        const int u = __ldg(p.src + e);
        minv = u < minv ? u : minv;
        if (p.exact) { isc += __ldg(p.icost + e); ist += __ldg(p.itime + e); }
        else {
            double ce = __ldg(p.cost + e), te = __ldg(p.time + e);
            cv.scenario_patch(k, e, ce, te);
            sc += ce; stt += te;
            mabs = fmax(mabs, fabs(__ldcg(static_cast<const double *>(cv.dist()) + (ll)u * K + k)));
        }
    }
    for (int o = 16; o; o >>= 1) {
        sc += __shfl_down_sync(0xffffffffu, sc, o); stt += __shfl_down_sync(0xffffffffu, stt, o);
        mabs = fmax(mabs, __shfl_down_sync(0xffffffffu, mabs, o));
        isc += __shfl_down_sync(0xffffffffu, isc, o); ist += __shfl_down_sync(0xffffffffu, ist, o);
        const int mo = __shfl_down_sync(0xffffffffu, minv, o);
        minv = mo < minv ? mo : minv;
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_d[0][w] = sc; s_d[1][w] = stt; s_d[2][w] = mabs; s_l[0][w] = isc; s_l[1][w] = ist; s_m[w] = minv; }
    __syncthreads();
    if (threadIdx.x == 0) {
        sc = 0.0; stt = 0.0; mabs = 0.0; isc = 0; ist = 0; minv = 0x7fffffff;
        for (int j = 0; j < (int)(blockDim.x >> 5); j++) {
            sc += s_d[0][j]; stt += s_d[1][j]; mabs = fmax(mabs, s_d[2][j]);
            isc += s_l[0][j]; ist += s_l[1][j]; minv = s_m[j] < minv ? s_m[j] : minv;
        }
        int kind = 0;
        if (p.exact) {
            if ((__int128)cv.b(k) * isc - (__int128)cv.a(k) * ist < 0) kind = MRC_CYC_NEG;
        } else if (stt > 0) {
            // r < lambda: negative.  r >= lambda but (r - lambda)*T inside the rounding noise of the
            // walk (len * eps * max|dist|, x8): a rounding artefact that would relax forever -> TIE.
            const double rr = sc / stt;
            if (rr < cv.lam(k)) kind = MRC_CYC_NEG;
            else if ((rr - cv.lam(k)) * stt <= 8.0 * (double)len * 2.220446049250313e-16 * mabs) kind = MRC_CYC_TIE;
        }
        if (kind) {
            CycleOut o;
            o.found = kind; o.len = len; o.minv = minv; o.pad = x;   // pad = vertex the recorded walk starts from
            o.sumc = sc; o.sumt = stt; o.maxabs = mabs; o.isumc = isc; o.isumt = ist;
            st->out[k] = o;
            atomicOr(&st->decided, 1u << (k + cv.k0()));   // relaxation launches already enqueued skip this candidate
        } else {
            // false cycle: cut it for good (and mark its vertices dead for this launch of the doubling check)
            if (mark_jump) {
                int cur = minv;
                ll g = 0;
                do {
                    const int nxt = cycle_step<K>(cv, k, cur);
                    __stcg(cv.jump() + (ll)cur * K + k, -2);
                    cur = nxt;
                    g++;
                } while (cur != minv && cur >= 0 && g <= p.n);
            }
            __stcg(cv.pred() + (ll)minv * K + k, ~0ull);
            atomicAdd(&st->cuts, 1u);
            settled = false;
        }
    }
    __syncthreads();   // the shared arrays may be reused by the caller's next candidate
    return settled;
}

// Body of the check: every thread of a cooperative grid calls it (grid-uniform control flow); returns the mask of candidates
// whose predecessor graph held a cycle (results in st->out[], st->cyc_mask).  Shared by cycle_check_kernel and the
// persistent probe kernel (mrc_probe.cuh).
// Where a check finds its predecessor graph and its candidates: CheckParamView = the kernel parameters (host picked the
// columns), CheckRegView = the compact survivor column picked on the device by the persistent probe kernel.
struct CheckParamView {
    const CheckArgs &p;
    __device__ __forceinline__ ull *pred() const { return p.pred; }
    __device__ __forceinline__ const void *dist() const { return p.dist; }
    __device__ __forceinline__ int *jump() const { return p.jump; }
    __device__ __forceinline__ int *cyc_edges() const { return p.cyc_edges; }
    __device__ __forceinline__ unsigned active() const { return p.active; }
    __device__ __forceinline__ const unsigned *gate() const { return p.gate; }
    __device__ __forceinline__ int k0() const { return p.k0; }
    __device__ __forceinline__ int prune_above() const { return p.prune_above; }
    __device__ __forceinline__ double lam(int k) const { return p.lam[k]; }
    __device__ __forceinline__ ll a(int k) const { return p.a[k]; }
    __device__ __forceinline__ ll b(int k) const { return p.b[k]; }
    __device__ __forceinline__ void scenario_patch(int k, int e, double &ce, double &te) const {
        if (p.scen && e == p.pe[k]) { ce = __dadd_rn(ce, p.pdc[k]); te = __dadd_rn(te, p.pdt[k]); }
    }
};
struct CheckActiveView {   // the kernel parameters, but only the candidates in active_ (a mask read or kept on the device)
    const CheckArgs &p;
    unsigned active_;
    __device__ __forceinline__ ull *pred() const { return p.pred; }
    __device__ __forceinline__ const void *dist() const { return p.dist; }
    __device__ __forceinline__ int *jump() const { return p.jump; }
    __device__ __forceinline__ int *cyc_edges() const { return p.cyc_edges; }
    __device__ __forceinline__ unsigned active() const { return active_; }
    __device__ __forceinline__ const unsigned *gate() const { return nullptr; }
    __device__ __forceinline__ int k0() const { return p.k0; }
    __device__ __forceinline__ int prune_above() const { return p.prune_above; }
    __device__ __forceinline__ double lam(int k) const { return p.lam[k]; }
    __device__ __forceinline__ ll a(int k) const { return p.a[k]; }
    __device__ __forceinline__ ll b(int k) const { return p.b[k]; }
    __device__ __forceinline__ void scenario_patch(int k, int e, double &ce, double &te) const {
        if (p.scen && e == p.pe[k]) { ce = __dadd_rn(ce, p.pdc[k]); te = __dadd_rn(te, p.pdt[k]); }
    }
};
struct CheckRegView {   // one candidate (K = 1), results in slot 0
    ull *pred_;
    const void *dist_;
    int *jump_, *cyc_edges_;
    unsigned active_;
    int k0_, prune_;
    double lam_;
    ll a_, b_;
    __device__ __forceinline__ ull *pred() const { return pred_; }
    __device__ __forceinline__ const void *dist() const { return dist_; }
    __device__ __forceinline__ int *jump() const { return jump_; }
    __device__ __forceinline__ int *cyc_edges() const { return cyc_edges_; }
    __device__ __forceinline__ unsigned active() const { return active_; }
    __device__ __forceinline__ const unsigned *gate() const { return nullptr; }
    __device__ __forceinline__ int k0() const { return k0_; }
    __device__ __forceinline__ int prune_above() const { return prune_; }
    __device__ __forceinline__ double lam(int) const { return lam_; }
    __device__ __forceinline__ ll a(int) const { return a_; }
    __device__ __forceinline__ ll b(int) const { return b_; }
    __device__ __forceinline__ void scenario_patch(int, int, double &, double &) const {}
};
template <int K, typename CV>
__device__ __forceinline__ unsigned cycle_check_device(const CheckArgs &p, const CV &cv, cg::grid_group &grid) {
    const ll total = p.n * K;
    const ll tid0 = (ll)blockIdx.x * blockDim.x + threadIdx.x;
    const ll stride = (ll)gridDim.x * blockDim.x;
    DevStatus *st = p.st;

    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < MRC_MAX_ROUNDS * MRC_KMAX; i += blockDim.x) { st->cnt[i] = 0; st->minlive[i] = 0x7fffffff; }
    if (blockIdx.x == 0 && threadIdx.x < MRC_KMAX * MRC_MAX_EXTRACT) st->rep[threadIdx.x] = 0x7fffffff;
    const unsigned active = cv.gate() ? (cv.active() & (__ldcg(cv.gate()) >> cv.k0())) : cv.active();
    if (blockIdx.x == 0 && threadIdx.x < K && ((active >> threadIdx.x) & 1u)) {
        st->out[threadIdx.x].found = 0;
        st->out[threadIdx.x].len = 0;
    }
    if (tid0 == 0) { st->cyc_mask = 0; st->pend = 0; st->cuts = 0; }
    ull t_phase = 0;
    if (tid0 == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_phase));
    if (!active) return 0u;   // grid-uniform
    // jump = grandparent (two predecessor steps straight from pred, which nobody writes during this
    // phase): saves the separate init sweep, its grid sync and the src[] gather of an edge-only pred
    for (ll i = tid0; i < total; i += stride) {
        const int k = (int)(i % K);
        if (!((active >> k) & 1u)) continue;   // jump of a decided candidate is never read
        int j = -1;
        const int u = (int)(__ldcg(cv.pred() + i) >> 32);
        if (u >= 0) j = (int)(__ldcg(cv.pred() + (ll)u * K + k) >> 32);
        __stcg(cv.jump() + i, j);
    }
    grid.sync();

    // In a forest at least one live vertex reaches a root in every doubling round (the shallowest live
    // vertex of each chain), so a candidate whose live count did not drop in a round has only vertices
    // that lead into cycles left: stop doubling for it right there instead of running all log2(n) rounds.
    unsigned alive_mask = active;   // candidates with live vertices
    unsigned moving = active;       // ... whose live count is still dropping
    int r = 0;
    for (; r < p.rounds && moving; r++) {
        unsigned cnt[K];
        int low[K];   // smallest vertex still alive after this round: the representative of the first extraction below
#pragma unroll
        for (int k = 0; k < K; k++) { cnt[k] = 0; low[k] = 0x7fffffff; }
        for (ll i = tid0; i < total; i += stride) {
            const int k = (int)(i % K);
            if (!((alive_mask >> k) & 1u)) continue;
            const int j = __ldcg(cv.jump() + i);
            if (j >= 0) {
                int jj = j;
                if ((moving >> k) & 1u) {
                    jj = __ldcg(cv.jump() + (ll)j * K + k);
                    __stcg(cv.jump() + i, jj);
                }
                if (jj >= 0) {
                    const int v = (int)(i / K);
#pragma unroll
                    for (int kk = 0; kk < K; kk++)
                        if (kk == k) { cnt[kk]++; low[kk] = v < low[kk] ? v : low[kk]; }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (!((alive_mask >> k) & 1u)) continue;
            const unsigned c = __reduce_add_sync(0xffffffffu, cnt[k]);
            if (c) {   // warp-uniform
                const int lo = __reduce_min_sync(0xffffffffu, low[k]);
                if ((threadIdx.x & 31) == 0) { atomicAdd(&st->cnt[r * MRC_KMAX + k], c); atomicMin(&st->minlive[r * MRC_KMAX + k], lo); }
            }
        }
        grid.sync();
        unsigned na = 0, nm = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (!((alive_mask >> k) & 1u)) continue;
            const unsigned c = __ldcg(&st->cnt[r * MRC_KMAX + k]);
            const unsigned cp = r ? __ldcg(&st->cnt[(r - 1) * MRC_KMAX + k]) : 0xffffffffu;
            if (c) na |= 1u << k;
            if (c && c < cp && ((moving >> k) & 1u)) nm |= 1u << k;
        }
        if (cv.prune_above()) {
            // a candidate that is alive but no longer moving holds a cycle; a cycle negative at lambda_k is negative
            // at every larger lambda, so the (deeper, junk-cycle) predecessor graphs above it need no more rounds
#pragma unroll
            for (int k = 0; k < K; k++)
                if (((na >> k) & 1u) && !((nm >> k) & 1u)) { na &= (2u << k) - 1u; nm &= (2u << k) - 1u; break; }
        }
        alive_mask = na;
        moving = nm;
    }
    if (tid0 == 0) {
        st->rounds_run = (unsigned)r; st->pend = alive_mask;
        ull t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        st->ns_check[0] += t - t_phase; st->check_rounds += (ull)r;
        t_phase = t;
    }
    if (alive_mask == 0) return 0u;

    unsigned pending = alive_mask;   // candidates whose predecessor graph still holds an unexamined cycle
    const int r_last = r - 1;   // at least one doubling round ran
    for (int it = 0; it < MRC_MAX_EXTRACT && pending; it++) {
        if (it > 0) {   // grid-uniform.  it == 0: the smallest live vertex was found by the last doubling round itself
        int best[K];
#pragma unroll
        for (int k = 0; k < K; k++) best[k] = 0x7fffffff;
        for (ll i = tid0; i < total; i += stride) {
            const int k = (int)(i % K);
            if (!((pending >> k) & 1u)) continue;
            const int j = __ldcg(cv.jump() + i);
            if (j >= 0 && __ldcg(cv.jump() + (ll)j * K + k) != -2) {
                const int v = (int)(i / K);
#pragma unroll
                for (int kk = 0; kk < K; kk++)
                    if (kk == k && v < best[kk]) best[kk] = v;
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (!((pending >> k) & 1u)) continue;
            const int b = __reduce_min_sync(0xffffffffu, best[k]);
            if ((threadIdx.x & 31) == 0 && b != 0x7fffffff) atomicMin(&st->rep[it * MRC_KMAX + k], b);
        }
        grid.sync();
        }
        if (blockIdx.x < K && ((pending >> blockIdx.x) & 1u)) {   // block-uniform: block k serves candidate k
            const int k = blockIdx.x;
            __shared__ int s_x, s_len;
            if (threadIdx.x == 0) {
                const int rep = it == 0 ? __ldcg(&st->minlive[r_last * MRC_KMAX + k]) : __ldcg(&st->rep[it * MRC_KMAX + k]);
                int x = -1;
                ll len = 0;
                if (rep != 0x7fffffff) {
                    // rep leads into a cycle.  The doubled pointers land ON it in a few dozen hops (each hop is
                    // >= 2^(rounds+1) predecessor steps; doubling may have stalled after 2-3 rounds while the tail
                    // into the cycle is hundreds of steps long); a lap that returns to its start proves it.  If the
                    // tail is longer still, the point where the capped lap stopped is itself past the tail; only
                    // if even that fails Brent's cycle finding (one pointer step per iteration) runs from there.
                    int y = rep;
                    for (int t = 0; t < 96; t++) {   // (16 hops were measured: the capped laps of the fallback cost more, check 65 -> 80-93 us)
                        const int j = __ldcg(cv.jump() + (ll)y * K + k);
                        if (j < 0) break;
                        y = j;
                    }
                    int cur = cycle_lap<K>(p, cv, k, y, 2048, len);
                    if (cur == y) x = y;
                    else if (cur >= 0) {
                        y = cur;
                        cur = cycle_lap<K>(p, cv, k, y, 4096, len);
                        if (cur == y) x = y;
                        else if (cur >= 0) {
                            int hare = cycle_step<K>(cv, k, cur), tort = cur;
                            ll power = 1, lam_ = 1;
                            for (ll g = 0; hare >= 0 && hare != tort && g <= 3 * p.n; g++) {
                                if (power == lam_) { tort = hare; power <<= 1; lam_ = 0; }
                                hare = cycle_step<K>(cv, k, hare);
                                lam_++;
                            }
                            if (hare >= 0 && hare == tort && cycle_lap<K>(p, cv, k, hare, p.n + 1, len) == hare) x = hare;
                        }
                    }
                }
                s_x = x;
                s_len = x >= 0 ? (int)len : 0;
            }
            __syncthreads();
            const bool settled = cycle_verify_block<K>(p, cv, k, s_x, s_len, true);
            if (threadIdx.x == 0 && settled) atomicAnd(&st->pend, ~(1u << k));
        }
        grid.sync();
        pending = __ldcg(&st->pend);
    }
    if (tid0 == 0) {
        st->cyc_mask = alive_mask;
        ull t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        st->ns_check[1] += t - t_phase;
    }
    return alive_mask;
}

template <int K>
__global__ void __launch_bounds__(256) cycle_check_kernel(const __grid_constant__ CheckArgs p) {
    cg::grid_group grid = cg::this_grid();
    cycle_check_device<K>(p, CheckParamView{p}, grid);
}

// gather (src, cost, time) of a recorded cycle edge list for the host
__global__ void gather_cycle_kernel(const int *edges, ll len, const int *src, const double *cost,
                                    const double *time, const ll *icost, const ll *itime, int exact,
                                    int *o_src, double *o_c, double *o_t, ll *o_ic, ll *o_it) {
    // o_src[0..len) = source vertices, o_src[len..2len) = edge indices
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (ll)gridDim.x * blockDim.x) {
        const int e = edges[i];
        o_src[i] = src[e];
        o_src[len + i] = e;
        if (exact) { o_ic[i] = icost[e]; o_it[i] = itime[e]; }
        else { o_c[i] = cost[e]; o_t[i] = time[e]; }
    }
}

// ---- get_graph_properties (solver.py:339-387) on the resident arrays ----------------------------------------------
// acc layout (doubles): 0 cost sum, 1 time sum, 2 cost sum of squared deviations, 3 time ssd; keys (ull, order-preserving
// f64 keys): 0 cost min, 1 cost max, 2 time min, 3 time max; cnt (unsigned): 0 isolated, 1 sources, 2 sinks, 3 max in, 4 max out
__device__ __forceinline__ ull f64_key_of(double x) {
    ull b = (ull)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__global__ void props_edges_kernel(const int *src, const int *dst, const double *cost, const double *time, ll m,
                                   unsigned *indeg, unsigned *outdeg, double *acc, ull *keys) {
    double sc = 0.0, st = 0.0;
    ull cmin = ~0ull, cmax = 0ull, tmin = ~0ull, tmax = 0ull;
    for (ll e = (ll)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (ll)gridDim.x * blockDim.x) {
        atomicAdd(indeg + dst[e], 1u);
        atomicAdd(outdeg + src[e], 1u);
        const double c = cost[e], t = time[e];
        sc += c; st += t;
        const ull kc = f64_key_of(c), kt = f64_key_of(t);
        cmin = kc < cmin ? kc : cmin; cmax = kc > cmax ? kc : cmax;
        tmin = kt < tmin ? kt : tmin; tmax = kt > tmax ? kt : tmax;
    }
    for (int o = 16; o; o >>= 1) {
        sc += __shfl_xor_sync(0xffffffffu, sc, o); st += __shfl_xor_sync(0xffffffffu, st, o);
        ull x;
        x = __shfl_xor_sync(0xffffffffu, cmin, o); cmin = x < cmin ? x : cmin;
        x = __shfl_xor_sync(0xffffffffu, cmax, o); cmax = x > cmax ? x : cmax;
        x = __shfl_xor_sync(0xffffffffu, tmin, o); tmin = x < tmin ? x : tmin;
        x = __shfl_xor_sync(0xffffffffu, tmax, o); tmax = x > tmax ? x : tmax;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(acc + 0, sc); atomicAdd(acc + 1, st);
        atomicMin(keys + 0, cmin); atomicMax(keys + 1, cmax); atomicMin(keys + 2, tmin); atomicMax(keys + 3, tmax);
    }
}
__global__ void props_deviation_kernel(const double *cost, const double *time, ll m, double *acc) {
    const double mc = acc[0] / (double)m, mt = acc[1] / (double)m;
    double dc = 0.0, dt = 0.0;
    for (ll e = (ll)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (ll)gridDim.x * blockDim.x) {
        const double a = cost[e] - mc, b = time[e] - mt;
        dc += a * a; dt += b * b;
    }
    for (int o = 16; o; o >>= 1) { dc += __shfl_xor_sync(0xffffffffu, dc, o); dt += __shfl_xor_sync(0xffffffffu, dt, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(acc + 2, dc); atomicAdd(acc + 3, dt); }
}
__global__ void props_vertices_kernel(const unsigned *indeg, const unsigned *outdeg, ll n, unsigned *cnt) {
    unsigned iso = 0, so = 0, si = 0, mi = 0, mo = 0;
    for (ll v = (ll)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (ll)gridDim.x * blockDim.x) {
        const unsigned a = indeg[v], b = outdeg[v];
        iso += (a == 0 && b == 0); so += (a == 0 && b > 0); si += (a > 0 && b == 0);
        mi = a > mi ? a : mi; mo = b > mo ? b : mo;
    }
    iso = __reduce_add_sync(0xffffffffu, iso); so = __reduce_add_sync(0xffffffffu, so); si = __reduce_add_sync(0xffffffffu, si);
    mi = __reduce_max_sync(0xffffffffu, mi); mo = __reduce_max_sync(0xffffffffu, mo);
    if ((threadIdx.x & 31) == 0) {
        if (iso) atomicAdd(cnt + 0, iso);
        if (so) atomicAdd(cnt + 1, so);
        if (si) atomicAdd(cnt + 2, si);
        atomicMax(cnt + 3, mi); atomicMax(cnt + 4, mo);
    }
}

// equal_mask = (dist[U] + W) == dist[V], solver.py:878-879 (dist layout [n][1])
__global__ void tight_mask_kernel(const int *src, const int *dst, const ll *icost, const ll *itime,
                                  const ll *dist, ll a, ll b, ll m, uint8_t *mask) {
    for (ll e = (ll)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (ll)gridDim.x * blockDim.x) {
        const ll w = b * icost[e] - a * itime[e];
        mask[e] = (dist[src[e]] + w) == dist[dst[e]];
    }
}

// min / max of cost/time (default bracket, solver.py:1021-1023) via order-preserving keys
__device__ __forceinline__ ull f64_key(double x) {
    ull b = (ull)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__global__ void ratio_bounds_kernel(const double *cost, const double *time, ll m, ull *kmin, ull *kmax) {
    ull lo = ~0ull, hi = 0ull;
    for (ll e = (ll)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (ll)gridDim.x * blockDim.x) {
        const ull k = f64_key(__ddiv_rn(cost[e], time[e]));
        lo = k < lo ? k : lo;
        hi = k > hi ? k : hi;
    }
    for (int o = 16; o; o >>= 1) {
        const ull l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo;
        hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(kmin, lo); atomicMax(kmax, hi); }
}

// input checks of mrc_graph_create: bit0 endpoint out of range, bit1 not destination-sorted, bit2 time <= 0
__global__ void validate_edges_kernel(const int *src, const int *dst, const double *time, ll n, ll m, int *err) {
    int bad = 0;
    for (ll e = (ll)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (ll)gridDim.x * blockDim.x) {
        const int u = src[e], v = dst[e];
        if ((unsigned)u >= (unsigned)n || (unsigned)v >= (unsigned)n) bad |= 1;
        if (e > 0 && v < dst[e - 1]) bad |= 2;
        if (!(time[e] > 0.0)) bad |= 4;
    }
    bad = __reduce_or_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicOr(err, bad);
}

__global__ void patch_edges_kernel(double *cost, double *time, const ll *idx, const double *dc, const double *dt,
                                   ll k, double *old_c, double *old_t, int *err) {
    const ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) {
        const ll e = idx[i];
        old_c[i] = cost[e]; old_t[i] = time[e];
        cost[e] = __dadd_rn(cost[e], dc[i]);
        time[e] = __dadd_rn(time[e], dt[i]);
        if (!(time[e] > 0.0)) atomicOr(err, 1);
    }
}
__global__ void unpatch_edges_kernel(double *cost, double *time, const ll *idx, ll k, const double *old_c,
                                     const double *old_t) {
    const ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) { cost[idx[i]] = old_c[i]; time[idx[i]] = old_t[i]; }
}

// starts[v] = first position of the destination-sorted arrays whose destination is >= v, v in [0, n] (CSR by destination)
__global__ void in_offsets_kernel(const int *sorted_dst, ll m, ll n, int *starts) {
    for (ll v = (ll)blockIdx.x * blockDim.x + threadIdx.x; v <= n; v += (ll)gridDim.x * blockDim.x) {
        ll lo = 0, hi = m;
        while (lo < hi) {
            const ll mid = (lo + hi) >> 1;
            if (sorted_dst[mid] < (int)v) lo = mid + 1; else hi = mid;
        }
        starts[v] = (int)lo;
    }
}
