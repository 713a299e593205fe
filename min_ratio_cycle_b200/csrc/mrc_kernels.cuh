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
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mrc_b200.h"

namespace cg = cooperative_groups;
typedef long long ll;
typedef unsigned long long ull;

#define MRC_MAX_BURST 32
#ifndef MRC_CROWD_MIN
#define MRC_CROWD_MIN 6   /* improving lanes per row above which the segmented scan replaces the pairwise loop */
#endif
#define MRC_MAX_ROUNDS 40
#ifndef MRC_RELAX_THREADS
#define MRC_RELAX_THREADS 256
#endif

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
    // Convergence gate: the changed word of the PREVIOUS launch of the same burst (NULL for the first).  If that
    // launch lowered nothing for any candidate this launch serves, every one of them sits at its fixed point and
    // the launch returns at once -- the host enqueues whole bursts without looking, the device stops on the pass.
    const unsigned *gate;
    int reverse;         // sweep the edge stream back to front (L2 reuse between consecutive passes)
    // frontier mode (optional): one bit per vertex, "dist[v] was lowered for some candidate"
    unsigned *act_out;         // set by this launch (NULL = not tracked)
    const unsigned *act_in;    // set by the previous launch; read only by relax_kernel_frontier
    ull *examined;             // edges actually relaxed by relax_kernel_frontier launches
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
    ull examined;                              // frontier launches: edges whose source was active (cleared with `changed`)
    unsigned cnt[MRC_MAX_ROUNDS * MRC_KMAX];   // live vertices per doubling round and candidate
    unsigned cyc_mask, rounds_run, pend, cuts;
    int rep[MRC_KMAX * 6];   // [MRC_MAX_EXTRACT][MRC_KMAX]
    CycleOut out[MRC_KMAX];
    unsigned act_count;      // sparse mode: vertices lowered by the last relaxation pass (size of the next worklist)
    unsigned act_cnt[3];     // worklist sizes, rotating: pass p reads [p % 3], appends to [(p+1) % 3], zeroes [(p+2) % 3]
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
template <> struct DistVec<8> {
    template <typename T> static __device__ __forceinline__ void load(T *r, const T *p) {
        DistVec<4>::load(r, p); DistVec<4>::load(r + 4, p + 4);
    }
};

// ---- one row of 32 consecutive edges, K candidates ------------------------------------------------------
// numeric: w = cost - lam*time with a rounded product and a rounded difference (no FMA), exactly as
// NumPy evaluates `self._C - lam * self._T` (solver.py:1144); improve test `cand < dist[v] - slack`
// (:1161).  dist <= 0 always (starts at +0.0, only decreases), so for negative doubles
// "smaller value" == "larger bit pattern": atomicMax on the raw bits is a 64-bit atomic min.
// exact: w = b*icost - a*itime in int64 (solver.py:799-800), strict < without slack (:833).
__device__ __forceinline__ double mrc_cand(const RelaxArgs &p, int k, double du, double c, double t) {
    return __dadd_rn(du, __dsub_rn(c, __dmul_rn(p.lam[k], t)));
}
__device__ __forceinline__ ll mrc_cand(const RelaxArgs &p, int k, ll du, ll c, ll t) {
    return du + (p.b[k] * c - p.a[k] * t);
}
// scenario batching: candidate k re-solves the graph with ONE edge perturbed (sensitivity_analysis, solver.py:1716-1721)
__device__ __forceinline__ void mrc_scenario_patch(const RelaxArgs &p, int k, int e, double &c, double &t) {
    if (p.scen && e == p.pe[k]) { c = __dadd_rn(c, p.pdc[k]); t = __dadd_rn(t, p.pdt[k]); }
}
__device__ __forceinline__ void mrc_scenario_patch(const RelaxArgs &, int, int, ll &, ll &) {}
__device__ __forceinline__ bool mrc_improves(const RelaxArgs &p, double cand, double dv) {
    return cand < __dsub_rn(dv, p.slack);
}
__device__ __forceinline__ bool mrc_improves(const RelaxArgs &, ll cand, ll dv) { return cand < dv; }
__device__ __forceinline__ double mrc_identity(double) { return __longlong_as_double(0x7ff0000000000000ll); }
__device__ __forceinline__ ll mrc_identity(ll) { return 0x7fffffffffffffffll; }
__device__ __forceinline__ bool mrc_atomic_lower(double *addr, double cand) {
    const ull bits = (ull)__double_as_longlong(cand);
    return atomicMax(reinterpret_cast<ull *>(addr), bits) < bits;
}
__device__ __forceinline__ bool mrc_atomic_lower(ll *addr, ll cand) { return cand < atomicMin(addr, cand); }

// Crowded row (first passes: many lanes improve): segmented min over the runs of equal destination (prefix-min
// inside the run, read back from the run's last lane); the lowest lane holding the run minimum keeps `imp`.
template <typename T>
__device__ __forceinline__ bool mrc_resolve_crowded(bool imp, T cand, int v, int lane) {
    const int vprev = __shfl_up_sync(0xffffffffu, v, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || v != vprev);
    const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
    const unsigned above = (lane == 31) ? 0u : (heads & ~((2u << lane) - 1u));
    const int end = above ? (__ffs(above) - 2) : 31;
    // inclusive prefix-min inside the run (5 shuffle steps); the run's last lane then holds the run minimum
    T x = imp ? cand : mrc_identity(T(0));
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const T a = __shfl_up_sync(0xffffffffu, x, off);
        if (lane - off >= start && a < x) x = a;
    }
    const T runmin = __shfl_sync(0xffffffffu, x, end);
    const unsigned eq = __ballot_sync(0xffffffffu, imp && cand == runmin);
    const unsigned runmask = ((end == 31) ? 0xffffffffu : ((2u << end) - 1u)) & ~((1u << start) - 1u);
    return imp && (__ffs(eq & runmask) - 1 == lane);
}

// The row is warp-synchronous: every lane calls it (valid == false for lanes past the end).  Several
// in-edges of one destination sit in adjacent lanes, and in the first passes several of them would lower
// dist[v] in the same instruction; if they all issued their atomic, the pred[v] store of a worse
// candidate could land last, pred would name an edge that did NOT produce dist[v], and the predecessor
// graph could close cycles that are not negative.  So per destination only the lane with the smallest
// candidate of the row issues the atomic (which also removes most same-address atomics of the early
// passes).  Control flow never waits for an atomic's return value; rows in which nothing improves
// (almost all rows after the first passes) cost one ballot per candidate.
template <int K, typename T>
__device__ __forceinline__ void relax_row(const RelaxArgs &p, bool valid, int lane, int e, int u, int v, T c, T t,
                                          unsigned &chg) {
    // K = 8 is processed as two halves of 4 candidates (two 256-bit gathers from the same 64-B group, one after
    // the other), so its register footprint stays that of K = 4 and 4 CTAs remain resident per SM.
    constexpr int H = K > 4 ? 4 : K;
    T *dist = static_cast<T *>(p.dist);
#pragma unroll
    for (int h = 0; h < K; h += H) {
    if (!((p.active >> h) & ((1u << H) - 1u))) continue;   // warp-uniform: nothing active in this half
    T du[H], dv[H];
    if (valid) {
        DistVec<H>::load(du, dist + (size_t)u * p.ks + h);
        DistVec<H>::load(dv, dist + (size_t)v * p.ks + h);
    }
#pragma unroll
    for (int k = h; k < h + H; k++) {
        if (!((p.active >> k) & 1u)) continue;   // warp-uniform
        T cand = 0;
        bool imp = false;
        if (valid) {
            T ck = c, tk = t;
            mrc_scenario_patch(p, k, e, ck, tk);
            cand = mrc_cand(p, k, du[k - h], ck, tk);
            imp = mrc_improves(p, cand, dv[k - h]);
        }
        const unsigned im = __ballot_sync(0xffffffffu, imp);
        if (im == 0) continue;                   // warp-uniform
        const int nimp = __popc(im);
        if (nimp > MRC_CROWD_MIN) {
            imp = mrc_resolve_crowded<T>(imp, cand, v, lane);
        } else if (nimp > 1) {
            bool lose = false;
            unsigned rest = im;
            while (rest) {
                const int l = __ffs(rest) - 1;
                rest &= rest - 1;
                const int vl = __shfl_sync(0xffffffffu, v, l);
                const T cl = __shfl_sync(0xffffffffu, cand, l);
                if (l != lane && vl == v && (cl < cand || (cl == cand && l < lane))) lose = true;
            }
            imp = imp && !lose;
        }
        if (imp && mrc_atomic_lower(dist + (size_t)v * p.ks + k, cand)) {
            __stcg(p.pred + (size_t)v * p.ks + k, ((ull)(unsigned)u << 32) | (unsigned)e);
            if (p.act_out) atomicOr(p.act_out + (v >> 5), 1u << (v & 31));
            chg |= 1u << k;
        }
    }
    }
}

template <bool EXACT> struct WeightT { typedef double type; };
template <> struct WeightT<true> { typedef ll type; };

// Edge-parallel relaxation.  A warp owns MRC_ROWS rows of 32 CONSECUTIVE edges per step; lane i takes edge
// base + 32*j + i, so every edge-stream load is one fully coalesced 128/256-B request and -- because
// the edges are destination-sorted -- the 32 dist[dst] reads of a row collapse to one or two lines.
// What remains is exactly one random line per edge (dist[src]); the kernel is bound by that L1TEX
// gather rate and by the 24 B/edge HBM stream.  Persistent grid: occupancy x 148 SMs, warps walk
// consecutive chunks so DRAM pages are streamed in order.
// Tuned on B200 (tools/relax_variants.py, profiles/r01_relax_variants.txt): K <= 4 wants 2 rows per warp
// step and 4 resident CTAs per SM (64 registers); K = 8 carries 2 x 64 B of dist per edge and is faster
// with 4 rows and no register cap.
#ifndef MRC_TUNE_ROWS
#define MRC_TUNE_ROWS 2
#endif
#ifndef MRC_TUNE_MB
#define MRC_TUNE_MB 4
#endif
#ifndef MRC_TUNE_PF
#define MRC_TUNE_PF (K <= 2)
#endif
#ifndef MRC_TUNE_ROWS8
#define MRC_TUNE_ROWS8 2
#endif
#ifndef MRC_TUNE_MB8
#define MRC_TUNE_MB8 4
#endif
#ifndef MRC_TUNE_STAGE
#define MRC_TUNE_STAGE false   /* measured slower than plain loads at every K (profiles/r01_relax_variants.txt) */
#endif
#ifndef MRC_TUNE_DEPTH
#define MRC_TUNE_DEPTH 3
#endif
template <int K> struct RelaxTune {
    static constexpr int rows = MRC_TUNE_ROWS, minblocks = MRC_TUNE_MB;
    static constexpr bool prefetch = MRC_TUNE_PF;
    static constexpr bool stage = MRC_TUNE_STAGE;   // edge stream through a per-warp cp.async ring in shared memory
};
template <> struct RelaxTune<8> {
    static constexpr int rows = MRC_TUNE_ROWS8, minblocks = MRC_TUNE_MB8;
    static constexpr bool prefetch = false;
    static constexpr bool stage = false;
};
__device__ __forceinline__ void mrc_cp_async4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void mrc_cp_async8(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void mrc_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void mrc_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int K, bool EXACT>
__global__ void __launch_bounds__(MRC_RELAX_THREADS, RelaxTune<K>::minblocks) relax_kernel(const __grid_constant__ RelaxArgs p) {
    typedef typename WeightT<EXACT>::type W;
    const W *cost = EXACT ? reinterpret_cast<const W *>(p.icost) : reinterpret_cast<const W *>(p.cost);
    const W *time = EXACT ? reinterpret_cast<const W *>(p.itime) : reinterpret_cast<const W *>(p.time);
    if (p.gate && !((__ldcg(p.gate) >> p.k0) & p.active)) return;
    unsigned chg = 0;
    const int lane = threadIdx.x & 31;
    const ll warp = ((ll)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    constexpr int MRC_ROWS = RelaxTune<K>::rows;
    constexpr ll CH = 32 * MRC_ROWS;
    const ll nchunks = (p.m + CH - 1) / CH;
    // K <= 2 (registers to spare): the edge stream of the NEXT chunk is requested before this chunk's
    // gathers are issued, so the HBM latency of the stream overlaps the L2 latency of the dist gathers.
    // (Measured on B200, profiles/r01_relax_variants.txt: -8% per pass at K=1; at K>=4 the extra registers
    // cost more occupancy than the overlap wins.  Hoisting all gathers of a chunk ahead of the first use
    // was also measured and is slower at every K.)
    constexpr bool PREFETCH = RelaxTune<K>::prefetch;
    if constexpr (RelaxTune<K>::stage) {
        // EXPERIMENT, off by default.  Stall sampling put 2/3 of the memory stalls on the first use of the edge
        // stream, not on the gathers, so here every warp keeps its own DEPTH-deep ring of chunks in shared
        // memory, filled with cp.async (LDGSTS: no registers, no block barrier).  Measured: 88 us (K=4) and
        // 71 us (K=1) per launch against 74 / 58 us for plain loads on the same box -- the LDGSTS issue cost and
        // the extra shared-memory round trip outweigh the overlap.
        constexpr int D = MRC_TUNE_DEPTH;
        constexpr int WARPS = MRC_RELAX_THREADS / 32;
        __shared__ int s_src[WARPS][D][MRC_ROWS][32], s_dst[WARPS][D][MRC_ROWS][32];
        __shared__ W s_cost[WARPS][D][MRC_ROWS][32], s_time[WARPS][D][MRC_ROWS][32];
        const int w = threadIdx.x >> 5;
        auto issue = [&](ll c, int slot) {
            if (c < nchunks) {
                const ll e0 = (p.reverse ? (nchunks - 1 - c) : c) * CH + lane;
#pragma unroll
                for (int j = 0; j < MRC_ROWS; j++) {
                    const ll e = e0 + 32 * j;
                    if (e < p.m) {
                        mrc_cp_async4(&s_src[w][slot][j][lane], p.src + e);
                        mrc_cp_async4(&s_dst[w][slot][j][lane], p.dst + e);
                        mrc_cp_async8(&s_cost[w][slot][j][lane], cost + e);
                        mrc_cp_async8(&s_time[w][slot][j][lane], time + e);
                    }
                }
            }
            mrc_cp_async_commit();
        };
#pragma unroll
        for (int d = 0; d < D - 1; d++) issue(warp + (ll)d * nwarps, d);
        int slot = 0;
        for (ll c = warp; c < nchunks; c += nwarps) {
            int nslot = slot + D - 1;
            if (nslot >= D) nslot -= D;
            issue(c + (ll)(D - 1) * nwarps, nslot);
            mrc_cp_async_wait<D - 1>();
            __syncwarp();
            const ll e0 = (p.reverse ? (nchunks - 1 - c) : c) * CH + lane;
            if (e0 - lane + CH <= p.m) {
#pragma unroll
                for (int j = 0; j < MRC_ROWS; j++)
                    relax_row<K, W>(p, true, lane, (int)(e0 + 32 * j), s_src[w][slot][j][lane], s_dst[w][slot][j][lane],
                                    s_cost[w][slot][j][lane], s_time[w][slot][j][lane], chg);
            } else {
#pragma unroll
                for (int j = 0; j < MRC_ROWS; j++) {
                    const ll e = e0 + 32 * j;
                    const bool ok = e < p.m;
                    relax_row<K, W>(p, ok, lane, (int)e, ok ? s_src[w][slot][j][lane] : 0,
                                    ok ? s_dst[w][slot][j][lane] : -1 - lane, ok ? s_cost[w][slot][j][lane] : W(0),
                                    ok ? s_time[w][slot][j][lane] : W(0), chg);
                }
            }
            __syncwarp();
            slot = slot + 1 == D ? 0 : slot + 1;
        }
    } else if constexpr (!PREFETCH) {
        for (ll c = warp; c < nchunks; c += nwarps) {
            const ll e0 = (p.reverse ? (nchunks - 1 - c) : c) * CH + lane;
            if (e0 - lane + CH <= p.m) {
                int s[MRC_ROWS], d[MRC_ROWS];
                W cs[MRC_ROWS], tm[MRC_ROWS];
#pragma unroll
                for (int j = 0; j < MRC_ROWS; j++) {
                    const ll e = e0 + 32 * j;
                    s[j] = __ldcs(p.src + e); d[j] = __ldcs(p.dst + e);
                    cs[j] = __ldcs(cost + e); tm[j] = __ldcs(time + e);
                }
#pragma unroll
                for (int j = 0; j < MRC_ROWS; j++)
                    relax_row<K, W>(p, true, lane, (int)(e0 + 32 * j), s[j], d[j], cs[j], tm[j], chg);
            } else {
#pragma unroll
                for (int j = 0; j < MRC_ROWS; j++) {
                    const ll e = e0 + 32 * j;
                    const bool ok = e < p.m;
                    relax_row<K, W>(p, ok, lane, (int)e, ok ? p.src[e] : 0, ok ? p.dst[e] : -1 - lane,
                                    ok ? cost[e] : W(0), ok ? time[e] : W(0), chg);
                }
            }
        }
    } else {
    int s[MRC_ROWS], d[MRC_ROWS], s2[MRC_ROWS], d2[MRC_ROWS];
    W cs[MRC_ROWS], tm[MRC_ROWS], cs2[MRC_ROWS], tm2[MRC_ROWS];
    auto chunk_base = [&](ll c) { return (p.reverse ? (nchunks - 1 - c) : c) * CH; };
    auto is_full = [&](ll c) { return c < nchunks && chunk_base(c) + CH <= p.m; };
    ll c = warp;
    bool have = PREFETCH && is_full(c);
    if (have) {
        const ll e0 = chunk_base(c) + lane;
#pragma unroll
        for (int j = 0; j < MRC_ROWS; j++) {
            const ll e = e0 + 32 * j;
            s[j] = __ldcs(p.src + e); d[j] = __ldcs(p.dst + e);
            cs[j] = __ldcs(cost + e); tm[j] = __ldcs(time + e);
        }
    }
    while (c < nchunks) {
        const ll nx = c + nwarps;
        const bool have2 = PREFETCH && is_full(nx);
        if (have2) {
            const ll e0 = chunk_base(nx) + lane;
#pragma unroll
            for (int j = 0; j < MRC_ROWS; j++) {
                const ll e = e0 + 32 * j;
                s2[j] = __ldcs(p.src + e); d2[j] = __ldcs(p.dst + e);
                cs2[j] = __ldcs(cost + e); tm2[j] = __ldcs(time + e);
            }
        }
        const ll e0 = chunk_base(c) + lane;
        if (e0 - lane + CH <= p.m) {   // full chunk: no per-lane validity predicate anywhere
            if (!have) {
#pragma unroll
                for (int j = 0; j < MRC_ROWS; j++) {
                    const ll e = e0 + 32 * j;
                    s[j] = __ldcs(p.src + e); d[j] = __ldcs(p.dst + e);
                    cs[j] = __ldcs(cost + e); tm[j] = __ldcs(time + e);
                }
            }
#pragma unroll
            for (int j = 0; j < MRC_ROWS; j++)
                relax_row<K, W>(p, true, lane, (int)(e0 + 32 * j), s[j], d[j], cs[j], tm[j], chg);
        } else {
#pragma unroll
            for (int j = 0; j < MRC_ROWS; j++) {
                const ll e = e0 + 32 * j;
                const bool ok = e < p.m;
                relax_row<K, W>(p, ok, lane, (int)e, ok ? p.src[e] : 0, ok ? p.dst[e] : -1 - lane, ok ? cost[e] : W(0),
                                ok ? time[e] : W(0), chg);
            }
        }
        if (PREFETCH) {
#pragma unroll
            for (int j = 0; j < MRC_ROWS; j++) { s[j] = s2[j]; d[j] = d2[j]; cs[j] = cs2[j]; tm[j] = tm2[j]; }
        }
        have = have2;
        c = nx;
    }
    }
    // changed flag: warp vote, one atomic per warp at most
    chg = __reduce_or_sync(0xffffffffu, chg);
    if (lane == 0 && chg) atomicOr(p.changed, chg << p.k0);
}

// ---- frontier relaxation ------------------------------------------------------------------------------
// After the first passes only a few per cent of the vertices still change, and an edge can only improve
// its destination if its SOURCE was lowered since the edge was last looked at.  This variant keeps the
// edge-parallel sweep but consults a one-bit-per-vertex "lowered in the previous launch" map, held in
// shared memory (n/8 bytes, one 1024-thread CTA per SM): a row of 32 edges whose sources are all inactive
// costs 128 B of src stream and nothing else; dst/cost/time and the dist gathers are issued for active lanes
// only.  A launch that lowers nothing is still a true fixed point (an edge skipped now was examined after its
// source's last change), so verdicts, potentials and parity are unchanged.  Requires n <= 8 * (smem - 1 KB).
#define MRC_FRONTIER_THREADS 1024
#ifndef MRC_FRONTIER_ROWS
#define MRC_FRONTIER_ROWS 8
#endif
template <int K, bool EXACT>
__global__ void __launch_bounds__(MRC_FRONTIER_THREADS, 1) relax_kernel_frontier(const __grid_constant__ RelaxArgs p) {
    typedef typename WeightT<EXACT>::type W;
    extern __shared__ __align__(16) unsigned s_act[];
    const W *cost = EXACT ? reinterpret_cast<const W *>(p.icost) : reinterpret_cast<const W *>(p.cost);
    const W *time = EXACT ? reinterpret_cast<const W *>(p.itime) : reinterpret_cast<const W *>(p.time);
    const int words = (int)((p.n + 31) >> 5);
    for (int i = threadIdx.x; i < words; i += MRC_FRONTIER_THREADS) s_act[i] = __ldcg(p.act_in + i);
    __syncthreads();
    unsigned chg = 0, seen = 0;
    const int lane = threadIdx.x & 31;
    const ll warp = ((ll)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    constexpr int R = MRC_FRONTIER_ROWS;
    constexpr ll CH = 32 * R;
    const ll nchunks = (p.m + CH - 1) / CH;
    for (ll c = warp; c < nchunks; c += nwarps) {
        const ll e0 = (p.reverse ? (nchunks - 1 - c) : c) * CH + lane;
        int s[R];
        bool act[R];
#pragma unroll
        for (int j = 0; j < R; j++) {
            const ll e = e0 + 32 * j;
            s[j] = e < p.m ? __ldcs(p.src + e) : -1;
        }
#pragma unroll
        for (int j = 0; j < R; j++) act[j] = s[j] >= 0 && ((s_act[s[j] >> 5] >> (s[j] & 31)) & 1u);
#pragma unroll
        for (int j = 0; j < R; j++) {
            const unsigned am = __ballot_sync(0xffffffffu, act[j]);
            if (am == 0) continue;   // warp-uniform: the whole row is skipped
            seen += __popc(am);
            const ll e = e0 + 32 * j;
            int d = -1 - lane;
            W cs = W(0), tm = W(0);
            if (act[j]) { d = __ldcs(p.dst + e); cs = __ldcs(cost + e); tm = __ldcs(time + e); }
            relax_row<K, W>(p, act[j], lane, (int)e, s[j], d, cs, tm, chg);
        }
    }
    chg = __reduce_or_sync(0xffffffffu, chg);
    if (lane == 0) {
        if (chg) atomicOr(p.changed, chg << p.k0);
        if (seen) atomicAdd(p.examined, (ull)seen);
    }
}

// dist = 0, pred = none for the candidate slots in `mask` only (asynchronous multisection refills one slot while the
// others keep relaxing)
__global__ void reset_slots_kernel(void *dist, ull *pred, ll n, int KT, unsigned mask) {
    ull *d = static_cast<ull *>(dist);   // +0.0 and int64 0 are the same bits
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < n * KT; i += (ll)gridDim.x * blockDim.x) {
        if ((mask >> (int)(i % KT)) & 1u) { __stcg(d + i, 0ull); __stcg(pred + i, ~0ull); }
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

template <bool EXACT>
__global__ void __launch_bounds__(256) relax_sparse_kernel(const __grid_constant__ SparseArgs a) {
    typedef typename WeightT<EXACT>::type W;
    cg::grid_group grid = cg::this_grid();
    const RelaxArgs &p = a.r;
    const W *cost = EXACT ? reinterpret_cast<const W *>(p.icost) : reinterpret_cast<const W *>(p.cost);
    const W *time = EXACT ? reinterpret_cast<const W *>(p.itime) : reinterpret_cast<const W *>(p.time);
    W *dist = static_cast<W *>(p.dist);
    DevStatus *st = a.st;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int sub = tid & 7;                                   // 8 lanes share one worklist vertex
    const int grp = tid >> 3, ngrp = (gridDim.x * blockDim.x) >> 3;
    int in = a.first, ci = a.cidx;
    for (int pass = 0; pass < a.npasses; pass++) {
        const int out = in ^ 1, co = ci == 2 ? 0 : ci + 1, cz = co == 2 ? 0 : co + 1;
        const unsigned cin = __ldcg(&st->act_cnt[ci]);          // final since the previous grid sync
        if (cin == 0) break;                                    // grid-uniform: fixed point
        if (tid == 0) st->act_cnt[cz] = 0;                      // next pass appends there; last read one pass ago
        unsigned chg = 0, seen = 0;
        for (unsigned i = grp; i < cin; i += ngrp) {
            const int u = __ldcg(a.wl[in] + i);
            if (sub == 0) atomicAnd(a.bm[in] + (u >> 5), ~(1u << (u & 31)));   // leave the bitmap clean for its next use
            W du[MRC_KMAX];
#pragma unroll
            for (int k = 0; k < MRC_KMAX; k++)
                du[k] = (k < a.kc && ((p.active >> k) & 1u)) ? __ldcg(dist + (size_t)u * p.ks + k) : W(0);
            const int beg = __ldg(a.out_off + u), end = __ldg(a.out_off + u + 1);
            for (int j = beg + sub; j < end; j += 8) {
                const int e = __ldg(a.out_edge + j);
                const int v = __ldg(p.dst + e);
                const W c = __ldg(cost + e), t = __ldg(time + e);
                seen++;
                bool any = false;
#pragma unroll
                for (int k = 0; k < MRC_KMAX; k++) {
                    if (k >= a.kc || !((p.active >> k) & 1u)) continue;
                    W ck = c, tk = t;
                    mrc_scenario_patch(p, k, e, ck, tk);
                    const W cand = mrc_cand(p, k, du[k], ck, tk);
                    const W dv = __ldcg(dist + (size_t)v * p.ks + k);
                    if (mrc_improves(p, cand, dv) && mrc_atomic_lower(dist + (size_t)v * p.ks + k, cand)) {
                        __stcg(p.pred + (size_t)v * p.ks + k, ((ull)(unsigned)u << 32) | (unsigned)e);
                        chg |= 1u << k;
                        any = true;
                    }
                }
                if (any) {
                    const unsigned bit = 1u << (v & 31);
                    if (!(atomicOr(a.bm[out] + (v >> 5), bit) & bit)) a.wl[out][atomicAdd(&st->act_cnt[co], 1u)] = v;
                }
            }
        }
        chg = __reduce_or_sync(0xffffffffu, chg);
        seen = __reduce_add_sync(0xffffffffu, seen);
        if ((threadIdx.x & 31) == 0) {
            if (chg) atomicOr(p.changed + pass, chg << p.k0);
            if (seen) atomicAdd(p.examined, (ull)seen);
        }
        grid.sync();
        in = out; ci = co;
    }
    if (tid == 0) st->act_count = st->act_cnt[ci];
}

// ---- TMA-staged relaxation (sm_100a: cp.async.bulk + mbarrier) ------------------------------------------
// Same relaxation as relax_kernel, but the edge stream never touches registers or the LSU on its way in:
// one elected thread per CTA issues 1-D bulk copies (cp.async.bulk.shared::cluster.global, SASS UBLKCP) of
// the four edge arrays of a 1024-edge tile into a 3-stage shared-memory ring; completion is signalled on
// an mbarrier (complete_tx::bytes).  The copy engine runs 2 tiles (48 KB per CTA, 3 CTAs per SM) ahead of
// the warps, so HBM latency is fully decoupled from the dist gathers, and the threads keep all their
// registers for the K-wide gather/compare/atomic part.
#ifndef MRC_TMA_TILE
#define MRC_TMA_TILE 1024
#endif
#ifndef MRC_TMA_STAGES
#define MRC_TMA_STAGES 3
#endif
#ifndef MRC_TMA_MB
#define MRC_TMA_MB 3
#endif
#define MRC_TMA_THREADS 256
#define MRC_TMA_STAGE_BYTES (MRC_TMA_TILE * 24)
#define MRC_TMA_SMEM (MRC_TMA_STAGES * MRC_TMA_STAGE_BYTES + 64)

__device__ __forceinline__ unsigned mrc_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mrc_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mrc_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mrc_mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mrc_bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int K, bool EXACT>
__global__ void __launch_bounds__(MRC_TMA_THREADS, MRC_TMA_MB) relax_kernel_tma(const __grid_constant__ RelaxArgs p) {
    typedef typename WeightT<EXACT>::type W;
    extern __shared__ __align__(128) unsigned char smem[];
    const W *cost = EXACT ? reinterpret_cast<const W *>(p.icost) : reinterpret_cast<const W *>(p.cost);
    const W *time = EXACT ? reinterpret_cast<const W *>(p.itime) : reinterpret_cast<const W *>(p.time);
    constexpr int T = MRC_TMA_TILE, S = MRC_TMA_STAGES;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const ll ntiles = (p.m + T - 1) / T;
    const unsigned bar0 = mrc_smem_u32(smem + S * MRC_TMA_STAGE_BYTES);
    auto tile_of = [&](ll i) { return p.reverse ? (ntiles - 1 - i) : i; };
    auto issue = [&](ll i, int slot) {   // elected thread: 4 bulk copies, one barrier
        const ll e0 = tile_of(i) * T;
        ll cnt = p.m - e0;
        if (cnt > T) cnt = T;
        const unsigned c4 = (unsigned)((cnt + 3) & ~3ll);   // arrays are padded to a multiple of 4 edges
        const unsigned bar = bar0 + 8u * slot;
        const unsigned base = mrc_smem_u32(smem + slot * MRC_TMA_STAGE_BYTES);
        mrc_mbar_expect_tx(bar, c4 * 24u);
        mrc_bulk_g2s(base, p.src + e0, c4 * 4u, bar);
        mrc_bulk_g2s(base + T * 4, p.dst + e0, c4 * 4u, bar);
        mrc_bulk_g2s(base + T * 8, cost + e0, c4 * 8u, bar);
        mrc_bulk_g2s(base + T * 16, time + e0, c4 * 8u, bar);
    };
    if (tid == 0) {
        for (int s_ = 0; s_ < S; s_++) mrc_mbar_init(bar0 + 8u * s_, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int s_ = 0; s_ < S; s_++) {
            const ll i = (ll)blockIdx.x + (ll)s_ * gridDim.x;
            if (i < ntiles) issue(i, s_);
        }
    unsigned chg = 0;
    for (ll it = 0;; it++) {
        const ll i = (ll)blockIdx.x + it * gridDim.x;
        if (i >= ntiles) break;
        const int slot = (int)(it % S);
        mrc_mbar_wait(bar0 + 8u * slot, (unsigned)((it / S) & 1));
        const unsigned char *st = smem + slot * MRC_TMA_STAGE_BYTES;
        const int *ssrc = reinterpret_cast<const int *>(st);
        const int *sdst = reinterpret_cast<const int *>(st + T * 4);
        const W *scost = reinterpret_cast<const W *>(st + T * 8);
        const W *stime = reinterpret_cast<const W *>(st + T * 16);
        const ll e0 = tile_of(i) * T;
#pragma unroll
        for (int r = 0; r < T / 32 / (MRC_TMA_THREADS / 32); r++) {
            const int row = wid + r * (MRC_TMA_THREADS / 32);
            const int o = row * 32 + lane;
            const ll e = e0 + o;
            const bool ok = e < p.m;
            relax_row<K, W>(p, ok, lane, (int)e, ssrc[o], ok ? sdst[o] : -1 - lane, scost[o], stime[o], chg);
        }
        __syncthreads();   // every warp is done with this slot
        if (tid == 0) {
            const ll nx = i + (ll)S * gridDim.x;
            if (nx < ntiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(nx, slot);
            }
        }
    }
    chg = __reduce_or_sync(0xffffffffu, chg);
    if (lane == 0 && chg) atomicOr(p.changed, chg << p.k0);
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
template <int K>
__global__ void __launch_bounds__(256) cycle_check_kernel(const __grid_constant__ CheckArgs p) {
    cg::grid_group grid = cg::this_grid();
    const ll total = p.n * K;
    const ll tid0 = (ll)blockIdx.x * blockDim.x + threadIdx.x;
    const ll stride = (ll)gridDim.x * blockDim.x;
    DevStatus *st = p.st;

    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < MRC_MAX_ROUNDS * MRC_KMAX; i += blockDim.x) st->cnt[i] = 0;
    if (blockIdx.x == 0 && threadIdx.x < MRC_KMAX * MRC_MAX_EXTRACT) st->rep[threadIdx.x] = 0x7fffffff;
    if (blockIdx.x == 0 && threadIdx.x < MRC_KMAX) {
        st->out[threadIdx.x].found = 0;
        st->out[threadIdx.x].len = 0;
    }
    if (tid0 == 0) { st->cyc_mask = 0; st->pend = 0; st->cuts = 0; }
    const unsigned active = p.gate ? (p.active & __ldcg(p.gate)) : p.active;
    if (!active) return;   // grid-uniform
    // jump = grandparent (two predecessor steps straight from pred, which nobody writes during this
    // phase): saves the separate init sweep, its grid sync and the src[] gather of an edge-only pred
    for (ll i = tid0; i < total; i += stride) {
        const int k = (int)(i % K);
        if (!((active >> k) & 1u)) continue;   // jump of a decided candidate is never read
        int j = -1;
        const int u = (int)(__ldcg(p.pred + i) >> 32);
        if (u >= 0) j = (int)(__ldcg(p.pred + (ll)u * K + k) >> 32);
        __stcg(p.jump + i, j);
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
#pragma unroll
        for (int k = 0; k < K; k++) cnt[k] = 0;
        for (ll i = tid0; i < total; i += stride) {
            const int k = (int)(i % K);
            if (!((alive_mask >> k) & 1u)) continue;
            const int j = __ldcg(p.jump + i);
            if (j >= 0) {
                int jj = j;
                if ((moving >> k) & 1u) {
                    jj = __ldcg(p.jump + (ll)j * K + k);
                    __stcg(p.jump + i, jj);
                }
                if (jj >= 0) {
#pragma unroll
                    for (int kk = 0; kk < K; kk++) cnt[kk] += (kk == k);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (!((alive_mask >> k) & 1u)) continue;
            const unsigned c = __reduce_add_sync(0xffffffffu, cnt[k]);
            if ((threadIdx.x & 31) == 0 && c) atomicAdd(&st->cnt[r * MRC_KMAX + k], c);
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
        if (p.prune_above) {
            // a candidate that is alive but no longer moving holds a cycle; a cycle negative at lambda_k is negative
            // at every larger lambda, so the (deeper, junk-cycle) predecessor graphs above it need no more rounds
#pragma unroll
            for (int k = 0; k < K; k++)
                if (((na >> k) & 1u) && !((nm >> k) & 1u)) { na &= (2u << k) - 1u; nm &= (2u << k) - 1u; break; }
        }
        alive_mask = na;
        moving = nm;
    }
    if (tid0 == 0) { st->rounds_run = (unsigned)r; st->pend = alive_mask; }
    if (alive_mask == 0) return;

    unsigned pending = alive_mask;   // candidates whose predecessor graph still holds an unexamined cycle
    for (int it = 0; it < MRC_MAX_EXTRACT && pending; it++) {
        int best[K];
#pragma unroll
        for (int k = 0; k < K; k++) best[k] = 0x7fffffff;
        for (ll i = tid0; i < total; i += stride) {
            const int k = (int)(i % K);
            if (!((pending >> k) & 1u)) continue;
            const int j = __ldcg(p.jump + i);
            if (j >= 0 && __ldcg(p.jump + (ll)j * K + k) != -2) {
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
        if (blockIdx.x < K && ((pending >> blockIdx.x) & 1u)) {   // block-uniform: block k serves candidate k
            const int k = blockIdx.x;
            __shared__ int s_x, s_len;
            __shared__ double s_d[3][8];
            __shared__ ll s_l[2][8];
            __shared__ int s_m[8];
            auto step = [&](int y) { return y < 0 ? -1 : (int)(__ldcg(p.pred + (ll)y * K + k) >> 32); };
            // one lap over pred from y, recording the edge list; returns the vertex it stopped at
            auto lap = [&](int y, ll cap, ll &len) {
                int cur = y;
                ll i = 0;
                do {
                    const ull pe = __ldcg(p.pred + (ll)cur * K + k);
                    if (i < p.cap) p.cyc_edges[(ll)k * p.cap + i] = (int)(unsigned)pe;
                    cur = (int)(pe >> 32);
                    i++;
                } while (cur != y && cur >= 0 && i < cap);
                len = i;
                return cur;
            };
            if (threadIdx.x == 0) {
                const int rep = __ldcg(&st->rep[it * MRC_KMAX + k]);
                int x = -1;
                ll len = 0;
                if (rep != 0x7fffffff) {
                    // rep leads into a cycle.  The doubled pointers land ON it in a few dozen hops (each hop is
                    // >= 2^(rounds+1) predecessor steps; doubling may have stalled after 2-3 rounds while the tail
                    // into the cycle is hundreds of steps long); a lap that returns to its start proves it.  If the
                    // tail is longer still, the point where the capped lap stopped is itself past the tail; only
                    // if even that fails Brent's cycle finding (one pointer step per iteration) runs from there.
                    int y = rep;
                    for (int t = 0; t < 96; t++) {
                        const int j = __ldcg(p.jump + (ll)y * K + k);
                        if (j < 0) break;
                        y = j;
                    }
                    int cur = lap(y, 2048, len);
                    if (cur == y) x = y;
                    else if (cur >= 0) {
                        y = cur;
                        cur = lap(y, 4096, len);
                        if (cur == y) x = y;
                        else if (cur >= 0) {
                            int hare = step(cur), tort = cur;
                            ll power = 1, lam_ = 1;
                            for (ll g = 0; hare >= 0 && hare != tort && g <= 3 * p.n; g++) {
                                if (power == lam_) { tort = hare; power <<= 1; lam_ = 0; }
                                hare = step(hare);
                                lam_++;
                            }
                            if (hare >= 0 && hare == tort && lap(hare, p.n + 1, len) == hare) x = hare;
                        }
                    }
                }
                s_x = x;
                s_len = x >= 0 ? (int)len : 0;
            }
            __syncthreads();
            const int x = s_x, len = s_len;
            if (x >= 0) {   // block-uniform
                // sums, smallest vertex and max |dist| of the recorded lap, by the whole block
                double sc = 0.0, stt = 0.0, mabs = 0.0;
                ll isc = 0, ist = 0;
                int minv = 0x7fffffff;
                __threadfence_block();
                for (int i = threadIdx.x; i < len; i += blockDim.x) {
                    const int e = __ldcg(p.cyc_edges + (ll)k * p.cap + i);
                    const int u = __ldg(p.src + e);
                    minv = u < minv ? u : minv;
                    if (p.exact) { isc += __ldg(p.icost + e); ist += __ldg(p.itime + e); }
                    else {
                        double ce = __ldg(p.cost + e), te = __ldg(p.time + e);
                        if (p.scen && e == p.pe[k]) { ce = __dadd_rn(ce, p.pdc[k]); te = __dadd_rn(te, p.pdt[k]); }
                        sc += ce; stt += te;
                        mabs = fmax(mabs, fabs(__ldcg(static_cast<const double *>(p.dist) + (ll)u * K + k)));
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
                    for (int j = 0; j < 8; j++) {
                        sc += s_d[0][j]; stt += s_d[1][j]; mabs = fmax(mabs, s_d[2][j]);
                        isc += s_l[0][j]; ist += s_l[1][j]; minv = s_m[j] < minv ? s_m[j] : minv;
                    }
                }
                if (threadIdx.x == 0) {
                    bool settled = true;   // nothing left to examine for k unless a false cycle is cut below
                    int kind = 0;
                    if (p.exact) {
                        if ((__int128)p.b[k] * isc - (__int128)p.a[k] * ist < 0) kind = MRC_CYC_NEG;
                    } else if (stt > 0) {
                        // r < lambda: negative.  r >= lambda but (r - lambda)*T inside the rounding noise of the
                        // walk (len * eps * max|dist|, x8): a rounding artefact that would relax forever -> TIE.
                        const double rr = sc / stt;
                        if (rr < p.lam[k]) kind = MRC_CYC_NEG;
                        else if ((rr - p.lam[k]) * stt <= 8.0 * (double)len * 2.220446049250313e-16 * mabs) kind = MRC_CYC_TIE;
                    }
                    if (kind) {
                        CycleOut o;
                        o.found = kind; o.len = len; o.minv = minv; o.pad = x;   // pad = vertex the recorded walk starts from
                        o.sumc = sc; o.sumt = stt; o.maxabs = mabs; o.isumc = isc; o.isumt = ist;
                        st->out[k] = o;
                    } else {
                        // false cycle: mark its vertices dead for this launch and cut it for good
                        int cur = minv;
                        ll g = 0;
                        do {
                            const int nxt = step(cur);
                            __stcg(p.jump + (ll)cur * K + k, -2);
                            cur = nxt;
                            g++;
                        } while (cur != minv && cur >= 0 && g <= p.n);
                        __stcg(p.pred + (ll)minv * K + k, ~0ull);
                        atomicAdd(&st->cuts, 1u);
                        settled = false;
                    }
                    if (settled) atomicAnd(&st->pend, ~(1u << k));
                }
            } else if (threadIdx.x == 0) {
                // no representative left, or the chain broke while we walked (cannot happen between launches)
                atomicAnd(&st->pend, ~(1u << k));
            }
        }
        grid.sync();
        pending = __ldcg(&st->pend);
    }
    if (tid0 == 0) st->cyc_mask = alive_mask;
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
                                   ll k, double *old_c, double *old_t) {
    const ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) {
        const ll e = idx[i];
        old_c[i] = cost[e]; old_t[i] = time[e];
        cost[e] = __dadd_rn(cost[e], dc[i]);
        time[e] = __dadd_rn(time[e], dt[i]);
    }
}
__global__ void unpatch_edges_kernel(double *cost, double *time, const ll *idx, ll k, const double *old_c,
                                     const double *old_t) {
    const ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) { cost[idx[i]] = old_c[i]; time[idx[i]] = old_t[i]; }
}
