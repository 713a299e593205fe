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
//   pred     int32[n][KT]  edge index that last lowered dist (racy by design, verified on extraction)
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mrc_b200.h"

namespace cg = cooperative_groups;
typedef long long ll;
typedef unsigned long long ull;

#define MRC_MAX_BURST 8
#define MRC_MAX_ROUNDS 40
#ifndef MRC_RELAX_THREADS
#define MRC_RELAX_THREADS 256
#endif

struct RelaxArgs {
    const int *src, *dst;
    const double *cost, *time;
    const ll *icost, *itime;
    void *dist;
    int *pred;
    unsigned *changed;   // one word: bit k set when candidate k lowered some dist in this pass
    ll m;
    double lam[MRC_KMAX];
    ll a[MRC_KMAX], b[MRC_KMAX];
    double slack;
    unsigned active;     // bit k: candidate k still undecided
    int reverse;         // sweep the edge stream back to front (L2 reuse between consecutive passes)
};

struct CycleOut {
    int found, len, minv, pad;   // found: 0 none, 1 verified negative, 2 tie (rounding artefact)
    double sumc, sumt;   // numeric: sums along the predecessor walk
    double maxabs;       // numeric: max |dist| over the cycle's vertices (bounds the rounding noise of the walk)
    ll isumc, isumt;     // exact
};

struct DevStatus {
    unsigned changed[MRC_MAX_BURST];
    unsigned alive[MRC_MAX_ROUNDS];
    unsigned cyc_mask, rounds_run, pend, cuts;
    int rep[MRC_KMAX * 6];   // [MRC_MAX_EXTRACT][MRC_KMAX]
    CycleOut out[MRC_KMAX];
};

struct CheckArgs {
    const int *src;
    const double *cost, *time;
    const ll *icost, *itime;
    int *pred;           // written only to cut a false cycle
    const void *dist;
    int *jump;
    int *cyc_edges;      // [KT][cap]
    DevStatus *st;
    ll n, cap;
    int rounds;          // ceil(log2(n+1)) + 1
    unsigned active;
    int exact;
    double lam[MRC_KMAX];
    ll a[MRC_KMAX], b[MRC_KMAX];
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
__device__ __forceinline__ bool mrc_improves(const RelaxArgs &p, double cand, double dv) {
    return cand < __dsub_rn(dv, p.slack);
}
__device__ __forceinline__ bool mrc_improves(const RelaxArgs &, ll cand, ll dv) { return cand < dv; }
__device__ __forceinline__ bool mrc_atomic_lower(double *addr, double cand) {
    const ull bits = (ull)__double_as_longlong(cand);
    return atomicMax(reinterpret_cast<ull *>(addr), bits) < bits;
}
__device__ __forceinline__ bool mrc_atomic_lower(ll *addr, ll cand) { return cand < atomicMin(addr, cand); }

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
    T du[K], dv[K];
    T *dist = static_cast<T *>(p.dist);
    if (valid) {
        DistVec<K>::load(du, dist + (size_t)u * K);
        DistVec<K>::load(dv, dist + (size_t)v * K);
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
        if (!((p.active >> k) & 1u)) continue;   // warp-uniform
        T cand = 0;
        bool imp = false;
        if (valid) {
            cand = mrc_cand(p, k, du[k], c, t);
            imp = mrc_improves(p, cand, dv[k]);
        }
        const unsigned im = __ballot_sync(0xffffffffu, imp);
        if (im == 0) continue;                   // warp-uniform
        if (im & (im - 1)) {
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
        if (imp && mrc_atomic_lower(dist + (size_t)v * K + k, cand)) {
            __stcg(p.pred + (size_t)v * K + k, e);
            chg |= 1u << k;
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
template <int K> struct RelaxTune { static constexpr int rows = 2, minblocks = 4; };
template <> struct RelaxTune<8> { static constexpr int rows = 4, minblocks = 1; };
template <int K, bool EXACT>
__global__ void __launch_bounds__(MRC_RELAX_THREADS, RelaxTune<K>::minblocks) relax_kernel(const __grid_constant__ RelaxArgs p) {
    typedef typename WeightT<EXACT>::type W;
    const W *cost = EXACT ? reinterpret_cast<const W *>(p.icost) : reinterpret_cast<const W *>(p.cost);
    const W *time = EXACT ? reinterpret_cast<const W *>(p.itime) : reinterpret_cast<const W *>(p.time);
    unsigned chg = 0;
    const int lane = threadIdx.x & 31;
    const ll warp = ((ll)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const ll nwarps = ((ll)gridDim.x * blockDim.x) >> 5;
    constexpr int MRC_ROWS = RelaxTune<K>::rows;
    constexpr ll CH = 32 * MRC_ROWS;
    const ll nchunks = (p.m + CH - 1) / CH;
    for (ll c = warp; c < nchunks; c += nwarps) {
        const ll e0 = (p.reverse ? (nchunks - 1 - c) : c) * CH + lane;
        int s[MRC_ROWS], d[MRC_ROWS];
        W cs[MRC_ROWS], tm[MRC_ROWS];
        if (e0 - lane + CH <= p.m) {
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
                relax_row<K, W>(p, ok, lane, (int)e, ok ? p.src[e] : 0, ok ? p.dst[e] : 0, ok ? cost[e] : W(0),
                                ok ? time[e] : W(0), chg);
            }
        }
    }
    // changed flag: warp vote, one atomic per warp at most
    chg = __reduce_or_sync(0xffffffffu, chg);
    if (lane == 0 && chg) atomicOr(p.changed, chg);
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

    if (blockIdx.x == 0 && threadIdx.x < MRC_MAX_ROUNDS) st->alive[threadIdx.x] = 0;
    if (blockIdx.x == 0 && threadIdx.x < MRC_KMAX * MRC_MAX_EXTRACT) st->rep[threadIdx.x] = 0x7fffffff;
    if (blockIdx.x == 0 && threadIdx.x < MRC_KMAX) {
        st->out[threadIdx.x].found = 0;
        st->out[threadIdx.x].len = 0;
    }
    if (tid0 == 0) { st->cyc_mask = 0; st->pend = 0; st->cuts = 0; }
    for (ll i = tid0; i < total; i += stride) {
        const int k = (int)(i % K);
        int j = -1;
        if ((p.active >> k) & 1u) {
            const int e = __ldcg(p.pred + i);
            if (e >= 0) j = __ldg(p.src + e);
        }
        __stcg(p.jump + i, j);
    }
    grid.sync();

    unsigned alive_mask = p.active;
    int r = 0;
    for (; r < p.rounds; r++) {
        unsigned alive = 0;
        for (ll i = tid0; i < total; i += stride) {
            const int k = (int)(i % K);
            if (!((alive_mask >> k) & 1u)) continue;
            const int j = __ldcg(p.jump + i);
            if (j >= 0) {
                const int jj = __ldcg(p.jump + (ll)j * K + k);
                __stcg(p.jump + i, jj);
                if (jj >= 0) alive |= 1u << k;
            }
        }
        alive = __reduce_or_sync(0xffffffffu, alive);
        if ((threadIdx.x & 31) == 0 && alive) atomicOr(&st->alive[r], alive);
        grid.sync();
        alive_mask = __ldcg(&st->alive[r]);
        if (alive_mask == 0) break;
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
        if (threadIdx.x == 0 && blockIdx.x < K && ((pending >> blockIdx.x) & 1u)) {
            const int k = blockIdx.x;
            const int rep = __ldcg(&st->rep[it * MRC_KMAX + k]);
            bool settled = true;   // nothing left to examine for k unless a false cycle is cut below
            if (rep != 0x7fffffff) {
                const int x = __ldcg(p.jump + (ll)rep * K + k);   // on a cycle
                int cur = x, minv = x;
                ll len = 0;
                bool ok = true;
                do {
                    const int e = __ldcg(p.pred + (ll)cur * K + k);
                    if (e < 0) { ok = false; break; }
                    cur = __ldg(p.src + e);
                    len++;
                    if (cur < minv) minv = cur;
                } while (cur != x && len <= p.n);
                if (ok && cur == x) {
                    double sc = 0.0, stt = 0.0, mabs = 0.0;
                    ll isc = 0, ist = 0, i = 0;
                    cur = minv;
                    do {
                        const int e = __ldcg(p.pred + (ll)cur * K + k);
                        if (i < p.cap) p.cyc_edges[(ll)k * p.cap + i] = e;
                        if (p.exact) { isc += __ldg(p.icost + e); ist += __ldg(p.itime + e); }
                        else {
                            sc += __ldg(p.cost + e); stt += __ldg(p.time + e);
                            mabs = fmax(mabs, fabs(__ldcg(static_cast<const double *>(p.dist) + (ll)cur * K + k)));
                        }
                        cur = __ldg(p.src + e);
                        i++;
                    } while (cur != minv && i <= p.n);
                    int kind = 0;
                    if (p.exact) {
                        if ((__int128)p.b[k] * isc - (__int128)p.a[k] * ist < 0) kind = MRC_CYC_NEG;
                    } else if (stt > 0) {
                        // r < lambda: negative.  r >= lambda but (r - lambda)*T inside the rounding noise of the
                        // walk (len * eps * max|dist|, x8): a rounding artefact that would relax forever -> TIE.
                        const double rr = sc / stt;
                        if (rr < p.lam[k]) kind = MRC_CYC_NEG;
                        else if ((rr - p.lam[k]) * stt <= 8.0 * (double)i * 2.220446049250313e-16 * mabs) kind = MRC_CYC_TIE;
                    }
                    if (kind) {
                        CycleOut o;
                        o.found = kind; o.len = (int)i; o.minv = minv; o.pad = 0;
                        o.sumc = sc; o.sumt = stt; o.maxabs = mabs; o.isumc = isc; o.isumt = ist;
                        st->out[k] = o;
                    } else {
                        // false cycle: mark its vertices dead for this launch and cut it for good
                        cur = minv;
                        ll g = 0;
                        do {
                            const int e = __ldcg(p.pred + (ll)cur * K + k);
                            __stcg(p.jump + (ll)cur * K + k, -2);
                            cur = __ldg(p.src + e);
                            g++;
                        } while (cur != minv && g <= p.n);
                        __stcg(p.pred + (ll)minv * K + k, -1);
                        atomicAdd(&st->cuts, 1u);
                        settled = false;
                    }
                } else {
                    // the chain left the cycle while we walked (cannot happen between launches); give up on k
                }
            }
            if (settled) atomicAnd(&st->pend, ~(1u << k));
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
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (ll)gridDim.x * blockDim.x) {
        const int e = edges[i];
        o_src[i] = src[e];
        if (exact) { o_ic[i] = icost[e]; o_it[i] = itime[e]; }
        else { o_c[i] = cost[e]; o_t[i] = time[e]; }
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
