// mrc_kernels.cuh -- sm_100a device code of the parametric negative-cycle hot path.
//
// Replaces the NumPy inner body of the reference (min_ratio_cycle/solver.py):
//   relax_kernel      <- relax_once_kahan / relax_once_standard (:1156-1244) and the int64
//                        relax_once of _has_negative_cycle_exact (:831-849), K lambdas per edge read
//   cycle_check_kernel<- the predecessor walk of _detect_negative_cycle_numeric (:1262-1286), run as
//                        pointer doubling + a walk that sums cost/time over the edges actually used
//   tight_mask_kernel <- equal_mask of _find_zero_cycle_exact (:878-879)
//
// Data layout in HBM (all structure-of-arrays, 16-byte aligned):
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
#define MRC_RELAX_THREADS 256

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
    int found, len, minv, pad;
    double sumc, sumt;   // numeric: sums along the predecessor walk
    ll isumc, isumt;     // exact
};

struct DevStatus {
    unsigned changed[MRC_MAX_BURST];
    unsigned alive[MRC_MAX_ROUNDS];
    unsigned cyc_mask, rounds_run;
    int rep[MRC_KMAX];
    CycleOut out[MRC_KMAX];
};

struct CheckArgs {
    const int *src;
    const double *cost, *time;
    const ll *icost, *itime;
    const int *pred;
    int *jump;
    int *cyc_edges;      // [KT][cap]
    DevStatus *st;
    ll n, cap;
    int rounds;          // ceil(log2(n+1)) + 1
    unsigned active;
    int exact;
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
template <> struct DistVec<4> {
    template <typename T> static __device__ __forceinline__ void load(T *r, const T *p) {
        DistVec<2>::load(r, p); DistVec<2>::load(r + 2, p + 2);
    }
};
template <> struct DistVec<8> {
    template <typename T> static __device__ __forceinline__ void load(T *r, const T *p) {
        DistVec<4>::load(r, p); DistVec<4>::load(r + 4, p + 4);
    }
};

// ---- one edge, K candidates --------------------------------------------------------------------
// numeric: w = cost - lam*time with a rounded product and a rounded difference (no FMA), exactly as
// NumPy evaluates `self._C - lam * self._T` (solver.py:1144); improve test `cand < dist[v] - slack`
// (:1161).  dist <= 0 always (starts at +0.0, only decreases), so for negative doubles
// "smaller value" == "larger bit pattern": atomicMax on the raw bits is a 64-bit atomic min.
template <int K>
__device__ __forceinline__ void relax_edge(const RelaxArgs &p, int e, int u, int v, double c, double t,
                                           unsigned &chg) {
    double du[K], dv[K];
    double *dist = static_cast<double *>(p.dist);
    DistVec<K>::load(du, dist + (size_t)u * K);
    DistVec<K>::load(dv, dist + (size_t)v * K);
#pragma unroll
    for (int k = 0; k < K; k++) {
        if (!((p.active >> k) & 1u)) continue;
        double w = __dsub_rn(c, __dmul_rn(p.lam[k], t));
        double cand = __dadd_rn(du[k], w);
        if (cand < __dsub_rn(dv[k], p.slack)) {
            ull bits = (ull)__double_as_longlong(cand);
            ull old = atomicMax(reinterpret_cast<ull *>(dist + (size_t)v * K + k), bits);
            if (old < bits) {
                __stcg(p.pred + (size_t)v * K + k, e);
                chg |= 1u << k;
            }
        }
    }
}
// exact: w = b*icost - a*itime in int64 (solver.py:799-800), strict < without slack (:833)
template <int K>
__device__ __forceinline__ void relax_edge(const RelaxArgs &p, int e, int u, int v, ll c, ll t, unsigned &chg) {
    ll du[K], dv[K];
    ll *dist = static_cast<ll *>(p.dist);
    DistVec<K>::load(du, dist + (size_t)u * K);
    DistVec<K>::load(dv, dist + (size_t)v * K);
#pragma unroll
    for (int k = 0; k < K; k++) {
        if (!((p.active >> k) & 1u)) continue;
        ll cand = du[k] + (p.b[k] * c - p.a[k] * t);
        if (cand < dv[k]) {
            ll old = atomicMin(dist + (size_t)v * K + k, cand);
            if (cand < old) {
                __stcg(p.pred + (size_t)v * K + k, e);
                chg |= 1u << k;
            }
        }
    }
}

template <bool EXACT> struct WeightT { typedef double type; typedef double2 vec2; };
template <> struct WeightT<true> { typedef ll type; typedef longlong2 vec2; };

// Edge-parallel relaxation: each thread owns 4 consecutive edges per step (int4 src, int4 dst,
// 2 x 16-B cost, 2 x 16-B time = 96 B of coalesced stream per thread), persistent grid-stride loop.
template <int K, bool EXACT>
__global__ void __launch_bounds__(MRC_RELAX_THREADS) relax_kernel(const __grid_constant__ RelaxArgs p) {
    typedef typename WeightT<EXACT>::type W;
    typedef typename WeightT<EXACT>::vec2 W2;
    const W *cost = EXACT ? reinterpret_cast<const W *>(p.icost) : reinterpret_cast<const W *>(p.cost);
    const W *time = EXACT ? reinterpret_cast<const W *>(p.itime) : reinterpret_cast<const W *>(p.time);
    unsigned chg = 0;
    const ll nvec = p.m >> 2;
    const ll stride = (ll)gridDim.x * blockDim.x;
    for (ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const ll q = p.reverse ? (nvec - 1 - i) : i;
        int4 s = __ldcs(reinterpret_cast<const int4 *>(p.src) + q);
        int4 d = __ldcs(reinterpret_cast<const int4 *>(p.dst) + q);
        W2 c01 = __ldcs(reinterpret_cast<const W2 *>(cost) + 2 * q);
        W2 c23 = __ldcs(reinterpret_cast<const W2 *>(cost) + 2 * q + 1);
        W2 t01 = __ldcs(reinterpret_cast<const W2 *>(time) + 2 * q);
        W2 t23 = __ldcs(reinterpret_cast<const W2 *>(time) + 2 * q + 1);
        const int e = (int)(q << 2);
        relax_edge<K>(p, e + 0, s.x, d.x, c01.x, t01.x, chg);
        relax_edge<K>(p, e + 1, s.y, d.y, c01.y, t01.y, chg);
        relax_edge<K>(p, e + 2, s.z, d.z, c23.x, t23.x, chg);
        relax_edge<K>(p, e + 3, s.w, d.w, c23.y, t23.y, chg);
    }
    // tail (m % 4 edges)
    {
        const ll e = (nvec << 2) + (ll)blockIdx.x * blockDim.x + threadIdx.x;
        if (e < p.m) relax_edge<K>(p, (int)e, p.src[e], p.dst[e], cost[e], time[e], chg);
    }
    // changed flag: warp vote, one atomic per warp at most
    chg = __reduce_or_sync(0xffffffffu, chg);
    if ((threadIdx.x & 31) == 0 && chg) atomicOr(p.changed, chg);
}

// ---- predecessor-graph cycle check -----------------------------------------------------------------
// One cooperative launch:
//   0. jump[v][k] = src[pred[v][k]]  (or -1 at roots)
//   1. pointer doubling in place, <= ceil(log2 n)+1 rounds, stops as soon as every active candidate's
//      predecessor graph has collapsed to roots (it is a forest -> no cycle yet)
//   2. any vertex still alive leads into a cycle; its jump target lies ON the cycle.  Per candidate
//      the smallest such vertex picks the representative (deterministic for a given pred array)
//   3. one thread per candidate walks the cycle from its smallest vertex, records the edge list and
//      sums cost/time over exactly those edges.  The host verifies sum_cost - lam*sum_time < 0.
template <int K>
__global__ void __launch_bounds__(256) cycle_check_kernel(const __grid_constant__ CheckArgs p) {
    cg::grid_group grid = cg::this_grid();
    const ll total = p.n * K;
    const ll tid0 = (ll)blockIdx.x * blockDim.x + threadIdx.x;
    const ll stride = (ll)gridDim.x * blockDim.x;
    DevStatus *st = p.st;

    if (blockIdx.x == 0 && threadIdx.x < MRC_MAX_ROUNDS) st->alive[threadIdx.x] = 0;
    if (blockIdx.x == 0 && threadIdx.x < MRC_KMAX) {
        st->rep[threadIdx.x] = 0x7fffffff;
        st->out[threadIdx.x].found = 0;
        st->out[threadIdx.x].len = 0;
    }
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
    if (alive_mask == 0) {
        if (tid0 == 0) { st->cyc_mask = 0; st->rounds_run = (unsigned)r; }
        return;
    }
    // candidates in alive_mask have a cycle: smallest vertex that still has a live jump pointer
    {
        int best[K];
#pragma unroll
        for (int k = 0; k < K; k++) best[k] = 0x7fffffff;
        for (ll i = tid0; i < total; i += stride) {
            const int k = (int)(i % K);
            if (!((alive_mask >> k) & 1u)) continue;
            if (__ldcg(p.jump + i) >= 0) {
                const int v = (int)(i / K);
#pragma unroll
                for (int kk = 0; kk < K; kk++)
                    if (kk == k && v < best[kk]) best[kk] = v;
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (!((alive_mask >> k) & 1u)) continue;
            const int b = __reduce_min_sync(0xffffffffu, best[k]);
            if ((threadIdx.x & 31) == 0 && b != 0x7fffffff) atomicMin(&st->rep[k], b);
        }
    }
    grid.sync();
    if (threadIdx.x == 0 && blockIdx.x < K && ((alive_mask >> blockIdx.x) & 1u)) {
        const int k = blockIdx.x;
        const int rep = __ldcg(&st->rep[k]);
        CycleOut o;
        o.found = 0; o.len = 0; o.minv = -1; o.pad = 0; o.sumc = 0; o.sumt = 0; o.isumc = 0; o.isumt = 0;
        if (rep != 0x7fffffff) {
            const int x = __ldcg(p.jump + (ll)rep * K + k);   // on the cycle
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
                double sc = 0.0, stt = 0.0;
                ll isc = 0, ist = 0, i = 0;
                cur = minv;
                do {
                    const int e = __ldcg(p.pred + (ll)cur * K + k);
                    if (i < p.cap) p.cyc_edges[(ll)k * p.cap + i] = e;
                    if (p.exact) { isc += __ldg(p.icost + e); ist += __ldg(p.itime + e); }
                    else { sc += __ldg(p.cost + e); stt += __ldg(p.time + e); }
                    cur = __ldg(p.src + e);
                    i++;
                } while (cur != minv && i <= p.n);
                o.found = 1; o.len = (int)i; o.minv = minv;
                o.sumc = sc; o.sumt = stt; o.isumc = isc; o.isumt = ist;
            }
        }
        st->out[k] = o;
    }
    if (tid0 == 0) { st->cyc_mask = alive_mask; st->rounds_run = (unsigned)r; }
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
