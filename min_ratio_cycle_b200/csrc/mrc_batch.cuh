// mrc_batch.cuh -- whole-solve-on-device kernel for batches of SMALL graphs (BASELINE config 4: 4,096
// currency graphs n=64, m=4,032).  One CTA per graph: the graph's edges (20 B/edge: uint16 endpoints + f64
// cost/time), K distance vectors and predecessors live in shared memory for the entire lambda search, so
// every graph is read from HBM exactly once per solve() and the Bellman-Ford passes, cycle checks and
// bracket updates run without a single host round trip.  Replaces the loop over independent solve() calls
// (each _solve_numeric + _detect_negative_cycle_numeric, solver.py:1001-1294).
//
// Relaxation here is Jacobi with a deterministic arg-min (lowest edge index on ties), i.e. exactly the
// reference's pass semantics (solver.py:1160-1205): at this size the passes are latency-bound, so the
// second sweep that fixes the predecessor costs nothing and buys run-to-run determinism.
#pragma once
#include "mrc_kernels.cuh"
#include "mrc_search.h"

#define MRC_BATCH_THREADS 256

struct BatchArgs {
    int B;
    const ll *edge_off;     // [B+1] offsets into the concatenated edge arrays
    const int *nverts;      // [B]
    const int *src, *dst;   // concatenated, per graph destination-sorted, LOCAL vertex ids
    const double *cost, *time;
    double tol, slack;
    int max_iter, K;
    double *ratio, *sum_cost, *sum_time;   // [B]
    int *cyc_len, *cyc_buf;                // [B], [B][cyc_cap]  closed vertex lists
    ll cyc_cap;
    int *iters, *status;                   // [B]
    ull *relax_count;                      // total edge relaxations (all graphs)
    int n_max, m_max;                      // shared-memory carve-up
};

__host__ __device__ inline size_t mrc_batch_smem(int n_max, int m_max, int K) {
    size_t b = 0;
    b += (size_t)m_max * 16;                  // cost, time
    b += (size_t)K * n_max * 8 * 2;           // dist old/new
    b += (size_t)m_max * 4;                   // src, dst (uint16)
    b += (size_t)K * n_max * 4 * 3;           // pred, pcand, cyc_tmp
    b += (size_t)(n_max + 2) * 4;             // best cycle
    return b + 256;
}

struct BatchCtl {   // per-CTA control block in shared memory
    double lam[MRC_KMAX], cr[MRC_KMAX], sc[MRC_KMAX], st[MRC_KMAX];
    int verdict[MRC_KMAX], rep[MRC_KMAX], clen[MRC_KMAX], cminv[MRC_KMAX], kind[MRC_KMAX];
    unsigned active, changed;
    double lo, hi, best_sc, best_st;
    int have_cycle, have_best, best_len, passes, done;
    unsigned long long relax;
};

template <int K>
__global__ void __launch_bounds__(MRC_BATCH_THREADS) batch_solve_kernel(const __grid_constant__ BatchArgs p) {
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ BatchCtl ctl;
    const int g = blockIdx.x, tid = threadIdx.x;
    const ll e_lo = p.edge_off[g];
    const int m = (int)(p.edge_off[g + 1] - e_lo), n = p.nverts[g];
    double *s_cost = reinterpret_cast<double *>(sm);
    double *s_time = s_cost + p.m_max;
    double *d_old = s_time + p.m_max;                 // [K][n_max]
    double *d_new = d_old + (size_t)K * p.n_max;
    int *pred = reinterpret_cast<int *>(d_new + (size_t)K * p.n_max);
    int *pcand = pred + (size_t)K * p.n_max;
    int *ctmp = pcand + (size_t)K * p.n_max;          // backward edge lists of extracted cycles
    int *best_cyc = ctmp + (size_t)K * p.n_max;       // [n_max + 2]
    unsigned short *s_src = reinterpret_cast<unsigned short *>(best_cyc + p.n_max + 2);
    unsigned short *s_dst = s_src + p.m_max;
    const int NM = p.n_max;

    // ---- load the graph once; default bracket min/max(c/t) -/+ 1 (solver.py:1019-1023)
    double mn = INFINITY, mx = -INFINITY;
    for (int e = tid; e < m; e += MRC_BATCH_THREADS) {
        const double c = p.cost[e_lo + e], t = p.time[e_lo + e];
        s_cost[e] = c; s_time[e] = t;
        s_src[e] = (unsigned short)p.src[e_lo + e]; s_dst[e] = (unsigned short)p.dst[e_lo + e];
        const double r = __ddiv_rn(c, t);
        mn = fmin(mn, r); mx = fmax(mx, r);
    }
    for (int o = 16; o; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __shared__ double red[2][MRC_BATCH_THREADS / 32];
    if ((tid & 31) == 0) { red[0][tid >> 5] = mn; red[1][tid >> 5] = mx; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < MRC_BATCH_THREADS / 32; w++) { mn = fmin(mn, red[0][w]); mx = fmax(mx, red[1][w]); }
        ctl.lo = mn - 1.0; ctl.hi = mx + 1.0;
        ctl.have_cycle = 0; ctl.have_best = 0; ctl.best_len = 0; ctl.done = 0; ctl.relax = 0;
    }
    __syncthreads();

    // ---- one batched probe of Kc candidates ctl.lam[0..Kc) -> ctl.verdict / cycles in ctmp
    auto probe = [&](int Kc) {
        for (int i = tid; i < K * NM; i += MRC_BATCH_THREADS) { d_old[i] = 0.0; pred[i] = -1; }
        if (tid == 0) {
            ctl.active = (1u << Kc) - 1u; ctl.passes = 0;
            for (int k = 0; k < Kc; k++) { ctl.verdict[k] = -1; ctl.clen[k] = 0; }
        }
        __syncthreads();
        const int check_every = n <= 128 ? 1 : 8;
        while (true) {
            const unsigned active = ctl.active;
            if (!active) break;
            // pass, phase A: dist_new = min(dist_old, min over in-edges of dist_old[u] + w)
            for (int i = tid; i < K * NM; i += MRC_BATCH_THREADS) { d_new[i] = d_old[i]; pcand[i] = 0x7fffffff; }
            if (tid == 0) ctl.changed = 0;
            __syncthreads();
            for (int e = tid; e < m; e += MRC_BATCH_THREADS) {
                const int u = s_src[e], v = s_dst[e];
                const double c = s_cost[e], t = s_time[e];
#pragma unroll
                for (int k = 0; k < K; k++) {
                    if (!((active >> k) & 1u)) continue;
                    const double cand = __dadd_rn(d_old[k * NM + u], __dsub_rn(c, __dmul_rn(ctl.lam[k], t)));
                    if (cand < __dsub_rn(d_old[k * NM + v], p.slack))
                        atomicMax(reinterpret_cast<ull *>(&d_new[k * NM + v]), (ull)__double_as_longlong(cand));
                }
            }
            __syncthreads();
            // phase B: the lowest edge index that attains the new minimum becomes the predecessor
            for (int e = tid; e < m; e += MRC_BATCH_THREADS) {
                const int u = s_src[e], v = s_dst[e];
                const double c = s_cost[e], t = s_time[e];
#pragma unroll
                for (int k = 0; k < K; k++) {
                    if (!((active >> k) & 1u)) continue;
                    const double cand = __dadd_rn(d_old[k * NM + u], __dsub_rn(c, __dmul_rn(ctl.lam[k], t)));
                    if (cand < __dsub_rn(d_old[k * NM + v], p.slack) && cand == d_new[k * NM + v])
                        atomicMin(&pcand[k * NM + v], e);
                }
            }
            __syncthreads();
            unsigned chg = 0;
            for (int i = tid; i < K * NM; i += MRC_BATCH_THREADS) {
                const int k = i / NM;
                if (pcand[i] != 0x7fffffff) { pred[i] = pcand[i]; chg |= 1u << k; }
                d_old[i] = d_new[i];
            }
            chg = __reduce_or_sync(0xffffffffu, chg);
            if ((tid & 31) == 0 && chg) atomicOr(&ctl.changed, chg);
            if (tid == 0) { ctl.passes++; ctl.relax += (unsigned long long)m * __popc(active); }
            if (tid < MRC_KMAX) { ctl.rep[tid] = 0x7fffffff; ctl.kind[tid] = 0; }
            __syncthreads();
            const int passes = ctl.passes;
            const bool do_check = (passes % check_every) == 0 || passes >= n;
            if (do_check) {
                // a vertex whose predecessor chain survives n steps sits in front of (or on) a cycle
                for (int i = tid; i < K * n; i += MRC_BATCH_THREADS) {
                    const int k = i / n, v = i % n;
                    if (!((active >> k) & 1u) || !((ctl.changed >> k) & 1u)) continue;
                    int x = v, s_ = 0;
                    for (; s_ < n; s_++) {
                        const int e = pred[k * NM + x];
                        if (e < 0) break;
                        x = s_src[e];
                    }
                    if (s_ == n) atomicMin(&ctl.rep[k], x);
                }
                __syncthreads();
                if (tid < K && ctl.rep[tid] != 0x7fffffff) {
                    const int k = tid;
                    int x = ctl.rep[k], cur = x, minv = x, len = 0;
                    do { cur = s_src[pred[k * NM + cur]]; len++; if (cur < minv) minv = cur; } while (cur != x && len <= n);
                    double sc = 0.0, stt = 0.0, mabs = 0.0;
                    int i = 0;
                    cur = minv;
                    do {
                        const int e = pred[k * NM + cur];
                        ctmp[k * NM + i] = e;
                        sc += s_cost[e]; stt += s_time[e];
                        mabs = fmax(mabs, fabs(d_old[k * NM + cur]));
                        cur = s_src[e];
                        i++;
                    } while (cur != minv && i < n);
                    const double rr = sc / stt;
                    int kind = 0;
                    if (rr < ctl.lam[k]) kind = MRC_CYC_NEG;
                    else if ((rr - ctl.lam[k]) * stt <= 8.0 * (double)i * 2.220446049250313e-16 * mabs) kind = MRC_CYC_TIE;
                    else pred[k * NM + minv] = -1;   // not negative: cut it (cannot happen with a consistent arg-min)
                    ctl.kind[k] = kind; ctl.clen[k] = i; ctl.cminv[k] = minv; ctl.sc[k] = sc; ctl.st[k] = stt;
                }
                __syncthreads();
            }
            if (tid == 0) {
                unsigned act = ctl.active;
                for (int k = 0; k < Kc; k++) {
                    if (!((act >> k) & 1u)) continue;
                    if (!((ctl.changed >> k) & 1u)) ctl.verdict[k] = MRC_V_NONNEG;
                    else if (do_check && ctl.kind[k] == MRC_CYC_NEG) ctl.verdict[k] = MRC_V_NEG;
                    else if (do_check && ctl.kind[k] == MRC_CYC_TIE) ctl.verdict[k] = MRC_V_TIE;
                    else if (passes >= n) ctl.verdict[k] = MRC_V_NEG_NOCYCLE;   // reference rule, :1252-1257
                    if (ctl.verdict[k] >= 0) act &= ~(1u << k);
                }
                // verdicts are monotone in lambda; a verified cycle of ratio r is negative for every lambda > r
                for (int k = 0; k < Kc; k++) {
                    const int vk = ctl.verdict[k];
                    if (vk != MRC_V_NEG && vk != MRC_V_NONNEG) continue;
                    for (int j = 0; j < Kc; j++) {
                        if (!((act >> j) & 1u)) continue;
                        if (vk == MRC_V_NEG && (ctl.lam[j] > ctl.lam[k] || ctl.lam[j] > ctl.sc[k] / ctl.st[k])) {
                            ctl.verdict[j] = MRC_V_SKIPPED_NEG; act &= ~(1u << j);
                        } else if (vk == MRC_V_NONNEG && ctl.lam[j] < ctl.lam[k]) {
                            ctl.verdict[j] = MRC_V_SKIPPED_NONNEG; act &= ~(1u << j);
                        }
                    }
                }
                ctl.active = act;
            }
            __syncthreads();
        }
    };
    // forward copy of candidate k's cycle into best_cyc (closed, from its smallest vertex) with forward sums
    auto keep_cycle = [&](int k) {
        const int L = ctl.clen[k];
        double sc = 0.0, stt = 0.0;
        best_cyc[0] = ctl.cminv[k];
        for (int j = L - 1, w = 1; j >= 0; j--, w++) {
            const int e = ctmp[k * NM + j];
            sc += s_cost[e]; stt += s_time[e];
            best_cyc[w] = (j > 0) ? (int)s_src[ctmp[k * NM + j - 1]] : ctl.cminv[k];
        }
        ctl.best_len = L + 1; ctl.best_sc = sc; ctl.best_st = stt; ctl.have_best = 1;
    };

    // ---- lambda search (multisection, mrc_search.h)
    int iters = 0, it = -1;
    for (int i = 0; i < p.max_iter; i++) {
        it = i; iters++;
        if (tid == 0) mrc_plan(ctl.lo, ctl.hi, ctl.have_cycle, p.tol, p.K, ctl.lam);
        __syncthreads();
        probe(p.K);
        if (tid == 0) {
            for (int k = 0; k < p.K; k++)
                ctl.cr[k] = (ctl.verdict[k] == MRC_V_NEG || ctl.verdict[k] == MRC_V_TIE) ? ctl.sc[k] / ctl.st[k] : NAN;
            const int best = mrc_reduce(p.K, ctl.lam, ctl.verdict, ctl.cr, &ctl.lo, &ctl.hi, &ctl.have_cycle);
            if (best >= 0) keep_cycle(best);
            ctl.done = (ctl.hi - ctl.lo <= p.tol);   // solver.py:1075-1076
        }
        __syncthreads();
        if (ctl.done) break;
    }
    if (!ctl.have_best) {   // solver.py:1079-1085: one more try at the upper bound
        if (tid == 0) ctl.lam[0] = ctl.hi;
        __syncthreads();
        probe(1);
        if (tid == 0 && ctl.verdict[0] == MRC_V_NEG) keep_cycle(0);
        __syncthreads();
    }
    if (tid == 0) {
        p.iters[g] = iters;
        atomicAdd(p.relax_count, ctl.relax);
        if (!ctl.have_best) { p.status[g] = MRC_ERR_NO_CYCLE; p.cyc_len[g] = 0; p.ratio[g] = INFINITY; }
        else {
            p.status[g] = (p.max_iter > 0 && it == p.max_iter - 1 && ctl.hi - ctl.lo > p.tol) ? MRC_ERR_CONVERGENCE : MRC_OK;
            p.cyc_len[g] = ctl.best_len;
            p.sum_cost[g] = ctl.best_sc; p.sum_time[g] = ctl.best_st; p.ratio[g] = ctl.best_sc / ctl.best_st;
        }
    }
    __syncthreads();
    if (ctl.have_best)
        for (int i = tid; i < ctl.best_len && i < p.cyc_cap; i += MRC_BATCH_THREADS) p.cyc_buf[(ll)g * p.cyc_cap + i] = best_cyc[i];
}
