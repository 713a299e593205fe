// mrc_search.h -- bracket logic of the lambda search, shared by the host driver (mrc_solve_numeric), the
// multi-GPU driver (through mrc_plan_candidates / mrc_reduce_round) and the on-device small-graph batch
// solver.  Replaces the mid = (lo+hi)/2 / hi = mid / lo = mid bookkeeping of _solve_numeric
// (min_ratio_cycle/solver.py:1053-1076) by K-way multisection with cycle-ratio tightening.
#pragma once
#include <math.h>

#include "mrc_b200.h"

#ifdef __CUDACC__
#define MRC_HD __host__ __device__ __forceinline__
#else
#define MRC_HD inline
#endif

// K candidates in increasing order for the bracket [lo, hi].  have_cycle: hi is the ratio of a known cycle;
// then the last candidate is the certifier hi - min(0.75 tol, gap/2K): if it has no negative cycle the known
// cycle is optimal to within tol and the search ends.
MRC_HD void mrc_plan(double lo, double hi, int have_cycle, double tol, int K, double *lam) {
    const double gap = hi - lo;
    if (!have_cycle) {
        if (K == 1) { lam[0] = (lo + hi) / 2.0; return; }   // the reference's midpoint, solver.py:1053
        for (int j = 1; j <= K; j++) lam[j - 1] = lo + gap * ((double)j / (double)(K + 1));
        return;
    }
    const double a = tol * 0.75, b = gap / (2.0 * K);
    const double cert = hi - (a < b ? a : b);
    if (K == 1) { lam[0] = cert; return; }
    for (int j = 1; j < K; j++) lam[j - 1] = lo + gap * ((double)j / (double)K);
    lam[K - 1] = cert;
    if (lam[K - 2] >= cert) lam[K - 2] = lo + (cert - lo) * 0.5;
}

// Fold K (lambda, verdict, ratio of the extracted cycle) records into the bracket.  Returns the index of the
// record whose cycle is the new best, or -1.
MRC_HD int mrc_reduce(int K, const double *lam, const int *verdict, const double *cyc_ratio, double *lo, double *hi,
                      int *have_cycle) {
    double nlo = *lo, nhi = *hi;
    int best = -1, hc = *have_cycle;
    double best_ratio = hc ? *hi : INFINITY;
    for (int k = 0; k < K; k++) {
        switch (verdict[k]) {
            case MRC_V_NEG:
            case MRC_V_TIE:
                if (cyc_ratio[k] < best_ratio) { best_ratio = cyc_ratio[k]; best = k; }
                if (verdict[k] == MRC_V_TIE && lam[k] > nlo) nlo = lam[k];
                break;
            case MRC_V_NEG_NOCYCLE:   // reference: hi = mid without a cycle (solver.py:1057-1060)
                if (lam[k] < nhi) nhi = lam[k];
                break;
            case MRC_V_NONNEG:
                if (lam[k] > nlo) nlo = lam[k];
                break;
            default: break;
        }
    }
    if (best >= 0 && best_ratio <= nhi) { nhi = best_ratio; hc = 1; }
    else if (nhi < *hi) hc = 0;   // upper end moved without a cycle to show for it
    if (nlo > nhi) nlo = nhi;     // rounding noise near lambda*
    *lo = nlo; *hi = nhi; *have_cycle = hc;
    return best;
}
