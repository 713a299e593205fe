/*
 * mrc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the parametric negative-cycle hot path of
 * DiogoRibeiro7/min_ratio_cycle (min_ratio_cycle/solver.py).  It reproduces the
 * reference's arithmetic operation by operation (Jacobi Bellman-Ford, tie
 * breaks, Kahan-style update, bisection, Stern-Brocot, tight-subgraph DFS), so
 * that float results are bit-identical to the NumPy reference and integer
 * results are exactly equal.  Each function cites the reference lines it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may load this library; the product (min_ratio_cycle_b200) never does.
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against fixtures in
 * tests/golden/ that were produced by importing the Python reference
 * (tests/golden/make_golden.py) -- parity is pinned for both numeric and exact.
 *
 * Build: see oracle/Makefile.  MUST be compiled with -ffp-contract=off (NumPy
 * evaluates C - lam*T as a rounded product followed by a rounded subtraction).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_ERR_ALLOC 1
#define ORC_ERR_NO_CYCLE 2     /* NumericalInstabilityError("No negative cycle found"), solver.py:1087-1090 */
#define ORC_ERR_CONVERGENCE 3  /* ConvergenceError, solver.py:1102-1109 */
#define ORC_ERR_REDUCEAT 4     /* IndexError from np.minimum.reduceat when starts[v]==m, solver.py:839 */
#define ORC_ERR_LOWER_NEG 5    /* RuntimeError("Lower bound has negative cycle"), solver.py:759-760 */
#define ORC_ERR_UPPER_NONNEG 6 /* RuntimeError("Upper bound has no negative cycle"), solver.py:761-762 */
#define ORC_ERR_SB_STEPS 7     /* RuntimeError("Stern-Brocot search did not converge"), solver.py:797 */
#define ORC_ERR_EDGE_MISSING 8 /* RuntimeError("Edge u -> v not found"), solver.py:993,1325 */
#define ORC_ERR_OVERFLOW 9     /* Python big-int territory the int64 restatement does not follow */

typedef struct {
    int64_t n, m;
    int64_t *U, *V;          /* dst-sorted endpoints, solver.py:476-478 */
    double *C, *T;           /* float64 weights */
    int64_t *Ci, *Ti;        /* int64 weights, NULL unless all_int, solver.py:491-498 */
    int64_t *starts, *counts;/* CSR by destination, solver.py:481-485 */
    int64_t *order;          /* sorted position -> insertion index */
    int nthreads;
    int ref_quirks;          /* 1 = reproduce reference defects that change verdicts (default); 0 = textbook Jacobi */
} orc_graph;

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_graph_free(orc_graph *g) {
    if (!g) return;
    free(g->U); free(g->V); free(g->C); free(g->T); free(g->Ci); free(g->Ti);
    free(g->starts); free(g->counts); free(g->order); free(g);
}

/*
 * _remove_dominated_edges, solver.py:531-575.  keep[i]=0 when some other edge
 * with the same (u,v) has cost<= and time<= with one strict (float compares).
 * Returns the number of kept edges.
 */
typedef struct { int64_t u, v, idx; } orc_key;
static int orc_key_cmp(const void *a, const void *b) {
    const orc_key *x = (const orc_key *)a, *y = (const orc_key *)b;
    if (x->u != y->u) return x->u < y->u ? -1 : 1;
    if (x->v != y->v) return x->v < y->v ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}
int64_t orc_remove_dominated(int64_t m, const int64_t *u, const int64_t *v, const double *c,
                             const double *t, uint8_t *keep) {
    orc_key *k = (orc_key *)malloc(sizeof(orc_key) * (size_t)(m > 0 ? m : 1));
    if (!k) return -1;
    for (int64_t i = 0; i < m; i++) { k[i].u = u[i]; k[i].v = v[i]; k[i].idx = i; keep[i] = 1; }
    qsort(k, (size_t)m, sizeof(orc_key), orc_key_cmp);
    int64_t kept = m;
    for (int64_t s = 0; s < m;) {
        int64_t e = s + 1;
        while (e < m && k[e].u == k[s].u && k[e].v == k[s].v) e++;
        if (e - s > 1) {
            for (int64_t j = s; j < e; j++) {
                int64_t ej = k[j].idx;
                for (int64_t i = s; i < e; i++) {
                    if (i == j) continue;
                    int64_t ei = k[i].idx;
                    if (c[ei] <= c[ej] && t[ei] <= t[ej] && (c[ei] < c[ej] || t[ei] < t[ej])) {
                        if (keep[ej]) { keep[ej] = 0; kept--; }
                        break;
                    }
                }
            }
        }
        s = e;
    }
    free(k);
    return kept;
}

/*
 * _build_arrays_if_needed, solver.py:450-511: stable sort by destination
 * (np.argsort(V, kind="stable")), bincount, exclusive prefix sum.
 * ci/ti may be NULL (graph not all-integer).
 */
orc_graph *orc_graph_build(int64_t n, int64_t m, const int64_t *u, const int64_t *v,
                           const double *c, const double *t, const int64_t *ci, const int64_t *ti) {
    orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
    if (!g) return NULL;
    size_t mm = (size_t)(m > 0 ? m : 1), nn = (size_t)(n > 0 ? n : 1);
    g->n = n; g->m = m; g->nthreads = 1; g->ref_quirks = 1;
    g->U = (int64_t *)malloc(8 * mm); g->V = (int64_t *)malloc(8 * mm);
    g->C = (double *)malloc(8 * mm); g->T = (double *)malloc(8 * mm);
    g->order = (int64_t *)malloc(8 * mm);
    g->starts = (int64_t *)calloc(nn, 8); g->counts = (int64_t *)calloc(nn, 8);
    if (ci && ti) { g->Ci = (int64_t *)malloc(8 * mm); g->Ti = (int64_t *)malloc(8 * mm); }
    if (!g->U || !g->V || !g->C || !g->T || !g->order || !g->starts || !g->counts ||
        (ci && ti && (!g->Ci || !g->Ti))) { orc_graph_free(g); return NULL; }
    for (int64_t i = 0; i < m; i++) g->counts[v[i]]++;
    int64_t acc = 0;
    for (int64_t x = 0; x < n; x++) { g->starts[x] = acc; acc += g->counts[x]; }
    int64_t *cursor = (int64_t *)malloc(8 * nn);
    if (!cursor) { orc_graph_free(g); return NULL; }
    memcpy(cursor, g->starts, 8 * nn);
    for (int64_t i = 0; i < m; i++) {
        int64_t p = cursor[v[i]]++;
        g->order[p] = i; g->U[p] = u[i]; g->V[p] = v[i]; g->C[p] = c[i]; g->T[p] = t[i];
        if (g->Ci) { g->Ci[p] = ci[i]; g->Ti[p] = ti[i]; }
    }
    free(cursor);
    return g;
}

void orc_graph_set_threads(orc_graph *g, int nthreads) { g->nthreads = nthreads > 0 ? nthreads : 1; }
/* quirks off = the certifying checker of SURVEY 7 step 2(b): same arithmetic, defects removed */
void orc_graph_set_quirks(orc_graph *g, int on) { g->ref_quirks = on ? 1 : 0; }
int64_t orc_graph_m(const orc_graph *g) { return g->m; }
void orc_graph_arrays(const orc_graph *g, int64_t *U, int64_t *V, double *C, double *T, int64_t *starts,
                      int64_t *counts, int64_t *order) {
    memcpy(U, g->U, 8 * (size_t)g->m); memcpy(V, g->V, 8 * (size_t)g->m);
    memcpy(C, g->C, 8 * (size_t)g->m); memcpy(T, g->T, 8 * (size_t)g->m);
    memcpy(starts, g->starts, 8 * (size_t)g->n); memcpy(counts, g->counts, 8 * (size_t)g->n);
    memcpy(order, g->order, 8 * (size_t)g->m);
}

/* Bounds of _solve_numeric, solver.py:1019-1023: min/max of C/T, -/+ 1.0 */
void orc_numeric_bounds(const orc_graph *g, double *lo, double *hi) {
    double mn = INFINITY, mx = -INFINITY;
    for (int64_t e = 0; e < g->m; e++) {
        double r = g->C[e] / g->T[e];
        if (r < mn) mn = r;
        if (r > mx) mx = r;
    }
    *lo = mn - 1.0; *hi = mx + 1.0;
}

/*
 * _compute_cycle_weights_float, solver.py:1296-1327 (+ _build_edge_map :1329-1337):
 * per hop (u,v) take, among parallel edges in insertion order, the first one with
 * minimal cost/time; accumulate from 0.0 in cycle order.  `cycle` is OPEN.
 */
static int orc_cycle_weights_float(const orc_graph *g, const int64_t *cycle, int64_t len, double *sc,
                                   double *st) {
    double sum_cost = 0.0, sum_time = 0.0;
    for (int64_t i = 0; i < len; i++) {
        int64_t u = cycle[i], v = cycle[(i + 1) % len];
        int64_t s = g->starts[v], e = s + g->counts[v];
        int found = 0; double bc = 0, bt = 0, br = 0;
        for (int64_t k = s; k < e; k++) {
            if (g->U[k] != u) continue;
            double r = g->C[k] / g->T[k];
            if (!found || r < br) { found = 1; bc = g->C[k]; bt = g->T[k]; br = r; }
        }
        if (!found) return ORC_ERR_EDGE_MISSING;
        sum_cost += bc; sum_time += bt;
    }
    *sc = sum_cost; *st = sum_time;
    return ORC_OK;
}
int orc_cycle_weights_float_pub(const orc_graph *g, const int64_t *cycle, int64_t len, double *sc, double *st) {
    return orc_cycle_weights_float(g, cycle, len, sc, st);
}

/*
 * One Jacobi relaxation pass of relax_once_kahan / relax_once_standard,
 * solver.py:1156-1205 / :1207-1244.  All candidates come from the previous
 * pass's dist; per destination the minimal improving candidate wins, first
 * occurrence in dst-sorted order on ties (:1177-1183).  Returns 1 if changed.
 */
static int orc_relax_numeric(const orc_graph *g, const double *W, double slack, double *dist, double *comp,
                             int64_t *pred, double *best, int64_t *bidx) {
    const int64_t n = g->n;
    int any = 0;
#pragma omp parallel for schedule(static) reduction(| : any) num_threads(g->nthreads) if (g->nthreads > 1)
    for (int64_t v = 0; v < n; v++) {
        int64_t s = g->starts[v], e = s + g->counts[v];
        double thr = dist[v] - slack, b = INFINITY; int64_t bi = -1;
        for (int64_t k = s; k < e; k++) {
            double cand = dist[g->U[k]] + W[k];      /* :1160 */
            if (cand < thr && cand < b) { b = cand; bi = k; }   /* :1161,:1173-1183 */
        }
        /* Reference quirk (:1179-1183): base[1:] = cs[starts[1:] - 1] wraps to cs[-1] for every v
         * with starts[v] == 0, so when vertex 0 has no in-edge the FIRST non-empty segment never
         * gets a first_mask hit and that vertex is never updated. */
        if (g->ref_quirks && s == 0 && v > 0) bi = -1;
        best[v] = b; bidx[v] = bi;
        if (bi >= 0) any = 1;
    }
    if (!any) return 0;                                /* :1163-1164 / :1185-1186 */
#pragma omp parallel for schedule(static) num_threads(g->nthreads) if (g->nthreads > 1)
    for (int64_t v = 0; v < n; v++) {
        int64_t bi = bidx[v];
        if (bi < 0) continue;
        if (comp) {                                  /* :1193-1198 */
            double y = (best[v] - dist[v]) - comp[v];
            double t = dist[v] + y;
            comp[v] = (t - dist[v]) - y;
            dist[v] = t;
        } else {
            dist[v] = best[v];                       /* :1200 / :1239 */
        }
        pred[v] = g->U[bi];                          /* :1203 / :1242 */
    }
    return 1;
}

/*
 * _detect_negative_cycle_numeric, solver.py:1124-1294.
 * pass_limit < 0: reference semantics (n-1 passes + detection pass).
 * pass_limit >= 0: stop after that many passes and report *has_neg = -1 (undecided) if
 *   still changing -- used only by bounded CPU-baseline timing and full-size certificates.
 * Outputs: has_neg (0/1), has_cycle (cycle_data is not None), OPEN cycle in cyc_buf (cap n),
 *   sums, ratio, passes executed.
 */
int orc_detect_numeric(const orc_graph *g, double lam, double slack, int use_kahan, int64_t pass_limit,
                       int *has_neg, int *has_cycle, int64_t *cyc_buf, int64_t *cyc_len, double *sum_cost,
                       double *sum_time, double *ratio, int64_t *passes_out) {
    const int64_t n = g->n, m = g->m;
    *has_neg = 0; *has_cycle = 0; *cyc_len = 0; *passes_out = 0;
    double *W = (double *)malloc(8 * (size_t)(m > 0 ? m : 1));
    double *dist = (double *)calloc((size_t)n, 8);
    double *comp = use_kahan ? (double *)calloc((size_t)n, 8) : NULL;
    double *best = (double *)malloc(8 * (size_t)n);
    int64_t *bidx = (int64_t *)malloc(8 * (size_t)n);
    int64_t *pred = (int64_t *)malloc(8 * (size_t)n);
    int rc = ORC_OK;
    if (!W || !dist || (use_kahan && !comp) || !best || !bidx || !pred) { rc = ORC_ERR_ALLOC; goto done; }
    for (int64_t e = 0; e < m; e++) W[e] = g->C[e] - lam * g->T[e];   /* :1144 */
    for (int64_t v = 0; v < n; v++) pred[v] = -1;                      /* :1154 */

    int64_t passes = 0; int changed = 1;
    for (int64_t it = 0; it < n - 1; it++) {                           /* :1252-1254 */
        if (pass_limit >= 0 && passes >= pass_limit) { *has_neg = -1; *passes_out = passes; goto done; }
        changed = orc_relax_numeric(g, W, slack, dist, comp, pred, best, bidx);
        passes++;
        if (!changed) break;
    }
    if (pass_limit >= 0 && passes >= pass_limit && changed) { *has_neg = -1; *passes_out = passes; goto done; }
    changed = orc_relax_numeric(g, W, slack, dist, comp, pred, best, bidx);   /* :1257 */
    passes++;
    *passes_out = passes;
    if (!changed) goto done;                                          /* :1258-1259 */
    *has_neg = 1;

    /* Extraction :1262-1286 */
    int64_t x = -1;
    for (int64_t v = n - 1; v >= 0; v--) if (pred[v] != -1) { x = v; break; }   /* touched[-1] */
    if (x < 0) goto done;                                             /* :1263-1264 */
    for (int64_t it = 0; it < n; it++) {                               /* :1268-1271 */
        if (pred[x] == -1) break;
        x = pred[x];
    }
    {
        /* seen/sequence walk :1274-1281; bidx reused as "position in sequence" (-1 = unseen) */
        for (int64_t v = 0; v < n; v++) bidx[v] = -1;
        int64_t *seq = (int64_t *)best;   /* reuse storage: n int64 slots */
        int64_t len = 0, cur = x;
        while (cur != -1 && bidx[cur] < 0) {
            bidx[cur] = len; seq[len++] = cur;
            cur = pred[cur];              /* pred == -1 ends the walk */
        }
        if (cur == -1) goto done;                                     /* :1283-1284 */
        int64_t start = bidx[cur], clen = len - start;
        for (int64_t i = 0; i < clen; i++) cyc_buf[i] = seq[len - 1 - i];   /* [::-1], :1286 */
        double sc, st;
        rc = orc_cycle_weights_float(g, cyc_buf, clen, &sc, &st);     /* :1289 */
        if (rc != ORC_OK) goto done;
        if (st <= 0) goto done;                                       /* :1290-1291 */
        *has_cycle = 1; *cyc_len = clen; *sum_cost = sc; *sum_time = st; *ratio = sc / st;   /* :1293 */
    }
done:
    free(W); free(dist); free(comp); free(best); free(bidx); free(pred);
    return rc;
}

/*
 * Bounded timing helper for the CPU baseline: `passes` Jacobi passes at lam from
 * dist == 0 (same inner body as above), returns how many passes actually changed.
 */
int orc_time_passes_numeric(const orc_graph *g, double lam, double slack, int use_kahan, int64_t passes,
                            int64_t *done_out) {
    const int64_t n = g->n, m = g->m;
    double *W = (double *)malloc(8 * (size_t)(m > 0 ? m : 1));
    double *dist = (double *)calloc((size_t)n, 8);
    double *comp = use_kahan ? (double *)calloc((size_t)n, 8) : NULL;
    double *best = (double *)malloc(8 * (size_t)n);
    int64_t *bidx = (int64_t *)malloc(8 * (size_t)n);
    int64_t *pred = (int64_t *)malloc(8 * (size_t)n);
    if (!W || !dist || (use_kahan && !comp) || !best || !bidx || !pred) {
        free(W); free(dist); free(comp); free(best); free(bidx); free(pred); return ORC_ERR_ALLOC;
    }
#pragma omp parallel for schedule(static) num_threads(g->nthreads) if (g->nthreads > 1)
    for (int64_t e = 0; e < m; e++) W[e] = g->C[e] - lam * g->T[e];
    for (int64_t v = 0; v < n; v++) pred[v] = -1;
    int64_t done = 0;
    for (int64_t p = 0; p < passes; p++) {
        if (!orc_relax_numeric(g, W, slack, dist, comp, pred, best, bidx)) break;
        done++;
    }
    *done_out = done;
    free(W); free(dist); free(comp); free(best); free(bidx); free(pred);
    return ORC_OK;
}

/*
 * _solve_numeric, solver.py:1001-1122 (bisection body :1031-1117).  lo/hi are the
 * resolved bounds (the Python wrapper applies the `lambda_lo or lo` quirk of :1024-1025).
 * target_ratio: NaN = None.  The early-exit test compares sum_time (cycle_data[2]) to
 * target_ratio exactly as the reference does (:1063-1071).
 * Outputs the CLOSED cycle (cap n+1).
 */
int orc_solve_numeric(const orc_graph *g, double lo, double hi, int64_t max_iter, double tol, double slack,
                      int use_kahan, double target_ratio, int early_term, int64_t *cyc_buf, int64_t *cyc_len,
                      double *sum_cost, double *sum_time, double *ratio, int64_t *iterations_out,
                      int64_t *passes_out) {
    const int64_t n = g->n;
    int64_t *tmp = (int64_t *)malloc(8 * (size_t)(n + 1));
    if (!tmp) return ORC_ERR_ALLOC;
    int have_best = 0; int64_t best_len = 0; double bsc = 0, bst = 0, br = 0;
    int64_t iterations = 0, iteration = -1, total_passes = 0;
    int rc = ORC_OK;
    for (int64_t it = 0; it < max_iter; it++) {
        iteration = it;
        iterations++;
        double mid = (lo + hi) / 2.0;                                 /* :1053 */
        int hn, hc; int64_t cl, ps; double sc, st, r;
        rc = orc_detect_numeric(g, mid, slack, use_kahan, -1, &hn, &hc, tmp, &cl, &sc, &st, &r, &ps);
        if (rc != ORC_OK) { free(tmp); return rc; }
        total_passes += ps;
        if (hn) {
            hi = mid;                                                 /* :1058 */
            if (hc) {
                have_best = 1; best_len = cl; bsc = sc; bst = st; br = r;
                memcpy(cyc_buf, tmp, 8 * (size_t)cl);
                if (!isnan(target_ratio) && early_term && st < target_ratio) break;   /* :1063-1071 */
            }
        } else {
            lo = mid;                                                 /* :1073 */
        }
        if (hi - lo <= tol) break;                                    /* :1075-1076 */
    }
    if (!have_best) {                                                 /* :1079-1085 */
        int hn, hc; int64_t cl, ps; double sc, st, r;
        rc = orc_detect_numeric(g, hi, slack, use_kahan, -1, &hn, &hc, tmp, &cl, &sc, &st, &r, &ps);
        if (rc != ORC_OK) { free(tmp); return rc; }
        total_passes += ps;
        if (hn && hc) {
            have_best = 1; best_len = cl; bsc = sc; bst = st; br = r;
            memcpy(cyc_buf, tmp, 8 * (size_t)cl);
        }
    }
    free(tmp);
    *iterations_out = iterations; *passes_out = total_passes;
    if (!have_best) return ORC_ERR_NO_CYCLE;                          /* :1087-1090 */
    /* close :1094-1096 */
    if (best_len == 1 || cyc_buf[0] != cyc_buf[best_len - 1]) { cyc_buf[best_len] = cyc_buf[0]; best_len++; }
    *cyc_len = best_len; *sum_cost = bsc; *sum_time = bst; *ratio = br;
    if (max_iter > 0 && iteration == max_iter - 1 && hi - lo > tol) return ORC_ERR_CONVERGENCE;   /* :1102 */
    return ORC_OK;
}

/* ------------------------------------------------------------------ exact mode */

/* _combine_weights_exact, solver.py:799-809: int64 b*Ci - a*Ti with silent wrap-around. */
static void orc_combine_exact(const orc_graph *g, int64_t a, int64_t b, int64_t *W) {
    for (int64_t e = 0; e < g->m; e++)
        W[e] = (int64_t)((uint64_t)b * (uint64_t)g->Ci[e] - (uint64_t)a * (uint64_t)g->Ti[e]);
}

/*
 * relax_once of _has_negative_cycle_exact / _compute_potentials_exact,
 * solver.py:831-849 / :942-958.  Jacobi; every tied minimum scatters the same value.
 * Returns 1 changed, 0 unchanged, -1 = the reference would raise IndexError in
 * np.minimum.reduceat because starts[n-1] == m.
 */
static int orc_relax_exact(const orc_graph *g, const int64_t *W, int64_t *dist, int64_t *best) {
    const int64_t n = g->n;
    int any = 0;
#pragma omp parallel for schedule(static) reduction(| : any) num_threads(g->nthreads) if (g->nthreads > 1)
    for (int64_t v = 0; v < n; v++) {
        int64_t s = g->starts[v], e = s + g->counts[v];
        int64_t dv = dist[v], b = INT64_MAX; int imp = 0;
        for (int64_t k = s; k < e; k++) {
            int64_t cand = (int64_t)((uint64_t)dist[g->U[k]] + (uint64_t)W[k]);   /* :832 */
            if (cand < dv) { imp = 1; if (cand < b) b = cand; }                   /* :833,:838-839 */
        }
        best[v] = imp ? b : INT64_MAX;
        if (imp) any = 1;
    }
    if (!any) return 0;                                                /* :834-835 */
    if (g->ref_quirks && g->counts[n - 1] == 0) return -1;             /* reduceat IndexError, :839 */
    for (int64_t v = 0; v < n; v++) if (best[v] != INT64_MAX) dist[v] = best[v];   /* :847-848 */
    return 1;
}

/* shared BF loop: solver.py:852-856 and :960-969 are the same loop */
static int orc_bf_exact(const orc_graph *g, int64_t a, int64_t b, int64_t *dist, int *has_neg,
                        int64_t *passes_out) {
    const int64_t n = g->n, m = g->m;
    int64_t *W = (int64_t *)malloc(8 * (size_t)(m > 0 ? m : 1));
    int64_t *best = (int64_t *)malloc(8 * (size_t)n);
    if (!W || !best) { free(W); free(best); return ORC_ERR_ALLOC; }
    orc_combine_exact(g, a, b, W);
    memset(dist, 0, 8 * (size_t)n);
    int64_t passes = 0; int rc = ORC_OK, ch = 0;
    for (int64_t it = 0; it < n - 1; it++) {
        ch = orc_relax_exact(g, W, dist, best); passes++;
        if (ch < 0) { rc = ORC_ERR_REDUCEAT; goto done; }
        if (!ch) break;
    }
    ch = orc_relax_exact(g, W, dist, best); passes++;
    if (ch < 0) { rc = ORC_ERR_REDUCEAT; goto done; }
    *has_neg = ch;
done:
    if (passes_out) *passes_out += passes;
    free(W); free(best);
    return rc;
}

/* _has_negative_cycle_exact, solver.py:811-856 */
int orc_has_neg_exact(const orc_graph *g, int64_t a, int64_t b, int *has_neg, int64_t *passes_out) {
    if (!g->Ci) return ORC_ERR_ALLOC;
    int64_t *dist = (int64_t *)malloc(8 * (size_t)g->n);
    if (!dist) return ORC_ERR_ALLOC;
    int64_t p = 0;
    int rc = orc_bf_exact(g, a, b, dist, has_neg, &p);
    if (passes_out) *passes_out = p;
    free(dist);
    return rc;
}

/* _compute_potentials_exact, solver.py:922-969: ok = no negative cycle; dist returned */
int orc_potentials_exact(const orc_graph *g, int64_t a, int64_t b, int *ok, int64_t *dist) {
    if (!g->Ci) return ORC_ERR_ALLOC;
    int hn = 0; int64_t p = 0;
    int rc = orc_bf_exact(g, a, b, dist, &hn, &p);
    *ok = !hn;
    return rc;
}

/*
 * _compute_cycle_weights_exact, solver.py:971-995: per hop the FIRST inserted (u,v) edge.
 * In the stable dst-sorted layout that is the first k in v's segment with U[k]==u.
 */
static int orc_cycle_weights_exact(const orc_graph *g, const int64_t *cycle, int64_t len, int64_t *sc,
                                   int64_t *st) {
    int64_t sum_c = 0, sum_t = 0;
    for (int64_t i = 0; i < len; i++) {
        int64_t u = cycle[i], v = cycle[(i + 1) % len];
        int64_t s = g->starts[v], e = s + g->counts[v], k;
        for (k = s; k < e; k++) if (g->U[k] == u) break;
        if (k == e) return ORC_ERR_EDGE_MISSING;
        sum_c += g->Ci[k]; sum_t += g->Ti[k];
    }
    *sc = sum_c; *st = sum_t;
    return ORC_OK;
}

/*
 * _find_zero_cycle_exact, solver.py:858-920.  The recursive DFS (:894-918) is replayed
 * with an explicit stack: adjacency of u = tight edges (dist[U]+W == dist[V]) in
 * ascending dst-sorted edge index (:886-888); roots 0..n-1 (:914); first GRAY hit wins;
 * cycle = [child-of-v ... u, v] (:903-910).  found=0 -> None.  OPEN cycle in cyc_buf.
 */
int orc_find_zero_cycle_exact(const orc_graph *g, int64_t a, int64_t b, int *found, int64_t *cyc_buf,
                              int64_t *cyc_len, int64_t *sum_c, int64_t *sum_t) {
    const int64_t n = g->n, m = g->m;
    *found = 0; *cyc_len = 0;
    if (!g->Ci) return ORC_ERR_ALLOC;
    int64_t *dist = (int64_t *)malloc(8 * (size_t)n);
    int64_t *W = (int64_t *)malloc(8 * (size_t)(m > 0 ? m : 1));
    int64_t *adj_start = (int64_t *)calloc((size_t)n + 1, 8);
    int64_t *adj = NULL, *cursor = NULL, *parent = NULL, *stack = NULL;
    int8_t *color = NULL;
    int rc = ORC_OK;
    if (!dist || !W || !adj_start) { rc = ORC_ERR_ALLOC; goto done; }
    int ok = 0;
    rc = orc_potentials_exact(g, a, b, &ok, dist);                     /* :865 */
    if (rc != ORC_OK || !ok) goto done;                               /* :866-867 */
    orc_combine_exact(g, a, b, W);                                    /* :878 */
    int64_t tight = 0;
    for (int64_t e = 0; e < m; e++) {
        int eq = ((int64_t)((uint64_t)dist[g->U[e]] + (uint64_t)W[e]) == dist[g->V[e]]);   /* :879 */
        if (eq) { adj_start[g->U[e] + 1]++; tight++; }
        W[e] = eq;   /* reuse W as the mask */
    }
    if (!tight) goto done;                                            /* :881-882 */
    for (int64_t v = 0; v < n; v++) adj_start[v + 1] += adj_start[v];
    adj = (int64_t *)malloc(8 * (size_t)tight);
    cursor = (int64_t *)malloc(8 * (size_t)n);
    parent = (int64_t *)malloc(8 * (size_t)n);
    stack = (int64_t *)malloc(8 * (size_t)n);
    color = (int8_t *)calloc((size_t)n, 1);
    if (!adj || !cursor || !parent || !stack || !color) { rc = ORC_ERR_ALLOC; goto done; }
    memcpy(cursor, adj_start, 8 * (size_t)n);
    for (int64_t e = 0; e < m; e++) if (W[e]) adj[cursor[g->U[e]]++] = g->V[e];   /* :887-888 */
    memcpy(cursor, adj_start, 8 * (size_t)n);
    for (int64_t v = 0; v < n; v++) parent[v] = -1;
    for (int64_t start = 0; start < n && !*found; start++) {           /* :914-918 */
        if (color[start] != 0) continue;
        int64_t sp = 0;
        stack[sp++] = start; color[start] = 1;
        while (sp > 0 && !*found) {
            int64_t u = stack[sp - 1];
            if (cursor[u] < adj_start[u + 1]) {
                int64_t v = adj[cursor[u]++];
                if (color[v] == 0) { parent[v] = u; color[v] = 1; stack[sp++] = v; }
                else if (color[v] == 1) {
                    /* back edge: [v, u, parent[u], ...] reversed */
                    int64_t len = 0;
                    cyc_buf[len++] = v;
                    int64_t x = u;
                    while (x != v && x != -1) { cyc_buf[len++] = x; x = parent[x]; }
                    for (int64_t i = 0, j = len - 1; i < j; i++, j--) {
                        int64_t t = cyc_buf[i]; cyc_buf[i] = cyc_buf[j]; cyc_buf[j] = t;
                    }
                    *cyc_len = len; *found = 1;
                }
            } else { color[u] = 2; sp--; }
        }
    }
    if (*found) rc = orc_cycle_weights_exact(g, cyc_buf, *cyc_len, sum_c, sum_t);   /* :918 */
done:
    free(dist); free(W); free(adj_start); free(adj); free(cursor); free(parent); free(stack); free(color);
    return rc;
}

static int64_t orc_gcd(int64_t a, int64_t b) {
    if (a < 0) a = -a;
    if (b < 0) b = -b;
    while (b) { int64_t t = a % b; a = b; b = t; }
    return a;
}
static int64_t orc_floordiv(int64_t a, int64_t b) {   /* Python // for b > 0 */
    int64_t q = a / b, r = a % b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

/*
 * _stern_brocot_search, solver.py:734-797.  max_den/max_steps < 0 = None (defaults :746-751).
 * probes_ab (cap probes_cap pairs) records every (a,b) handed to _has_negative_cycle_exact,
 * in call order, so tests can pin the trajectory.  OPEN cycle out.
 */
int orc_stern_brocot(const orc_graph *g, int64_t max_den, int64_t max_steps, int64_t *cyc_buf,
                     int64_t *cyc_len, int64_t *sum_c, int64_t *sum_t, int64_t *a_out, int64_t *b_out,
                     int64_t *iterations_out, int64_t *probes_ab, int64_t probes_cap, int64_t *nprobes_out,
                     int64_t *passes_out) {
    if (!g->Ci) return ORC_ERR_ALLOC;
    const int64_t m = g->m;
    int64_t np = 0, passes = 0;
    if (nprobes_out) *nprobes_out = 0;
    if (max_den < 0) {                                                /* :746-748 */
        int64_t tmax = g->Ti[0];
        for (int64_t e = 1; e < m; e++) if (g->Ti[e] > tmax) tmax = g->Ti[e];
        max_den = g->n * (tmax > 1 ? tmax : 1);
    }
    if (max_steps < 0) max_steps = 2 * max_den + 10;                  /* :750-751 */
    double mn = INFINITY, mx = -INFINITY;                              /* :754-756 */
    for (int64_t e = 0; e < m; e++) {
        double r = (double)g->Ci[e] / (double)g->Ti[e];
        if (r < mn) mn = r;
        if (r > mx) mx = r;
    }
    int64_t aL = (int64_t)floor(mn) - 1, bL = 1, aR = (int64_t)ceil(mx) + 1, bR = 1;
    int hn, rc;
#define ORC_PROBE(A, B)                                                              \
    do {                                                                             \
        if (probes_ab && np < probes_cap) { probes_ab[2 * np] = (A); probes_ab[2 * np + 1] = (B); } \
        np++;                                                                        \
        int64_t p_ = 0;                                                              \
        rc = orc_has_neg_exact(g, (A), (B), &hn, &p_);                               \
        passes += p_;                                                                \
        if (nprobes_out) *nprobes_out = np;                                          \
        if (passes_out) *passes_out = passes;                                        \
        if (rc != ORC_OK) return rc;                                                 \
    } while (0)
    ORC_PROBE(aL, bL);
    if (hn) return ORC_ERR_LOWER_NEG;                                 /* :759-760 */
    ORC_PROBE(aR, bR);
    if (!hn) return ORC_ERR_UPPER_NONNEG;                             /* :761-762 */
    int64_t iterations = 0;
    for (int64_t step = 0; step < max_steps; step++) {                 /* :766 */
        iterations++;
        int64_t aM = aL + aR, bM = bL + bR;                           /* :768 */
        int found;
        if (bM > max_den) {                                           /* :770-780 */
            rc = orc_find_zero_cycle_exact(g, aL, bL, &found, cyc_buf, cyc_len, sum_c, sum_t);
            if (rc != ORC_OK) return rc;
            if (found) {
                int64_t gg = orc_gcd(aL, bL);
                *a_out = orc_floordiv(aL, gg); *b_out = orc_floordiv(bL, gg);
                *iterations_out = -1;   /* reference does not record iterations on this path (:772-776) */
                return ORC_OK;
            }
            int64_t k = orc_floordiv(max_den - bL, bR);
            if (k < 1) k = 1;
            if (llabs(aR) > INT64_MAX / (k + 1) || bR > INT64_MAX / (k + 1)) return ORC_ERR_OVERFLOW;
            aM = aL + k * aR; bM = bL + k * bR;
        }
        ORC_PROBE(aM, bM);                                            /* :782 */
        if (hn) { aR = aM; bR = bM; continue; }                       /* :783-784 */
        rc = orc_find_zero_cycle_exact(g, aM, bM, &found, cyc_buf, cyc_len, sum_c, sum_t);   /* :787 */
        if (rc != ORC_OK) return rc;
        if (found) {                                                  /* :788-793 */
            int64_t gg = orc_gcd(aM, bM);
            *a_out = orc_floordiv(aM, gg); *b_out = orc_floordiv(bM, gg);
            *iterations_out = iterations;
            return ORC_OK;
        }
        aL = aM; bL = bM;                                             /* :795 */
    }
#undef ORC_PROBE
    return ORC_ERR_SB_STEPS;                                          /* :797 */
}
