/*
 * mrc_b200.h -- C ABI of libmrc_b200.so: the B200 (sm_100a) implementation of the
 * parametric negative-cycle hot path of DiogoRibeiro7/min_ratio_cycle.
 *
 * The reference has no FFI: its "plugin boundary" for this path is the set of private
 * methods that MinRatioCycleSolver.solve() calls (min_ratio_cycle/solver.py).  Every entry
 * point below names the reference method it replaces (file:line).  The reference-side
 * binding a maintainer would add is a ctypes stub; see INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every function returns an int status
 * (MRC_OK == 0), never throws; mrc_last_error() gives the message for the calling thread.
 * Host pointers unless a name starts with d_.  Edge arrays handed to mrc_graph_create are
 * ALREADY destination-sorted in the reference's stable order (solver.py:476-478), so an
 * edge index here is the reference's index into _U/_V/_C/_T.
 *
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * MRC_ERR_NO_DEVICE.
 */
#ifndef MRC_B200_H
#define MRC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRC_ABI_VERSION 4   /* 2: mrc_stats grew (sparse_passes, sparse_ms); 3: mrc_comm_*, sharded solve, options pruned; 4: persistent probe kernel (mrc_stats grew) */
#define MRC_KMAX 8 /* candidate lambdas evaluated by one relaxation launch */

/* status codes */
#define MRC_OK 0
#define MRC_ERR_CUDA 1
#define MRC_ERR_ARG 2
#define MRC_ERR_NO_DEVICE 3
#define MRC_ERR_OVERFLOW 4     /* int64 q*cost - p*time (or its n-fold sum) would wrap; reference wraps silently (solver.py:799-809) */
#define MRC_ERR_NO_CYCLE 5     /* NumericalInstabilityError("No negative cycle found"), solver.py:1087-1090 */
#define MRC_ERR_CONVERGENCE 6  /* ConvergenceError, solver.py:1102-1109 */
#define MRC_ERR_LOWER_NEG 7    /* RuntimeError("Lower bound has negative cycle"), solver.py:759-760 */
#define MRC_ERR_UPPER_NONNEG 8 /* RuntimeError("Upper bound has no negative cycle"), solver.py:761-762 */
#define MRC_ERR_SB_STEPS 9     /* RuntimeError("Stern-Brocot search did not converge"), solver.py:797 */
#define MRC_ERR_NOT_INTEGER 10 /* "Exact mode requires integer weights", solver.py:689-697 */
#define MRC_ERR_TIMEOUT 11     /* TimeoutError, solver.py:1032-1041 */
#define MRC_ERR_EDGE_MISSING 12
#define MRC_ERR_COMM 13         /* NCCL could not be loaded / a collective failed (multi-GPU entry points only) */

/* per-candidate verdicts of a probe */
#define MRC_V_NONNEG 0         /* relaxation reached a fixed point: no negative cycle at this lambda */
#define MRC_V_NEG 1            /* a cycle of the predecessor graph was extracted and verified negative */
#define MRC_V_NEG_NOCYCLE 2    /* still relaxing after n passes (reference: has_neg=True, cycle_data=None) */
#define MRC_V_TIE 3            /* extracted cycle has ratio >= lambda within rounding: lambda ~ ratio of that cycle */
#define MRC_V_SKIPPED_NEG 4    /* not evaluated to the end: a smaller lambda of the batch was already negative */
#define MRC_V_SKIPPED_NONNEG 5 /* not evaluated to the end: a larger lambda of the batch already had no negative cycle */
#define MRC_V_ABANDONED 6      /* not evaluated to the end and nothing implied (MRC_F_STOP_ON_NEG / MRC_F_PATIENCE) */

/* probe flags */
#define MRC_F_MONOTONE_PRUNE 1u /* stop candidates whose verdict is implied by monotonicity in lambda */
#define MRC_F_STOP_ON_NEG 2u     /* end the probe at the first cycle check that verifies a negative cycle: the search only
                                   needs that cycle's ratio to tighten the bracket; the candidates still relaxing are
                                   reported MRC_V_ABANDONED */
#define MRC_F_PATIENCE 4u        /* once a negative cycle is verified, end the probe at the first cycle check whose whole
                                   interval (as long as everything before it) brought no further verdict; the open
                                   candidates -- slow ones next to lambda* -- are reported MRC_V_ABANDONED */

#define MRC_F_CERT_FIRST 16u     /* the LAST candidate is the certifier of a cycle the caller already holds (lambda = its ratio minus a
                                   sliver): the one probe that can end the search.  Once the second cycle check (pass 6) has
                                   passed without any candidate turning negative, the other candidates -- whose verdicts could
                                   only raise the lower end -- are reported MRC_V_ABANDONED and the certifier continues alone */
#define MRC_F_NO_CYCLES 8u       /* do not bring the cycles to the host: cyc_len is the length each would have, the sums are the
                                   device's (accumulated along the predecessor walk); the one cycle the caller wants is
                                   fetched afterwards with mrc_fetch_cycle_sums.  Ignored by mrc_probe_numeric_scenarios. */

typedef struct mrc_graph mrc_graph;

typedef struct mrc_stats {
    int64_t relax_launches;       /* dense relaxation launches that did work */
    int64_t relax_edge_visits;    /* edge relaxations actually evaluated (dense launch: m per candidate; frontier launch: edges with an active source) */
    int64_t check_launches;       /* predecessor-graph cycle checks */
    int64_t probes;               /* candidate lambdas evaluated */
    int64_t rounds;               /* outer search rounds */
    int64_t other_launches;       /* init / tight-mask / gather / bounds kernels */
    double relax_ms;              /* CUDA-event time spent in relaxation launches */
    double check_ms;              /* CUDA-event time spent in cycle checks */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    int64_t false_cycles_cut;     /* predecessor-graph cycles that did not verify negative and were cut */
    int64_t relax_edge_slots;     /* sum over launches of m * (active candidates), whether or not an edge was skipped */
    int64_t sparse_passes;        /* worklist passes (relax_sparse_kernel: only out-edges of the vertices the previous pass lowered) */
    double sparse_ms;             /* CUDA-event time spent in them (not part of relax_ms) */
    int64_t kernel_launches;      /* CUDA kernel launches of this library on the graph's stream (a persistent probe is ONE launch however many
                                     sweeps it runs; relax_launches / check_launches / sparse_passes keep counting sweeps, checks and passes) */
    int64_t probe_launches;       /* persistent probe kernels (csrc/mrc_probe.cuh) */
    double probe_ms;              /* device time inside them (globaltimer): sweeps + checks + worklist passes + grid syncs */
    int64_t crowded_sweeps;       /* persistent probes: of relax_launches, the first two sweeps of each probe (every vertex is lowered) */
    double crowded_ms;            /* ... and their share of relax_ms */
    double check_doubling_ms;     /* inside check_ms (device globaltimer): predecessor-graph pointer doubling ... */
    double check_extract_ms;      /* ... and cycle extraction + verification */
    int64_t check_rounds;         /* pointer-doubling rounds over all checks */
    int64_t relax_stream_bytes;   /* edge-stream bytes the dense sweeps read: 24 B x edges swept (a rank of a distributed probe sweeps its slice only) */
} mrc_stats;

int mrc_abi_version(void);
const char *mrc_last_error(void);
int mrc_device_count(int *count);

/* ---- device-resident graph -------------------------------------------------------------
 * Replaces the array build + implicit "upload" of _build_arrays_if_needed (solver.py:450-511).
 * src/dst: int32[m], dst non-decreasing.  cost/time: f64[m].  icost/itime: int64[m] or both NULL
 * (graph not all-integer, solver.py:491-498).  The library copies everything to `device`.
 */
int mrc_graph_create(int64_t n, int64_t m, const int32_t *src, const int32_t *dst, const double *cost,
                     const double *time, const int64_t *icost, const int64_t *itime, int device,
                     mrc_graph **out);
int mrc_graph_destroy(mrc_graph *g);
int mrc_graph_info(const mrc_graph *g, int64_t *n, int64_t *m, int *device, int *has_int);
/* min/max of cost/time over edges -/+ 1.0: the default bracket of _solve_numeric (solver.py:1019-1023) */
int mrc_numeric_bounds(mrc_graph *g, double *lo, double *hi);
/* get_graph_properties (solver.py:339-387) on the resident arrays.  out[13] = isolated, source, sink vertices, max
 * in-degree, max out-degree, cost min / max / mean / std, time min / max / mean / std (population std, as np.std).
 * solve() stores this dictionary in result.metrics on every call; at m = 10^7 the NumPy version costs 160 ms. */
int mrc_graph_properties(mrc_graph *g, double *out);
int mrc_get_stats(const mrc_graph *g, mrc_stats *out);
int mrc_reset_stats(mrc_graph *g);

/* Perturb edges in place on the device: replaces the per-scenario Edge replacement + array rebuild of
 * sensitivity_analysis (solver.py:1716-1722).  idx are dst-sorted edge indices; deltas are added to the
 * f64 weights.  mrc_graph_restore undoes all outstanding patches (solver.py:1726-1727). */
int mrc_graph_patch_edges(mrc_graph *g, int64_t k, const int64_t *idx, const double *dcost, const double *dtime);
int mrc_graph_restore(mrc_graph *g);

/* ---- numeric probes -----------------------------------------------------------------------
 * K candidate lambdas in ONE batched Bellman-Ford (1 <= K <= MRC_KMAX).  K = 1 is the GPU analogue
 * of _detect_negative_cycle_numeric(lam, slack) (solver.py:1124-1294): same arithmetic per edge
 * (w = cost - lam*time rounded twice, cand = dist[src] + w, strict < with absolute slack), but
 * asynchronous relaxation and early exit as soon as a verified negative cycle closes in the
 * predecessor graph, instead of n passes.
 * Outputs (arrays of K): verdict; cyc_len (0 if none); cycles as CLOSED vertex lists, candidate k at
 * cyc_buf[k*cyc_cap .. ), starting at the smallest vertex of the cycle, truncated to cyc_cap entries;
 * sum_cost/sum_time accumulated in that forward order over the edges Bellman-Ford actually used;
 * passes run for that candidate.  Any output pointer may be NULL.
 */
int mrc_probe_numeric(mrc_graph *g, const double *lam, int K, double slack, uint32_t flags, int32_t *verdict,
                      int32_t *cyc_len, int32_t *cyc_buf, int64_t cyc_cap, double *sum_cost, double *sum_time,
                      int64_t *passes);

/* Scenario batching for sensitivity_analysis (solver.py:1699-1729): candidate k is evaluated on the graph with
 * ONE edge perturbed, patch_edge[k] (dst-sorted index, -1 = none) with cost + dcost[k], time + dtime[k]; the K
 * scenarios share every edge read.  Same outputs as mrc_probe_numeric (sums include the perturbation).  The
 * device-resident edge arrays are not modified. */
int mrc_probe_numeric_scenarios(mrc_graph *g, const double *lam, int K, double slack, uint32_t flags,
                                const int64_t *patch_edge, const double *dcost, const double *dtime, int32_t *verdict,
                                int32_t *cyc_len, int32_t *cyc_buf, int64_t cyc_cap, double *sum_cost,
                                double *sum_time, int64_t *passes);

/* sensitivity_analysis (solver.py:1699-1729) for S single-edge scenarios in one call: verdict[s] (MRC_V_*) of the probe at
 * lam[s] on the graph with edge[s] (dst-sorted index, -1 = none) perturbed by (dcost[s], dtime[s]); K scenarios share every
 * edge sweep, and every group starts from the converged potentials of ONE base probe at lam_base on the unperturbed graph
 * (normally the certifier of the caller's optimum), so a scenario that leaves the optimum alone costs one or two sweeps.
 * NONNEG = no negative cycle at lam[s] in that scenario.  The device-resident arrays are not modified. */
int mrc_scenarios_certify(mrc_graph *g, int64_t S, const int64_t *edge, const double *dcost, const double *dtime,
                          const double *lam, double lam_base, double slack, int K, int32_t *verdict, int64_t *passes);

/* Full cycle of candidate k of the LAST probe (numeric or exact), for callers that probed with a small
 * cyc_cap: CLOSED vertex list; *cycle_len = 0 if that candidate extracted none. */
int mrc_fetch_cycle(mrc_graph *g, int k, int32_t *cycle, int64_t cycle_cap, int64_t *cycle_len);
/* The same plus the cost / time sums accumulated in forward cycle order from the smallest vertex -- what a probe
 * without MRC_F_NO_CYCLES reports (numeric probes; after an exact probe the integer sums converted to double). */
int mrc_fetch_cycle_sums(mrc_graph *g, int k, int32_t *cycle, int64_t cycle_cap, int64_t *cycle_len, double *sum_cost,
                         double *sum_time);

/* ---- numeric lambda search -----------------------------------------------------------------
 * Replaces the bisection of _solve_numeric (solver.py:1017-1117) by K-way multisection: every round
 * evaluates K lambdas of the bracket in one batched probe; a verified negative cycle tightens the upper
 * end to that cycle's own ratio, the largest lambda without negative cycle becomes the lower end; stops
 * when hi - lo <= tol (same absolute test as solver.py:1075).  K = 1 degenerates to bisection with the
 * same midpoints as the reference until the first cycle is found.
 * Returns MRC_ERR_NO_CYCLE / MRC_ERR_CONVERGENCE / MRC_ERR_TIMEOUT like the reference raises.
 * cycle: CLOSED vertex list (cap entries), max_seconds <= 0 = no limit (solver.py:1032-1041).
 * hi_is_cycle != 0: the caller already holds a cycle whose ratio is `hi` (warm start, used by
 * sensitivity_analysis re-solves); MRC_OK with *cycle_len == 0 then means "no better cycle exists".
 */
int mrc_solve_numeric(mrc_graph *g, double lo, double hi, int hi_is_cycle, int max_iter, double tol, double slack,
                      int K, double max_seconds, double *ratio, double *sum_cost, double *sum_time, int32_t *cycle,
                      int64_t cycle_cap, int64_t *cycle_len, int32_t *iterations, double *final_lo,
                      double *final_hi);

/* Measurement aid (bench/profiling only): `passes` relaxation launches for K lambdas from dist == 0, no
 * cycle checks, each launch timed with CUDA events on the library's stream. */
int mrc_bench_relax(mrc_graph *g, const double *lam, int K, int passes, float *ms_per_pass);
/* Tuning knobs (defaults in brackets; measurements: profiles/):
 *   "burst_init" [2], "burst_max" [16]  passes per host round trip (bursts of 2, 4, 8, 16, 16, ...)
 *   "gate" [1]        a relaxation launch skips the candidates the previous launch of its burst did not move and returns at
 *                     once when that is all of them, so bursts can be long without wasting passes after convergence
 *   "narrow" [1]      once some of the K candidates of a probe are decided, the remaining passes run the narrowest kernel
 *                     (1, 2 or 4 candidates wide) that covers the undecided slots of the dist/pred layout
 *   "compact" [1], "compact_min_edges" [2 M]  a probe narrowed to ONE live candidate continues on a compact [n][1] copy of
 *                     its dist/pred column (8 MB of contiguous doubles instead of every K-th double of 8K MB)
 *   "sparse" [0; env MRC_B200_SPARSE; the Python facade switches it on], "sparse_div" [32], "sparse_min_edges" [1.5 M],
 *   "sparse_look" [0 = by graph size]  once fewer than n / sparse_div vertices were lowered by the last pass of a burst, the
 *                     next bursts run as worklist passes over the out-edges of just those vertices (same fixed point, same
 *                     verdicts); smaller graphs keep sweeping; dense bursts are at most sparse_look long while the worklist
 *                     waits to take over
 *   "prune_above" [1] with MRC_F_MONOTONE_PRUNE and candidates in increasing order a cycle check stops working on the
 *                     candidates above the lowest one whose predecessor graph holds a cycle
 *   "cert_first" [1]  mrc_solve_numeric passes MRC_F_CERT_FIRST to the probes of rounds that already hold a cycle
 *   "stop_on_neg" [0] minimum K from which mrc_solve_numeric passes MRC_F_STOP_ON_NEG; "patience" [0]: MRC_F_PATIENCE
 *   "check_growth" [50]  cycle checks after 2 passes and then every max(4, 50 % of the passes so far): 2, 6, 10, 15, 22, 33, 49, ...
 *                     (0: the doubling schedule 2, 6, 14, 30, ...).  A negative probe ends at the first check after its cycle closed;
 *                     with doubling a cycle that closes at pass 36 is noticed at 62
 *   "check_every" [0] experiments: cycle checks every so many passes after the first
 *   "tick_max" [8]    longest tick (passes between two exchanges) of mrc_solve_numeric_sharded
 *   "sparse_poison"   test hook
 * A graph handle is not thread-safe: one probe / solve at a time per mrc_graph (different handles, also on different
 * devices, may be used from different threads). */
int mrc_set_option(mrc_graph *g, const char *name, int value);

/* Quantile levels of the FIRST round of a search that holds no cycle (csrc/mrc_b200.cu, plan_first_round): K increasing levels
 * in (0, 1); candidate j is the q[j]-quantile of a sample of cost/time over the edges instead of a uniform point of
 * [min c/t - 1, max c/t + 1] (solver.py:1019-1023, :1053).  Exposed so that a caller (or a test) can see where a round will look. */
int mrc_first_round_levels(int K, double *q);

/* Pure host helpers shared by mrc_solve_numeric and the multi-GPU driver (one process per GPU, the
 * candidates of a round split across ranks, verdicts all-gathered): no device work.
 * plan: writes K_total candidates in increasing order for the bracket [lo, hi]; have_cycle says whether
 *       hi is the ratio of a known cycle (then the last candidate is the certifier hi - min(0.75*tol, gap/(2K))).
 * reduce: folds K_total (lambda, verdict, ratio-of-extracted-cycle) records into the new bracket; returns
 *       in *best_idx the record whose cycle is the new best (or -1). */
int mrc_plan_candidates(double lo, double hi, int have_cycle, double tol, int K_total, double *lam_out);
int mrc_reduce_round(int K_total, const double *lam, const int32_t *verdict, const double *cyc_ratio, double *lo,
                     double *hi, int *have_cycle, int *best_idx);

/* ---- exact (int64) probes ---------------------------------------------------------------------
 * lambda_k = a[k]/b[k], b > 0; weights w = b*icost - a*itime in int64 (solver.py:799-809), guarded
 * against overflow (MRC_ERR_OVERFLOW) where the reference would wrap.  Verdicts as above (no TIE:
 * integer arithmetic is exact).  K = 1 replaces _has_negative_cycle_exact (solver.py:811-856).
 * sum_cost/sum_time: exact int64 sums over the extracted negative cycle.
 */
int mrc_probe_exact(mrc_graph *g, const int64_t *a, const int64_t *b, int K, uint32_t flags, int32_t *verdict,
                    int32_t *cyc_len, int32_t *cyc_buf, int64_t cyc_cap, int64_t *sum_cost, int64_t *sum_time,
                    int64_t *passes);

/* _compute_potentials_exact (solver.py:922-969): ok = 1 and dist[n] = fixed point of the relaxation from
 * dist == 0 when there is no negative cycle at a/b; ok = 0 otherwise (dist undefined). */
int mrc_potentials_exact(mrc_graph *g, int64_t a, int64_t b, uint8_t *ok, int64_t *dist);

/* (dist[U] + W) == dist[V] of _find_zero_cycle_exact (solver.py:878-879), one byte per edge */
int mrc_tight_mask_exact(mrc_graph *g, int64_t a, int64_t b, uint8_t *ok, uint8_t *mask);

/* _find_zero_cycle_exact (solver.py:858-920) incl. _compute_cycle_weights_exact (:971-995):
 * found = 0 -> None.  cycle is the OPEN list the reference's DFS returns (same order). */
int mrc_zero_cycle_exact(mrc_graph *g, int64_t a, int64_t b, uint8_t *found, int32_t *cycle, int64_t cycle_cap,
                         int64_t *cycle_len, int64_t *sum_cost, int64_t *sum_time);

/* _stern_brocot_search (solver.py:734-797).  max_den / max_steps < 0 = None (defaults :746-751).
 * strategy 0: every Stern-Brocot step probes the GPU (same (a,b) sequence as the reference).
 * strategy 1: lambda* is first found by exact Newton steps on cycle ratios (int64 probes), then the
 *             reference's Stern-Brocot trajectory is replayed on the host with verdicts decided by exact
 *             comparison with lambda*; identical outputs, far fewer probes.
 * a_out/b_out: reduced fraction; iterations: Stern-Brocot steps as the reference counts them (-1 on the
 * path where the reference does not record them, :772-776); probes_ab (2*probes_cap int64, may be NULL)
 * receives the (a,b) sequence the reference would hand to _has_negative_cycle_exact.
 */
int mrc_solve_exact(mrc_graph *g, int64_t max_den, int64_t max_steps, int strategy, int32_t *cycle,
                    int64_t cycle_cap, int64_t *cycle_len, int64_t *sum_cost, int64_t *sum_time, int64_t *a_out,
                    int64_t *b_out, int64_t *iterations, int64_t *probes_ab, int64_t probes_cap,
                    int64_t *nprobes);

/* ---- several GPUs of one box (SURVEY 8b "mrc_comm_init / mrc_comm_destroy", 8e) -----------------------------------
 * The reference has no counterpart: its lambda search is the sequential loop solver.py:1031-1076.  Here every GPU
 * holds a full replica of the edge arrays and evaluates its own lambda candidates; the only traffic between GPUs is,
 * per tick, ONE all-gather of 32 B per candidate (lambda, verdict, cycle ratio, cycle length) enqueued on the
 * library's stream straight from the status block the cycle check wrote (NCCL over NVLink/NVSwitch), and one
 * broadcast of the winning cycle at the end.  NCCL is loaded at run time (libnccl.so.2 already in the process --
 * e.g. PyTorch's -- or the system one); nothing else in this header needs it.
 *
 * Two ways to get a communicator:
 *   one process per GPU (torchrun): rank 0 calls mrc_comm_unique_id, ships the 128 bytes to the other ranks through
 *     whatever channel the job already has (torch.distributed broadcast, MPI, a file), every rank then calls
 *     mrc_comm_init_rank.  Collective.
 *   one process, several GPUs: mrc_comm_init(ndev, devs) (ncclCommInitAll); the library runs one host thread per
 *     device inside mrc_mgraph_* calls.  This is what SolverConfig.gpu_devices uses.
 */
typedef struct mrc_comm mrc_comm;
#define MRC_UNIQUE_ID_BYTES 128
int mrc_comm_unique_id(uint8_t *id /* [MRC_UNIQUE_ID_BYTES] */);
int mrc_comm_init_rank(int nranks, int rank, const uint8_t *id, int device, mrc_comm **out);
int mrc_comm_init(int ndev, const int *devs, mrc_comm **out);
int mrc_comm_destroy(mrc_comm *c);
int mrc_comm_info(const mrc_comm *c, int *nranks, int *rank /* -1: single-process group */, int *device);

/* Replica of a graph on every rank: `root` passes the (destination-sorted) host arrays, the other ranks NULL; ONE
 * host-to-device copy on the root, then ncclBroadcast over NVLink (n = 1 M, m = 10 M: 240 MB, a fraction of a
 * millisecond per GPU) instead of one H2D copy per GPU.  Numeric graphs only.  Collective (rank communicators). */
int mrc_graph_create_replicated(mrc_comm *c, int root, int64_t n, int64_t m, const int32_t *src, const int32_t *dst,
                                const double *cost, const double *time, mrc_graph **out);

/* Numeric lambda search with the candidates sharded over the ranks of `c` (replaces the loop solver.py:1031-1076;
 * single-GPU form: mrc_solve_numeric).  Every rank runs probes of K_local candidates of its own on its replica `g`;
 * after every burst the per-candidate records of all ranks are all-gathered on the device, every rank folds them
 * into the same bracket [lo, hi] (a verified cycle lowers hi to its ratio, a converged candidate raises lo), abandons
 * the candidates of its own that the bracket now implies (lambda >= hi: negative; lambda <= lo: not negative) and,
 * when none is left, starts K_local fresh ones from the current bracket without waiting for the other ranks.  Ends
 * when hi - lo <= tol (same test as solver.py:1075) on the tick every rank sees it.  Outputs as mrc_solve_numeric,
 * identical on every rank (the owner of the best cycle broadcasts it); *iterations = probes started by this rank.
 * Collective: every rank of `c` calls it with the same lo/hi/max_iter/tol/slack/K_local. */
int mrc_solve_numeric_sharded(mrc_comm *c, mrc_graph *g, double lo, double hi, int max_iter, double tol, double slack,
                              int K_local, double max_seconds, double *ratio, double *sum_cost, double *sum_time,
                              int32_t *cycle, int64_t cycle_cap, int64_t *cycle_len, int32_t *iterations,
                              double *final_lo, double *final_hi, int32_t *ticks);

/* Single-process form: replicas of one graph on every device of a group communicator (mrc_comm_init), and the same
 * sharded search driven by one host thread per device.  mrc_mgraph_graph(mg, i) is the replica on device i (for
 * statistics / options; do not destroy it). */
typedef struct mrc_mgraph mrc_mgraph;
int mrc_mgraph_create(mrc_comm *c, int64_t n, int64_t m, const int32_t *src, const int32_t *dst, const double *cost,
                      const double *time, mrc_mgraph **out);
int mrc_mgraph_destroy(mrc_mgraph *mg);
mrc_graph *mrc_mgraph_graph(mrc_mgraph *mg, int i);
int mrc_mgraph_solve_numeric(mrc_mgraph *mg, double lo, double hi, int max_iter, double tol, double slack, int K_local,
                             double max_seconds, double *ratio, double *sum_cost, double *sum_time, int32_t *cycle,
                             int64_t cycle_cap, int64_t *cycle_len, int32_t *iterations, double *final_lo,
                             double *final_hi, int32_t *ticks);

/* Pure host helper of the sharded search (no device, no NCCL; the multi-rank logic is tested with it on CPU):
 * folds the records of one tick -- rec[4*i .. 4*i+3] = (lambda, state, cycle ratio, cycle length), state = -1 while the
 * candidate is still running, else its MRC_V_* verdict -- into the bracket.  *best_idx = index of the record whose
 * cycle became the new best (or -1).  implied[i] (may be NULL) = 1 where a running candidate's verdict now follows
 * from the bracket. */
int mrc_shard_fold(int nrec, const double *rec, double *lo, double *hi, int *have_cycle, int *best_idx, uint8_t *implied);

/* ---- device-side array build -------------------------------------------------------------------
 * The stable sort by destination of _build_arrays_if_needed (np.argsort(V, kind="stable") + gathers,
 * solver.py:476-478) as one radix sort + one gather sweep on the device.  Host arrays in insertion order in,
 * destination-sorted host arrays out (ready for mrc_graph_create); order_out[i] (may be NULL) = insertion index of
 * sorted edge i; dup_pairs_out (may be NULL) = number of edges whose (u,v) also occurs earlier (0 means
 * _remove_dominated_edges, solver.py:531-575, has nothing to examine). */
int mrc_sort_edges_by_dst(int64_t n, int64_t m, const int32_t *src, const int32_t *dst, const double *cost,
                          const double *time, int device, int32_t *src_out, int32_t *dst_out, double *cost_out,
                          double *time_out, int32_t *order_out, int64_t *dup_pairs_out);

/* _preprocess_graph + _build_arrays_if_needed + upload in ONE step (solver.py:450-575): host arrays in INSERTION order
 * (vertex ids as int32 or int64: index_bytes 4 / 8) are uploaded by four host threads through pinned staging buffers;
 * range checks, _remove_dominated_edges (:531-575; remove_dominated != 0), the stable sort by destination (:476-478) and
 * the CSR offsets (:480-489) run on the device and the sorted arrays stay there -- no device -> host -> device round trip
 * of the 24 B/edge.  *m_out = edges left after preprocessing, *dup_pairs = edges whose (u,v) also occurs earlier,
 * *removed = dominated edges dropped.  Numeric graphs only (exact mode needs host mirrors: mrc_graph_create). */
int mrc_graph_build(int64_t n, int64_t m, const void *src, const void *dst, int index_bytes, const double *cost,
                    const double *time, int remove_dominated, int device, mrc_graph **out, int64_t *m_out,
                    int64_t *dup_pairs, int64_t *removed);
/* The destination-sorted arrays of a graph back on the host (any pointer may be NULL): the reference's _U/_V/_C/_T
 * (solver.py:488-491).  order[i] = insertion index, among the edges that survived preprocessing, of sorted edge i and
 * keep[j] (one byte per INPUT edge) = 1 where input edge j survived: graphs made by mrc_graph_build only. */
int mrc_graph_fetch_arrays(mrc_graph *g, int32_t *src, int32_t *dst, double *cost, double *time, int32_t *order,
                           uint8_t *keep);
/* _compute_cycle_weights_float + _verify_cycle_exists (solver.py:1296-1327, :1563-1582) on the resident arrays: closed
 * cycle of L hops (L + 1 vertex ids) in; per hop the first edge with minimal cost/time among the parallel (u,v) edges
 * (:1309-1313); sums accumulated in cycle order from 0.0.  *missing_hop = first hop without an edge, or -1;
 * hop_edges (L ints, may be NULL) = dst-sorted index of each hop's edge (-1 where missing). */
int mrc_cycle_weights(mrc_graph *g, int64_t L, const int32_t *closed_cycle, double *sum_cost, double *sum_time,
                      int32_t *missing_hop, int32_t *hop_edges);

/* ---- batches of small graphs -------------------------------------------------------------------
 * Replaces a loop of B independent solve() calls (each _solve_numeric, solver.py:1001-1122) for graphs that
 * fit one SM's shared memory (n <= 65535, 20*m + 28*K*n bytes <= ~220 KB; BASELINE config 4: n=64, m=4032):
 * one CTA per graph runs the whole multisection search on the device.  Graph g owns edges
 * [edge_off[g], edge_off[g+1]) of the concatenated arrays, destination-sorted, with LOCAL vertex ids.
 * Outputs per graph: status (MRC_OK / MRC_ERR_NO_CYCLE / MRC_ERR_CONVERGENCE), ratio, sums, closed cycle
 * (cycle_buf[g*cycle_cap ..)), rounds.  stats_out (may be NULL): kernel time and relaxation count. */
int mrc_batch_solve_numeric(int64_t B, const int32_t *nverts, const int64_t *edge_off, const int32_t *src,
                            const int32_t *dst, const double *cost, const double *time, int max_iter, double tol,
                            double slack, int K, int device, double *ratio, double *sum_cost, double *sum_time,
                            int32_t *cycle_len, int32_t *cycle_buf, int64_t cycle_cap, int32_t *iterations,
                            int32_t *status, mrc_stats *stats_out);

#ifdef __cplusplus
}
#endif
#endif /* MRC_B200_H */
