// mrc_probe.cuh -- one probe = ONE cooperative launch (persistent probe kernel).
//
// The burst driver (ProbeRun in mrc_b200.cu) enqueues a few relaxation launches + a cycle check, reads a status block back,
// and decides on the host what to launch next: narrowing, the compact survivor, the switch to worklist passes, pruning by
// monotonicity.  At C3 every such round trip costs 30-40 us and a solve makes 15-20 of them: a quarter of a 5 ms solve was
// neither relaxation nor check.  Here the same decisions are taken ON THE DEVICE at grid-sync boundaries: the relaxation
// sweeps (relax_sweep), the worklist passes (sparse_pass_device) and the cycle check (cycle_check_device) are the device
// functions the stand-alone kernels are made of, called from one persistent grid of occupancy x 148 SMs CTAs; thread 0 of
// every CTA replays the (identical) control logic from the same global words, so no extra synchronisation is needed to
// agree on what comes next.  The host launches once per probe and reads one result block.
//
// Replaces, like ProbeRun, the pass loop of _detect_negative_cycle_numeric (min_ratio_cycle/solver.py:1124-1294) and of
// _has_negative_cycle_exact (:811-856).
#pragma once
#include "mrc_kernels.cuh"

#define MRC_PW 3   /* sweep widths a probe kernel can run: 1, 2, 4 candidates */

struct ProbeDev {              // result block of one persistent probe
    int verdict[MRC_KMAX];
    ll passes[MRC_KMAX];
    CycleOut cyc[MRC_KMAX];
    // where the time went (globaltimer of thread 0 of CTA 0 at the phase boundaries), ns
    ull ns_sweep[MRC_PW];      // dense sweeps by width (1, 2, 4 candidates)
    ull ns_crowded;            // ... of which the first two sweeps of the probe (every vertex is lowered)
    ull ns_sparse, ns_check, ns_other, ns_total;
    unsigned n_sweep[MRC_PW], n_crowded, n_sparse, n_check, n_compaction, n_colcopy, cuts, pad;
    ull slots;                 // sum over dense sweeps of the number of active candidates
    ull sparse_visits;         // sum over worklist passes of edges examined x active candidates
    ull sparse_examined;       // ... edges examined
    ull ns_chk_phase[2], check_rounds;   // DevStatus::ns_check / check_rounds summed over the probe's checks
};

struct ProbeArgs {
    RelaxArgs r;               // full-width view of the [n][KT] layout: arrays, K candidates, slack (r.changed / r.gate unused)
    CheckArgs c;               // full-width check arguments
    int K;
    unsigned flags;            // MRC_F_MONOTONE_PRUNE | MRC_F_CERT_FIRST
    int asc;                   // candidates in increasing lambda order (pruning by monotonicity is meaningful)
    ll max_passes;
    int first_check, check_growth, check_every;
    // compact survivor
    int may_compact;
    void *cdist;
    ull *cpred;
    int *cjump;
    // worklist passes
    int use_sparse, sparse_div, look0;
    const int *out_off, *out_edge;
    unsigned *bm[2];
    int *wl[2];
    int bm_words;
    const void *warm;          // [n] start potentials, or NULL = zero
    ProbeDev *out;
    // distributed probe (dW > 1; K = 1, numeric): the ranks of a communicator run ONE probe together.  This rank sweeps the
    // slice r.src .. r.src + r.m of the edge stream (global edge ids = local + e_base) = the in-edges of the vertices
    // [v_lo, v_hi) it owns, keeps a full replica of dist (r.dist: peers push what they lower) and of pred (r.pred: own slice
    // live, the others gathered before every check), and agrees with its peers through mailboxes in peer memory.
    int dW, dme;
    int dG;                    // sweeps of the slice between two exchanges (the sweeps of a group run back to back, no barrier of any kind)
    unsigned e_base;
    ll v_lo, v_hi;
    PeerPush pp;               // the other ranks' dist replicas
    ull *peer_pred[MRC_KMAX];  // ... and pred replicas (same order)
    ull *peer_mbox[MRC_KMAX];  // every rank's mailbox, mine included (index = rank)
    ull *mbox;                 // mine: [dW][MRC_MBOX_SLOTS] words (step << 8 | payload), slot = step % MRC_MBOX_SLOTS
    ull seq0;                  // step number of this probe's first message (steps never repeat over the life of a communicator)
};
#define MRC_MBOX_SLOTS 16      /* a rank is never more than one step ahead of the slowest */
#define MRC_X_TIMEOUT 0x80u    /* payload bit: a peer did not answer within MRC_X_TIMEOUT_NS (its process died?) */
#define MRC_X_TIMEOUT_NS 4000000000ull

// The arguments of the probe in flight.  They live in constant memory rather than in the kernel's parameter block so that
// the sweeps can be SEPARATE (noinline) functions with their own register allocation and still take every array pointer and
// lambda as a constant-bank operand, exactly like the stand-alone relax_kernel: inlined into one body together with the
// check and the control code the K = 4 sweep spilled ~1 KB per thread.  One probe per device at a time (a cooperative grid
// fills the device anyway); the host serialises [copy to the symbol, launch] per device.
__constant__ ProbeArgs c_probe;

__device__ __forceinline__ ull mrc_globaltimer() {
    ull t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

struct ProbeCtl {              // per-CTA copy of the control state (identical in every CTA)
    unsigned active;
    int ck;                    // slot served from the compact column, or -1
    int sparse_mode, sp_in, sp_ci;
    unsigned cin;              // size of the worklist the next worklist pass reads
    ll pass, next_check, next_look;
    unsigned step;             // iterations of the pass loop so far (a step is one sweep, or one group of sweeps of a distributed probe)
    int track;                 // the next dense sweep records the vertices it lowers
    int do_compaction, do_check, do_colcopy, do_bmclear;
    int any_neg;
    ull xstep;                 // distributed probe: messages exchanged so far
    int xfail;                 // ... a peer timed out: the probe is abandoned
    int verdict[MRC_KMAX];
    ll passes[MRC_KMAX];
    double sumc[MRC_KMAX], sumt[MRC_KMAX];   // cost / time sums of the verified cycles (pruning by their ratio)
    ll isumc[MRC_KMAX], isumt[MRC_KMAX];
};

// narrowest aligned sub-range of the KT slots that covers every active one (narrow_candidates of the host driver)
__device__ __forceinline__ void mrc_narrow(int KT, unsigned active, int &kc, int &k0) {
    kc = KT; k0 = 0;
    for (int w = 1; w < KT; w *= 2)
        for (int o = 0; o < KT; o += w)
            if (active && !(active & ~(((1u << w) - 1u) << o))) { kc = w; k0 = o; return; }
}

template <int W, int KT, bool EXACT>
__device__ __forceinline__ RegView<W> mrc_reg_view(int k0, bool compact) {
    const ProbeArgs &p = c_probe;
    RegView<W> v;
#pragma unroll
    for (int j = 0; j < W; j++) {
        v.lam_[j] = p.r.lam[k0 + j];
        v.a_[j] = p.r.a[k0 + j];
        v.b_[j] = p.r.b[k0 + j];
    }
    if (compact) { v.dist_ = p.cdist; v.pred_ = p.cpred; v.ks_ = 1; }
    else { v.dist_ = (char *)p.r.dist + (size_t)k0 * 8; v.pred_ = p.r.pred + k0; v.ks_ = KT; }
    v.k0_ = k0;
    return v;
}

// is lambda_j > lambda_k (numeric) / a_j/b_j > a_k/b_k (exact)?
template <bool EXACT> __device__ __forceinline__ bool mrc_above(int j, int k) {
    const ProbeArgs &p = c_probe;
    if (EXACT) return (__int128)p.r.a[j] * p.r.b[k] > (__int128)p.r.a[k] * p.r.b[j];
    return p.r.lam[j] > p.r.lam[k];
}

struct ProbeClock {            // chief's clock and counters (thread 0 of CTA 0)
    ProbeDev acc;
    ull t_start, t_prev, examined_prev;
};
__device__ __forceinline__ ull mrc_lap(ProbeClock *c) {
    const ull t = mrc_globaltimer(), d = t - c->t_prev;
    c->t_prev = t;
    return d;
}

// Every part of the probe is a separate function with its own register allocation; the kernel body is the loop that calls them.

// ---- start: dist = 0 (or the warm potentials), pred = none, bitmaps clean, status words zero
template <int KT>
__device__ __noinline__ void probe_start(ProbeCtl *s, ProbeClock *clk) {
    const ProbeArgs &p = c_probe;
    DevStatus *st = p.c.st;
    const ll n = p.r.n;
    const ll tid0 = (ll)blockIdx.x * blockDim.x + threadIdx.x, stride = (ll)gridDim.x * blockDim.x;
    ull *dist = reinterpret_cast<ull *>(p.r.dist);
    const ull *warm = reinterpret_cast<const ull *>(p.warm);
    for (ll i = tid0; i < n * KT; i += stride) {
        __stcg(dist + i, warm ? __ldg(warm + i / KT) : 0ull);
        __stcg(p.r.pred + i, ~0ull);
    }
    if (p.use_sparse)
        for (ll i = tid0; i < p.bm_words; i += stride) { p.bm[0][i] = 0u; p.bm[1][i] = 0u; }
    if (threadIdx.x != 0) return;
    if (blockIdx.x == 0) {
        for (int i = 0; i < MRC_MAX_BURST; i++) st->changed[i] = 0u;
        st->examined = 0ull;
        st->ns_check[0] = st->ns_check[1] = st->check_rounds = 0ull;
        st->decided = 0u;
        st->act_count = 0u;
        st->act_cnt[0] = st->act_cnt[1] = st->act_cnt[2] = 0u;
        for (int i = 0; i < 4; i++) st->xword[i] = 0ull;   // (a graph may serve several communicators, whose step numbers repeat)
        ProbeDev &acc = clk->acc;
        for (int i = 0; i < MRC_PW; i++) { acc.ns_sweep[i] = 0; acc.n_sweep[i] = 0; }
        acc.ns_crowded = acc.ns_sparse = acc.ns_check = acc.ns_other = 0;
        acc.n_crowded = acc.n_sparse = acc.n_check = acc.n_compaction = acc.n_colcopy = acc.cuts = 0;
        acc.slots = acc.sparse_visits = 0;
        clk->examined_prev = 0;
    }
    s->active = (1u << p.K) - 1u;
    s->ck = -1;
    s->sparse_mode = 0; s->sp_in = 0; s->sp_ci = 0; s->cin = 0u;
    s->step = 0u;
    s->pass = 0; s->next_check = p.first_check; s->next_look = p.look0 - 1;
    s->track = p.use_sparse && p.look0 <= 1;
    s->do_compaction = s->do_check = s->do_colcopy = s->do_bmclear = 0;
    s->any_neg = 0;
    s->xstep = 0; s->xfail = 0;
    for (int k = 0; k < MRC_KMAX; k++) { s->verdict[k] = -1; s->passes[k] = 0; }
}

__device__ __forceinline__ void probe_drop(ProbeCtl *s, int k, int verdict) {
    if ((s->active >> k) & 1u) { s->verdict[k] = verdict; s->passes[k] = s->pass; s->active &= ~(1u << k); }
}
// verdicts are monotone in lambda: NEG at x implies NEG above x (and above the ratio of x's cycle), NONNEG at x implies
// NONNEG below x (ProbeRun::collect)
template <bool EXACT> __device__ __forceinline__ void probe_prune(ProbeCtl *s) {
    const ProbeArgs &p = c_probe;
    if (!(p.flags & MRC_F_MONOTONE_PRUNE)) return;
    const int K = p.K;
    for (int k = 0; k < K; k++) {
        const int vk = s->verdict[k];
        if (vk != MRC_V_NEG && vk != MRC_V_NONNEG) continue;
        for (int j = 0; j < K; j++) {
            if (!((s->active >> j) & 1u)) continue;
            bool above = mrc_above<EXACT>(j, k);
            const bool below = mrc_above<EXACT>(k, j);
            if (vk == MRC_V_NEG) {
                if (!EXACT && s->sumt[k] > 0 && p.r.lam[j] > s->sumc[k] / s->sumt[k]) above = true;
                if (EXACT && (__int128)p.r.a[j] * s->isumt[k] > (__int128)s->isumc[k] * p.r.b[j]) above = true;
                if (above) probe_drop(s, j, MRC_V_SKIPPED_NEG);
            } else if (below) probe_drop(s, j, MRC_V_SKIPPED_NONNEG);
        }
    }
}

// ---- distributed probe: one word per rank and step through mailboxes in peer memory (NVLink P2P stores; no host, no NCCL).
// Called by thread 0 of every CTA after a grid sync.  The chief (CTA 0) sends this rank's payload to every rank, waits for
// all of them and publishes the OR in st->xword; the other leaders wait for that word, so every CTA takes the same decision
// even when a peer times out.  Everything this GPU wrote to peer memory before the grid sync is visible to a peer once it
// has received the step's message (fence.sys on both sides).
__device__ __noinline__ unsigned probe_exchange(ProbeCtl *s, unsigned payload) {
    const ProbeArgs &p = c_probe;
    DevStatus *st = p.c.st;
    const ull step = p.seq0 + s->xstep;
    s->xstep++;
    volatile ull *xw = &st->xword[step & 3];
    if (blockIdx.x == 0) {
        __threadfence_system();
        const ull msg = (step << 8) | (ull)(payload & 0x7fu);
        for (int r = 0; r < p.dW; r++) {
            volatile ull *slot = p.peer_mbox[r] + (size_t)p.dme * MRC_MBOX_SLOTS + (step % MRC_MBOX_SLOTS);
            *slot = msg;
        }
        __threadfence_system();
        unsigned acc = 0;
        const ull t0 = mrc_globaltimer();
        for (int r = 0; r < p.dW && !(acc & MRC_X_TIMEOUT); r++) {
            volatile ull *slot = p.mbox + (size_t)r * MRC_MBOX_SLOTS + (step % MRC_MBOX_SLOTS);
            for (;;) {
                const ull v = *slot;
                if ((v >> 8) == step) { acc |= (unsigned)(v & 0xffu); break; }
                if (mrc_globaltimer() - t0 > MRC_X_TIMEOUT_NS) { acc |= MRC_X_TIMEOUT; break; }
            }
        }
        __threadfence_system();
        *xw = (step << 8) | acc;
        __threadfence();
        if (acc & MRC_X_TIMEOUT) s->xfail = 1;
        return acc;
    }
    ull v;
    do { v = *xw; } while ((v >> 8) != step);
    __threadfence();
    if (v & MRC_X_TIMEOUT) s->xfail = 1;
    return (unsigned)(v & 0xffu);
}

// ---- control point after a pass (every CTA's thread 0: same inputs, same answer): what did it decide, what comes next
template <int KT, bool EXACT>
__device__ __noinline__ void probe_after_pass(ProbeCtl *s, ProbeClock *clk) {
    const ProbeArgs &p = c_probe;
    if (threadIdx.x != 0) return;
    DevStatus *st = p.c.st;
    const unsigned active = s->active;
    const ll pass = s->pass;
    const bool sparse_now = s->sparse_mode != 0;
    unsigned *chw = &st->changed[s->step & 3u];
    if (blockIdx.x == 0) {
        const ull d = mrc_lap(clk);
        ProbeDev &acc = clk->acc;
        if (sparse_now) {
            acc.ns_sparse += d; acc.n_sparse++;
            const ull ex = __ldcg(&st->examined);
            acc.sparse_visits += (ex - clk->examined_prev) * (ull)__popc(active);
            clk->examined_prev = ex;
        } else {
            int kc = KT, k0 = 0;
            if (s->ck >= 0) kc = 1; else mrc_narrow(KT, active, kc, k0);
            const int wi = kc == 1 ? 0 : kc == 2 ? 1 : 2;
            const unsigned G = p.dW > 1 ? (unsigned)p.dG : 1u;   // sweeps of this step
            acc.ns_sweep[wi] += d; acc.n_sweep[wi] += G;
            if (pass < 2) { acc.ns_crowded += d; acc.n_crowded += G; }
            acc.slots += (ull)__popc(active) * G;
        }
        st->changed[(s->step + 2u) & 3u] = 0u;   // last read two control points ago, next written two sweeps from now
    }
    unsigned chg = __ldcg(chw);
    ll done = 1;   // sweeps of this step
    if (p.dW > 1) {   // distributed probe: a candidate is at its fixed point only if NO rank lowered anything in this group of sweeps
        done = p.dG;
        chg = probe_exchange(s, chg & 1u);
        if (chg & MRC_X_TIMEOUT) { s->step++; s->pass = pass + done; s->active = 0u; s->do_compaction = s->do_check = s->do_colcopy = s->do_bmclear = 0; return; }
        chg &= 1u;
    }
    s->step++;
    s->pass = pass + done;
    for (int k = 0; k < p.K; k++)
        if (((active >> k) & 1u) && !((chg >> k) & 1u)) {   // nothing moved: a true fixed point
            s->verdict[k] = MRC_V_NONNEG; s->passes[k] = pass + done;
        }
    s->active = active & chg;
    probe_prune<EXACT>(s);
    s->do_compaction = s->do_check = s->do_colcopy = s->do_bmclear = 0;
    if (s->active) {
        if (sparse_now) {
            const int co = s->sp_ci == 2 ? 0 : s->sp_ci + 1;
            s->sp_in ^= 1; s->sp_ci = co;
            s->cin = __ldcg(&st->act_cnt[co]);
            if ((ll)s->cin * p.sparse_div > 2 * p.r.n) { s->sparse_mode = 0; s->do_bmclear = 1; s->next_look = s->pass + p.look0; }   // the worklist grew back
        } else if (s->track) s->do_compaction = 1;
        s->do_check = s->pass >= s->next_check || s->pass >= p.max_passes;
    }
}

// ---- worklist of the vertices the last sweep lowered (compact_active_kernel); the bitmap is left clean
__device__ __noinline__ void probe_compaction(ProbeCtl *s, ProbeClock *clk, cg::grid_group &grid) {
    const ProbeArgs &p = c_probe;
    DevStatus *st = p.c.st;
    const ll n = p.r.n;
    const ll tid0 = (ll)blockIdx.x * blockDim.x + threadIdx.x, stride = (ll)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    for (ll w0 = tid0 - lane; w0 < p.bm_words; w0 += stride) {
        const ll w = w0 + lane;
        const unsigned bits = w < p.bm_words ? __ldcg(p.bm[0] + w) : 0u;
        if (bits) p.bm[0][w] = 0u;
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
        if (lane == 0) base = atomicAdd(&st->act_cnt[0], (unsigned)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        int pos = (int)base + incl - c;
        unsigned b = bits;
        while (b) {
            const int i = __ffs(b) - 1;
            b &= b - 1;
            p.wl[0][pos++] = (int)((w << 5) | i);
        }
    }
    grid.sync();
    if (threadIdx.x == 0) {
        const unsigned cnt = __ldcg(&st->act_cnt[0]);
        const ll scaled = (ll)cnt * p.sparse_div;
        if (scaled < n) { s->sparse_mode = 1; s->sp_in = 0; s->sp_ci = 0; s->cin = cnt; }
        else s->next_look = s->pass + (scaled > 8 * n ? 8 : scaled > 4 * n ? 4 : scaled > 2 * n ? 2 : 1) * (ll)p.look0 - 1;
        if (blockIdx.x == 0) { clk->acc.ns_other += mrc_lap(clk); clk->acc.n_compaction++; }
    }
    __syncthreads();
    if (!s->sparse_mode) {   // grid-uniform: not yet -- the counter must be zero for the next compaction
        grid.sync();         // (every leader has read it)
        if (tid0 == 0) st->act_cnt[0] = 0u;
    }
}

// ---- rare: back to dense sweeps, the pending worklist and its membership bits are dropped
__device__ __noinline__ void probe_bmclear(cg::grid_group &grid) {
    const ProbeArgs &p = c_probe;
    DevStatus *st = p.c.st;
    const ll tid0 = (ll)blockIdx.x * blockDim.x + threadIdx.x, stride = (ll)gridDim.x * blockDim.x;
    for (ll i = tid0; i < p.bm_words; i += stride) { p.bm[0][i] = 0u; p.bm[1][i] = 0u; }
    if (tid0 == 0) st->act_cnt[0] = st->act_cnt[1] = st->act_cnt[2] = 0u;
    grid.sync();
}

// ---- cycle check of the undecided candidates + what its verdicts imply
template <int KT, bool EXACT>
__device__ __noinline__ void probe_check(ProbeCtl *s, ProbeClock *clk, cg::grid_group &grid) {
    const ProbeArgs &p = c_probe;
    DevStatus *st = p.c.st;
    const unsigned act_now = s->active;
    const int ck = s->ck, K = p.K;
    if (p.dW > 1) {
        // every rank brings its slice of pred to every peer, then all of them run the same check on identical copies
        const ll tid0 = (ll)blockIdx.x * blockDim.x + threadIdx.x, stride = (ll)gridDim.x * blockDim.x;
        for (ll i = p.v_lo + tid0; i < p.v_hi; i += stride) {
            const ull x = __ldcg(p.r.pred + i);
            for (int r = 0; r < p.pp.n_peers; r++) p.peer_pred[r][i] = x;
        }
        __threadfence_system();
        grid.sync();
        if (threadIdx.x == 0) probe_exchange(s, 0u);
        __syncthreads();
        if (s->xfail) { if (threadIdx.x == 0) s->active = 0u; return; }   // grid-uniform
    }
    unsigned cyc;
    if (KT > 1 && ck >= 0) {
        CheckRegView cv;
        cv.pred_ = p.cpred; cv.dist_ = p.cdist; cv.jump_ = p.cjump; cv.cyc_edges_ = p.c.cyc_edges + (ll)ck * p.c.cap;
        cv.active_ = 1u; cv.k0_ = ck; cv.prune_ = 0;
        cv.lam_ = p.r.lam[ck]; cv.a_ = p.r.a[ck]; cv.b_ = p.r.b[ck];
        cyc = cycle_check_device<1>(p.c, cv, grid);
    } else {
        cyc = cycle_check_device<KT>(p.c, CheckActiveView{p.c, act_now}, grid);
    }
    grid.sync();   // st->out[] of the last extraction round is visible to every leader
    if (threadIdx.x != 0) return;
    for (int k = 0; k < K; k++) {
        if (!((act_now >> k) & 1u)) continue;
        const int ko = ck >= 0 ? 0 : k;
        const int found = ((cyc >> ko) & 1u) ? __ldcg(&st->out[ko].found) : 0;
        if (!found) continue;
        s->verdict[k] = found == MRC_CYC_NEG ? MRC_V_NEG : MRC_V_TIE;
        s->passes[k] = s->pass;
        s->active &= ~(1u << k);
        s->any_neg |= found == MRC_CYC_NEG;
        s->sumc[k] = __ldcg(&st->out[ko].sumc); s->sumt[k] = __ldcg(&st->out[ko].sumt);
        s->isumc[k] = __ldcg(&st->out[ko].isumc); s->isumt[k] = __ldcg(&st->out[ko].isumt);
        if (blockIdx.x == 0) p.out->cyc[k] = st->out[ko];
    }
    probe_prune<EXACT>(s);
    // MRC_F_CERT_FIRST: nothing turned negative by the second check and the certifier (last candidate) is still relaxing --
    // its companions can only raise lo; it continues alone (ProbeRun::collect)
    if ((p.flags & MRC_F_CERT_FIRST) && s->pass >= 6 && K > 1 && ((s->active >> (K - 1)) & 1u) && (s->active & (s->active - 1))) {
        bool neg_seen = false;
        for (int k = 0; k < K; k++)
            neg_seen |= s->verdict[k] == MRC_V_NEG || s->verdict[k] == MRC_V_TIE || s->verdict[k] == MRC_V_SKIPPED_NEG;
        if (!neg_seen) for (int k = 0; k + 1 < K; k++) probe_drop(s, k, MRC_V_ABANDONED);
    }
    if (s->pass >= p.max_passes)   // still relaxing after n passes (reference rule :1252-1257)
        for (int k = 0; k < K; k++) probe_drop(s, k, MRC_V_NEG_NOCYCLE);
    const ll grow = s->next_check * p.check_growth / 100;
    s->next_check = p.check_every > 0 ? s->next_check + p.check_every : p.check_growth > 0 ? s->next_check + (grow > 4 ? grow : 4) : 2 * s->next_check + 2;
    if (blockIdx.x == 0) { clk->acc.ns_check += mrc_lap(clk); clk->acc.n_check++; clk->acc.cuts += __ldcg(&st->cuts); }
}

// ---- the next pass: does it record what it lowers (worklist seed)?  does the survivor move to the compact column?
template <int KT>
__device__ __noinline__ void probe_next(ProbeCtl *s) {
    const ProbeArgs &p = c_probe;
    if (threadIdx.x != 0) return;
    s->track = p.use_sparse && !s->sparse_mode && s->active && s->pass >= s->next_look;
    s->do_colcopy = p.may_compact && KT > 1 && s->ck < 0 && s->active && !(s->active & (s->active - 1));
    if (s->do_colcopy) s->ck = 31 - __clz(s->active);
}
template <int KT>
__device__ __noinline__ void probe_colcopy(ProbeCtl *s, ProbeClock *clk, cg::grid_group &grid) {
    const ProbeArgs &p = c_probe;
    const ll n = p.r.n;
    const ll tid0 = (ll)blockIdx.x * blockDim.x + threadIdx.x, stride = (ll)gridDim.x * blockDim.x;
    const int c1 = s->ck;
    const ull *sd = reinterpret_cast<const ull *>(p.r.dist) + c1, *sp = p.r.pred + c1;
    ull *dd = reinterpret_cast<ull *>(p.cdist), *dp = p.cpred;
    for (ll i = tid0; i < n; i += stride) {
        dd[i] = __ldcg(sd + i * KT);
        dp[i] = __ldcg(sp + i * KT);
    }
    grid.sync();
    if (tid0 == 0) { clk->acc.ns_other += mrc_lap(clk); clk->acc.n_colcopy++; }
}

__device__ __noinline__ void probe_finish(const ProbeCtl *s, const ProbeClock *clk) {
    const ProbeArgs &p = c_probe;
    ProbeDev *o = p.out;
    const ProbeDev &acc = clk->acc;
    for (int k = 0; k < MRC_KMAX; k++) { o->verdict[k] = s->verdict[k]; o->passes[k] = s->passes[k]; }
    for (int i = 0; i < MRC_PW; i++) { o->ns_sweep[i] = acc.ns_sweep[i]; o->n_sweep[i] = acc.n_sweep[i]; }
    o->ns_crowded = acc.ns_crowded; o->ns_sparse = acc.ns_sparse; o->ns_check = acc.ns_check; o->ns_other = acc.ns_other;
    o->ns_total = mrc_globaltimer() - clk->t_start;
    o->n_crowded = acc.n_crowded; o->n_sparse = acc.n_sparse; o->n_check = acc.n_check; o->n_compaction = acc.n_compaction;
    o->n_colcopy = acc.n_colcopy; o->cuts = acc.cuts;
    o->slots = acc.slots; o->sparse_visits = acc.sparse_visits; o->sparse_examined = clk->examined_prev;
    const DevStatus *st = p.c.st;
    o->ns_chk_phase[0] = st->ns_check[0]; o->ns_chk_phase[1] = st->ns_check[1]; o->check_rounds = st->check_rounds;
}

// ---- the hot parts: one function per sweep width
#ifndef MRC_PROBE_PF
#define MRC_PROBE_PF 2         /* prefetch mode of the probe's sweeps (relax_sweep): 2 = prefetch.global.L2 one chunk ahead, no register double buffer */
#endif
#ifndef MRC_PROBE_MINBLOCKS
#define MRC_PROBE_MINBLOCKS 4  /* CTAs per SM the probe kernel is compiled for (3 was measured: slower, profiles/r02_persist.txt) */
#endif
template <int KT, bool EXACT>
__device__ __noinline__ void probe_sweep_full(unsigned active, unsigned *chw, unsigned *act_out) {
    relax_sweep<KT, EXACT, false, ParamView, MRC_PROBE_PF>(c_probe.r, ParamView{c_probe.r}, active, chw, act_out);
}
template <int W, int KT, bool EXACT>
__device__ __noinline__ void probe_sweep_narrow(int k0, bool compact, unsigned act, unsigned *chw, unsigned *act_out) {
    const RegView<W> v = mrc_reg_view<W, KT, EXACT>(k0, compact);
    relax_sweep<W, EXACT, false, RegView<W>, MRC_PROBE_PF>(c_probe.r, v, act, chw, act_out);
}
// the distributed sweep: this rank's slice of the edge stream, lowered values pushed to the peers' replicas
__device__ __noinline__ void probe_sweep_dist(unsigned *chw) {
    const ProbeArgs &p = c_probe;
    DistView v;
    v.lam_ = p.r.lam[0]; v.a_ = p.r.a[0]; v.b_ = p.r.b[0];
    v.dist_ = p.r.dist; v.pred_ = p.r.pred; v.e_base_ = p.e_base; v.pp_ = &p.pp;
    // A group of sweeps between two exchanges: at 8 ranks a sweep of the slice takes ~7 us and an exchange (grid sync, one word
    // per rank over NVLink, control) ~15 us.  Relaxation needs no barrier between sweeps; only the fixed-point test does, and
    // it stays exact: the candidate is declared converged when NO rank lowered anything during a WHOLE group.
    for (int i = 0; i < p.dG; i++) relax_sweep<1, false, false, DistView, MRC_PROBE_PF>(p.r, v, 1u, chw, nullptr);
    __threadfence_system();   // my pushes are performed before the grid sync that precedes this step's message
}

template <int KT, bool EXACT>
__device__ __noinline__ void probe_sparse_pass(unsigned active, int ck, int in, unsigned cin, int co, unsigned *chw) {
    const ProbeArgs &p = c_probe;
    const int out = in ^ 1;
    if (KT > 1 && ck >= 0) {
        const RegView<1> v = mrc_reg_view<1, KT, EXACT>(ck, true);
        sparse_pass_device<EXACT>(p.r, v, 1, 1u, p.out_off, p.out_edge, p.wl[in], p.wl[out], p.bm[in], p.bm[out], p.c.st, cin, co, chw);
    } else {
        sparse_pass_device<EXACT>(p.r, ParamView{p.r}, KT, active, p.out_off, p.out_edge, p.wl[in], p.wl[out], p.bm[in], p.bm[out], p.c.st,
                                  cin, co, chw);
    }
}

template <int KT, bool EXACT>
__global__ void __launch_bounds__(MRC_RELAX_THREADS, MRC_PROBE_MINBLOCKS) probe_kernel() {
    cg::grid_group grid = cg::this_grid();
    __shared__ ProbeCtl s;
    __shared__ ProbeClock clk;
    volatile ProbeCtl &vs = s;

    probe_start<KT>(&s, &clk);
    grid.sync();
    if constexpr (KT == 1 && !EXACT) {
        if (c_probe.dW > 1) {   // no rank pushes into a replica that its owner is still initialising
            if (threadIdx.x == 0) probe_exchange(&s, 0u);
            __syncthreads();
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) clk.t_start = clk.t_prev = mrc_globaltimer();

    for (;;) {
        __syncthreads();
        const unsigned active = vs.active;
        if (!active) break;
        const int ck = vs.ck;
        unsigned *chw = &c_probe.c.st->changed[vs.step & 3u];
        if (vs.sparse_mode) {
            // ---- one worklist pass
            const int ci = vs.sp_ci, co = ci == 2 ? 0 : ci + 1, cz = co == 2 ? 0 : co + 1;
            if (blockIdx.x == 0 && threadIdx.x == 0) c_probe.c.st->act_cnt[cz] = 0u;   // the NEXT pass appends there; last read one pass ago
            probe_sparse_pass<KT, EXACT>(active, ck, vs.sp_in, vs.cin, co, chw);
        } else {
            // ---- one dense sweep, as narrow as the undecided candidates allow
            unsigned *act_out = vs.track ? c_probe.bm[0] : nullptr;
            int kc = KT, k0 = 0;
            if (ck >= 0) { kc = 1; k0 = ck; } else mrc_narrow(KT, active, kc, k0);
            if (KT == 1 && !EXACT && c_probe.dW > 1) {
                if constexpr (KT == 1 && !EXACT) probe_sweep_dist(chw);
            } else if (kc == KT && ck < 0) probe_sweep_full<KT, EXACT>(active, chw, act_out);
            else if constexpr (KT > 1) {
                if (kc == 1) probe_sweep_narrow<1, KT, EXACT>(k0, ck >= 0, 1u, chw, act_out);
                else if constexpr (KT > 2) probe_sweep_narrow<2, KT, EXACT>(k0, false, active >> k0, chw, act_out);
            }
        }
        grid.sync();
        probe_after_pass<KT, EXACT>(&s, &clk);
        __syncthreads();
        if (vs.do_compaction) probe_compaction(&s, &clk, grid);   // grid-uniform, like every branch below
        if (vs.do_bmclear) probe_bmclear(grid);
        if (vs.do_check) { probe_check<KT, EXACT>(&s, &clk, grid); __syncthreads(); }
        probe_next<KT>(&s);
        __syncthreads();
        if (vs.do_colcopy) probe_colcopy<KT>(&s, &clk, grid);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) probe_finish(&s, &clk);
}
