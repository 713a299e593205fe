"""Per-pass relaxation timing and whole-solve timing for one library build (MRC_B200_LIB selects it); C3 graph by default."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from min_ratio_cycle_b200 import _native as N, synthetic

n = int(os.environ.get("N", 1_000_000)); m = int(os.environ.get("M", 10_000_000)); seeds = [int(x) for x in os.environ.get("SEEDS", "0").split(",")]
tag = os.environ.get("MRC_B200_LIB", "default").split("/")[-1]
for seed in seeds:
    _, U, V, C, T = synthetic.random_float(n, m, seed=seed)
    o = np.argsort(V, kind="stable")
    g = N.DeviceGraph(n, U[o].astype(np.int32), V[o].astype(np.int32), C[o], T[o])
    lo, hi = g.bounds()
    for kv in filter(None, os.environ.get("OPTS", "").split(",")):
        k, v = kv.split("="); g.set_option(k, int(v)); tag = tag + " " + kv if kv not in tag else tag
    g.set_option("sparse", 0)
    base = g.solve_numeric(lo, hi, K=1)
    r = base["ratio"]
    if seed == seeds[0] and not os.environ.get("NO_PASSES"):
        for K in (1, 2, 4, 8):
            ms = g.bench_relax([r - 0.5 - 0.1 * k for k in range(K)], 30)
            print(f"[{tag}] relax K={K}: pass1 {ms[0]*1e3:6.1f} pass2 {ms[1]*1e3:6.1f} pass3 {ms[2]*1e3:6.1f} late(21-30) {ms[20:].mean()*1e3:6.1f} us", flush=True)
    modes = [tuple(int(x) for x in md.split(":")) for md in os.environ.get("MODES", "4:0,4:1,1:0,1:1").split(",")]
    for K, sparse in (() if os.environ.get("ONLY_PASSES") else modes):
        g.set_option("sparse", sparse)
        for _ in range(2):
            g.solve_numeric(lo, hi, K=K)
        g.reset_stats()
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            out = g.solve_numeric(lo, hi, K=K)
        dt = (time.perf_counter() - t0) / reps
        st = g.stats()
        assert abs(out["ratio"] - r) <= 1e-9 * max(1, abs(r)), (out["ratio"], r)
        print(f"[{tag}] seed {seed} K={K} sparse={sparse}: {dt*1e3:6.2f} ms/solve  rounds {out['iterations']}  relax {st['relax_launches']/reps:5.1f} x "
              f"{1e3*st['relax_ms']/max(st['relax_launches'],1):5.1f} us = {st['relax_ms']/reps:5.2f} ms  sparse {st['sparse_passes']/reps:5.1f} passes {st['sparse_ms']/reps:5.2f} ms  "
              f"check {st['check_launches']/reps:4.1f} x {1e3*st['check_ms']/max(st['check_launches'],1):5.1f} us = {st['check_ms']/reps:5.2f} ms "
              f"(doubling {st['check_doubling_ms']/reps:4.2f} in {st['check_rounds']/max(st['check_launches'],1):3.1f} rounds, extract {st['check_extract_ms']/reps:4.2f})  "
              f"crowded {st['crowded_sweeps']/reps:3.1f} x {1e3*st['crowded_ms']/max(st['crowded_sweeps'],1):5.1f} us  probes {st['probe_launches']/reps:3.1f} {st['probe_ms']/reps:5.2f} ms  "
              f"launches {st['kernel_launches']/reps:5.1f}  cyc_len {len(out['cycle'])-1}", flush=True)
    g.close()
