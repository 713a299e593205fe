"""
BASELINE config 1 statistics: reference solve() on 16 more seeds of the C1 generator (n=1000, m=5000 simple
random digraph, float weights), ~100 s per seed on the Python reference.  Writes tests/golden/c1_seeds.json.

    python tests/golden/make_golden_c1.py [first_seed last_seed]     (run several ranges in parallel, then --merge)
"""
import json, os, sys, glob
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, py_edges, run_solve, synthetic  # noqa: E402

def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--merge":
        out = {}
        for f in sorted(glob.glob(os.path.join(HERE, "c1_part_*.json"))):
            out.update(json.load(open(f)))
            os.remove(f)
        json.dump(dict(sorted(out.items(), key=lambda kv: int(kv[0]))), open(os.path.join(HERE, "c1_seeds.json"), "w"), indent=0)
        print("merged", len(out))
        return
    a, b = int(sys.argv[1]), int(sys.argv[2])
    S = import_reference()
    out = {}
    for seed in range(a, b + 1):
        n, U, V, C, T = synthetic.random_float(1000, 5000, seed=seed)
        out[str(seed)] = run_solve(S, n, py_edges(U, V, C, T, False))
        print(seed, out[str(seed)].get("ratio"), out[str(seed)].get("seconds"), flush=True)
    json.dump(out, open(os.path.join(HERE, f"c1_part_{a}_{b}.json"), "w"))

main()
