#!/bin/bash
# final single-GPU evidence of a round: tests, smoke, bench, launch list, ncu captures (tools/final_n1.sh <tag>)
cd "$(dirname "$0")/.."
tag=${1:-r02f}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/${tag}_tests.log 2>&1
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -8 ) > gpurun_out/${tag}_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2>/dev/null
CMD="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1
$CMD > gpurun_out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:relax_kernel -s 60 -c 3 -o gpurun_out/${tag}_relax $CMD > gpurun_out/${tag}_ncu2.log 2>&1
$CMD > gpurun_out/${tag}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:probe_kernel -s 3 -c 1 -o gpurun_out/${tag}_probe $CMD > gpurun_out/${tag}_ncu3.log 2>&1
tail -4 gpurun_out/${tag}_tests.log; tail -3 gpurun_out/${tag}_smoke.log; python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${tag}_n1.json").read().strip().splitlines()[-1])
print("N=1 value %.1f  ms %.2f  e2e %.1f (%.2f ms)  frac %.3f relax_us %.1f default %.2f ms" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["relax_us_per_pass"], d.get("solve_ms_default", -1)))
PY
ls -la gpurun_out/${tag}_*; tail -2 gpurun_out/${tag}_ncu2.log gpurun_out/${tag}_ncu3.log
