#!/bin/bash
# bench.py at N = 1, 2, 4, 8 on one box, the way the driver launches it (tools/scale_run.sh <tag> [Ns...])
tag=${1:-r02}; shift
Ns=${@:-"1 2 4 8"}
port=29600
for N in $Ns; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err
  else
    port=$((port+1))
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${tag}_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value %.1f G/s  ms_per_step %.2f  e2e %.1f G/s (%.2f ms)  frac %.3f  relax_us %.1f  parity %s ticks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["relax_us_per_pass"], d.get("parity_ok"), d.get("ticks_per_step")))
except Exception as e:
    print("N=$N failed:", e); print(open("gpurun_out/bench_${tag}_n$N.err").read()[-1500:])
PY
done
