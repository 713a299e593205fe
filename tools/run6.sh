#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== group mode"
TRACE_W=2 TRACE_K=4 REPS=5 SEEDS=0,1 timeout 600 python tools/multi_gpu_check.py group
echo "== torchrun x2"
REPS=5 SEEDS=0,1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/multi_gpu_check.py
echo "== torchrun x2 dprobe off"
MRC_B200_DPROBE=0 REPS=5 SEEDS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 tools/multi_gpu_check.py
} > gpurun_out/r6_mg.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_comm.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r6_tests.log 2>&1
cat gpurun_out/r6_mg.log | tail -60; tail -8 gpurun_out/r6_tests.log
