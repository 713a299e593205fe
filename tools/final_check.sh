#!/bin/bash
# last look at a finished build on one GPU: the whole GPU suite and smoke (tools/final_check.sh <tag>)
cd "$(dirname "$0")/.."
tag=${1:-final}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/${tag}_tests.log 2>&1
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -8 ) > gpurun_out/${tag}_smoke.log 2>&1
tail -4 gpurun_out/${tag}_tests.log; tail -3 gpurun_out/${tag}_smoke.log
