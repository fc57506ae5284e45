#!/bin/bash
tag=${1:-r2d}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${tag}_tests.log 2>&1
timeout 100 tools/microbench/l2_peak > gpurun_out/${tag}_l2_peak.json 2>&1
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
PB_TRACE=1 timeout 300 python tools/k1_quick.py --steps 2 > gpurun_out/${tag}_trace_C2.log 2>&1
tail -15 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_l2_peak.json; python tools/brief.py gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_trace_C2.log
