#!/bin/bash
# Round 2 one-GPU check: GPU tests, traces, L2 microbenchmark, a short bench line.
tag=${1:-r2b}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${tag}_tests.log 2>&1
PB_TRACE=1 timeout 300 python tools/k1_quick.py --steps 3 > gpurun_out/${tag}_trace_C2.log 2>&1
timeout 100 tools/microbench/l2_peak > gpurun_out/${tag}_l2_peak.json 2>&1
timeout 100 tools/microbench/h2d_peak 1024 > gpurun_out/${tag}_h2d_peak.json 2>&1
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -15 gpurun_out/${tag}_tests.log; tail -3 gpurun_out/${tag}_trace_C2.log; cat gpurun_out/${tag}_l2_peak.json gpurun_out/${tag}_h2d_peak.json; python tools/brief.py gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
