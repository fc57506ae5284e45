#!/bin/bash
# Round 2: ncu --set full captures of the K3 kernels (label generator, tight count) at C2 and the count kernels at C4,
# after a plain run of the same command; PB_TRACE timelines of both.
tag=${1:-r2a}
mkdir -p gpurun_out
PB_TRACE=1 timeout 300 python tools/k1_quick.py --steps 3 > gpurun_out/${tag}_trace_C2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k3_gen_kernel -c 1 -o gpurun_out/${tag}_k3gen -f python tools/k1_quick.py --steps 1 > gpurun_out/${tag}_ncu_gen.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k3_count_tight -c 1 -o gpurun_out/${tag}_k3tight -f python tools/k1_quick.py --steps 1 > gpurun_out/${tag}_ncu_tight.log 2>&1
PB_TRACE=1 timeout 300 python tools/k1_quick.py --config C4 --steps 2 > gpurun_out/${tag}_trace_C4.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k3_count_tight -c 1 -o gpurun_out/${tag}_k3tight_C4 -f python tools/k1_quick.py --config C4 --steps 1 > gpurun_out/${tag}_ncu_tight_C4.log 2>&1
tail -4 gpurun_out/${tag}_trace_C2.log; tail -3 gpurun_out/${tag}_trace_C4.log
