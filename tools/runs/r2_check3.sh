#!/bin/bash
tag=${1:-r2e}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${tag}_tests.log 2>&1
{ timeout 200 python tools/k1_quick.py --steps 10; timeout 200 python tools/k1_quick.py --steps 3 --config C4; timeout 200 python tools/k1_quick.py --steps 10 --config C1; timeout 300 python tools/k1_quick.py --steps 5 --config C3 --sites 625000; } > gpurun_out/${tag}_quick.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k3_gen_kernel -c 1 -o gpurun_out/${tag}_k3gen -f python tools/k1_quick.py --steps 1 > gpurun_out/${tag}_ncu_gen.log 2>&1
tail -15 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_quick.log
