#!/bin/bash
tag=${1:-r2f}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -6) > gpurun_out/${tag}_tests.log 2>&1
{ timeout 200 python tools/k1_quick.py --steps 10; timeout 300 python tools/k1_quick.py --steps 5 --config C3 --sites 625000; } > gpurun_out/${tag}_quick.log 2>&1
tail -6 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_quick.log
