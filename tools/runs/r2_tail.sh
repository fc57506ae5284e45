#!/bin/bash
# eager tail + early per-site records: parity tests, then the end-to-end bench with and without it
tag=${1:-r2n}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_shapes.py tests/test_gpu_parity.py tests/test_multi_gpu.py tests/test_host_cpp.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${tag}_tests.log 2>&1
tail -15 gpurun_out/${tag}_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-exact-null > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-exact-null --no-early-output > gpurun_out/${tag}_bench_noearly.json 2> gpurun_out/${tag}_bench_noearly.err
PB_NO_EAGER_TAIL=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-exact-null > gpurun_out/${tag}_bench_notail.json 2> gpurun_out/${tag}_bench_notail.err
python tools/brief.py gpurun_out/${tag}_bench*.json
for f in gpurun_out/${tag}_bench*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $f | tail -4; done
