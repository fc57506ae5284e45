#!/bin/bash
tag=${1:-r2q}; n=${2:-2}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/${tag}_tests.log 2>&1
{ timeout 200 python tools/k1_quick.py --steps 10; timeout 300 python tools/k1_quick.py --steps 5 --config C3 --sites 625000; } > gpurun_out/${tag}_quick.log 2>&1
run_t() { name=$1; port=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err; }
run_t bench_n${n} 29561 --steps 5 --warmup 3
timeout 600 python bench.py --gpus $n --inproc --steps 5 --warmup 3 > gpurun_out/${tag}_bench_inproc_n${n}.json 2> gpurun_out/${tag}_bench_inproc_n${n}.err
run_t bench_C3dry_n${n} 29564 --config C3 --sites 1250000 --steps 3 --warmup 3
tail -8 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_quick.log
python tools/brief.py gpurun_out/${tag}_bench*.json
for f in gpurun_out/${tag}_bench*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $f | tail -4; done
