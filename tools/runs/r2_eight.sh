#!/bin/bash
# Round 2, the one full-box call (gpurun --gpus 8): multi-GPU parity on 8 GPUs, the BASELINE configs that need the box,
# the upload ceiling.  Ordered by importance; every leg has its own timeout.
tag=${1:-r2x}; n=${2:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
(timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_host_cpp.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${tag}_tests.log 2>&1
run_t() { name=$1; port=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err; }
run_t bench_n${n} 29531 --steps 5 --warmup 3
timeout 600 python bench.py --gpus $n --inproc --steps 5 --warmup 3 > gpurun_out/${tag}_bench_inproc_n${n}.json 2> gpurun_out/${tag}_bench_inproc_n${n}.err
run_t bench_C3_n${n} 29532 --config C3 --steps 5 --warmup 3
run_t bench_northstar_n${n} 29533 --replicates 1000000 --steps 5 --warmup 3
timeout 200 tools/microbench/h2d_peak 1024 > gpurun_out/${tag}_h2d_peak.json 2>&1
PB_BENCH_EXCHANGE=nccl run_t bench_n${n}_nccl 29534 --steps 5 --warmup 3
timeout 600 python bench.py --gpus $n --inproc --config C3 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_C3_inproc_n${n}.json 2> gpurun_out/${tag}_bench_C3_inproc_n${n}.err
tail -15 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_h2d_peak.json
python tools/brief.py gpurun_out/${tag}_bench*.json
for f in gpurun_out/${tag}_bench*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $f | tail -4; done
