#!/bin/bash
# 2-GPU check of the final build: sharded parity tests (pb_group, peer memory, NCCL, label sharing, native host --devices), bench lines
tag=${1:-r2m}; n=${2:-2}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_host_cpp.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/${tag}_tests.log 2>&1
run_t() { name=$1; port=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err; }
run_t bench_n${n} 29561 --steps 5 --warmup 3
timeout 600 python bench.py --gpus $n --inproc --steps 5 --warmup 3 > gpurun_out/${tag}_bench_inproc_n${n}.json 2> gpurun_out/${tag}_bench_inproc_n${n}.err
run_t bench_northstar_n${n} 29564 --replicates 1000000 --scaling strong --steps 3 --warmup 3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29565 bench.py --gpus $n --impl reference --steps 1 --warmup 1 > gpurun_out/${tag}_bench_ref_n${n}.json 2> gpurun_out/${tag}_bench_ref_n${n}.err
tail -8 gpurun_out/${tag}_tests.log
python tools/brief.py gpurun_out/${tag}_bench_n*.json gpurun_out/${tag}_bench_inproc*.json gpurun_out/${tag}_bench_north*.json
head -c 300 gpurun_out/${tag}_bench_ref_n${n}.json; echo
for f in gpurun_out/${tag}_bench*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $f | tail -4; done
