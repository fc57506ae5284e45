#!/bin/bash
# 4-GPU check: the sharded label generator with 64-thread blocks (few replicates per rank), gather kernel, parity
tag=${1:-r2z}; n=${2:-4}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_host_cpp.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/${tag}_tests.log 2>&1
run_t() { name=$1; port=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err; }
run_t bench_n${n} 29571 --steps 5 --warmup 3
PB_GEN_BIG_BLOCKS=1 run_t bench_n${n}_bigblocks 29572 --steps 5 --warmup 3
run_t bench_C3_n${n} 29573 --config C3 --steps 3 --warmup 3
run_t bench_northstar_n${n} 29574 --replicates 1000000 --scaling strong --steps 3 --warmup 3
tail -8 gpurun_out/${tag}_tests.log
python tools/brief.py gpurun_out/${tag}_bench*.json
for f in gpurun_out/${tag}_bench*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $f | tail -4; done
