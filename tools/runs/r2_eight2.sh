#!/bin/bash
# Round 2, second full-box call: label sharing at 8 GPUs — parity, weak C2 (both launch shapes), C3 at full size (strong),
# the north_star target (1M sites x 5k strains, m = 1e6, strong).
tag=${1:-r2y}; n=${2:-8}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_host_cpp.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${tag}_tests.log 2>&1
run_t() { name=$1; port=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err; }
run_t bench_n${n} 29551 --steps 5 --warmup 3
timeout 600 python bench.py --gpus $n --inproc --steps 5 --warmup 3 > gpurun_out/${tag}_bench_inproc_n${n}.json 2> gpurun_out/${tag}_bench_inproc_n${n}.err
run_t bench_C3_n${n} 29552 --config C3 --steps 5 --warmup 3
run_t bench_northstar_n${n} 29553 --replicates 1000000 --scaling strong --steps 5 --warmup 3
timeout 600 python bench.py --gpus $n --inproc --replicates 1000000 --scaling strong --steps 5 --warmup 3 > gpurun_out/${tag}_bench_northstar_inproc_n${n}.json 2> gpurun_out/${tag}_bench_northstar_inproc_n${n}.err
tail -15 gpurun_out/${tag}_tests.log
python tools/brief.py gpurun_out/${tag}_bench*.json
for f in gpurun_out/${tag}_bench*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $f | tail -4; done
