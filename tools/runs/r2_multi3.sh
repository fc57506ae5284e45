#!/bin/bash
# 2-GPU validation of label sharing + the strong-scaling bench lines (scaled)
tag=${1:-r2p}; n=${2:-2}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_host_cpp.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${tag}_tests.log 2>&1
run_t() { name=$1; port=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err; }
run_t bench_n${n} 29541 --steps 5 --warmup 3
PB_BENCH_LABEL_SHARING=0 run_t bench_n${n}_noshare 29542 --steps 5 --warmup 3
timeout 600 python bench.py --gpus $n --inproc --steps 5 --warmup 3 > gpurun_out/${tag}_bench_inproc_n${n}.json 2> gpurun_out/${tag}_bench_inproc_n${n}.err
run_t bench_northstar_n${n} 29543 --replicates 1000000 --scaling strong --steps 3 --warmup 3
run_t bench_C3dry_n${n} 29544 --config C3 --sites 1250000 --steps 3 --warmup 3
tail -15 gpurun_out/${tag}_tests.log
python tools/brief.py gpurun_out/${tag}_bench*.json
for f in gpurun_out/${tag}_bench*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $f | tail -4; done
