#!/bin/bash
# 2-GPU dry run of the expensive configurations (scaled down) + single-GPU experiments on GPU 0
tag=${1:-r2n}; n=${2:-2}
mkdir -p gpurun_out
{ for v in default g64 g128; do
    if [ $v = default ]; then unset PB_LIB; else export PB_LIB=exp/libpb_$v.so; fi
    timeout 200 python tools/k1_quick.py --steps 10
  done; unset PB_LIB; } > gpurun_out/${tag}_gen_variants.log 2>&1
(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "philox or multi_tile" 2>&1 | tail -3) > gpurun_out/${tag}_tests.log 2>&1
PB_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_fused.json 2> gpurun_out/${tag}_bench_fused.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --two-calls > gpurun_out/${tag}_bench_twocalls.json 2> gpurun_out/${tag}_bench_twocalls.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --config C3 --sites 1000000 --steps 3 --warmup 3 > gpurun_out/${tag}_bench_C3dry_n${n}.json 2> gpurun_out/${tag}_bench_C3dry_n${n}.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $n --replicates 1000000 --steps 3 --warmup 3 > gpurun_out/${tag}_bench_northstar_n${n}.json 2> gpurun_out/${tag}_bench_northstar_n${n}.err
timeout 900 python bench.py --gpus $n --inproc --config C3 --sites 1000000 --steps 3 --warmup 3 > gpurun_out/${tag}_bench_C3dry_inproc_n${n}.json 2> gpurun_out/${tag}_bench_C3dry_inproc_n${n}.err
cat gpurun_out/${tag}_gen_variants.log; tail -3 gpurun_out/${tag}_tests.log
python tools/brief.py gpurun_out/${tag}_bench*.json
for f in gpurun_out/${tag}_bench*.err; do echo "== $f"; grep -v "^\[poutine_b200 trace\] \(events\|assoc\):" $f | tail -6; done
