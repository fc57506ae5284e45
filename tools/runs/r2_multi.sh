#!/bin/bash
# Round 2 multi-GPU check (gpurun --gpus N): sharded == single parity (pb_group in one process, process-per-GPU over peer
# memory and over NCCL), the native host with --devices, weak-scaling bench lines in both launch shapes.
tag=${1:-r2m}; n=${2:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
(timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_host_cpp.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${tag}_tests.log 2>&1
{ echo "--- pregen early + low priority"; PB_PREGEN_EARLY=1 PB_GEN_LOW_PRIORITY=1 timeout 200 python tools/k1_quick.py --steps 10
  echo "--- default"; timeout 200 python tools/k1_quick.py --steps 10
  echo "--- C4"; timeout 200 python tools/k1_quick.py --steps 3 --config C4; } > gpurun_out/${tag}_quick.log 2>&1
(timeout 600 python -m pytest tests/test_gpu_shapes.py -m gpu -x -q -k "sweep_under or fused" 2>&1 | tail -5) >> gpurun_out/${tag}_tests.log 2>&1
timeout 100 tools/microbench/l2_peak > gpurun_out/${tag}_l2_peak.json 2>&1
timeout 200 tools/microbench/h2d_peak 1024 > gpurun_out/${tag}_h2d_peak.json 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n${n}.json 2> gpurun_out/${tag}_bench_n${n}.err
PB_BENCH_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/${tag}_bench_n${n}_nccl.json 2> gpurun_out/${tag}_bench_n${n}_nccl.err
timeout 600 python bench.py --gpus $n --inproc --steps 5 --warmup 3 > gpurun_out/${tag}_bench_inproc_n${n}.json 2> gpurun_out/${tag}_bench_inproc_n${n}.err
tail -25 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_quick.log; cat gpurun_out/${tag}_l2_peak.json gpurun_out/${tag}_h2d_peak.json
python tools/brief.py gpurun_out/${tag}_bench*n${n}*.json; for f in gpurun_out/${tag}_bench*.err; do tail -3 $f; done
