#!/bin/bash
# One GPU-box pass for a round's evidence (run as: gpurun --timeout 1500 -- 'bash tools/round_measure.sh TAG [all]'):
# GPU tests, the bench line, the reference arm, the ncu launch list and --set full captures of the event sweep and the K3 kernels.
tag=${1:-vX}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/${tag}_tests.log 2>&1
timeout 300 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k1x_events -c 1 -o gpurun_out/${tag}_k1x -f python tools/k1_quick.py --steps 1 > gpurun_out/${tag}_ncu2.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k3_gen_kernel -c 1 -o gpurun_out/${tag}_k3gen -f python tools/k1_quick.py --steps 1 > gpurun_out/${tag}_ncu3.log 2>&1
timeout 120 python tools/bench_native_host.py 100000 16384 > gpurun_out/${tag}_native_host.json 2>> gpurun_out/${tag}_bench.err   # files-in/files-out, native host
PB_TRACE=1 timeout 200 python tools/k1_quick.py --steps 2 > gpurun_out/${tag}_trace_C2.log 2>&1
if [ -n "$2" ]; then
  timeout 120 python bench.py --config C1 --no-cpu-baseline > gpurun_out/${tag}_bench_C1.json 2>> gpurun_out/${tag}_bench.err
  timeout 200 python bench.py --config C3 --sites 625000 --scaling weak --no-cpu-baseline > gpurun_out/${tag}_bench_C3shard.json 2>> gpurun_out/${tag}_bench.err
  timeout 200 python bench.py --config C4 --no-cpu-baseline > gpurun_out/${tag}_bench_C4.json 2>> gpurun_out/${tag}_bench.err
  timeout 200 python bench.py --config C5 --no-cpu-baseline > gpurun_out/${tag}_bench_C5.json 2>> gpurun_out/${tag}_bench.err
fi
tail -8 gpurun_out/${tag}_tests.log
python tools/brief.py gpurun_out/${tag}_bench*.json 2>/dev/null | tail -8
tail -2 gpurun_out/${tag}_trace_C2.log; tail -5 gpurun_out/${tag}_bench.err
