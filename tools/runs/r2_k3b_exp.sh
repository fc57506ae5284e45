#!/bin/bash
# K3b 16-byte-load variants + tail / generator scheduling knobs; parity first.
tag=${1:-r2c}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${tag}_tests.log 2>&1
{
for v in default u2 u8 b3 b5; do
  if [ $v = default ]; then unset PB_LIB; else export PB_LIB=exp/libpb_$v.so; fi
  timeout 200 python tools/k1_quick.py --steps 10
  timeout 200 python tools/k1_quick.py --steps 3 --config C4
  timeout 200 python tools/k1_quick.py --steps 10 --config C1
done
unset PB_LIB
echo "--- headstart 0"; PB_GEN_HEADSTART_US=0 timeout 200 python tools/k1_quick.py --steps 10
echo "--- headstart 3"; PB_GEN_HEADSTART_US=3 timeout 200 python tools/k1_quick.py --steps 10
echo "--- gen low priority"; PB_GEN_LOW_PRIORITY=1 timeout 200 python tools/k1_quick.py --steps 10
echo "--- no pregen"; PB_NO_PREGEN=1 timeout 200 python tools/k1_quick.py --steps 10
echo "--- C4 gen low priority"; PB_GEN_LOW_PRIORITY=1 timeout 200 python tools/k1_quick.py --steps 3 --config C4
} > gpurun_out/${tag}_k3b.log 2>&1
PB_TRACE=1 timeout 300 python tools/k1_quick.py --steps 2 > gpurun_out/${tag}_trace_C2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k3_count_tight -c 1 -o gpurun_out/${tag}_k3tight -f python tools/k1_quick.py --steps 1 > gpurun_out/${tag}_ncu_tight.log 2>&1
tail -3 gpurun_out/${tag}_tests.log; cat gpurun_out/${tag}_k3b.log; tail -2 gpurun_out/${tag}_trace_C2.log
