#!/bin/bash
# spec v3 label generator: Philox-mode parity tests, then short bench lines
tag=${1:-r2h}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -12) > gpurun_out/${tag}_tests.log 2>&1
tail -12 gpurun_out/${tag}_tests.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --config C3 --sites 625000 --scaling weak --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_C3shard.json 2> gpurun_out/${tag}_bench_C3shard.err
timeout 300 python bench.py --config C4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_C4.json 2> gpurun_out/${tag}_bench_C4.err
python tools/brief.py gpurun_out/${tag}_bench*.json
for f in gpurun_out/${tag}_bench*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', 'gen_kernel', d['breakdown_ms'].get('ms_gen_kernel'), 'exact_null', json.dumps({k:v for k,v in (d.get('exact_null_check') or {}).items() if k!='what'}), 'parity', d.get('parity_checked'))"; done
for f in gpurun_out/${tag}_bench*.err; do echo "== $f"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $f | tail -4; done
