mkdir -p gpurun_out
(timeout 250 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/v14_tests.log 2>&1
timeout 200 python bench.py > gpurun_out/v14_bench.json 2> gpurun_out/v14_bench.err
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/v14_bench_ref.json 2>> gpurun_out/v14_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v14_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/v14_ncu1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k1x_events -c 1 -o gpurun_out/v14_k1x -f python tools/k1_quick.py --steps 1 > gpurun_out/v14_ncu2.log 2>&1
tail -3 gpurun_out/v14_tests.log; python -c "
import json
d=json.loads(open('gpurun_out/v14_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['breakdown_ms'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['three_planes']['ms_per_step'], d['three_planes']['roofline']['frac'], d['gpu_launches'], d['clocks'])
print(d['breakdown_ms'])
"; tail -2 gpurun_out/v14_ncu2.log; ls -la gpurun_out/v14*
