L=$PWD/multiview_motion_capture_b200/lib
MVMC_LIBRARY=$L/variants/libmvmc_c3.so timeout 300 python -m pytest tests/test_gpu_stages.py -x -q -m gpu -k als 2>&1 | tail -1
for c in 444 1332; do
MVMC_LIBRARY=$L/variants/libmvmc_c3.so timeout 600 python bench.py --clips $c --groups 1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --als-phases > gpurun_out/bench_c3_$c.json 2> gpurun_out/bench_c3_$c.err
tail -1 gpurun_out/bench_c3_$c.err | cut -c1-330
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c3_$c.json').read().strip().splitlines()[-1])
print($c, d['value'], d['roofline']['frac'], d['roofline']['stage_ms_per_step'])
PY
done
