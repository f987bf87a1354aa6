L=$PWD/multiview_motion_capture_b200/lib
for v in w4; do
MVMC_LIBRARY=$L/variants/libmvmc_$v.so timeout 300 python -m pytest tests/test_gpu_stages.py -x -q -m gpu -k als 2>&1 | tail -2
MVMC_LIBRARY=$L/variants/libmvmc_$v.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --als-phases > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err
tail -1 gpurun_out/bench_var_$v.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_var_$v.json').read().strip().splitlines()[-1])
print("$v", d['value'], d['roofline']['frac'], d['roofline']['stage_ms_per_step'])
PY
done
