for v in ns8; do
MVMC_LIBRARY=$PWD/multiview_motion_capture_b200/lib/variants/libmvmc_$v.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --als-phases > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err
echo $v; tail -1 gpurun_out/bench_var_$v.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_var_$v.json').read().strip().splitlines()[-1])
print("$v", d['value'], d['roofline']['frac'], d['roofline']['stage_ms_per_step'])
PY
done
