L=$PWD/multiview_motion_capture_b200/lib
run() {
MVMC_LIBRARY=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --als-phases > gpurun_out/bench_var_$1.json 2> gpurun_out/bench_var_$1.err
tail -1 gpurun_out/bench_var_$1.err | cut -c200-330
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_var_$1.json').read().strip().splitlines()[-1])
print("$1", d['value'], d['roofline']['frac'], d['roofline']['stage_ms_per_step'])
PY
}
run a $L/variants/libmvmc_a.so
run b $L/variants/libmvmc_b.so
run c $L/variants/libmvmc_c.so
