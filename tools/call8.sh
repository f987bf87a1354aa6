L=$PWD/multiview_motion_capture_b200/lib
MVMC_LIBRARY=$L/variants/libmvmc_passonly.so timeout 600 python bench.py --clips 296 --steps 1 --warmup 3 --preroll 2 --no-cpu-baseline --no-e2e --als-phases > gpurun_out/bench_var_passonly.json 2> gpurun_out/bench_var_passonly.err
tail -1 gpurun_out/bench_var_passonly.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_var_passonly.json').read().strip().splitlines()[-1])
print(d['roofline']['stage_ms_per_step'], d['config']['als_iters_per_clip_frame'])
PY
MVMC_LIBRARY=$L/variants/libmvmc_passonly.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_als$ --launch-skip 4 --launch-count 1 -o gpurun_out/prof_als_passonly -f python bench.py --clips 296 --steps 1 --warmup 3 --preroll 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_als_passonly.log 2>&1
