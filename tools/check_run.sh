timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --als-phases > gpurun_out/bench_s3_v26.json 2> gpurun_out/bench_s3_v26.err
tail -1 gpurun_out/bench_s3_v26.err | cut -c1-330
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s3_v26.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['roofline']['stage_ms_per_step'])
PY
