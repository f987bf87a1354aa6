timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_s3_v30.json 2> gpurun_out/bench_s3_v30.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s3_v30.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['roofline']['stage_ms_per_step'])
PY
