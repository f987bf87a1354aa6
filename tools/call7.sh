timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s3_v23.json 2> gpurun_out/bench_s3_v23.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s3_v23.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['stage_ms_per_step'])
PY
