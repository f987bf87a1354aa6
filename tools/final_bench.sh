timeout 900 python bench.py > gpurun_out/bench_r01j.json 2> gpurun_out/bench_r01j.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01j.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['stage_ms_per_step'], d['gpu_launches'], d['clocks'])
PY
