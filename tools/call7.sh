set -x

timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --als-phases > gpurun_out/bench_s3_v21.json 2> gpurun_out/bench_s3_v21.err
tail -2 gpurun_out/bench_s3_v21.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s3_v21.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['roofline']['stage_ms_per_step'])
PY
