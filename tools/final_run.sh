timeout 900 python bench.py > gpurun_out/bench_r01g.json 2> gpurun_out/bench_r01g.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01g_ref.json 2> gpurun_out/bench_r01g_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01g_launches.csv python bench.py --steps 2 --warmup 3 --clips 592 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_g.log 2>&1
python - <<'PY'
import json
for f in ['gpurun_out/bench_r01g.json','gpurun_out/bench_r01g_ref.json']:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'), d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('stage_ms_per_step'), d.get('gpu_launches'), d.get('clocks'))
PY
