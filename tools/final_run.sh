timeout 900 python bench.py > gpurun_out/bench_r01i.json 2> gpurun_out/bench_r01i.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01i_ref.json 2> gpurun_out/bench_r01i_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01i_launches.csv python bench.py --steps 2 --warmup 3 --clips 592 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_g.log 2>&1
python - <<'PY'
import json
for f in ['gpurun_out/bench_r01i.json','gpurun_out/bench_r01i_ref.json']:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d['value'], (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'), d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('stage_ms_per_step'), d.get('gpu_launches'), d.get('clocks'))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_als$ --launch-skip 7 --launch-count 1 -o gpurun_out/prof_als_r01i -f python bench.py --steps 2 --warmup 3 --clips 296 --groups 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_als_r01i.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --als-phases > /dev/null 2> gpurun_out/r01i_k_als_phases.txt
