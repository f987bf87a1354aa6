set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stages.py -x -q -m gpu -k "als or assign or affinity" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/memcheck.log
timeout 900 python bench.py > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err; tail -2 gpurun_out/bench_r01e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r01e.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['cpu_baseline'], d['roofline']['frac'], d['roofline']['kernel_dram'], d['roofline']['stage_ms_per_step'], d['gpu_launches'], d['clocks'])
PY
