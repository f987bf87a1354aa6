timeout 600 python bench.py --views 5 --people 4 --clips 9472 --distinct 148 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_shape_5x4.json 2> gpurun_out/bench_shape_5x4.err; tail -2 gpurun_out/bench_shape_5x4.err
timeout 600 python bench.py --views 8 --people 16 --clips 2368 --distinct 148 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_shape_8x16.json 2> gpurun_out/bench_shape_8x16.err; tail -2 gpurun_out/bench_shape_8x16.err
python - <<'PY'
import json
for f in ['gpurun_out/bench_shape_5x4.json','gpurun_out/bench_shape_8x16.json']:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['value'], d['e2e']['value'], d['roofline']['stage_ms_per_step'], d['config']['als_iters_per_clip_frame'], d['e2e'].get('alive_tracks_per_clip'), d['e2e'].get('capacity_errors'))
    except Exception as e: print(f, 'failed', e)
PY
