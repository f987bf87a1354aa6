set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s3_v2.json 2> gpurun_out/bench_s3_v2.err
tail -c 2500 gpurun_out/bench_s3_v2.json; tail -5 gpurun_out/bench_s3_v2.err
