set -x
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
./tools/micro/dmma_probe > gpurun_out/dmma_probe2.log 2>&1
cat gpurun_out/dmma_probe2.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s3_base.json 2> gpurun_out/bench_s3_base.err
tail -c 3000 gpurun_out/bench_s3_base.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_als --launch-skip 7 --launch-count 1 -o gpurun_out/prof_als_s3a -f python bench.py --steps 2 --warmup 3 --clips 296 --no-e2e --no-cpu-baseline > gpurun_out/ncu_als.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ik_solve --launch-skip 15 --launch-count 1 -o gpurun_out/prof_ik_s3a -f python bench.py --steps 2 --warmup 3 --clips 296 --no-e2e --no-cpu-baseline > gpurun_out/ncu_ik.log 2>&1
ls -la gpurun_out
