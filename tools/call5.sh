set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_als$ --launch-skip 7 --launch-count 1 -o gpurun_out/prof_als_s3d -f python bench.py --steps 2 --warmup 3 --clips 296 --no-e2e --no-cpu-baseline > gpurun_out/ncu_als_d.log 2>&1
tail -3 gpurun_out/ncu_als_d.log | cut -c1-300
