set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_als$ --launch-skip 7 --launch-count 1 -o gpurun_out/prof_als_r01f -f python bench.py --steps 2 --warmup 3 --clips 296 --groups 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_als_r01f.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ik_solve --launch-skip 14 --launch-count 1 -o gpurun_out/prof_ik_r01f -f python bench.py --steps 2 --warmup 3 --clips 296 --groups 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_ik_r01f.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01f_launches_g1.csv python bench.py --steps 2 --warmup 3 --clips 592 --groups 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_f.log 2>&1
