set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err
tail -c 600 gpurun_out/bench_r01c.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01c_ref.json 2> gpurun_out/bench_r01c_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01c_launches.csv python bench.py --steps 2 --warmup 3 --clips 592 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_als$ --launch-skip 7 --launch-count 1 -o gpurun_out/prof_als_r01c -f python bench.py --steps 2 --warmup 3 --clips 296 --no-e2e --no-cpu-baseline > gpurun_out/ncu_als_r01c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ik_solve --launch-skip 15 --launch-count 1 -o gpurun_out/prof_ik_r01c -f python bench.py --steps 2 --warmup 3 --clips 296 --no-e2e --no-cpu-baseline > gpurun_out/ncu_ik_r01c.log 2>&1
ls -la gpurun_out | tail -12
