set -x
timeout 900 python bench.py > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01f_ref.json 2> gpurun_out/bench_r01f_ref.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --als-phases > /dev/null 2> gpurun_out/r01f_k_als_phases.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01f_launches.csv python bench.py --steps 2 --warmup 3 --clips 592 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches_d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_als$ --launch-skip 7 --launch-count 1 -o gpurun_out/prof_als_r01f -f python bench.py --steps 2 --warmup 3 --clips 296 --no-e2e --no-cpu-baseline > gpurun_out/ncu_als_r01f.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ik_solve --launch-skip 14 --launch-count 1 -o gpurun_out/prof_ik_r01f -f python bench.py --steps 2 --warmup 3 --clips 296 --no-e2e --no-cpu-baseline > gpurun_out/ncu_ik_r01f.log 2>&1
