# Final evidence at HEAD of round 2 (one B200): full GPU test log, bench, reference arm, launch list, ncu of the kernels
# that changed after tools/profile_run_r02.sh was run (k_ik_solve, k_triangulate), kernel studies.
set -x
T=r02z
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${T}_gputests.txt 2>&1; tail -2 gpurun_out/${T}_gputests.txt
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --groups 1 --clips 1184 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ik_solve --launch-skip 7 --launch-count 1 -o gpurun_out/prof_ik_${T} -f python bench.py --steps 2 --warmup 3 --groups 1 --clips 1184 --no-e2e --no-cpu-baseline > gpurun_out/ncu_ik.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_triangulate --launch-skip 3 --launch-count 1 -o gpurun_out/prof_tri_${T} -f python bench.py --config kernels --frames 300 > gpurun_out/ncu_tri.log 2>&1
timeout 900 python bench.py --config kernels > gpurun_out/${T}_kernels.json 2> gpurun_out/${T}_kernels.err
