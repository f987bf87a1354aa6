# IK solver changes late in round 2 (six CTAs per SM, then the load/store regrouping): full GPU tests, the bench, one ncu capture of the update kernel.
set -x
T=r02v
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${T}_gputests.txt 2>&1; tail -2 gpurun_out/${T}_gputests.txt
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ik_solve --launch-skip 7 --launch-count 1 -o gpurun_out/prof_ik_${T} -f python bench.py --steps 2 --warmup 3 --groups 1 --clips 1184 --no-e2e --no-cpu-baseline > gpurun_out/ncu_ik.log 2>&1; tail -3 gpurun_out/ncu_ik.log
