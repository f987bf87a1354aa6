# Round-2 evidence run (one B200): bench at HEAD, reference arm, launch list, ncu --set full of the top kernels, memcheck.
set -x
T=${1:-r02}
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --als-phases > /dev/null 2> gpurun_out/${T}_k_als_phases.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --groups 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_als$ --launch-skip 3 --launch-count 1 -o gpurun_out/prof_als_${T} -f python bench.py --steps 2 --warmup 3 --groups 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_als.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ik_solve --launch-skip 7 --launch-count 1 -o gpurun_out/prof_ik_${T} -f python bench.py --steps 2 --warmup 3 --groups 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_ik.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_affinity --launch-skip 3 --launch-count 1 -o gpurun_out/prof_aff_${T} -f python bench.py --steps 2 --warmup 3 --groups 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_aff.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_triangulate --launch-skip 3 --launch-count 1 -o gpurun_out/prof_tri_${T} -f python bench.py --config kernels --frames 300 > gpurun_out/ncu_tri.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_als$ --launch-skip 3 --launch-count 1 -o gpurun_out/prof_als_small_${T} -f python bench.py --views 8 --people 16 --clips 2368 --steps 2 --warmup 3 --groups 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_als_small.log 2>&1
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_stages.py tests/test_gpu_seams.py -m gpu -q -x -k "alternative or linear_sum or ik_3d or affinity_synthetic or body25 or torch_extension or assign_synthetic or solver_and_fk" > gpurun_out/${T}_memcheck.txt 2>&1
tail -5 gpurun_out/${T}_memcheck.txt
ls -la gpurun_out/*.ncu-rep
