for v in nomath; do
MVMC_LIBRARY=$PWD/multiview_motion_capture_b200/lib/variants/libmvmc_$v.so timeout 600 python bench.py --clips 296 --steps 1 --warmup 3 --preroll 2 --no-cpu-baseline --no-e2e --als-phases > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err
echo $v; tail -1 gpurun_out/bench_var_$v.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_als$ --launch-skip 7 --launch-count 1 -o gpurun_out/prof_als_s3e -f python bench.py --steps 2 --warmup 3 --clips 296 --no-e2e --no-cpu-baseline > gpurun_out/ncu_als_e.log 2>&1
