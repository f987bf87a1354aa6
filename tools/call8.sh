for v in nostore noload; do
MVMC_LIBRARY=$PWD/multiview_motion_capture_b200/lib/variants/libmvmc_$v.so timeout 600 python bench.py --clips 296 --steps 1 --warmup 3 --preroll 2 --no-cpu-baseline --no-e2e --als-phases > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err
echo $v; tail -1 gpurun_out/bench_var_$v.err
done
