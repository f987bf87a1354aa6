L=$PWD/multiview_motion_capture_b200/lib
for v in notma; do
MVMC_LIBRARY=$L/variants/libmvmc_$v.so timeout 600 python bench.py --clips 296 --steps 1 --warmup 3 --preroll 2 --no-cpu-baseline --no-e2e --als-phases > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err
echo $v; tail -1 gpurun_out/bench_var_$v.err
done
MVMC_LIBRARY=$L/libmvmc.so timeout 600 python bench.py --clips 296 --steps 1 --warmup 3 --preroll 2 --no-cpu-baseline --no-e2e --als-phases > gpurun_out/bench_var_base296.json 2> gpurun_out/bench_var_base296.err
echo base; tail -1 gpurun_out/bench_var_base296.err
