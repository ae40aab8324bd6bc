set -x
# round 2, second half (role-split render backward): bench line, ncu captures of the render kernels at batch 16 and 32, launch list
timeout 900 python bench.py > gpurun_out/r02n_bench_1gpu.json 2> gpurun_out/r02n_bench_1gpu.err
for B in 16 32; do
SC_PROFILE_BATCH=$B timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"render_tc_(fwd|bwd)_kernel<\(int\)0" --launch-skip 4 -c 2 -o gpurun_out/r02n_render_tc_b$B python scripts/profile_render.py > /dev/null 2>&1
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02n_launches_bench_eager.csv python bench.py --eager --steps 2 --warmup 3 --batch 16 --no-configs --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/r02n_*
tail -c 600 gpurun_out/r02n_bench_1gpu.err
