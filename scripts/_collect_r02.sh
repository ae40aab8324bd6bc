set -x
# 1. CLIP tower: per-phase launch list (both modes) and one full capture of the cooperative launch (fp16, batch 64)
for m in fp16 split; do
ncu --kernel-name regex:clip_tower --launch-skip 65 --launch-count 65 --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r02_tower_phases_$m.csv python scripts/profile_clip_tower.py 64 $m > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on --kernel-name regex:clip_tower --launch-skip 3 --launch-count 1 -o gpurun_out/r02_tower_coop_fp16 python scripts/profile_clip_tower_coop.py 64 fp16 > /dev/null 2>&1
ncu --set full --clock-control none --kernel-name regex:clip_tower --launch-skip 3 --launch-count 1 -o gpurun_out/r02_tower_coop_split python scripts/profile_clip_tower_coop.py 64 split > /dev/null 2>&1
# 2. render kernels at configs[1] shape (default split mode) and the bf16 mode
ncu --set full --import-source on --clock-control none -k regex:render_tc_\(fwd\|bwd\)_kernel --launch-skip 4 -c 2 -o gpurun_out/r02_render_tc python scripts/profile_render.py > /dev/null 2>&1
SC_RENDER_FORWARD=bf16 SC_RENDER_BACKWARD=bf16 ncu --set full --clock-control none -k regex:render_tc_\(fwd\|bwd\)_kernel --launch-skip 4 -c 2 -o gpurun_out/r02_render_tc_bf16 python scripts/profile_render.py > /dev/null 2>&1
# 3. launch list of the step (eager: ncu cannot attach to kernels launched during capture)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_eager.csv python bench.py --eager --steps 2 --warmup 3 --batch 16 --no-configs --no-cpu-baseline > /dev/null 2>&1
# 4. marching cubes + chamfer kernels of one evaluate.py shape
ls -la gpurun_out/r02_*
