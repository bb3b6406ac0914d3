mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:union_drive_kernel -s 3 -c 1 -o gpurun_out/union_prof_r1e -f \
    python bench.py --op wand --no-also --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu_union2.log 2>&1
tail -2 gpurun_out/ncu_union2.log | cut -c1-300
