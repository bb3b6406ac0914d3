# one ncu --set full capture of the conjunctive kernel (ranked_and, full-size batch) + the launch list of a bench run
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:and_block_kernel -s 3 -c 1 -o gpurun_out/and_prof_r1d -f \
    python bench.py --no-also --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/ncu_and.log 2>&1
tail -3 gpurun_out/ncu_and.log
