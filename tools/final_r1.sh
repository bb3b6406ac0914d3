# round-1 final measurements: bench lines, ncu launch lists, one ncu --set full capture per dominant kernel, microbenchmarks
mkdir -p gpurun_out/final
O=gpurun_out/final
python bench.py > $O/bench_ranked_and.json 2> $O/bench_ranked_and.err
python bench.py --op maxscore --no-also > $O/bench_maxscore.json 2>/dev/null
python bench.py --op and --no-also > $O/bench_and.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_ranked_and.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:and_block_kernel -s 3 -c 1 -o $O/and_prof -f python bench.py --no-also --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:union_drive_kernel -s 3 -c 1 -o $O/union_prof -f python bench.py --op wand --no-also --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
bash tools/prof_decode.sh
python tools/microbench.py decode 2>/dev/null > $O/micro_decode.jsonl
python tools/microbench.py pef 2>/dev/null > $O/micro_pef.jsonl
cp ds2i_b200/lib/libds2i_gpu.so $O/libds2i_gpu.so.profiled
ls -la $O
cut -c1-400 $O/bench_ranked_and.json
