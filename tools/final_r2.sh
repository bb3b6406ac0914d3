#!/bin/bash
# round-2 final measurements.  tools/final_r2.sh            -> 1 GPU: tests, bench lines, ncu launch list, one ncu --set full capture per dominant kernel
#                              tools/final_r2.sh N (2|4|8)  -> the bench line under torch.distributed.run on N GPUs
mkdir -p gpurun_out/final2
O=gpurun_out/final2
if [ -n "$1" ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $1 > $O/bench_n$1.json 2> $O/bench_n$1.err
  echo "N=$1 rc=$?"; tail -c 300 $O/bench_n$1.err
  exit 0
fi
python -m pytest tests -q -m gpu 2>&1 | tail -3 > $O/tests.log; cat $O/tests.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > /dev/null 2>&1
K="python tools/kbench.py --child --no-stats --steps 1"
ncu --set full --clock-control none --import-source on -k regex:and_block_kernel -s 3 -c 1 -o $O/and_prof -f $K --ops ranked_and > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:union_drive_kernel -s 3 -c 1 -o $O/union_prof -f $K --ops wand > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:and_block_kernel -s 3 -c 1 -o $O/pef_and_prof -f $K --itype opt --ops ranked_and > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_full_blocks_kernel -s 2 -c 1 -o $O/decode_full_prof -f python tools/microbench.py decode --steps 1 --warmup 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_serial_blocks_kernel -s 4 -c 1 -o $O/decode_serial_prof -f python tools/microbench.py decode --steps 1 --warmup 2 > /dev/null 2>&1
cp ds2i_b200/lib/libds2i_gpu.so $O/libds2i_gpu.so.profiled
ls -la $O
