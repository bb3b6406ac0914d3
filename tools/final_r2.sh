#!/bin/bash
# round-2 final measurements.  tools/final_r2.sh            -> 1 GPU: tests, bench lines, ncu launch list, one ncu --set full capture per dominant kernel
#                              tools/final_r2.sh N (2|4|8)  -> the bench line under torch.distributed.run on N GPUs
#                              tools/final_r2.sh captures   -> 1 GPU: only the ncu captures of the three query kernels + a short bench line
# The ncu reports are summarised on the box (tools/ncu_summary.py, tools/hot_lines.py); only the summaries travel back (gpurun_out <= 64 MiB).
mkdir -p gpurun_out/final2
O=gpurun_out/final2
if [ -n "$1" ] && [ "$1" != captures ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $1 > $O/bench_n$1.json 2> $O/bench_n$1.err
  echo "N=$1 rc=$?"; tail -c 300 $O/bench_n$1.err
  exit 0
fi
if [ "$1" != captures ]; then
python -m pytest tests -q -m gpu 2>&1 | tail -3 > $O/tests.log; cat $O/tests.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also > /dev/null 2>&1
else
python bench.py --no-also --no-cpu-baseline > $O/bench_n1_short.json 2> $O/bench_n1_short.err; echo "short bench rc=$?"
fi
K="python tools/kbench.py --child --no-stats --steps 1"
capture() {   # name, kernel regex, mangled-name substring for hot_lines, launch skip, command...
  local name=$1 regex=$2 sub=$3 skip=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o /tmp/$name -f "$@" > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep > $O/${name}_ncu.json 2>/dev/null
  python tools/hot_lines.py /tmp/$name.ncu-rep $sub 40 > $O/${name}_hot_lines.txt 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin))
h,v=rows[0],rows[2]
keep={k:x for k,x in zip(h,v) if k in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','smsp__inst_executed.sum','lts__t_bytes.sum','l1tex__t_bytes.sum')}
print(json.dumps(keep))" > $O/${name}_traffic.json
  rm -f /tmp/$name.ncu-rep
}
capture and_block_kernel and_block_kernel and_block_kernelILi0ELb1ELi7ELb0 3 $K --ops ranked_and
capture union_drive_kernel union_drive_kernel union_drive_kernelILi0ELi7ELb0ELi0 3 $K --ops wand
capture pef_and_block_kernel and_block_kernel and_block_kernelILi5ELb1ELi4ELb0 3 $K --itype opt --ops ranked_and
[ "$1" = captures ] && { ls -la $O; exit 0; }
capture decode_full_blocks_kernel decode_full_blocks_kernel decode_full_blocks_kernelILi0 2 python tools/microbench.py decode --steps 1 --warmup 2
capture decode_serial_blocks_kernel decode_serial_blocks_kernel decode_serial_blocks_kernel 4 python tools/microbench.py decode --steps 1 --warmup 2
ls -la $O
