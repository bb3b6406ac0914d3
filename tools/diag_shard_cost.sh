#!/bin/bash
# 1 GPU: the kernel time of the 10k-query batch of N = 1 next to the shards ranks 0, 3 and 5 of an 8-GPU weak-scaling run hold
# (each alone on the GPU) -- separates "the shards cost more" from "the machine scales worse" in DESIGN.md 6.
mkdir -p gpurun_out/diag
for e in "" 0/8 3/8 5/8; do
  DS2I_BENCH_EMULATE_SHARD=$e python bench.py --no-also --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('shard', '$e' or 'N=1', d['ms_per_step'], d['config']['queries_per_step'])" | tee -a gpurun_out/diag/shard_cost.txt
done
