#!/bin/bash
# 4 GPUs: what the weak-scaling step costs beyond the slowest rank's kernel -- the in-line gather (default), the overlapped
# double-buffered gather (DS2I_BENCH_OVERLAP_GATHER=1), each with and without the NVML clock sampler thread.
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 4 --no-also --steps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().splitlines()[-1]); print('$2', d['ms_per_step'], d['kernel_ms_per_rank'], d['clocks'])"; }
run 29531 inline
DS2I_BENCH_NO_CLOCKS=1 run 29532 inline_noclocks
DS2I_BENCH_OVERLAP_GATHER=1 run 29533 overlapped
DS2I_BENCH_NO_CLOCKS=1 DS2I_BENCH_OVERLAP_GATHER=1 run 29534 overlapped_noclocks
