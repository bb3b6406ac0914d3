run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 4 --no-also --steps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().splitlines()[-1]); print('$2', d['ms_per_step'], d['kernel_ms_per_rank'], d['clocks'])"; }
run 29531 default
DS2I_BENCH_NO_CLOCKS=1 run 29532 noclocks
DS2I_BENCH_SYNC_GATHER=1 run 29533 syncgather
DS2I_BENCH_NO_CLOCKS=1 DS2I_BENCH_SYNC_GATHER=1 run 29534 noclocks_syncgather
