mkdir -p gpurun_out/final
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/final/bench_ranked_and_${N}gpu.json 2> gpurun_out/final/bench_${N}gpu.err
tail -2 gpurun_out/final/bench_${N}gpu.err | cut -c1-300
cut -c1-300 gpurun_out/final/bench_ranked_and_${N}gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/sharded_demo.py > gpurun_out/final/sharded_${N}gpu.json 2> gpurun_out/final/sharded_${N}gpu.err
tail -3 gpurun_out/final/sharded_${N}gpu.err | cut -c1-300
cat gpurun_out/final/sharded_${N}gpu.json
