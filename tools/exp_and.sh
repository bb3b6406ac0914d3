timeout 300 python -m pytest tests -m gpu -x -q --timeout 60 --timeout-method=thread 2>&1 | tail -3
for ch in 32 16 8; do
DS2I_GPU_AND_CHUNK_BLOCKS=$ch python bench.py --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('AND chunk $ch', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])"
done
python bench.py --op and --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('and', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])"
