python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for p in 2048 4096 8192; do
DS2I_GPU_UNION_ITEM_POSTINGS=$p python bench.py --op wand --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('ITEM $p', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['roofline']['counters'])"
done
DS2I_GPU_TRACE=1 python bench.py --op wand --no-also --no-cpu-baseline --steps 3 2>&1 | grep "ds2i_gpu\]" | tail -4
