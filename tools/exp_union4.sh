python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for p in 8192 16384; do
DS2I_GPU_UNION_ITEM_POSTINGS=$p python bench.py --op wand --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('ITEM $p', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['roofline']['counters'])"
done
DS2I_NVCC_EXTRA="-DDS2I_UNION_MIN_CTAS=5" python -m ds2i_b200.build --force
for p in 8192 16384; do
DS2I_GPU_UNION_ITEM_POSTINGS=$p python bench.py --op wand --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('CTAS5 ITEM $p', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])"
done
python -m ds2i_b200.build --force
