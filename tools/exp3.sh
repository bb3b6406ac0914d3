python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 6 4 5 8; do
DS2I_GPU_AND_MIN_CTAS=$v python bench.py --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('CTAS$v', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['counters'])"
done
python bench.py --op and --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('AND', d['ms_per_step'], d['roofline']['kernel_ms'])"
python bench.py --op wand --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('WAND', d['ms_per_step'], d['roofline']['kernel_ms'])"
