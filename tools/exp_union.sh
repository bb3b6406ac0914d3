python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for op in wand maxscore; do
python bench.py --op $op --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$op', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['roofline']['counters'])"
done
