python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --op wand --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('WAND', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['counters'])"
python bench.py --op maxscore --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('MAXSCORE', d['ms_per_step'], d['roofline']['kernel_ms'])"
