python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('BASE', d['ms_per_step'], d['roofline']['kernel_ms'])"
DS2I_GPU_SPECIALIZE=1 python bench.py --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('SPEC', d['ms_per_step'], d['roofline']['kernel_ms'])"
python bench.py --op wand --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('WAND', d['ms_per_step'], d['roofline']['kernel_ms'])"
python tools/microbench.py decode 2>/dev/null | tail -3
