set -x
python bench.py --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('BASE', d['ms_per_step'], d['roofline']['kernel_ms'])"
DS2I_GPU_SPECIALIZE=1 python bench.py --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('SPEC', d['ms_per_step'], d['roofline']['kernel_ms'])"
DS2I_GPU_SLOTS_OVERRIDE=16 python bench.py --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('SLOTS16', d['ms_per_step'], d['roofline']['kernel_ms'])"
DS2I_GPU_SLOTS_OVERRIDE=11 python bench.py --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('SLOTS11', d['ms_per_step'], d['roofline']['kernel_ms'])"
