# end-of-round sanity: the GPU parity suite, the decode microbenchmark, the default bench line
timeout 300 python -m pytest tests -m gpu -q -x --timeout 60 --timeout-method=thread 2>&1 | tail -3
python tools/microbench.py decode 2>/dev/null | cut -c1-200
python bench.py --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('ranked_and', d['ms_per_step'], d['e2e']['value'], 'wand', d['also']['wand']['ms_per_step'], d['also']['wand']['e2e']['value'])"
