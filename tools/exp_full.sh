python -m pytest tests -m gpu -x -q 2>&1 | tail -3
DS2I_GPU_TRACE=1 python bench.py 2> gpurun_out/bench.err > gpurun_out/bench.json
grep "ds2i_gpu\]" gpurun_out/bench.err | tail -4
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').readline())
print('ranked_and', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'frac', d['roofline']['frac'], d.get('parity'))
w=d['also']['wand']; print('wand', w['value'], w['ms_per_step'], 'e2e', w['e2e']['value'], 'cpu', w['cpu_baseline']['value'])
PY
