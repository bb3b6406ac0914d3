DS2I_NVCC_EXTRA="-DUNION_PROFILE" python -m ds2i_b200.build --force
DS2I_GPU_UNION_REPRIME=1 python bench.py --op wand --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); c=d['roofline']['counters']; print('PROFILE-REPRIME', d['ms_per_step'], c)
print('docs_blocks/launch', c['docs_blocks'], 'p1', c['block_maxs_read']//10**9, 'p2', c['docs_scored']//10**9, 'drv0', c['aux']%10**9, 'skipped items', c['aux']//10**9)"
python -m ds2i_b200.build --force
