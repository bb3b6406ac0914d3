# experiment: where do the union kernel's block decodes come from (UNION_PROFILE build), then item-size sweep with the normal build
DS2I_NVCC_EXTRA="-DUNION_PROFILE" python -m ds2i_b200.build --force
python bench.py --op wand --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.readline()); c=d['roofline']['counters']; print('PROFILE', d['ms_per_step'], c)
print('docs_blocks/launch', c['docs_blocks'], 'p1', c['block_maxs_read']//10**9, 'p2', c['docs_scored']//10**9, 'drv0', c['aux']%10**9, 'skipped items', c['aux']//10**9)"
python -m ds2i_b200.build --force
for p in 4096 8192 32768 65536; do
DS2I_GPU_UNION_ITEM_POSTINGS=$p python bench.py --op wand --no-also --no-cpu-baseline --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('ITEM $p', d['ms_per_step'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['roofline']['counters'])"
done
