#!/usr/bin/env python3
"""Join ncu per-SASS counters (.ncu-rep, --page source) with nvdisasm -g line info of the built cubin ->
executed-instruction and PC-sample shares per CUDA source line.  usage: hot_lines.py <rep> <mangled kernel substring>"""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
td = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 'ds2i_b200', 'lib', 'libds2i_gpu.so')], cwd=td, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(td) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(td, cubin)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout.splitlines()
rows = list(csv.reader(io.StringIO(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout)))
hdr, data = rows[1], rows[2:]
ia, isamp = hdr.index('Instructions Executed'), hdr.index('# Samples')
# every .text section whose name holds the pattern (template instances share a prefix); the one with as many SASS
# instructions as the report is the captured kernel
sections, cur, name = {}, None, None
for line in dis:
    if line.startswith('\t.section\t.text.') or line.startswith('.section'):
        name = line.split()[1].split(',')[0] if pat in line and '.text.' in line else None
        if name:
            sections[name] = []
        continue
    if name is None:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);', line)
    if m:
        sections[name].append(cur)
exact = [k for k, v in sections.items() if len(v) == len(data)]
pick = exact[0] if exact else (next(iter(sections)) if sections else None)
seq = sections.get(pick, [])
print('kernel section %s (%d candidates for the pattern)' % (pick, len(sections)))
if len(seq) != len(data):
    print('WARNING: %d SASS instructions in the cubin vs %d in the report (rebuild mismatch?)' % (len(seq), len(data)))
n = min(len(seq), len(data))
agg, samp = collections.Counter(), collections.Counter()
for i in range(n):
    agg[seq[i]] += int(data[i][ia]); samp[seq[i]] += int(data[i][isamp])
tot = sum(int(r[ia]) for r in data) or 1; tots = sum(int(r[isamp]) for r in data) or 1
print('total %.2f G warp instructions, %d samples' % (tot / 1e9, tots))
src_cache = {}
order = samp.most_common(top) if os.environ.get('BY_SAMPLES') else agg.most_common(top)
for k, _ in order:
    v = agg[k]
    text = ''
    if k:
        p = os.path.join(ROOT, 'ds2i_b200', 'csrc', k[0])
        if os.path.exists(p):
            src_cache.setdefault(p, open(p).read().splitlines())
            if k[1] - 1 < len(src_cache[p]):
                text = src_cache[p][k[1] - 1].strip()[:90]
    print('%5.1f%% inst %5.1f%% samp  %s:%s  %s' % (100 * v / tot, 100 * samp[k] / tots, k[0] if k else None, k[1] if k else None, text))
