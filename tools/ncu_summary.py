#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here with `ncu -i`): headline metrics + hot SASS regions -> JSON on stdout."""
import csv, io, json, subprocess, sys

KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'l1tex__t_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']


def page(rep, name):
    return subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main(rep):
    rows = list(csv.reader(io.StringIO(page(rep, 'raw'))))
    hdr, units, val = rows[0], rows[1], rows[2]
    out = {'report': rep, 'kernel': val[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else None, 'metrics': {}}
    for h, u, v in zip(hdr, units, val):
        if h in KEEP:
            out['metrics'][h] = v + (' ' + u if u else '')
    rows = list(csv.reader(io.StringIO(page(rep, 'source'))))
    hdr, data = rows[1], rows[2:]
    ia, isrc, isamp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
    tot = sum(int(r[ia]) for r in data) or 1
    tots = sum(int(r[isamp]) for r in data) or 1
    out['sass_instructions'] = len(data)
    out['hot_regions'] = []
    B = 100
    for g in range(0, len(data), B):
        grp = data[g:g + B]
        s = sum(int(r[ia]) for r in grp); sm = sum(int(r[isamp]) for r in grp)
        if s / tot < 0.03 and sm / tots < 0.03:
            continue
        ops = {}
        for r in grp:
            t = r[isrc].split()
            op = t[0] if not t[0].startswith('@') else t[1]
            ops[op] = ops.get(op, 0) + int(r[ia])
        top = sorted(ops.items(), key=lambda x: -x[1])[:6]
        out['hot_regions'].append({'sass_index': g, 'pct_instructions': round(100 * s / tot, 1), 'pct_samples': round(100 * sm / tots, 1),
                                   'top_opcodes': {k: v for k, v in top}})
    # stall reasons over all samples
    stall_cols = [c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
    stalls = {c: sum(int(r[hdr.index(c)] or 0) for r in data) for c in stall_cols}
    out['stall_samples'] = dict(sorted(stalls.items(), key=lambda x: -x[1])[:8])
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main(sys.argv[1])
