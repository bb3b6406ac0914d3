"""bench.py legs over the `opt` (partitioned Elias-Fano) index of the benchmark collection: BASELINE config 3 (next / next_geq
microbenchmark against the HBM roofline, the reference's enumerator timed beside it) and the query operators on `opt`."""
import json
import os
import struct

import numpy as np


def legs(d, args, paths, queries, qidx, wdata, peak, measure, line_of, cpu_baseline_for, parity_block, ref_tool):
    out = {}
    idx = d.Index(paths["opt"], "opt", 0)
    cores = os.cpu_count() or 1
    ddir = os.path.dirname(paths["opt"])
    # ---- config 3 (i): next(): full sequential scan of the 4096 longest lists (term ids are ranks: the first 4096) ----
    NL = 4096
    terms = np.arange(NL, dtype=np.uint32)
    nbytes = int(idx.list_bytes(terms).sum())
    for _ in range(args.warmup):
        idx.decode_lists_checksum(terms)
    ms = []
    for _ in range(args.steps):
        postings, sd, sf, m = idx.decode_lists_checksum(terms)
        ms.append(m)
    m = sum(ms) / len(ms)
    ref_all = json.loads(ref_tool("scan", "opt", paths["opt"], "first:%d" % NL, cores, 1).strip().splitlines()[-1])
    ref_one = json.loads(ref_tool("scan", "opt", paths["opt"], "first:256", 1, 1).strip().splitlines()[-1])
    ach = (nbytes + 8 * postings) / (m * 1e-3) / 1e9
    out["pef_next"] = {
        "metric": "decoded ints/sec (opt index, next()/docid()/freq() over the %d longest lists)" % NL, "value": 2 * postings / (m * 1e-3), "unit": "ints/s",
        "ms_per_step": m, "postings": postings, "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "bytes_in": nbytes, "bytes_out": 8 * postings,
                     "algorithmic_bytes_per_launch": nbytes + 8 * postings,
                     "note": "in = bits of the scanned lists in the docs and freqs bit vectors / 8 (every partition payload + its header), out = 4 B docid + 4 B freq per posting"},
        "cpu_baseline": {"value": 2 * ref_all["postings_per_s"], "unit": "ints/s", "cores": cores, "kind": "reference",
                         "sample": "the reference's enumerator over the same %d lists, list i -> thread i %% n" % NL,
                         "single_thread": {"value": 2 * ref_one["postings_per_s"], "unit": "ints/s", "cores": 1, "sample": "the 256 longest lists"}},
        "parity": {"postings_equal": postings == ref_all["postings"], "sum_docids_equal": sd == ref_all["sum_docids"], "sum_freqs_equal": sf == ref_all["sum_freqs"],
                   "ok": bool(postings == ref_all["postings"] and sd == ref_all["sum_docids"] and sf == ref_all["sum_freqs"])}}
    # ---- config 3 (ii): next_geq sweeps, lower bounds = value of every 2^j-th element + 1 ----
    GL = 1024
    gterms = np.arange(GL, dtype=np.uint32)
    offs, docs, _, _ = idx.decode_lists(gterms)
    sweeps = {}
    for j in (0, 3, 6, 9, 12):
        # every list's bound sequence is cut into runs of CHUNK calls, each run driven through its own enumerator (opened at
        # the list start): one warp per list alone would leave most of the 148 SMs idle
        total_calls = sum((int(offs[i + 1]) - int(offs[i]) + (1 << j) - 1) >> j for i in range(GL))
        CHUNK = max(8, min(512, total_calls // 16384))       # >= ~16k independent cursors whenever the sweep has that many calls
        bounds, which = [], []
        for i in range(GL):
            bb = docs[int(offs[i]):int(offs[i + 1])][::1 << j].astype(np.uint64) + 1
            for c0 in range(0, len(bb), CHUNK):
                bounds.append(bb[c0:c0 + CHUNK]); which.append(i)
        calls = sum(len(b) for b in bounds)
        which = np.asarray(which, dtype=np.uint32)
        t = []
        for it in range(2 + max(2, args.steps - 2)):
            gd, gf, m = idx.next_geq_batch(which, bounds)
            t.append(m)
        m = float(np.mean(t[2:]))
        spec = os.path.join(ddir, "geq_spec_%d.bin" % j)
        with open(spec, "wb") as f:
            f.write(struct.pack("<Q", len(which)))
            for tt, b in zip(which, bounds):
                f.write(struct.pack("<QQ", int(tt), len(b)))
                f.write(np.asarray(b, dtype="<u8").tobytes())
        ref = json.loads(ref_tool("geqbench", "opt", paths["opt"], spec, cores, 1).strip().splitlines()[-1])
        dev_sum = int(gd.sum() + gf.sum())
        sweeps["skip_%d" % (1 << j)] = {"calls": calls, "enumerators": len(bounds), "calls_per_enumerator": CHUNK, "ms": m, "calls_per_s": calls / (m * 1e-3),
                                        "postings_skipped_per_s": calls * (1 << j) / (m * 1e-3),
                                        "cpu_calls_per_s": ref["calls_per_s"], "cpu_cores": cores,
                                        "parity_checksum_equal": dev_sum == ref["checksum"]}
        os.remove(spec)
    out["pef_next_geq"] = {"metric": "next_geq calls/sec (opt index, %d longest lists, lower bound = every 2^j-th docid + 1)" % GL, "unit": "calls/s",
                           "value": sweeps["skip_1"]["calls_per_s"], "sweeps": sweeps,
                           "parity": {"ok": all(s["parity_checksum_equal"] for s in sweeps.values()),
                                      "what": "sum of docid() + freq() after every call, device vs the reference enumerator over the same bound sequences"}}
    # ---- the query operators over `opt` ----
    for op in ("ranked_and", "wand"):
        mm = measure(op, queries, [len(queries)], idx=idx)
        leg = line_of(mm, op, len(queries), "weak", itype="opt")
        leg["roofline"]["note"] = "the Elias-Fano kernels do not count decoded bytes: achieved / frac are not meaningful here, kernel_ms is"
        leg["cpu_baseline"] = cpu_baseline_for(paths, op, len(queries), itype="opt", single_prefix=300)
        leg["parity"] = parity_block(d, paths, op, args.k, qidx, mm["counts"], mm["scores"], itype="opt")
        out["opt_" + op] = leg
    idx.close()
    return out
