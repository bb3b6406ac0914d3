#!/usr/bin/env python3
"""Decode / enumerator microbenchmarks (BASELINE.json configs[1] and configs[2]); run on the GPU box:

  python tools/microbench.py decode   [--docs N --terms T]     batched block decode of EVERY list of the
        synthetic index (block_optpfor and block_interpolative, built by ds2i_build): decoded ints/s,
        compressed-in + decoded-out bytes/s against the measured HBM peak.
  python tools/microbench.py pef      full sequential decode (next) and next_geq sweeps (skip 2^j postings)
        over `opt` indexes built by the reference (oracle/_ref/data: S10 = 1M docs / 100k terms, T).
Writes one JSON line per measurement (stdout) — copied into profiles/ by the round notes.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ds2i_b200 as d          # noqa: E402
from ds2i_b200 import build    # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def synth(docs, terms, seed, types):
    build.build()
    base = os.environ.get("DS2I_BENCH_DATA", "/tmp/ds2i_b200_data")
    dname = os.path.join(base, "M_%d_%d_%d" % (docs, terms, seed))
    if not os.path.exists(os.path.join(dname, "DONE")):
        os.makedirs(dname, exist_ok=True)
        subprocess.run([build.BUILDER, "synth", os.path.join(dname, "S"), str(docs), str(terms), str(seed), "0", "1000", ":".join(types)], check=True)
        open(os.path.join(dname, "DONE"), "w").write("ok")
    return dname


def decode_bench(args):
    types = ["block_optpfor", "block_interpolative"]
    dname = synth(args.docs, args.terms, args.seed, types)
    for t in types:
        path = os.path.join(dname, "S.%s.idx" % t)
        idx = d.Index(path, t)
        terms = np.arange(idx.size(), dtype=np.uint32)
        times = []
        for _ in range(args.warmup + args.steps):
            n, ms = idx.decode_lists_device(terms)
            times.append(ms)
        ms = float(np.mean(times[args.warmup:]))
        comp = os.path.getsize(path)
        nblocks = int(np.sum((idx.list_sizes(terms) + 127) // 128))
        alg_in = comp + 8 * nblocks * 0      # block_max + endpoints are part of the file bytes
        alg_out = 8 * n
        print(json.dumps({"bench": "decode_all_lists", "index_type": t, "num_docs": args.docs, "num_terms": args.terms, "postings": n,
                          "blocks": nblocks, "kernel_ms": ms, "ints_per_s": 2 * n / (ms * 1e-3), "postings_per_s": n / (ms * 1e-3),
                          "compressed_bytes": comp, "read_GBps": alg_in / (ms * 1e-3) / 1e9, "read_plus_write_GBps": (alg_in + alg_out) / (ms * 1e-3) / 1e9,
                          "hbm_peak_GBps": PEAK, "frac_of_measured_peak": (alg_in + alg_out) / (ms * 1e-3) / 1e9 / PEAK,
                          "note": "docids + freqs materialised to HBM (8 B per posting written)"}), flush=True)
        idx.close()


def pef_bench(args):
    data = os.path.join(ROOT, "oracle", "_ref", "data")
    for name in ("S10", "T"):
        for t in ("opt", "block_optpfor"):
            path = os.path.join(data, "%s.%s.idx" % (name, t))
            if not os.path.exists(path):
                continue
            idx = d.Index(path, t)
            sizes = idx.list_sizes(np.arange(idx.size(), dtype=np.uint32))
            longest = np.argsort(sizes)[::-1][:4096].astype(np.uint32)
            times = []
            for _ in range(args.warmup + args.steps):
                n, ms = idx.decode_lists_device(longest)
                times.append(ms)
            ms = float(np.mean(times[args.warmup:]))
            print(json.dumps({"bench": "next_full_scan_4096_longest_lists", "collection": name, "index_type": t, "postings": n, "kernel_ms": ms,
                              "postings_per_s": n / (ms * 1e-3), "index_bytes": os.path.getsize(path),
                              "note": "index smaller than the 126 MB L2: this measures L2-resident decode"}), flush=True)
            # next_geq sweeps: lower bounds = docid of every 2^j-th posting + 1
            NL = min(4096, len(longest))
            offs, docs, _, _ = idx.decode_lists(longest[:NL])
            for j in (0, 3, 6, 9, 12):
                # every list's bound sequence is cut into runs of CHUNK calls, each run driven through its own
                # enumerator (opened at the list start): one warp per list alone would leave most of the 148 SMs idle
                CHUNK = 512
                bounds, which = [], []
                for i in range(NL):
                    dd = docs[int(offs[i]):int(offs[i + 1])]
                    bb = dd[::1 << j].astype(np.uint64) + 1
                    for c0 in range(0, len(bb), CHUNK):
                        bounds.append(bb[c0:c0 + CHUNK]); which.append(longest[i])
                calls = sum(len(b) for b in bounds)
                which = np.asarray(which, dtype=np.uint32)
                times = []
                for _ in range(args.warmup + args.steps):
                    _, _, ms = idx.next_geq_batch(which, bounds)
                    times.append(ms)
                ms = float(np.mean(times[args.warmup:]))
                print(json.dumps({"bench": "next_geq_sweep", "collection": name, "index_type": t, "skip_postings": 1 << j, "lists": NL, "enumerators": len(bounds), "calls": calls,
                                  "kernel_ms": ms, "calls_per_s": calls / (ms * 1e-3), "postings_skipped_per_s": calls * (1 << j) / (ms * 1e-3)}), flush=True)
            idx.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["decode", "pef"])
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--terms", type=int, default=1_000_000)
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    decode_bench(a) if a.what == "decode" else pef_bench(a)
