#!/usr/bin/env python3
"""Kernel-experiment harness: time the resident-batch query kernels of several builds of libds2i_gpu.so on the
benchmark index and check every build's results against the first one (which the parity tests pin to the reference).

  python tools/kbench.py [--ops ranked_and,and,wand,maxscore] [--steps 5] lib_a.so lib_b.so ...

Each library runs in its own process (DS2I_GPU_LIB); one JSON line per (library, op) on stdout."""
import argparse
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(args):
    import numpy as np
    import bench
    import ds2i_b200 as d
    a = argparse.Namespace(docs=args.docs, terms=args.terms, seed=20261017, queries=args.queries, queries_total=80000)
    paths = bench.ensure_data(a, 0, (args.itype,))
    index = d.Index(paths[args.itype], args.itype, 0)
    wdata = d.WandData(paths["wand"], 0)
    queries = d.read_queries(paths["queries"], args.queries)
    batch = d.QueryBatch(index, wdata, queries)
    for op in args.ops.split(","):
        for _ in range(3):
            batch.run(op, 10, stats=not args.no_stats)
        ms = [batch.run(op, 10, stats=not args.no_stats) for _ in range(args.steps)]
        batch.run(op, 10)
        st = batch.stats()
        counts, scores = batch.fetch()
        h = hashlib.sha1(counts.tobytes() + (scores.tobytes() if op in d.RANKED else b"")).hexdigest()[:16]
        print(json.dumps({"lib": os.path.basename(os.environ.get("DS2I_GPU_LIB", "default")), "op": op, "ms_min": min(ms), "ms_mean": sum(ms) / len(ms),
                          "qps": len(queries) / (min(ms) * 1e-3), "hash": h, "docs_blocks": st["docs_blocks"], "freqs_blocks": st["freqs_blocks"],
                          "scored": st["docs_scored"]}), flush=True)
        np.save("/tmp/kbench_%s_%s.npy" % (os.path.basename(os.environ.get("DS2I_GPU_LIB", "default")), op), scores)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="ranked_and,and,wand,maxscore")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--terms", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--itype", default="block_optpfor")
    ap.add_argument("--no-stats", action="store_true", help="time the kernel instances without the work counters (DS2I_RUN_NO_STATS)")
    ap.add_argument("libs", nargs="*")
    args = ap.parse_args()
    if args.child:
        return child(args)
    base = {}
    for lib in args.libs or [None]:
        env = dict(os.environ)
        if lib:
            env["DS2I_GPU_LIB"] = os.path.abspath(lib)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--ops", args.ops, "--steps", str(args.steps), "--docs", str(args.docs),
                            "--terms", str(args.terms), "--queries", str(args.queries), "--itype", args.itype] + (["--no-stats"] if args.no_stats else []), env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        sys.stderr.write(r.stderr[-4000:])
        if r.returncode != 0:
            print(json.dumps({"lib": lib, "failed": r.stderr[-600:]}), flush=True)
            continue
        for line in r.stdout.splitlines():
            if not line.startswith("{"):
                continue
            rec = json.loads(line)
            key = rec["op"]
            if key not in base:
                base[key] = (rec["hash"], rec["lib"])
            # wand / maxscore: the last score bit depends on the pruning history, so only and / ranked_and hashes must agree
            rec["same_as_first"] = rec["hash"] == base[key][0]
            if not rec["same_as_first"]:
                import numpy as np
                a = np.load("/tmp/kbench_%s_%s.npy" % (base[key][1], key)).astype(np.float64)
                b = np.load("/tmp/kbench_%s_%s.npy" % (rec["lib"], key)).astype(np.float64)
                rec["max_rel_diff_vs_first"] = float(np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-30))) if a.shape == b.shape else None
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
