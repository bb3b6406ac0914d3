#!/usr/bin/env python3
"""Where does the batch time go?  Times each operator on subsets of the bench query log grouped by number of terms."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ds2i_b200 as d
base = "/tmp/ds2i_b200_data/S_10000000_1000000_20261017_q80000/S"
idx = d.Index(base + ".block_optpfor.idx", "block_optpfor"); wd = d.WandData(base + ".wand")
qs = d.read_queries(base + ".queries", 10000)
ops = sys.argv[1:] or ["ranked_and", "wand"]
groups = {"all": qs}
for n in (1, 2, 3, 4):
    groups["nt=%d" % n] = [q for q in qs if len(set(q)) == n]
groups["nt>=5"] = [q for q in qs if len(set(q)) >= 5]
for op in ops:
    for name, g in groups.items():
        b = d.QueryBatch(idx, wd, g)
        for _ in range(2): b.run(op, 10)
        ms = np.mean([b.run(op, 10) for _ in range(3)])
        st = b.stats()
        print(json.dumps({"op": op, "group": name, "queries": len(g), "ms": round(float(ms), 3), "docs_blocks": st["docs_blocks"], "freqs_blocks": st["freqs_blocks"], "scored": st["docs_scored"]}), flush=True)
        b.close()
