import os, sys, time, argparse
sys.path.insert(0, "/root/repo")
import bench, ds2i_b200 as d
a = argparse.Namespace(docs=10_000_000, terms=1_000_000, seed=20261017, queries=10000, queries_total=80000)
paths = bench.ensure_data(a)
index = d.Index(paths["index"], "block_optpfor", 0); wd = d.WandData(paths["wand"], 0)
qs = d.read_queries(paths["queries"], 10000)
hb = d.flatten_queries(qs)
for op in ("ranked_and", "wand"):
    for i in range(5):
        t0 = time.perf_counter(); c, s, ms = d.query_batch(index, wd, op, hb, 10); t1 = time.perf_counter()
        print(op, "wall %.2f ms kernel %.2f ms" % ((t1 - t0) * 1e3, ms), flush=True)
