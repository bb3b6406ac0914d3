#!/usr/bin/env python3
"""Document-partitioned deployment over the GPUs of one box, one process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tools/sharded_demo.py

Rank 0 generates a synthetic collection (1/10 of the benchmark's scale by default), cuts it into G shards with
`ds2i_build shard` and builds the unsharded index as the checker; every rank loads its shard, the ranks exchange
document frequencies (one NCCL all_reduce), every rank evaluates the whole query batch on its shard, the per-shard
results are all-gathered over NVLink and merged on the device.  Rank 0 compares with the unsharded index on its own GPU
and prints one JSON line."""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=1_000_000)
    ap.add_argument("--terms", type=int, default=100_000)
    ap.add_argument("--queries", type=int, default=2000)
    ap.add_argument("--k", type=int, default=10)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import ds2i_b200 as d
    from ds2i_b200 import build, sharding
    base = os.path.join(os.environ.get("DS2I_BENCH_DATA", "/tmp/ds2i_b200_data"), "SH_%d_%d_%d" % (args.docs, args.terms, world))
    coll, out = os.path.join(base, "C"), os.path.join(base, "sh")
    if rank == 0 and not os.path.exists(os.path.join(base, "DONE")):
        build.build()
        os.makedirs(base, exist_ok=True)
        subprocess.run([build.BUILDER, "gen", coll, str(args.docs), str(args.terms), "20261017", "0.35", str(args.queries)], check=True)
        subprocess.run([build.BUILDER, "shard", "block_optpfor", coll, out, str(world)], check=True)
        subprocess.run([build.BUILDER, "index", "block_optpfor", coll, coll + ".idx"], check=True)
        subprocess.run([build.BUILDER, "wand", coll, coll + ".wand"], check=True)
        open(os.path.join(base, "DONE"), "w").write("ok")
    if world > 1:
        dist.barrier()
    queries = d.read_queries(coll + ".queries", args.queries)
    lo = sharding.shard_ranges(args.docs, world)[rank][0]
    shard = sharding.Shard(out, "block_optpfor", rank, lo, local)
    sharding.exchange_global_stats([shard], args.terms, world)
    line = {"layout": "documents partitioned over %d GPU(s), every shard evaluates the whole batch, NCCL all_gather + device merge" % world,
            "num_docs": args.docs, "num_terms": args.terms, "queries": len(queries), "k": args.k, "ops": {}}
    whole = None
    if rank == 0:
        idx = d.Index(coll + ".idx", "block_optpfor", local); wd = d.WandData(coll + ".wand", local)
        whole = d.QueryBatch(idx, wd, queries)
    for op in ("and", "ranked_and", "wand", "maxscore"):
        sharding.query_sharded([shard], op, queries, args.k, world)          # warm-up
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        counts, scores, docids = sharding.query_sharded([shard], op, queries, args.k, world)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            whole.run(op, args.k)
            ec, es = whole.fetch()
            ok_counts = bool(np.array_equal(counts, ec))
            rel = float(np.max(np.abs(scores.astype(np.float64) - es) / np.maximum(np.abs(es), 1e-30))) if op in d.RANKED else 0.0
            bit = bool(np.array_equal(scores.view(np.uint32), es.view(np.uint32))) if op in d.RANKED else None
            line["ops"][op] = {"queries_per_s": len(queries) / dt, "ms": dt * 1e3, "counts_equal_unsharded": ok_counts,
                               "scores_bit_exact": bit, "scores_max_rel_err": rel}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
