"""Multi-process plumbing on CPU (gloo, world_size 2): query sharding and the top-k gather that the
GPU ranks run over NCCL (ds2i_b200/parallel.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ds2i_b200.parallel import gather_topk, shard_queries


def test_shard_queries_partitions_the_batch():
    qs = [[i] for i in range(10)]
    assert shard_queries(qs, 0, 2) + shard_queries(qs, 1, 2) == qs
    parts = [shard_queries(qs, r, 3) for r in range(3)]
    assert sum(parts, []) == qs and max(map(len, parts)) - min(map(len, parts)) <= 1
    assert shard_queries(qs, 1, 2, per_rank=4) == qs[4:8]            # weak scaling: fixed work per rank


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nq, k = 5, 3
    counts = torch.arange(nq, dtype=torch.int64) + 100 * rank
    scores = torch.arange(nq * k, dtype=torch.float32).reshape(nq, k) + 1000 * rank
    all_counts, all_scores = gather_topk(counts, scores, world)
    ok = all_counts.shape == (world * nq,) and all_scores.shape == (world * nq, k)
    for r in range(world):
        ok = ok and torch.equal(all_counts[r * nq:(r + 1) * nq], torch.arange(nq, dtype=torch.int64) + 100 * r)
        ok = ok and torch.equal(all_scores[r * nq:(r + 1) * nq], torch.arange(nq * k, dtype=torch.float32).reshape(nq, k) + 1000 * r)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_gather_topk_world_size_2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]
