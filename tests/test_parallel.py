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


# ---- document-partitioned shards (ds2i_b200/sharding.py): the host-side logic, without a GPU ----------------------
def test_shard_ranges_and_term_mapping():
    from ds2i_b200.sharding import map_queries, shard_ranges
    assert shard_ranges(10, 3) == [(0, 3), (3, 6), (6, 10)]
    lookup = {5: 0, 9: 1}                                   # this shard holds lists for the global terms 5 and 9
    qs = [[5, 9], [5, 7], [7], []]
    assert map_queries(lookup, qs, "ranked_and") == [[0, 1], [], [], []]      # a missing term empties a conjunction
    assert map_queries(lookup, qs, "wand") == [[0, 1], [0], [], []]           # ... and just drops out of a disjunction
    assert map_queries(lookup, qs, "or") == [[0, 1], [0], [], []]


class _FakeIndex:
    def __init__(self, n):
        self.n = n

    def num_docs(self):
        return self.n


class _FakeShard:
    """what exchange_global_stats needs of a Shard"""

    def __init__(self, terms, sizes, ndocs):
        import numpy as np
        self.terms, self.sizes, self.index, self.got = np.asarray(terms), np.asarray(sizes), _FakeIndex(ndocs), None

    def local_df(self, num_terms_global):
        import numpy as np
        df = np.zeros(num_terms_global, dtype=np.int64)
        df[self.terms] = self.sizes
        return df

    def set_global_stats(self, df, ndocs):
        self.got = (df.copy(), ndocs)


def _stats_worker(rank, world, port, out):
    import numpy as np
    from ds2i_b200.sharding import exchange_global_stats, gather_shard_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = _FakeShard([0, 2] if rank == 0 else [1, 2], [3, 4] if rank == 0 else [5, 6], 100 + rank)
    df, ndocs = exchange_global_stats([sh], 4, world)
    ok = ndocs == 201 and df.tolist() == [3, 5, 10, 0] and sh.got[1] == 201 and sh.got[0].tolist() == [3, 5, 10, 0]
    rows = gather_shard_rows(torch.full((1, 2, 3), float(rank)), world)       # [S_local=1, nq=2, k=3] per rank
    ok = ok and rows.shape == (world, 2, 3) and all(bool((rows[r] == r).all()) for r in range(world))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_global_stats_exchange_world_size_2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_stats_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


# ---- the fused gather and the shard balancer (ds2i_b200/parallel.py), world_size 2 on gloo ----------------------------------
class _SizesIndex:
    def __init__(self, sizes):
        import numpy as np
        self.sizes = np.asarray(sizes, dtype=np.uint64)

    def list_sizes(self, terms):
        return self.sizes[terms]


def test_balanced_shards_and_cost_models():
    import numpy as np
    from ds2i_b200.parallel import balanced_shards, query_costs
    idx = _SizesIndex([100, 128 * 40, 128 * 5000, 300, 7])
    qs = [[0, 2], [1, 2], [2], [3, 3, 4], []]
    post = query_costs(idx, qs, "postings")
    assert post.tolist() == [100 + 640000, 5120 + 640000, 640000, 300 + 300 + 7, 0]
    conj = query_costs(idx, qs, "conjunctive")
    # [0,2]: shortest list 1 block -> 2 + min(5000, 128); [1,2]: 40 blocks -> 80 + min(5000, 5120); [2]: 2 * 5000; [3,3,4]: distinct {3,4}: 1 block, 3 blocks
    assert conj.tolist() == [2 + 128, 80 + 5000, 10000, 2 + 3, 0]
    shards = balanced_shards(conj, 2)
    assert sorted(np.concatenate(shards).tolist()) == list(range(5)) and abs(len(shards[0]) - len(shards[1])) <= 1
    assert shards[0][0] == 2 and shards[1][0] == 1            # the two heaviest queries go to different ranks


def _fused_worker(rank, world, port, out):
    import numpy as np
    from ds2i_b200.parallel import fused_row_bytes, gather_fused, split_fused
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nq, k = 4, 3
    counts = np.arange(nq, dtype=np.uint64) + 10 * rank
    scores = (np.arange(nq * k, dtype=np.float32) + 100 * rank).reshape(nq, k)
    docids = (np.arange(nq * k, dtype=np.uint32) + 1000 * rank).reshape(nq, k)
    fused = torch.from_numpy(np.concatenate([counts.view(np.uint8), scores.reshape(-1).view(np.uint8), docids.reshape(-1).view(np.uint8)]))
    assert fused.numel() == fused_row_bytes(nq, k)
    allr = gather_fused(fused, world)                         # ONE collective for counts + scores + docids
    ok = allr.shape == (world, fused_row_bytes(nq, k))
    for r in range(world):
        c, s, dd = split_fused(allr[r].numpy(), nq, k)
        ok = ok and c.tolist() == (np.arange(nq) + 10 * r).tolist() and float(s[1, 2]) == 5 + 100 * r and int(dd[3, 0]) == 9 + 1000 * r
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_fused_gather_world_size_2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_fused_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]
