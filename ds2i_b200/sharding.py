"""Document-partitioned deployment: every GPU holds the index of a contiguous range of documents
(an ordinary ds2i index with local docids, built by `ds2i_build shard`), every shard evaluates the whole query batch,
and the per-shard top-k lists are gathered (NCCL all_gather over NVLink) and merged on the device.  This is the layout
for a collection that does not fit one GPU's HBM; bench.py's default (index replicated, queries sharded) is the one for
a collection that does.

BM25 uses collection-wide statistics — bm25::query_term_weight(qtf, df, N) (bm25.hpp:17-24), norm_len = len / average
length (wand_data.hpp:24-33) — so the shard builder normalises with the global average length and the shards exchange
their document frequencies once at load time (one all_reduce); scores then equal the unsharded reference's."""
import ctypes as C
import os

import numpy as np

from . import _native
from .api import Index, QueryBatch, WandData, RANKED

CONJUNCTIVE = ("and", "and_freq", "ranked_and")


class Shard:
    """Shard `g` of `<prefix>.<g>.{idx,wand,terms}` resident on `device`."""

    def __init__(self, prefix, index_type, g, doc_lo, device=0):
        self.g = g
        self.doc_lo = int(doc_lo)                   # global docid of local document 0
        self.index = Index("%s.%d.idx" % (prefix, g), index_type, device)
        self.wand = WandData("%s.%d.wand" % (prefix, g), device)
        self.terms = np.fromfile("%s.%d.terms" % (prefix, g), dtype=np.uint32)     # global term id of every local list
        if len(self.terms) != self.index.size():
            raise ValueError("term map and index disagree for shard %d" % g)

    def local_df(self, num_terms_global):
        """Dense vector over the global term ids: this shard's posting count of every term."""
        df = np.zeros(num_terms_global, dtype=np.int64)
        df[self.terms] = self.index.list_sizes(np.arange(self.index.size(), dtype=np.uint32)).astype(np.int64)
        return df

    def set_global_stats(self, df_global, num_docs_total):
        self.index.set_global_stats(np.asarray(df_global)[self.terms].astype(np.uint64), num_docs_total)
        self._lookup = None

    def map_queries(self, queries, op):
        if getattr(self, "_lookup", None) is None:
            self._lookup = {int(t): i for i, t in enumerate(self.terms)}
        return map_queries(self._lookup, queries, op)

    def run(self, op, queries, k):
        """-> (counts [nq] i64, scores [nq,k] f32, docids [nq,k] i64, global) as torch CUDA tensors owned by the caller."""
        import torch
        batch = QueryBatch(self.index, self.wand if op in RANKED else None, self.map_queries(queries, op))
        batch.run(op, k)
        if op in RANKED:
            c, s, d = batch.device_results(k, with_docids=True)
            d = d.to(torch.int64) & 0xFFFFFFFF
            d = torch.where(d == 0xFFFFFFFF, d, d + self.doc_lo)
            res = (c.clone(), s.clone(), d)
        else:
            c, _ = batch.device_results(1)
            res = (c.clone(), torch.zeros((len(queries), k), dtype=torch.float32, device=c.device),
                   torch.full((len(queries), k), 0xFFFFFFFF, dtype=torch.int64, device=c.device))
        batch.close()
        return res


def map_queries(lookup, queries, op):
    """Global term ids -> list numbers of a shard (`lookup`: global id -> local list).  A term without postings in the
    shard has no list (ds2i lists cannot be empty): a conjunctive query then matches nothing there, a disjunctive one
    just loses the term."""
    out = []
    for q in queries:
        local = [lookup.get(int(t), -1) for t in q]
        if op in CONJUNCTIVE and any(t < 0 for t in local):
            out.append([])
        else:
            out.append([t for t in local if t >= 0])
    return out


def shard_ranges(num_docs_total, num_shards):
    """[lo, hi) of every shard — the same split `ds2i_build shard` makes."""
    return [(num_docs_total * g // num_shards, num_docs_total * (g + 1) // num_shards) for g in range(num_shards)]


def exchange_global_stats(shards, num_terms_global, group_world=1):
    """Sum the per-shard document frequencies over the shards of this process and, under torch.distributed, over all
    ranks (one all_reduce at load time); hand every shard the collection-wide statistics."""
    df = np.zeros(num_terms_global, dtype=np.int64)
    ndocs = 0
    for sh in shards:
        df += sh.local_df(num_terms_global)
        ndocs += sh.index.num_docs()
    if group_world > 1:
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.from_numpy(np.concatenate([df, [ndocs]])).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t = t.cpu().numpy()
        df, ndocs = t[:-1], int(t[-1])
    for sh in shards:
        sh.set_global_stats(df, ndocs)
    return df, ndocs


def merge_shard_results(counts, scores, docids, k, ranked):
    """counts [S, nq] i64, scores [S, nq, k] f32, docids [S, nq, k] i64 (CUDA tensors, rows = shards) -> merged
    (counts [nq], scores [nq, k], docids [nq, k]) through ds2i_gpu_merge_shards."""
    import torch
    S, nq = counts.shape
    dev = counts.device
    c_in = counts.contiguous()
    s_in = scores.contiguous()
    d_in = docids.to(torch.int32).contiguous()          # the C ABI carries docids as u32 bit patterns
    oc = torch.zeros((nq,), dtype=torch.int64, device=dev)
    os_ = torch.zeros((nq, k), dtype=torch.float32, device=dev)
    od = torch.full((nq, k), -1, dtype=torch.int32, device=dev)
    torch.cuda.current_stream().synchronize()
    _native.check(_native.lib().ds2i_gpu_merge_shards(C.c_void_p(c_in.data_ptr()), C.c_void_p(s_in.data_ptr()), C.c_void_p(d_in.data_ptr()),
                                                      S, nq, k, 1 if ranked else 0, C.c_void_p(oc.data_ptr()), C.c_void_p(os_.data_ptr()),
                                                      C.c_void_p(od.data_ptr())))
    return oc, os_, od.to(torch.int64) & 0xFFFFFFFF


def gather_shard_rows(t, world):
    """[S_local, ...] on every rank -> [world * S_local, ...], rank-major: the row layout ds2i_gpu_merge_shards reads."""
    if world == 1:
        return t
    import torch
    import torch.distributed as dist
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1))
    return out.view((world * t.shape[0],) + tuple(t.shape[1:]))


def query_sharded(shards, op, queries, k=10, world=1):
    """Evaluate `queries` on the shards of this process, gather the other ranks' results when world > 1 (every rank
    ends up with the merged answer), merge on the device.  -> numpy (counts, scores, docids)."""
    import torch
    res = [sh.run(op, queries, k) for sh in shards]
    counts = torch.stack([r[0] for r in res])
    scores = torch.stack([r[1] for r in res])
    docids = torch.stack([r[2] for r in res])
    counts, scores, docids = gather_shard_rows(counts, world), gather_shard_rows(scores, world), gather_shard_rows(docids, world)
    oc, os_, od = merge_shard_results(counts, scores, docids, k, op in RANKED)
    return oc.cpu().numpy().astype(np.uint64), os_.cpu().numpy(), od.cpu().numpy().astype(np.uint32)
