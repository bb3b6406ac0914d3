"""Multi-GPU plumbing: one process per GPU, the index replicated in every GPU's HBM, the query
batch sharded across ranks (queries are independent — no collective on the data path); NCCL is
used only to gather the per-shard top-k.

A shard's results are ONE fused device buffer ([counts u64][scores f32 x k][docids u32 x k],
ds2i_gpu_batch_device_fused), so the gather is a single all_gather_into_tensor, enqueued behind the
query kernels on the same stream without a host synchronisation (QueryBatch.run(..., wait=False))."""
import numpy as np


def shard_queries(all_queries, rank, world, per_rank=None):
    """Contiguous shard of the query list for `rank`.  With per_rank given every rank gets exactly
    that many queries (weak scaling: the file holds world*per_rank queries); otherwise the list is
    split as evenly as possible (strong scaling)."""
    if per_rank is not None:
        return all_queries[rank * per_rank:(rank + 1) * per_rank]
    n = len(all_queries)
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return all_queries[lo:hi]


def balanced_shards(costs, world):
    """Cost-balanced shards: queries sorted by decreasing cost (ties keep their order) and dealt round-robin, so shard
    sizes differ by at most one and every shard gets the same mix of heavy and light queries.  Returns one index array
    per rank; np.concatenate of them is a permutation of range(len(costs)).  The C ABI's multi-GPU entry point
    (ds2i_gpu_group_query_batch) cuts its batch the same way."""
    order = np.argsort(-np.asarray(costs, dtype=np.int64), kind="stable")
    return [order[r::world] for r in range(world)]


def query_costs(index, queries, model="postings"):
    """Cost of each query for the shard balancer.
    model "postings": postings in the lists of the query (what the library's own scheduler orders work by);
    model "conjunctive": block decodes of a conjunctive evaluation — two per block of the shortest list (docs + freqs) plus, for
    every other list, the blocks it can be probed in: min(its blocks, 128 candidates x blocks of the shortest list).  Dealing the
    queries by this cost spreads the heavy conjunctions (a mid-sized list against a huge one) evenly over the ranks."""
    flat = np.fromiter((t for q in queries for t in q), dtype=np.uint32)
    sizes = index.list_sizes(flat).astype(np.int64) if len(flat) else np.zeros(0, np.int64)
    bounds = np.cumsum([0] + [len(q) for q in queries])
    if model == "postings":
        csum = np.concatenate([[0], np.cumsum(sizes)])
        return csum[bounds[1:]] - csum[bounds[:-1]]
    blocks = (sizes + 127) // 128
    out = np.zeros(len(queries), dtype=np.int64)
    for i in range(len(queries)):
        b = np.unique(flat[bounds[i]:bounds[i + 1]], return_index=True)[1]          # distinct terms
        nb = np.sort(blocks[bounds[i]:bounds[i + 1]][b])
        if len(nb):
            out[i] = 2 * nb[0] + int(np.minimum(nb[1:], 128 * nb[0]).sum())
    return out


def gather_topk(counts, scores, world):
    """all_gather of the per-shard results (device tensors) -> ([world*nq], [world*nq, k]) on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return counts, scores
    all_counts = torch.empty((world * counts.shape[0],), dtype=counts.dtype, device=counts.device)
    all_scores = torch.empty((world * scores.shape[0], scores.shape[1]), dtype=scores.dtype, device=scores.device)
    dist.all_gather_into_tensor(all_counts, counts.contiguous())
    dist.all_gather_into_tensor(all_scores, scores.contiguous())
    return all_counts, all_scores


def fused_row_bytes(nq, k):
    return nq * (8 + 8 * k)


def gather_fused(fused, world, out=None):
    """ONE all_gather_into_tensor of the fused per-shard result buffers (uint8 tensors of equal length; a shard with fewer
    queries pads).  Returns a [world, bytes] tensor on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return fused.reshape(1, -1)
    if out is None:
        out = torch.empty((world * fused.shape[0],), dtype=torch.uint8, device=fused.device)
    dist.all_gather_into_tensor(out.view(-1), fused)          # a flat output: gloo (the CPU tests) insists on it
    return out.view(world, fused.shape[0])


def split_fused(buf, nq, k):
    """Views of one shard's fused buffer (a 1-D uint8 tensor or numpy array): counts [nq] i64/u64, scores [nq,k] f32, docids [nq,k] i32/u32."""
    c_end, s_end = nq * 8, nq * 8 + nq * k * 4
    if isinstance(buf, np.ndarray):
        return (buf[:c_end].view(np.uint64), buf[c_end:s_end].view(np.float32).reshape(nq, k), buf[s_end:s_end + nq * k * 4].view(np.uint32).reshape(nq, k))
    import torch
    return (buf[:c_end].view(torch.int64), buf[c_end:s_end].view(torch.float32).reshape(nq, k), buf[s_end:s_end + nq * k * 4].view(torch.int32).reshape(nq, k))
