"""Multi-GPU plumbing: one process per GPU, the index replicated in every GPU's HBM, the query
batch sharded across ranks (queries are independent — no collective on the data path); NCCL is
used only to gather the per-shard top-k."""


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
