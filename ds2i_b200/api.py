"""Host-side mirror of ds2i's Index / wand_data / query-operator interface over the C ABI.

Names and argument meaning follow the reference (queries.hpp, block_freq_index.hpp, wand_data.hpp):
an operator object is called as ``op(index, terms)`` and returns what the reference operator
returns (match count for and/or, ``topk().size()`` for the ranked ones); ``op.topk()`` gives the
scores.  ``op.batch(index, queries)`` evaluates a whole list of queries in one device launch —
that is the product path; the single-query call is a batch of one.
"""
import ctypes as C

import numpy as np

from . import _native

OPS = ("and", "and_freq", "or", "or_freq", "ranked_and", "wand", "maxscore", "ranked_or")
RANKED = ("ranked_and", "wand", "maxscore", "ranked_or")


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _flatten(queries):
    offs = np.zeros(len(queries) + 1, dtype=np.uint64)
    if len(queries):
        offs[1:] = np.cumsum([len(q) for q in queries], dtype=np.uint64)
    flat = np.fromiter((t for q in queries for t in q), dtype=np.uint32, count=int(offs[-1])) if len(queries) else np.zeros(0, np.uint32)
    if flat.size == 0:
        flat = np.zeros(1, np.uint32)[:0]
    return np.ascontiguousarray(flat), offs


class Index:
    """An index file of type `index_type` (index_types.hpp:41) resident in HBM of `device`."""

    def __init__(self, path, index_type, device=0):
        self._h = C.c_void_p()
        self.index_type = index_type
        self.device = device
        L = _native.lib()
        _native.check(L.ds2i_gpu_index_open_file(str(path).encode(), index_type.encode(), device, C.byref(self._h)))

    @classmethod
    def from_bytes(cls, data, index_type, device=0):
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        self.index_type = index_type
        self.device = device
        buf = (C.c_char * len(data)).from_buffer_copy(data)
        _native.check(_native.lib().ds2i_gpu_index_open(C.cast(buf, C.c_void_p), len(data), index_type.encode(), device, C.byref(self._h)))
        return self

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _native.lib().ds2i_gpu_index_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        return int(_native.lib().ds2i_gpu_index_size(self._h))

    def num_docs(self):
        return int(_native.lib().ds2i_gpu_index_num_docs(self._h))

    def device_bytes(self):
        return int(_native.lib().ds2i_gpu_index_device_bytes(self._h))

    def set_global_stats(self, df, num_docs_total):
        """Collection-wide statistics for a document-partitioned shard: df[i] = documents of the WHOLE collection that
        hold the term of list i (None clears them).  See ds2i_gpu_index_set_global_stats."""
        if df is None:
            _native.check(_native.lib().ds2i_gpu_index_set_global_stats(self._h, None, 0, 0))
            return
        df = np.ascontiguousarray(df, dtype=np.uint64)
        _native.check(_native.lib().ds2i_gpu_index_set_global_stats(self._h, _p(df, C.c_uint64), len(df), int(num_docs_total)))

    def list_sizes(self, terms):
        terms = np.ascontiguousarray(terms, dtype=np.uint32)
        out = np.zeros(len(terms), dtype=np.uint64)
        _native.check(_native.lib().ds2i_gpu_index_list_sizes(self._h, _p(terms, C.c_uint32), len(terms), _p(out, C.c_uint64)))
        return out

    def list_bytes(self, terms):
        """Compressed bytes of each list in the index file (ds2i_gpu_index_list_bytes)."""
        terms = np.ascontiguousarray(terms, dtype=np.uint32)
        out = np.zeros(len(terms), dtype=np.uint64)
        _native.check(_native.lib().ds2i_gpu_index_list_bytes(self._h, _p(terms, C.c_uint32), len(terms), _p(out, C.c_uint64)))
        return out

    def decode_lists(self, terms):
        """docid()/freq() of every posting of index[term], as next() delivers them.
        Returns (offsets, docs, freqs, elapsed_ms)."""
        terms = np.ascontiguousarray(terms, dtype=np.uint32)
        sizes = self.list_sizes(terms)
        offs = np.zeros(len(terms) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum(sizes, dtype=np.uint64)
        total = int(offs[-1])
        docs = np.zeros(max(total, 1), dtype=np.uint32)
        freqs = np.zeros(max(total, 1), dtype=np.uint32)
        ms = C.c_float()
        _native.check(_native.lib().ds2i_gpu_decode_lists(self._h, _p(terms, C.c_uint32), len(terms), _p(offs, C.c_uint64),
                                                       _p(docs, C.c_uint32), _p(freqs, C.c_uint32), C.byref(ms)))
        return offs, docs[:total], freqs[:total], ms.value

    def decode_lists_device(self, terms):
        """Same decode, outputs left in HBM (no D2H): returns (postings, kernel_ms) — the batched block
        decode microbenchmark (BASELINE config 2)."""
        terms = np.ascontiguousarray(terms, dtype=np.uint32)
        sizes = self.list_sizes(terms)
        offs = np.zeros(len(terms) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum(sizes, dtype=np.uint64)
        ms = C.c_float()
        _native.check(_native.lib().ds2i_gpu_decode_lists(self._h, _p(terms, C.c_uint32), len(terms), _p(offs, C.c_uint64),
                                                       None, None, C.byref(ms)))
        return int(offs[-1]), ms.value

    def decode_lists_checksum(self, terms):
        """Full decode left in HBM and reduced there: (postings, sum of docids, sum of freqs, kernel_ms)."""
        terms = np.ascontiguousarray(terms, dtype=np.uint32)
        sizes = self.list_sizes(terms)
        offs = np.zeros(len(terms) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum(sizes, dtype=np.uint64)
        ms, sd, sf = C.c_float(), C.c_uint64(), C.c_uint64()
        _native.check(_native.lib().ds2i_gpu_decode_lists_checksum(self._h, _p(terms, C.c_uint32), len(terms), _p(offs, C.c_uint64),
                                                                C.byref(sd), C.byref(sf), C.byref(ms)))
        return int(offs[-1]), int(sd.value), int(sf.value), ms.value

    def next_geq_batch(self, terms, bounds_per_list):
        """index[term] opened, then next_geq(b) for each b (non-decreasing).  Returns (docids, freqs, ms)."""
        terms = np.ascontiguousarray(terms, dtype=np.uint32)
        flat, offs = _flatten64(bounds_per_list)
        n = int(offs[-1])
        docids = np.zeros(max(n, 1), dtype=np.uint64)
        freqs = np.zeros(max(n, 1), dtype=np.uint64)
        ms = C.c_float()
        _native.check(_native.lib().ds2i_gpu_next_geq_batch(self._h, _p(terms, C.c_uint32), len(terms), _p(flat, C.c_uint64),
                                                         _p(offs, C.c_uint64), _p(docids, C.c_uint64), _p(freqs, C.c_uint64), C.byref(ms)))
        return docids[:n], freqs[:n], ms.value


def _flatten64(lists):
    offs = np.zeros(len(lists) + 1, dtype=np.uint64)
    if len(lists):
        offs[1:] = np.cumsum([len(b) for b in lists], dtype=np.uint64)
    flat = np.concatenate([np.asarray(b, dtype=np.uint64) for b in lists]) if len(lists) and int(offs[-1]) else np.zeros(1, np.uint64)
    return np.ascontiguousarray(flat), offs


class WandData:
    """wand_data<bm25> (wand_data.hpp) resident in HBM."""

    def __init__(self, path, device=0):
        self._h = C.c_void_p()
        _native.check(_native.lib().ds2i_gpu_wand_open_file(str(path).encode(), device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _native.lib().ds2i_gpu_wand_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class QueryBatch:
    """A batch of queries resident in HBM (ds2i_gpu_batch_*): prepare once, run several operators."""

    def __init__(self, index, wdata, queries):
        self._h = C.c_void_p()
        self.nq = len(queries)
        self._keep = (index, wdata)
        self._device = getattr(index, "device", 0)
        flat, offs = _flatten(queries)
        wh = wdata._h if wdata is not None else C.c_void_p()
        _native.check(_native.lib().ds2i_gpu_batch_prepare(index._h, wh, _p(flat, C.c_uint32), _p(offs, C.c_uint64), self.nq, C.byref(self._h)))
        self.h2d_bytes = flat.nbytes + offs.nbytes

    def run(self, op, k=10, faithful=False, wait=True, stats=True):
        """Evaluate `op` over the resident batch; returns the CUDA-event time in ms.  faithful=True
        selects the literal one-candidate-at-a-time kernels (DS2I_RUN_FAITHFUL).  wait=False launches
        without a host synchronisation (DS2I_RUN_ASYNC): the work is ordered on the device's default
        stream, so a collective enqueued afterwards runs behind it; wait() returns the kernel time."""
        ms = C.c_float()
        self._k = k
        self._op = op
        flags = (1 if faithful else 0) | (0 if wait else 2) | (0 if stats else 4)        # stats=False: DS2I_RUN_NO_STATS
        _native.check(_native.lib().ds2i_gpu_batch_run_ex(self._h, OPS.index(op), k, flags, C.byref(ms)))
        return ms.value

    def wait(self):
        ms = C.c_float()
        _native.check(_native.lib().ds2i_gpu_batch_wait(self._h, C.byref(ms)))
        return ms.value

    def device_fused(self, pad_to=None):
        """The results of the last run as ONE torch uint8 CUDA tensor aliasing the library's buffer:
        [counts nq x u64][scores nq*k x f32][docids nq*k x u32] (unranked operators: counts only).  pad_to: expose
        that many bytes (the buffer is allocated for the largest k, so a shard can pad to the size of its peers)."""
        import torch
        p, n = C.c_void_p(), C.c_size_t()
        _native.check(_native.lib().ds2i_gpu_batch_device_fused(self._h, C.byref(p), C.byref(n)))
        nbytes = int(n.value) if pad_to is None else int(pad_to)
        if nbytes > max(self.nq, 1) * (8 + 8 * 32):
            raise ValueError("pad_to beyond the fused buffer")

        class _Cai:
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": "|u1", "data": (ptr, False), "version": 3}

        return torch.as_tensor(_Cai(p.value, (nbytes,)), device=torch.device("cuda", self._device))

    def fetch(self):
        counts = np.zeros(max(self.nq, 1), dtype=np.uint64)
        scores = np.zeros((max(self.nq, 1), self._k), dtype=np.float32)
        _native.check(_native.lib().ds2i_gpu_batch_fetch(self._h, _p(counts, C.c_uint64), _p(scores, C.c_float)))
        return counts[:self.nq], scores[:self.nq]

    def fetch_docids(self):
        """docids of the scores of the last ranked run (nq x k, 0xffffffff padding) — an extension: the reference keeps scores only."""
        ids = np.full((max(self.nq, 1), self._k), 0xFFFFFFFF, dtype=np.uint32)
        _native.check(_native.lib().ds2i_gpu_batch_fetch_docids(self._h, _p(ids, C.c_uint32)))
        return ids[:self.nq]

    def device_results(self, k=None, with_docids=False):
        """(counts, scores) of the last run as torch CUDA tensors that alias the library's device buffers."""
        import torch
        k = self._k if k is None else k
        pc, ps = C.c_void_p(), C.c_void_p()
        _native.check(_native.lib().ds2i_gpu_batch_device_results(self._h, C.byref(pc), C.byref(ps)))

        class _Cai:
            def __init__(self, ptr, shape, typestr):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}

        dev = torch.device("cuda", self._device)          # the index's device, whatever torch's current device is
        counts = torch.as_tensor(_Cai(pc.value, (self.nq,), "<i8"), device=dev)
        scores = torch.as_tensor(_Cai(ps.value, (self.nq, k), "<f4"), device=dev)
        if with_docids:
            pd = C.c_void_p()
            _native.check(_native.lib().ds2i_gpu_batch_device_docids(self._h, C.byref(pd)))
            return counts, scores, torch.as_tensor(_Cai(pd.value, (self.nq, k), "<i4"), device=dev)
        return counts, scores

    def stats(self):
        s = np.zeros(8, dtype=np.uint64)
        _native.check(_native.lib().ds2i_gpu_batch_stats(self._h, _p(s, C.c_uint64)))
        keys = ("docs_blocks", "freqs_blocks", "docs_bytes", "freqs_bytes", "block_maxs_read", "docs_scored", "launches", "aux")
        return {k: int(v) for k, v in zip(keys, s)}

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _native.lib().ds2i_gpu_batch_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Group:
    """The index (and wand data) replicated over `ngpus` devices of this process; a batch is cut into cost-balanced shards,
    evaluated concurrently and gathered over NCCL on the first device (ds2i_gpu_group_*)."""

    def __init__(self, index_path, index_type, wand_path=None, ngpus=1, devices=None):
        self._h = C.c_void_p()
        dev = None
        if devices is not None:
            ngpus = len(devices)
            dev = (C.c_int * ngpus)(*devices)
        _native.check(_native.lib().ds2i_gpu_group_open(str(index_path).encode(), index_type.encode(),
                                                     str(wand_path).encode() if wand_path else None, dev, ngpus, C.byref(self._h)))

    def size(self):
        return int(_native.lib().ds2i_gpu_group_size(self._h))

    def query_batch(self, op, queries, k=10):
        """Returns (counts, scores, docids, max per-GPU kernel ms) in the caller's query order."""
        flat, offs = queries if isinstance(queries, tuple) else _flatten(queries)
        nq = len(offs) - 1
        counts = np.zeros(max(nq, 1), dtype=np.uint64)
        scores = np.zeros((max(nq, 1), k), dtype=np.float32)
        docids = np.full((max(nq, 1), k), 0xFFFFFFFF, dtype=np.uint32)
        ms = C.c_float()
        _native.check(_native.lib().ds2i_gpu_group_query_batch(self._h, OPS.index(op), k, _p(flat, C.c_uint32), _p(offs, C.c_uint64), nq,
                                                            _p(counts, C.c_uint64), _p(scores, C.c_float), _p(docids, C.c_uint32), C.byref(ms)))
        return counts[:nq], scores[:nq], docids[:nq], ms.value

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _native.lib().ds2i_gpu_group_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def flatten_queries(queries):
    """list of term-id lists -> (terms u32[], offsets u64[nq+1]): the host buffers of the C ABI."""
    return _flatten(queries)


def query_batch(index, wdata, op, queries, k=10):
    """One call, host buffers in and out (ds2i_gpu_query_batch).  `queries` is a list of term-id
    lists or the (terms, offsets) pair of flatten_queries.  Returns (counts, scores, elapsed_ms)."""
    if isinstance(queries, tuple):
        flat, offs = queries
    else:
        flat, offs = _flatten(queries)
    nq = len(offs) - 1
    counts = np.zeros(max(nq, 1), dtype=np.uint64)
    scores = np.zeros((max(nq, 1), k), dtype=np.float32)
    ms = C.c_float()
    wh = wdata._h if wdata is not None else C.c_void_p()
    _native.check(_native.lib().ds2i_gpu_query_batch(index._h, wh, OPS.index(op), k, _p(flat, C.c_uint32), _p(offs, C.c_uint64), nq,
                                                  _p(counts, C.c_uint64), _p(scores, C.c_float), C.byref(ms)))
    return counts[:nq], scores[:nq], ms.value


class _Operator:
    name = None

    def __init__(self, wdata=None, k=10):
        self._wdata = wdata
        self._k = k
        self._topk = np.zeros(0, dtype=np.float32)

    def __call__(self, index, terms):
        counts, scores, _ = query_batch(index, self._wdata, self.name, [list(terms)], self._k)
        if self.name in RANKED:
            self._topk = scores[0][: int(counts[0])].copy()
        return int(counts[0])

    def batch(self, index, queries):
        return query_batch(index, self._wdata, self.name, queries, self._k)

    def topk(self):
        return self._topk


def _mk(opname):
    return type(opname + "_query", (_Operator,), {"name": opname})


and_query = _mk("and")
and_freq_query = _mk("and_freq")
or_query = _mk("or")
or_freq_query = _mk("or_freq")
ranked_and_query = _mk("ranked_and")
wand_query = _mk("wand")
maxscore_query = _mk("maxscore")
ranked_or_query = _mk("ranked_or")


def read_queries(path, limit=None):
    """read_query (queries.hpp:15-27): one query per line, whitespace-separated term ids."""
    out = []
    with open(path) as f:
        for line in f:
            out.append([int(t) for t in line.split()])
            if limit is not None and len(out) >= limit:
                break
    return out
