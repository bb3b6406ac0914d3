"""Document-partitioned shards: `ds2i_build shard` writes ordinary ds2i indexes over disjoint docid
ranges; evaluated shard by shard and merged, a query batch must give what the unsharded collection gives — the
reference's own results (tests/golden/mini.expected.*)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from util import GOLDEN, ROOT, load_dump, read_queries, rel_close

ORACLE_C = os.path.join(ROOT, "oracle", "_build", "ds2i_oracle")


@pytest.fixture(scope="module")
def builder():
    from ds2i_b200 import build
    build.build()
    return build.BUILDER


@pytest.fixture(scope="module")
def mini():
    z = np.load(os.path.join(GOLDEN, "mini.collection.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def mini_prefix(tmp_path_factory, mini):
    d = tmp_path_factory.mktemp("shardcoll")
    prefix = str(d / "mini")
    starts = np.concatenate([[0], np.cumsum(mini["lens"])]).astype(np.int64)
    with open(prefix + ".docs", "wb") as fd, open(prefix + ".freqs", "wb") as ff:
        np.array([1, int(mini["num_docs"])], dtype=np.uint32).tofile(fd)
        for i in range(len(mini["lens"])):
            n = np.array([int(mini["lens"][i])], dtype=np.uint32)
            n.tofile(fd); mini["docs"][starts[i]:starts[i + 1]].astype(np.uint32).tofile(fd)
            n.tofile(ff); mini["freqs"][starts[i]:starts[i + 1]].astype(np.uint32).tofile(ff)
    with open(prefix + ".sizes", "wb") as fs:
        np.array([len(mini["sizes"])], dtype=np.uint32).tofile(fs)
        mini["sizes"].astype(np.uint32).tofile(fs)
    return prefix


def _shards(builder, prefix, out, G, itype="block_optpfor"):
    subprocess.run([builder, "shard", itype, prefix, out, str(G)], check=True)


@pytest.mark.skipif(not os.access(ORACLE_C, os.X_OK), reason="oracle/_build/ds2i_oracle not built (python __graft_entry__.py)")
@pytest.mark.parametrize("G", [2, 3])
def test_shards_partition_the_collection(builder, mini, mini_prefix, tmp_path, G):
    """Decoded by the CPU oracle, the shards hold every posting of the collection exactly once, with local docids,
    and the wand data is the matching slice of the collection-wide norm_lens."""
    out = str(tmp_path / "sh")
    _shards(builder, mini_prefix, out, G)
    N = int(mini["num_docs"])
    starts = np.concatenate([[0], np.cumsum(mini["lens"])]).astype(np.int64)
    seen = np.zeros(len(mini["docs"]), dtype=np.int64)
    full_wand = open(os.path.join(GOLDEN, "mini.wand"), "rb").read()
    full_norm = np.frombuffer(full_wand, dtype="<f4", count=N, offset=16)
    for g in range(G):
        lo, hi = N * g // G, N * (g + 1) // G
        terms = np.fromfile("%s.%d.terms" % (out, g), dtype=np.uint32)
        assert np.all(np.diff(terms.astype(np.int64)) > 0)
        dump = str(tmp_path / ("lists.%d.bin" % g))
        subprocess.run([ORACLE_C, "lists", "block_optpfor", "%s.%d.idx" % (out, g), dump], check=True)
        raw = open(dump, "rb").read()
        off = 0
        for t in terms:
            n = struct.unpack_from("<Q", raw, off)[0]; off += 8
            d = np.frombuffer(raw, dtype="<u4", count=n, offset=off); off += 4 * n
            f = np.frombuffer(raw, dtype="<u4", count=n, offset=off); off += 4 * n
            gd = mini["docs"][starts[t]:starts[t + 1]]
            gf = mini["freqs"][starts[t]:starts[t + 1]]
            a, b = np.searchsorted(gd, lo), np.searchsorted(gd, hi)
            assert b - a == n and np.array_equal(d + lo, gd[a:b]) and np.array_equal(f, gf[a:b])
            seen[starts[t] + a:starts[t] + b] += 1
        assert off == len(raw)
        w = open("%s.%d.wand" % (out, g), "rb").read()
        assert struct.unpack_from("<Q", w, 8)[0] == hi - lo
        assert np.array_equal(np.frombuffer(w, dtype="<f4", count=hi - lo, offset=16), full_norm[lo:hi])
    assert np.all(seen == 1)


@pytest.mark.gpu
@pytest.mark.parametrize("G", [2, 3])
def test_sharded_queries_equal_the_unsharded_reference(native_lib, builder, mini, mini_prefix, tmp_path, G):
    import ds2i_b200 as d
    from ds2i_b200 import sharding
    out = str(tmp_path / "sh")
    _shards(builder, mini_prefix, out, G)
    N = int(mini["num_docs"])
    ranges = sharding.shard_ranges(N, G)
    shards = [sharding.Shard(out, "block_optpfor", g, ranges[g][0]) for g in range(G)]
    df, ndocs = sharding.exchange_global_stats(shards, len(mini["lens"]))
    assert ndocs == N and np.array_equal(df, mini["lens"].astype(np.int64))
    queries = read_queries(os.path.join(GOLDEN, "mini.queries"))
    strict = load_dump(os.path.join(GOLDEN, "mini.expected.strict.bin"))
    # the unsharded device results, for the docids
    idx = d.Index(os.path.join(GOLDEN, "mini.block_optpfor.idx"), "block_optpfor")
    wd = d.WandData(os.path.join(GOLDEN, "mini.wand"))
    whole = d.QueryBatch(idx, wd, queries)
    for op in ("and", "or", "ranked_and", "wand", "maxscore", "ranked_or"):
        counts, scores, docids = sharding.query_sharded(shards, op, queries, 10)
        ec, es = strict[op]
        assert np.array_equal(counts, ec), op
        if op in d.RANKED:
            assert rel_close(scores, es), op
            if op == "ranked_and":       # same statistics, same summation order: bit-identical (ranked_or runs on the union kernel: 1e-5)
                assert np.array_equal(scores.view(np.uint32), es.view(np.uint32)), op
            whole.run(op, 10)
            wids = whole.fetch_docids()
            for qi in range(len(queries)):
                n = int(ec[qi])
                s = es[qi][:n]
                # documents whose score is unique inside the top-k and above its last entry are determined uniquely
                sure = [j for j in range(n) if s[j] > s[n - 1] and (j == 0 or s[j - 1] > s[j]) and (j + 1 >= n or s[j] > s[j + 1])]
                assert np.array_equal(docids[qi][sure], wids[qi][sure]), (op, qi)
                assert np.all(docids[qi][n:] == 0xFFFFFFFF)
