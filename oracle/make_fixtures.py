#!/usr/bin/env python3
"""ORACLE — test infrastructure.  Generates parity fixtures by RUNNING THE REFERENCE ITSELF
(oracle/_ref binaries compiled from /root/reference by oracle/Makefile).  Run in the build
container (needs /root/reference for the test collection); the outputs travel to the GPU box.

  tests/golden/mini.*      (committed, small)  a sub-collection of the reference's own
                           test/test_data/test_collection: the posting lists of the first 60 test
                           queries plus edge-case lists, indexes built by the reference's
                           create_freq_index, results dumped by oracle/drivers/ref_tool.cpp.
  oracle/_ref/data/T.*     (git-ignored, travels) the full test collection: all index types,
                           wand data, per-query results of every operator (known answers of
                           SURVEY.md Appendix C).
"""
import os
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("DS2I_REFERENCE", "/root/reference")
BIN = os.path.join(HERE, "_ref")
DATA = os.path.join(BIN, "data")
GOLDEN = os.path.join(REPO, "tests", "golden")
TYPES_FULL = ["block_optpfor", "block_interpolative", "block_varint", "block_qmx", "opt", "uniform", "single", "ef", "block_mixed"]
OPS = "and:or:ranked_and:wand:maxscore:ranked_or"
FREQ_OPS = "and_freq:or_freq"          # and_query<true> / or_query<true> (queries.hpp:73-76,116-118): separate dump, added in round 2


def run(*cmd, **kw):
    r = subprocess.run(list(cmd), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("failed: " + " ".join(cmd))
    return r.stdout


def read_collection(prefix):
    d = np.fromfile(prefix + ".docs", dtype=np.uint32)
    f = np.fromfile(prefix + ".freqs", dtype=np.uint32)
    s = np.fromfile(prefix + ".sizes", dtype=np.uint32)
    num_docs = int(d[1])
    docs, freqs = [], []
    pos, fpos = 2, 0
    while pos < len(d):
        n = int(d[pos])
        docs.append(d[pos + 1:pos + 1 + n])
        freqs.append(f[fpos + 1:fpos + 1 + n])
        pos += 1 + n
        fpos += 1 + n
    return num_docs, docs, freqs, s[1:]


def write_collection(prefix, num_docs, docs, freqs, sizes):
    with open(prefix + ".docs", "wb") as fd, open(prefix + ".freqs", "wb") as ff:
        np.array([1, num_docs], dtype=np.uint32).tofile(fd)
        for d, f in zip(docs, freqs):
            np.array([len(d)], dtype=np.uint32).tofile(fd)
            np.asarray(d, dtype=np.uint32).tofile(fd)
            np.array([len(f)], dtype=np.uint32).tofile(ff)
            np.asarray(f, dtype=np.uint32).tofile(ff)
    with open(prefix + ".sizes", "wb") as fs:
        np.array([len(sizes)], dtype=np.uint32).tofile(fs)
        np.asarray(sizes, dtype=np.uint32).tofile(fs)


def build_all(prefix, out_prefix, types, queries_path):
    for t in types:
        if t == "block_mixed":
            # only creatable by transformation (mixed_block.hpp:32-36): ref_tool mkmixed re-codes the block_optpfor
            # index through the reference's own block_transformer / write_blocks with seeded per-block types
            run(os.path.join(BIN, "ref_tool"), "mkmixed", "block_optpfor", out_prefix + ".block_optpfor.idx", out_prefix + ".block_mixed.idx", "20261017")
        else:
            run(os.path.join(BIN, "create_freq_index"), t, prefix, out_prefix + "." + t + ".idx", "--check")
    run(os.path.join(BIN, "create_wand_data"), prefix, out_prefix + ".wand")
    for flavour, tool in (("stock", "ref_tool"), ("strict", "ref_tool_strict")):
        run(os.path.join(BIN, tool), "dump", types[0], out_prefix + "." + types[0] + ".idx", out_prefix + ".wand",
            queries_path, out_prefix + ".expected." + flavour + ".bin", OPS)
    dump_freq_ops(out_prefix, types, queries_path)
    # every index type holds the same postings, so the reference returns the same bytes for each of them
    want = open(out_prefix + ".expected.strict.bin", "rb").read()
    for t in types[1:]:
        tmp = out_prefix + ".check." + t + ".bin"
        run(os.path.join(BIN, "ref_tool_strict"), "dump", t, out_prefix + "." + t + ".idx", out_prefix + ".wand", queries_path, tmp, OPS)
        same = open(tmp, "rb").read() == want
        os.remove(tmp)
        if not same:
            raise RuntimeError("reference results differ between %s and %s" % (types[0], t))


def dump_freq_ops(out_prefix, types, queries_path):
    """and_freq / or_freq results of the reference (the operators also touch freq() of every match; the return value is the
    match count) from the first index type, cross-checked on one Elias-Fano type."""
    out = out_prefix + ".expected.freqops.bin"
    run(os.path.join(BIN, "ref_tool_strict"), "dump", types[0], out_prefix + "." + types[0] + ".idx", out_prefix + ".wand", queries_path, out, FREQ_OPS)
    if "opt" in types:
        tmp = out_prefix + ".check.freqops.bin"
        run(os.path.join(BIN, "ref_tool_strict"), "dump", "opt", out_prefix + ".opt.idx", out_prefix + ".wand", queries_path, tmp, FREQ_OPS)
        same = open(tmp, "rb").read() == open(out, "rb").read()
        os.remove(tmp)
        if not same:
            raise RuntimeError("reference and_freq / or_freq results differ between %s and opt" % types[0])


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "freqops":      # only the round-2 addition, from the indexes already built
        dump_freq_ops(os.path.join(DATA, "T"), TYPES_FULL, os.path.join(DATA, "T.queries"))
        dump_freq_ops(os.path.join(GOLDEN, "mini"), TYPES_FULL, os.path.join(GOLDEN, "mini.queries"))
        return
    os.makedirs(DATA, exist_ok=True)
    os.makedirs(GOLDEN, exist_ok=True)
    tcoll = os.path.join(REF, "test", "test_data", "test_collection")
    tq = os.path.join(REF, "test", "test_data", "queries")

    # ---- full test collection (git-ignored) ----
    qdst = os.path.join(DATA, "T.queries")
    with open(tq) as f, open(qdst, "w") as g:
        g.write(f.read())
    build_all(tcoll, os.path.join(DATA, "T"), TYPES_FULL, qdst)
    num_docs, docs, freqs, sizes = read_collection(tcoll)
    # every posting of every list, as the collection has them (what verify_collection checks)
    np.savez_compressed(os.path.join(DATA, "T.lists.npz"), lens=np.array([len(d) for d in docs], dtype=np.uint64),
                        docs=np.concatenate(docs), freqs=np.concatenate(freqs))

    # ---- mini collection (committed) ----
    queries = [[int(t) for t in l.split()] for l in open(tq)]
    lens = np.array([len(d) for d in docs])
    terms = set(t for q in queries[:60] for t in q)
    for want in (1, 2, 127, 128, 129, 255, 256, 257, 384):   # block-size edge cases
        hit = np.nonzero(lens == want)[0]
        if len(hit):
            terms.add(int(hit[0]))
    terms.add(int(np.argmax(lens)))
    terms = sorted(terms)
    remap = {t: i for i, t in enumerate(terms)}
    mq = [[remap[t] for t in q] for q in queries[:60]]
    extra = [remap[int(np.argmax(lens))], 0, 1]
    mq.append(extra)                  # a query over the longest list
    mq.append([mq[0][0], mq[0][0]])   # duplicate term (query_freqs path)
    mq.append([])                     # empty query
    tmp = os.path.join(DATA, "mini")
    write_collection(tmp, num_docs, [docs[t] for t in terms], [freqs[t] for t in terms], sizes)
    mqp = os.path.join(GOLDEN, "mini.queries")
    with open(mqp, "w") as g:
        for q in mq:
            g.write("\t".join(str(t) for t in q) + "\n")
    build_all(tmp, os.path.join(GOLDEN, "mini"), TYPES_FULL, mqp)
    np.savez_compressed(os.path.join(GOLDEN, "mini.collection.npz"), num_docs=np.uint64(num_docs),
                        lens=np.array([len(docs[t]) for t in terms], dtype=np.uint64),
                        docs=np.concatenate([docs[t] for t in terms]), freqs=np.concatenate([freqs[t] for t in terms]),
                        sizes=np.asarray(sizes, dtype=np.uint32), source_terms=np.array(terms, dtype=np.uint32))
    # ---- 1/10-scale synthetic collection (git-ignored): reference-built opt and block_optpfor indexes ----
    builder = os.path.join(REPO, "ds2i_b200", "lib", "ds2i_build")
    if os.access(builder, os.X_OK) and not os.path.exists(os.path.join(DATA, "S10.opt.idx")):
        s10 = os.path.join(DATA, "S10")
        run(builder, "gen", s10, "1000000", "100000", "20261017", "0.35", "2000")
        # only `opt` is kept: the block_optpfor index of the same collection is rebuilt at test time by
        # ds2i_build (byte-identical to the reference's, tests/test_builder.py), which keeps the snapshot small
        run(os.path.join(BIN, "create_freq_index"), "opt", s10, s10 + ".opt.idx")
        run(os.path.join(BIN, "create_wand_data"), s10, s10 + ".wand")
        run(os.path.join(BIN, "ref_tool_strict"), "dump", "opt", s10 + ".opt.idx", s10 + ".wand", s10 + ".queries",
            s10 + ".expected.strict.bin", "and:ranked_and:wand:maxscore", "10", "300")
        for ext in (".docs", ".freqs", ".sizes"):
            os.remove(s10 + ext)
    print("mini: %d lists, %d postings, %d queries" % (len(terms), sum(len(docs[t]) for t in terms), len(mq)))
    for f in sorted(os.listdir(GOLDEN)):
        print("  golden/%s %d bytes" % (f, os.path.getsize(os.path.join(GOLDEN, f))))


if __name__ == "__main__":
    main()
