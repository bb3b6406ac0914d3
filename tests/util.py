"""Shared helpers of the parity tests: locations of the fixtures, parser of the result dumps that
oracle/drivers/ref_tool.cpp writes (the reference's own operators run over the same index file)."""
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_BIN = os.path.join(ROOT, "oracle", "_ref")
ORACLE_DATA = os.path.join(ORACLE_BIN, "data")
DUMP_OPS = ("and", "or", "ranked_and", "wand", "maxscore", "ranked_or")   # order used by make_fixtures.py


def have_oracle_bin(name="ref_tool"):
    return os.access(os.path.join(ORACLE_BIN, name), os.X_OK)


def load_dump(path, ops=DUMP_OPS):
    """{op: (counts[nq] u64, scores[nq,k] f32)}"""
    b = open(path, "rb").read()
    nq, k, nops = struct.unpack("<QQQ", b[:24])
    assert nops == len(ops)
    rec = np.dtype([("c", "<u8"), ("s", "<f4", (k,))])
    out, off = {}, 24
    for op in ops:
        a = np.frombuffer(b, dtype=rec, count=nq, offset=off)
        off += nq * rec.itemsize
        out[op] = (a["c"].copy(), a["s"].copy())
    return out


def read_queries(path):
    return [[int(t) for t in line.split()] for line in open(path)]


def fnv1a64_words(words):
    h = 1469598103934665603
    for w in words:
        h = ((h ^ int(w)) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def run_oracle(tool, *args):
    r = subprocess.run([os.path.join(ORACLE_BIN, tool)] + [str(a) for a in args], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    return r.stdout


def rel_close(a, b, tol=1e-5):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= tol * np.maximum(np.abs(b), 1e-30))
