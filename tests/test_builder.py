"""CPU tests of the host-side pieces: the format-compatible builder (ds2i_build) must write index
and wand files BYTE-IDENTICAL to what the reference's create_freq_index / create_wand_data wrote for
the same collection (tests/golden/mini.*, built by the reference — oracle/make_fixtures.py)."""
import filecmp
import os
import subprocess

import numpy as np
import pytest

from util import GOLDEN, ORACLE_BIN, have_oracle_bin


@pytest.fixture(scope="module")
def builder():
    from ds2i_b200 import build
    build.build()
    assert os.access(build.BUILDER, os.X_OK)
    return build.BUILDER


@pytest.fixture(scope="module")
def mini_collection(tmp_path_factory):
    """the committed mini collection, written back in ds2i's collection format (README.md:159-181)"""
    d = tmp_path_factory.mktemp("coll")
    z = np.load(os.path.join(GOLDEN, "mini.collection.npz"))
    prefix = str(d / "mini")
    starts = np.concatenate([[0], np.cumsum(z["lens"])]).astype(np.int64)
    with open(prefix + ".docs", "wb") as fd, open(prefix + ".freqs", "wb") as ff:
        np.array([1, int(z["num_docs"])], dtype=np.uint32).tofile(fd)
        for i in range(len(z["lens"])):
            n = np.array([int(z["lens"][i])], dtype=np.uint32)
            n.tofile(fd); z["docs"][starts[i]:starts[i + 1]].astype(np.uint32).tofile(fd)
            n.tofile(ff); z["freqs"][starts[i]:starts[i + 1]].astype(np.uint32).tofile(ff)
    with open(prefix + ".sizes", "wb") as fs:
        np.array([len(z["sizes"])], dtype=np.uint32).tofile(fs)
        z["sizes"].astype(np.uint32).tofile(fs)
    return prefix


@pytest.mark.parametrize("itype", ["block_optpfor", "block_interpolative", "opt"])
@pytest.mark.parametrize("threads", [1, 5])
def test_index_byte_identical_to_reference_build(builder, mini_collection, tmp_path, itype, threads):
    out = str(tmp_path / ("my.%s.idx" % itype))
    subprocess.run([builder, "index", itype, mini_collection, out, str(threads)], check=True)
    assert filecmp.cmp(out, os.path.join(GOLDEN, "mini.%s.idx" % itype), shallow=False)


def test_wand_data_byte_identical_to_reference_build(builder, mini_collection, tmp_path):
    out = str(tmp_path / "my.wand")
    subprocess.run([builder, "wand", mini_collection, out], check=True)
    assert filecmp.cmp(out, os.path.join(GOLDEN, "mini.wand"), shallow=False)


def test_synth_is_deterministic_and_matches_gen(builder, tmp_path):
    """`synth` (no collection on disk, any thread count) == `gen` + `index` + `wand`."""
    a, b, g = str(tmp_path / "a"), str(tmp_path / "b"), str(tmp_path / "g")
    subprocess.run([builder, "synth", a, "20000", "3000", "7", "1", "50"], check=True)
    subprocess.run([builder, "synth", b, "20000", "3000", "7", "6", "50"], check=True)
    for ext in (".block_optpfor.idx", ".wand", ".queries"):
        assert filecmp.cmp(a + ext, b + ext, shallow=False), ext
    subprocess.run([builder, "gen", g, "20000", "3000", "7", "0.35", "50"], check=True)
    subprocess.run([builder, "index", "block_optpfor", g, g + ".idx"], check=True)
    subprocess.run([builder, "wand", g, g + ".wand"], check=True)
    assert filecmp.cmp(g + ".idx", a + ".block_optpfor.idx", shallow=False)
    assert filecmp.cmp(g + ".wand", a + ".wand", shallow=False)
    assert filecmp.cmp(g + ".queries", a + ".queries", shallow=False)
    d = np.fromfile(g + ".docs", dtype=np.uint32)
    assert d[0] == 1 and d[1] == 20000
    # every list strictly increasing and below num_docs
    pos = 2
    while pos < len(d):
        n = int(d[pos]); lst = d[pos + 1:pos + 1 + n]
        assert n >= 1 and np.all(np.diff(lst.astype(np.int64)) > 0) and lst[-1] < 20000
        pos += 1 + n


@pytest.mark.skipif(not have_oracle_bin("create_freq_index"), reason="oracle/_ref not built")
def test_synth_index_byte_identical_to_reference_builder(builder, tmp_path):
    g = str(tmp_path / "g")
    subprocess.run([builder, "gen", g, "50000", "5000", "11", "0.35", "20"], check=True)
    subprocess.run([builder, "index", "block_optpfor", g, g + ".mine.idx"], check=True)
    subprocess.run([os.path.join(ORACLE_BIN, "create_freq_index"), "block_optpfor", g, g + ".ref.idx"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    assert filecmp.cmp(g + ".mine.idx", g + ".ref.idx", shallow=False)
