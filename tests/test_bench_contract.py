"""CPU-side checks of bench.py's contract: the reference arm (`--impl reference`, ds2i's own operators on the host cores)
prints one JSON line with the agreed keys, on a tiny synthetic configuration; the reference block-profiler leg parses."""
import json
import os
import subprocess
import sys

import pytest

from util import GOLDEN, ROOT, have_oracle_bin


@pytest.mark.skipif(not have_oracle_bin("ref_tool"), reason="oracle/_ref/ref_tool not built")
def test_reference_arm_prints_one_json_line(tmp_path):
    env = dict(os.environ, DS2I_BENCH_DATA=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--docs", "20000",
                        "--terms", "2000", "--queries", "200", "--ref-sample", "100"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True and d["value"] > 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.skipif(not have_oracle_bin("ref_tool"), reason="oracle/_ref/ref_tool not built")
def test_reference_arm_other_ranks_stay_silent(tmp_path):
    env = dict(os.environ, DS2I_BENCH_DATA=str(tmp_path), RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(not have_oracle_bin("ref_tool"), reason="oracle/_ref/ref_tool not built")
def test_reference_block_profile_parses(tmp_path):
    sys.path.insert(0, ROOT)
    import bench
    idx = tmp_path / "mini.block_optpfor.idx"
    idx.write_bytes(open(os.path.join(GOLDEN, "mini.block_optpfor.idx"), "rb").read())      # the leg writes next to the index
    paths = {"index": str(idx), "wand": os.path.join(GOLDEN, "mini.wand"), "queries": os.path.join(GOLDEN, "mini.queries")}
    p = bench.reference_block_profile(paths, "ranked_and", 50)
    assert p["queries"] == 50 and p["docs_blocks_per_query"] > 0 and p["freqs_blocks_per_query"] <= p["docs_blocks_per_query"] * 1.0 + 1e-9
    assert 0 < p["bytes_per_query"] <= p["list_bytes_per_query"]
    w = bench.reference_block_profile(paths, "wand", 50)
    assert w["bytes_per_query"] >= p["bytes_per_query"]          # a disjunction decodes at least what the conjunction does
