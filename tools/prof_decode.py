"""Decode every list of the synthetic block_optpfor index three times (the workload tools/prof_decode.sh captures with ncu)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ds2i_b200 as d
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from microbench import synth              # builds the index if the box does not have it yet
dname = synth(10_000_000, 1_000_000, 20261017, ["block_optpfor", "block_interpolative"])
idx = d.Index(os.path.join(dname, "S.block_optpfor.idx"), "block_optpfor")
terms = np.arange(idx.size(), dtype=np.uint32)
for _ in range(3):
    print(idx.decode_lists_device(terms))
