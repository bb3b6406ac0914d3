import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, ds2i_b200 as d
p = "/tmp/ds2i_b200_data/M_10000000_1000000_20261017/S.block_optpfor.idx"
idx = d.Index(p, "block_optpfor")
terms = np.arange(idx.size(), dtype=np.uint32)
for _ in range(3):
    print(idx.decode_lists_device(terms))
