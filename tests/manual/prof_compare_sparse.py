"""Driver for ncu: the cfg3 all-vs-all matrix (10^4 x 10^4, num = 500) through the default (hash-grouped join) path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import sourmash_rust_b200 as smb
from bench import planted_sketches
N, NUM = 10000, 500
rows = planted_sketches(N, NUM, 0x5EED0100)
offs = np.arange(N + 1, dtype=np.uint64) * np.uint64(NUM)
dev = torch.device("cuda", 0)
common = torch.empty((N, N), dtype=torch.int32, device=dev); size = torch.empty_like(common)
ratio = torch.empty((N, N), dtype=torch.float64, device=dev)
coll = smb.SketchCollection.from_csr(rows.reshape(-1), offs, N, NUM, 31)
for i in range(2):
    smb.compare_matrix_device(coll, coll, "compare", 0, N, 0, N, common.data_ptr(), size.data_ptr(), ratio.data_ptr(), N)
print(int(common.sum().item()))
