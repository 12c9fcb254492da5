import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import sourmash_rust_b200 as smb
from bench import planted_sketches
N, NUM = 4096, 500
rows = planted_sketches(N, NUM, 5)
offs = np.arange(N + 1, dtype=np.uint64) * np.uint64(NUM)
dev = torch.device("cuda", 0)
common = torch.empty((N, N), dtype=torch.int32, device=dev); size = torch.empty_like(common)
ratio = torch.empty((N, N), dtype=torch.float64, device=dev)
coll = smb.SketchCollection.from_csr(rows.reshape(-1), offs, N, NUM, 31)
smb.compare_path(sys.argv[1] if len(sys.argv) > 1 else "dense")
for i in range(2):
    smb.compare_matrix_device(coll, coll, "compare", 0, N, 0, N, common.data_ptr(), size.data_ptr(), ratio.data_ptr(), N)
print(int(common.sum().item()))
