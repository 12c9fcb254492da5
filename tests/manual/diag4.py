import os, sys, time
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
for path in ("sparse", "dense", "auto"):
    smb.compare_path(path)
    smb.profile_enable(True)
    for k in ("compare", "join_sort"): smb.profile_read(k, reset=True)
    for i in range(4):
        t0 = time.perf_counter()
        smb.compare_matrix_device(coll, coll, "compare", 0, N, 0, N, common.data_ptr(), size.data_ptr(), ratio.data_ptr(), N)
        dt = time.perf_counter() - t0
    print(path, "%.2f ms -> %.3g pairs/s" % (dt * 1e3, N * N / dt), "compare kernels %.2f ms, sort stage %.2f ms (4 calls)" % (smb.profile_read("compare")[0], smb.profile_read("join_sort")[0]), int(common.sum().item()))
