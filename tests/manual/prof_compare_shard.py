"""One rank's share of the cfg3 matrix on ONE GPU, for ncu / timing: rows = an eighth of the 10^4 sketches (a collection of
its own, as a rank holds it), columns = all of them (as gathered).  usage: prof_compare_shard.py [world=8] [reps=5]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import sourmash_rust_b200 as smb
from bench import planted_sketches
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
N, NUM = 10000, 500
rows = planted_sketches(N, NUM, 0x5EED0100)
nr = N // world
dev = torch.device("cuda", 0)
local = smb.SketchCollection.from_csr(rows[:nr].reshape(-1), np.arange(nr + 1, dtype=np.uint64) * np.uint64(NUM), nr, NUM, 31)
allc = smb.SketchCollection.from_csr(rows.reshape(-1), np.arange(N + 1, dtype=np.uint64) * np.uint64(NUM), N, NUM, 31)
common = torch.empty((nr, N), dtype=torch.int32, device=dev); size = torch.empty_like(common)
ratio = torch.empty((nr, N), dtype=torch.float64, device=dev)
lib_stream = torch.cuda.ExternalStream(smb.stream_handle(), device=dev)
for i in range(2):
    smb.compare_matrix_device(local, allc, "compare", 0, nr, 0, N, common.data_ptr(), size.data_ptr(), ratio.data_ptr(), N)
smb.profile_enable(True)
for k in smb.PROFILE_KINDS:
    smb.profile_read(k, reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(lib_stream)
for i in range(reps):
    smb.compare_matrix_device(local, allc, "compare", 0, nr, 0, N, common.data_ptr(), size.data_ptr(), ratio.data_ptr(), N)
e1.record(lib_stream)
torch.cuda.synchronize()
print("shard 1/%d of cfg3: %.3f ms per block" % (world, e0.elapsed_time(e1) / reps),
      {k: round(smb.profile_read(k, reset=True)[0] / reps, 4) for k in ("join_sort", "compare", "compare_probe", "compare_fill", "compare_walk")},
      "related cells", int((common > 0).sum().item()))
