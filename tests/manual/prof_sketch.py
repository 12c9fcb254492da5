"""Small driver for ncu: a few multi-k read batches through the sketch kernels, then one compare."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import sourmash_rust_b200 as smb
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from bench import torch_reads, planted_sketches, MAX_HASH_1000
dev = torch.device("cuda", 0)
R = 1 << 20
g = torch.Generator(device=dev); g.manual_seed(7)
genome = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[torch.randint(0, 4, (20_000_000,), generator=g, device=dev)]
reads = torch_reads(genome, R, 11, dev)
mhs = [smb.KmerMinHash(0, k, False, 42, MAX_HASH_1000, True) for k in (21, 31, 51)]
# launches 1-9: three passes with one kernel per k-size; launches 10-12: three passes of the fused kernel
smb.fuse_multi_k(False)
for i in range(3):
    smb.add_reads(mhs, reads.data_ptr(), R, 150, force=False, on_device=True)
smb.fuse_multi_k(True)
for i in range(3):
    smb.add_reads(mhs, reads.data_ptr(), R, 150, force=False, on_device=True)
print([m.size() for m in mhs])
if "--compare" in sys.argv:
    N = 2048
    rows = planted_sketches(N, 500, 5)
    offs = np.arange(N + 1, dtype=np.uint64) * np.uint64(500)
    coll = smb.SketchCollection.from_csr(rows.reshape(-1), offs, N, 500, 31)
    c, s, r = smb.compare_matrix(coll, coll)
    print(int(c.sum()))
