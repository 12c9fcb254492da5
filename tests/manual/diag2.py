import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
import sourmash_rust_b200 as smb
from bench import torch_reads, MAX_HASH_1000
dev = torch.device("cuda", 0)
R, L = 1 << 21, 150
g = torch.Generator(device=dev); g.manual_seed(1)
genome = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[torch.randint(0, 4, (100_000_000,), generator=g, device=dev)]
batches = [torch_reads(genome, R, 5 + b, dev) for b in range(3)]
torch.cuda.synchronize()
def new(): return [smb.KmerMinHash(0, k, False, 42, MAX_HASH_1000, True) for k in (21, 31, 51)]
for prof in (False, True, False):
    mhs = new()
    for w in range(3):
        smb.add_reads(mhs, batches[w % 3].data_ptr(), R, L, force=False, on_device=True)
    [m.size() for m in mhs]
    mhs = new()
    smb.profile_enable(prof)
    ts = []
    t0 = time.perf_counter()
    for s in range(5):
        t1 = time.perf_counter()
        smb.add_reads(mhs, batches[s % 3].data_ptr(), R, L, force=False, on_device=True)
        ts.append((time.perf_counter() - t1) * 1e3)
    t1 = time.perf_counter()
    sizes = [m.size() for m in mhs]
    tf = (time.perf_counter() - t1) * 1e3
    print("prof=%s steps %s flush %.2f ms total %.2f ms" % (prof, ["%.2f" % t for t in ts], tf, (time.perf_counter() - t0) * 1e3))
    smb.profile_enable(False)
