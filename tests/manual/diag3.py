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
hosts = []
for b in batches:
    h = torch.empty(R * L, dtype=torch.uint8, pin_memory=True); h.copy_(b); hosts.append(h)
torch.cuda.synchronize()
def new(): return [smb.KmerMinHash(0, k, False, 42, MAX_HASH_1000, True) for k in (21, 31, 51)]
out_m = [np.zeros(1 << 22, dtype=np.uint64) for _ in range(3)]
out_a = [np.zeros(1 << 22, dtype=np.uint64) for _ in range(3)]
for rep in range(2):
    mhs = new()
    for s in range(6):
        t0 = time.perf_counter()
        smb.add_reads(mhs, hosts[s % 3].data_ptr(), R, L, force=False, on_device=False)
        t1 = time.perf_counter()
        sz = [m.size() for m in mhs]
        t2 = time.perf_counter()
        for i, m in enumerate(mhs):
            smb._call("kmerminhash_copy_mins", m._p, smb._vp(out_m[i]), smb._vp(out_a[i]), False)
        t3 = time.perf_counter()
        print("rep %d step %d: add_reads(host) %.2f ms, flush %.2f ms, copy out %.2f ms" % (rep, s, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
