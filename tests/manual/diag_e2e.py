"""Where does an end-to-end sketch step spend its time?  Host wall clock around each API call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import sourmash_rust_b200 as smb
from bench import torch_reads, MAX_HASH_1000
dev = torch.device("cuda", 0)
R = 1 << 21
g = torch.Generator(device=dev); g.manual_seed(7)
genome = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[torch.randint(0, 4, (100_000_000,), generator=g, device=dev)]
batches = [torch_reads(genome, R, 11 + b, dev) for b in range(3)]
host = []
for b in batches:
    h = torch.empty(R * 150, dtype=torch.uint8, pin_memory=True); h.copy_(b); host.append(h)
torch.cuda.synchronize()
# raw pinned copy rates
d = torch.empty(R * 150, dtype=torch.uint8, device=dev)
for _ in range(2):
    t0 = time.perf_counter(); d.copy_(host[0], non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print("H2D 315 MB pinned: %.2f ms = %.1f GB/s" % ((t1 - t0) * 1e3, R * 150 / (t1 - t0) / 1e9))
big = torch.empty(100_000_000, dtype=torch.float64, device=dev); hb = torch.empty(100_000_000, dtype=torch.float64, pin_memory=True)
for _ in range(2):
    t0 = time.perf_counter(); hb.copy_(big, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print("D2H 800 MB pinned: %.2f ms = %.1f GB/s" % ((t1 - t0) * 1e3, 8e8 / (t1 - t0) / 1e9))
out_m = [torch.zeros(1 << 22, dtype=torch.int64, pin_memory=True) for _ in range(3)]
out_a = [torch.zeros(1 << 22, dtype=torch.int64, pin_memory=True) for _ in range(3)]
for on_device in (True, False):
    mhs = [smb.KmerMinHash(0, k, False, 42, MAX_HASH_1000, True) for k in (21, 31, 51)]
    src = batches if on_device else host
    acc = [0.0, 0.0]
    for s in range(8):
        t0 = time.perf_counter()
        smb.add_reads(mhs, src[s % 3].data_ptr(), R, 150, force=False, on_device=on_device)
        t1 = time.perf_counter()
        for i, m in enumerate(mhs):
            smb._call("kmerminhash_copy_mins", m._p, smb._vp(out_m[i].data_ptr()), smb._vp(out_a[i].data_ptr()), False)
        t2 = time.perf_counter()
        if s >= 2:
            acc[0] += t1 - t0; acc[1] += t2 - t1
    print("on_device=%s: add_reads %.2f ms, 3 x copy_mins %.2f ms per step" % (on_device, acc[0] / 6 * 1e3, acc[1] / 6 * 1e3))
