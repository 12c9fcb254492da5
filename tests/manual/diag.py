import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, numpy as np
import sourmash_rust_b200 as smb
MAXH = 18446744073709552
dev = torch.device("cuda", 0)
R, L = 1 << 21, 150
n = R * L
g = torch.Generator(device=dev); g.manual_seed(1)
buf = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[torch.randint(0, 4, (n,), generator=g, device=dev)]
host = torch.empty(n, dtype=torch.uint8, pin_memory=True); host.copy_(buf); torch.cuda.synchronize()
# (a) H2D bandwidth
d2 = torch.empty_like(buf)
for _ in range(2):
    t0 = time.perf_counter(); d2.copy_(host, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("torch pinned H2D: %.1f GB/s" % (n / dt / 1e9))
hp = torch.empty(n, dtype=torch.uint8); hp.copy_(host)
t0 = time.perf_counter(); d2.copy_(hp); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("torch pageable H2D: %.1f GB/s" % (n / dt / 1e9))
# (b) per-call wall time, device-resident
for ks in ((31,), (21, 31, 51)):
    mhs = [smb.KmerMinHash(0, k, False, 42, MAXH, True) for k in ks]
    for i in range(6):
        t0 = time.perf_counter()
        smb.add_reads(mhs, buf.data_ptr(), R, L, force=False, on_device=True)
        dt = time.perf_counter() - t0
        print("ks=%s call %d: %.2f ms" % (ks, i, dt * 1e3))
    t0 = time.perf_counter(); s = [m.size() for m in mhs]; dt = time.perf_counter() - t0
    print("flush: %.2f ms sizes %s" % (dt * 1e3, s))
# (c) host-fed
mhs = [smb.KmerMinHash(0, k, False, 42, MAXH, True) for k in (21, 31, 51)]
for i in range(4):
    t0 = time.perf_counter()
    smb.add_reads(mhs, host.data_ptr(), R, L, force=False, on_device=False)
    dt = time.perf_counter() - t0
    print("host-fed call %d: %.2f ms" % (i, dt * 1e3))
# force=True variant
mhs = [smb.KmerMinHash(0, k, False, 42, MAXH, True) for k in (21, 31, 51)]
for i in range(3):
    t0 = time.perf_counter()
    smb.add_reads(mhs, buf.data_ptr(), R, L, force=True, on_device=True)
    dt = time.perf_counter() - t0
    print("force=True call %d: %.2f ms" % (i, dt * 1e3))
