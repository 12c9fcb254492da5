import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import sourmash_rust_b200 as smb
MAXH = 18446744073709552
dev = torch.device("cuda", 0)
n = (1 << 21) * 150
g = torch.Generator(device=dev); g.manual_seed(1)
buf = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[torch.randint(0, 4, (n + 4096,), generator=g, device=dev)]
for L in (150, 151, 152, 153):
    R = n // L
    for k in (31,):
        mh = smb.KmerMinHash(0, k, False, 42, MAXH, False)
        ts = []
        for i in range(5):
            t0 = time.perf_counter()
            smb.add_reads([mh], buf.data_ptr(), R, L, force=True, on_device=True)
            ts.append((time.perf_counter() - t0) * 1e3)
        print("read_len %d k=%d: %.2f ms" % (L, k, min(ts)))
