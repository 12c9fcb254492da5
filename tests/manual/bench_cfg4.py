"""cfg4 shape on one GPU: containment search of 1,000 scaled queries against a linear index shard
(BASELINE.json configs[3]; the full 1M-sketch index is 8 such shards of 125,000).  Prints the search
time and the pair rate; --index N to change the shard size."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import sourmash_rust_b200 as smb
MAX_HASH = 18446744073709552
N_INDEX = int(sys.argv[sys.argv.index("--index") + 1]) if "--index" in sys.argv else 125_000
NQ, L = 1000, 5000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0x5EED1000)
# index: N_INDEX rows of L sorted distinct hashes <= MAX_HASH (distinctness: random 54-bit values, duplicates are ~1e-9)
idx = torch.randint(0, MAX_HASH, (N_INDEX, L), generator=g, device=dev, dtype=torch.int64)
idx, _ = torch.sort(idx, dim=1)
dup = (idx[:, 1:] == idx[:, :-1]).any().item()
assert not dup
# queries: half of the hashes of index sketch (q * 37 % N_INDEX), half fresh
src = (torch.arange(NQ, device=dev) * 37) % N_INDEX
q = torch.cat([idx[src][:, ::2], torch.randint(0, MAX_HASH, (NQ, L - L // 2), generator=g, device=dev, dtype=torch.int64)], dim=1)
q, _ = torch.sort(q, dim=1)
assert not (q[:, 1:] == q[:, :-1]).any().item()
offs_i = (torch.arange(N_INDEX + 1, device=dev, dtype=torch.int64) * L)
offs_q = (torch.arange(NQ + 1, device=dev, dtype=torch.int64) * L)
index = smb.SketchCollection.from_csr(idx.data_ptr(), offs_i.data_ptr(), N_INDEX, 0, 31, 42, MAX_HASH, on_device=True)
queries = smb.SketchCollection.from_csr(q.data_ptr(), offs_q.data_ptr(), NQ, 0, 31, 42, MAX_HASH, on_device=True)
del idx
for mode in ("containment", "similarity"):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        hits = smb.linear_find(index, queries, mode, 0.1, hits_cap=4 * NQ * 64)
        torch.cuda.synchronize(); t1 = time.perf_counter()
    n_hits = sum(len(h) for h in hits)
    ok = all(int(src[j]) in hits[j] for j in range(NQ))
    print("%s: %d index x %d queries in %.1f ms = %.3g pairs/s, %.3g index sketches/s; %d hits, planted sources found: %s" % (
        mode, N_INDEX, NQ, (t1 - t0) * 1e3, N_INDEX * NQ / (t1 - t0), N_INDEX / (t1 - t0), n_hits, ok))
