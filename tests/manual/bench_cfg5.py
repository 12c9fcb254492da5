"""cfg5 / cfg3 shape on one GPU (BASELINE.json configs[4] reduced, configs[2] in full with --genomes 10000 --num 500
--cluster 100): N synthetic 5 Mbp genomes in clusters -> scaled=1000 (or num) k=31 sketches -> all-vs-all Jaccard.  One-pass sketching (smgpu_sketch_collection) against the
per-genome loop through the reference ABI."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import sourmash_rust_b200 as smb
MAX_HASH = 18446744073709552
N = int(sys.argv[sys.argv.index("--genomes") + 1]) if "--genomes" in sys.argv else 128
NUM = int(sys.argv[sys.argv.index("--num") + 1]) if "--num" in sys.argv else 0   # --num 500: cfg3's sketches instead of scaled=1000
CLUSTER = int(sys.argv[sys.argv.index("--cluster") + 1]) if "--cluster" in sys.argv else 16
if NUM:
    MAX_HASH = 0
L = 5_000_000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0x5EED2000)
acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
n_roots = max(1, N // CLUSTER)
roots = [torch.randint(0, 4, (L,), generator=g, device=dev, dtype=torch.uint8) for _ in range(n_roots)]
buf = torch.empty(N * L + 64, dtype=torch.uint8, device=dev)
for i in range(N):
    codes = roots[i % n_roots].clone()
    rate = (0.001, 0.005, 0.01, 0.02, 0.05)[i % 5]
    hit = torch.rand(L, generator=g, device=dev) < rate
    codes[hit] = (codes[hit] + torch.randint(1, 4, (int(hit.sum()),), generator=g, device=dev, dtype=torch.uint8)) % 4
    buf[i * L:(i + 1) * L] = acgt[codes.long()]
offs = torch.arange(N + 1, device=dev, dtype=torch.int64) * L
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    coll = smb.SketchCollection.sketch_sequences(buf.data_ptr(), offs.data_ptr(), NUM, 31, 42, MAX_HASH, on_device=True, n_seqs=N)
    torch.cuda.synchronize(); t1 = time.perf_counter()
print("one pass: %d genomes x %d bp sketched in %.1f ms = %.1f Gbp/s" % (N, L, (t1 - t0) * 1e3, N * L / (t1 - t0) / 1e9))
t0 = time.perf_counter()
sk = []
for i in range(min(N, 32)):
    m = smb.KmerMinHash(NUM, 31, False, 42, MAX_HASH, False)
    m.add_reads(buf.data_ptr() + i * L, 1, L, on_device=True)  # 5 MB slices are 16-byte aligned (L % 16 == 0)
    m.size()
    sk.append(m)
torch.cuda.synchronize(); t1 = time.perf_counter()
print("per-genome loop: %d genomes in %.1f ms = %.1f Gbp/s" % (len(sk), (t1 - t0) * 1e3, len(sk) * L / (t1 - t0) / 1e9))
rows = coll.rows_np()
assert all(np.array_equal(rows[i], sk[i].mins_np()) for i in range(len(sk)))
common = torch.empty((N, N), dtype=torch.int32, device=dev); size = torch.empty_like(common)
ratio = torch.empty((N, N), dtype=torch.float64, device=dev)
for rep in range(2):
    t0 = time.perf_counter()
    smb.compare_matrix_device(coll, coll, "compare", 0, N, 0, N, common.data_ptr(), size.data_ptr(), ratio.data_ptr(), N)
    torch.cuda.synchronize(); t1 = time.perf_counter()
print("all-vs-all of %d sketches (%d hashes avg): %.2f ms = %.3g pairs/s; related pairs %d" % (
    N, sum(len(r) for r in rows) // N, (t1 - t0) * 1e3, N * N / (t1 - t0), int((ratio > 0.05).sum().item())))
