"""Cost of ONE kmerminhash_add_sequence call on a 150 bp read (the unmodified reference ABI, one read per call).
SMB200_DEFER_SEQ=0 gives the synchronous path (45 us per call in round 1)."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import sourmash_rust_b200 as smb
rng = random.Random(1)
reads = [bytes(rng.choice(b"ACGT") for _ in range(150)) for _ in range(256)]
mh = smb.KmerMinHash(0, 31, False, 42, 18446744073709552, True)
mh.add_sequence(reads[0])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
t = time.perf_counter()
for i in range(n):
    mh.add_sequence(reads[i & 255])
size = mh.size()
dt = time.perf_counter() - t
print("per-call add_sequence(150 bp): %.1f us/call = %.2f Mbp/s (sketch size %d)" % (dt / n * 1e6, n * 150 / dt / 1e6, size), flush=True)
