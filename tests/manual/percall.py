"""Cost of ONE kmerminhash_add_sequence call on a 150 bp read (the unmodified reference ABI, one read per call,
ffi.rs:55-70), timed from a C loop (sourmash_rust_b200/host/feed_reads.c) and from Python/ctypes.
SMB200_DEFER_SEQ=0 gives the synchronous path (45 us per call in round 1).
usage: percall.py [n_reads] [host_threads]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import sourmash_rust_b200 as smb
from util import MAX_HASH_1000, make_reads, random_dna

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if os.environ.get("SMB200_DEFER_SEQ") == "0":
    n = min(n, 20000)
genome = random_dna(2_000_000, 1)
reads = np.frombuffer(make_reads(genome, min(n, 200_000), 150, 2), dtype=np.uint8).reshape(-1, 150)
reads = np.tile(reads, ((n + len(reads) - 1) // len(reads), 1))[:n]
z = np.zeros((n, 151), dtype=np.uint8)
z[:, :150] = reads          # NUL-terminated strings, stride 151

for ks in ((31,), (21, 31, 51)):
    groups = [[smb.KmerMinHash(0, k, False, 42, MAX_HASH_1000, True) for k in ks] for _ in range(threads)]
    warm = min(n // threads // 4, 100_000)                # per thread, untimed: sets up the thread's stream and scratch
    dt = smb.feed_reads(groups, z, n, 151, warm_reads=warm)   # timed region ends with every sketch flushed
    sizes = [m.size() for g in groups for m in g]
    timed = n - warm * threads
    calls = timed * len(ks)
    print("C loop, k=%s, %d host thread(s): %.3f us/call (per thread: %.3f) incl. final flush = %.1f Mbp/s "
          "(%d timed reads, sketch sizes %s)" % (ks, threads, dt / calls * 1e6, dt / calls * 1e6 * threads, timed * 150 / dt / 1e6,
                                                 timed, sizes[:3]), flush=True)

mh = smb.KmerMinHash(0, 31, False, 42, MAX_HASH_1000, True)
rl = [bytes(r) for r in reads[:256]]
m = min(n, 100_000)
t = time.perf_counter()
for i in range(m):
    mh.add_sequence(rl[i & 255])
size = mh.size()
dt = time.perf_counter() - t
print("Python/ctypes loop, k=31: %.2f us/call = %.2f Mbp/s (sketch size %d)" % (dt / m * 1e6, m * 150 / dt / 1e6, size), flush=True)
