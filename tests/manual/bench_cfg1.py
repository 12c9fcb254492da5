"""BASELINE.json configs[0]: KmerMinHash num=500 k=31, add_sequence on a synthetic 5 Mbp genome and on a 1 %-mutated copy,
then compare -- through the reference ABI (host strings in, as a C caller would), next to the CPU port (oracle), with
the results checked against each other."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import sourmash_rust_b200 as smb
from oracle import oracle as orc
from util import mutate, random_dna
g0 = random_dna(5_000_000, 0x5EED0001)
g1 = mutate(g0, 0.01, 0x5EED0002)
for rep in range(3):
    t0 = time.perf_counter()
    a, b = smb.KmerMinHash(500, 31), smb.KmerMinHash(500, 31)
    a.add_sequence(g0); b.add_sequence(g1)
    j = a.compare(b)
    t1 = time.perf_counter()
print("GPU (kmerminhash_add_sequence x2 + kmerminhash_compare, host strings): %.2f ms, jaccard %.6f" % ((t1 - t0) * 1e3, j))
t0 = time.perf_counter()
oa, ob = orc.KmerMinHash(500, 31), orc.KmerMinHash(500, 31)
oa.add_sequence(g0); ob.add_sequence(g1)
oj = oa.compare(ob)
t1 = time.perf_counter()
print("CPU port, 1 thread: %.1f ms, jaccard %.6f" % ((t1 - t0) * 1e3, oj))
assert j == oj and np.array_equal(a.mins_np(), oa.mins_np()) and np.array_equal(b.mins_np(), ob.mins_np())
print("identical sketches and Jaccard")
