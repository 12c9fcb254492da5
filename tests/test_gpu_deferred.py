"""add_sequence on short sequences is collected on the host and sketched in batches (minhash.cu, "deferral of short
sequences").  Whatever the interleaving, every observable result must be that of sketching call by call: the oracle is
fed the same calls in the same order and the states are compared at every read point."""
import numpy as np
import pytest

import sourmash_rust_b200 as smb
from oracle import oracle as orc
from util import MAX_HASH_1000, random_dna

pytestmark = pytest.mark.gpu


def _same(g, o, ctx=None):
    assert np.array_equal(g.mins_np(), o.mins_np()), ctx
    ga, oa = g.abunds_np(), o.abunds_np()
    assert (ga is None) == (oa is None), ctx
    if ga is not None:
        assert np.array_equal(ga, oa), ctx


def _reads(n, length, seed):
    g = random_dna(n * length // 4 + length, seed)  # reads overlap: abundances above 1
    rng = np.random.Generator(np.random.PCG64(seed))
    starts = rng.integers(0, len(g) - length, n)
    return [g[int(s):int(s) + length] for s in starts]


@pytest.mark.parametrize("num,mx,abund", [(0, MAX_HASH_1000 * 20, True), (0, MAX_HASH_1000 * 20, False), (200, 0, True),
                                          (200, 0, False), (50, MAX_HASH_1000 * 100, True), (0, 0, True)])
@pytest.mark.parametrize("k", [21, 31, 51, 16])
def test_per_read_calls_equal_the_oracle(num, mx, abund, k):
    g, o = smb.KmerMinHash(num, k, False, 42, mx, abund), orc.KmerMinHash(num, k, False, 42, mx, abund)
    for i, r in enumerate(_reads(1500, 150, 7 + k)):
        g.add_sequence(r); o.add_sequence(r)
        if i in (0, 1, 17, 700):  # read points in the middle: the pending batch is flushed, then continues
            assert g.size() == o.size(), i
    _same(g, o)


def test_interleaving_with_every_other_entry_point():
    rng = np.random.Generator(np.random.PCG64(5))
    for num, mx in ((0, MAX_HASH_1000 * 50), (100, 0)):
        g, o = smb.KmerMinHash(num, 21, False, 42, mx, True), orc.KmerMinHash(num, 21, False, 42, mx, True)
        g2, o2 = smb.KmerMinHash(num, 21, False, 42, mx, True), orc.KmerMinHash(num, 21, False, 42, mx, True)
        reads = _reads(600, 100, 11)
        for i, r in enumerate(reads):
            op = int(rng.integers(0, 12))
            if op == 0:
                h = int(rng.integers(0, 1 << 62))
                g.add_hash(h); o.add_hash(h)
            elif op == 1:
                g.add_word(r[:21]); o.add_word(r[:21])
            elif op == 2:   # force flag flips between calls
                s = r[:40] + b"N" + r[40:]
                g.add_sequence(s, True); o.add_sequence(s, True)
            elif op == 3:   # lower case and shorter than k
                g.add_sequence(r.lower()); o.add_sequence(r.lower())
                g.add_sequence(r[:20]); o.add_sequence(r[:20])
                g.add_sequence(b""); o.add_sequence(b"")
            elif op == 4:   # the other sketch has pending sequences of its own when it is merged in / counted against
                g2.add_sequence(r); o2.add_sequence(r)
                assert g.count_common(g2) == o.count_common(o2)
            elif op == 5:
                g2.add_sequence(reads[(i * 7) % len(reads)]); o2.add_sequence(reads[(i * 7) % len(reads)])
                g.add_from(g2); o.add_from(o2)
            elif op == 6:   # long sequences take the synchronous path, in order
                s = random_dna(70000, 100 + i)
                g.add_sequence(s); o.add_sequence(s)
            else:
                g.add_sequence(r); o.add_sequence(r)
            if i % 97 == 0:
                _same(g, o, i)
        _same(g, o)
        _same(g2, o2)
        assert g.compare(g2) == o.compare(o2)
        c = g.clone() if hasattr(g, "clone") else None
        if c is not None:
            _same(c, o)


def test_error_belongs_to_the_call_that_caused_it():
    g, o = smb.KmerMinHash(0, 21, False, 42, MAX_HASH_1000 * 50, True), orc.KmerMinHash(0, 21, False, 42, MAX_HASH_1000 * 50, True)
    reads = _reads(300, 120, 3)
    for i, r in enumerate(reads):
        if i % 50 == 49:
            bad = r[:70] + b"N" + r[70:]
            with pytest.raises(smb.SourmashError) as ge:
                g.add_sequence(bad)
            with pytest.raises(orc.SourmashError) as oe:
                o.add_sequence(bad)
            assert ge.value.code == oe.value.code == 1101 and ge.value.message == oe.value.message
            _same(g, o, i)   # the windows before the bad one were added (lib.rs:268-273), the deferred reads before them
        else:
            g.add_sequence(r); o.add_sequence(r)
    _same(g, o)


def test_batches_larger_than_the_flush_size_and_three_sketches_per_read():
    ks = (21, 31, 51)
    gs = [smb.KmerMinHash(0, k, False, 42, MAX_HASH_1000, True) for k in ks]
    os_ = [orc.KmerMinHash(0, k, False, 42, MAX_HASH_1000, True) for k in ks]
    genome = random_dna(3_000_000, 99)
    step, L = 250, 1000   # 12 000 reads x 1 000 bp = 12 MB per sketch: crosses the 8 MiB flush size inside the loop
    for s in range(0, len(genome) - L, step):
        r = genome[s:s + L]
        for g in gs:        # the reference usage for a multi-k signature: every sketch gets every read
            g.add_sequence(r)
        for o in os_:
            o.add_sequence(r)
    for g, o in zip(gs, os_):
        _same(g, o)


def test_signature_and_collection_see_deferred_reads():
    g, o = smb.KmerMinHash(0, 31, False, 42, MAX_HASH_1000 * 20, False), orc.KmerMinHash(0, 31, False, 42, MAX_HASH_1000 * 20, False)
    for r in _reads(400, 150, 21):
        g.add_sequence(r); o.add_sequence(r)
    sig = smb.Signature()
    sig.push_mh(g)          # clones the sketch: pending reads included
    assert sig.save_json() == orc.signature_json([o])
    coll = smb.SketchCollection.from_sketches([g])
    assert np.array_equal(coll.rows_np()[0], o.mins_np())
