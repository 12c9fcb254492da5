"""Randomised differential test: random sketch parameters and random sequences of operations through
the C ABI (GPU) and through the oracle, states compared after every step.  Deterministic seeds."""
import numpy as np
import pytest

import sourmash_rust_b200 as smb
from oracle import oracle as orc
from util import random_dna

pytestmark = pytest.mark.gpu


def _dirty(rng, n):
    a = np.frombuffer(random_dna(n, int(rng.integers(1, 1 << 30))), dtype=np.uint8).copy()
    if n and rng.random() < 0.5:  # lower-case stretch
        s = int(rng.integers(0, n))
        a[s:s + int(rng.integers(1, 60))] |= 0x20
    if n and rng.random() < 0.5:  # a few invalid bytes
        for _ in range(int(rng.integers(1, 4))):
            a[int(rng.integers(0, n))] = rng.choice(np.frombuffer(b"NnRY-*\x80\xff ", dtype=np.uint8))
    if n > 20 and rng.random() < 0.3:  # a repeat (abundances)
        s = int(rng.integers(0, n // 2))
        L = int(rng.integers(1, n // 2))
        a[n - L:] = a[s:s + L]
    return a.tobytes().replace(b"\0", b"N")


def _same(g, o, ctx):
    gm, om = g.mins_np(), o.mins_np()
    assert np.array_equal(gm, om), (ctx, len(gm), len(om))
    ga, oa = g.abunds_np(), o.abunds_np()
    assert (ga is None) == (oa is None), ctx
    if ga is not None:
        assert np.array_equal(ga, oa), ctx


@pytest.mark.parametrize("seed", range(150))
def test_random_operation_sequences(seed):
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    k = int(rng.choice([21, 31, 51, 3, 5, 8, 15, 16, 17, 32, 33, 47, 48, 64, 65, 80]))
    kind = int(rng.integers(0, 5))
    num, mx = [(int(rng.integers(1, 300)), 0), (0, int(rng.integers(1, 1 << 63)) >> int(rng.integers(0, 8))),
               (int(rng.integers(1, 50)), int(rng.integers(1 << 58, 1 << 63))), (0, 0),
               (int(rng.integers(300, 2000)), 0)][kind]
    abund = bool(rng.integers(0, 2))
    mk = lambda mod: mod.KmerMinHash(num, k, False, 42, mx, abund)
    g, o = mk(smb), mk(orc)
    g2, o2 = mk(smb), mk(orc)
    merged = False
    for step in range(int(rng.integers(3, 9))):
        op = int(rng.integers(0, 7))
        ctx = (seed, step, op, k, num, mx, abund)
        if merged and op in (0, 1, 2, 3):
            op = 4  # after a merge the reference's abundance vector may be out of step with mins: compare only
        if op == 0:  # add_sequence, force
            s = _dirty(rng, int(rng.choice([0, 1, k - 1, k, k + 1, 200, 5000, 9000, 20000])))
            g.add_sequence(s, True); o.add_sequence(s, True)
        elif op == 1:  # add_sequence, strict: error parity + partial state
            s = _dirty(rng, int(rng.choice([k, 300, 6000])))
            ge = oe = None
            try:
                g.add_sequence(s, False)
            except smb.SourmashError as e:
                ge = (e.code, e.message)
            try:
                o.add_sequence(s, False)
            except orc.SourmashError as e:
                oe = (e.code, e.message)
            if oe and not oe[1].isascii():
                oe = ge  # the reference unwraps a UTF-8 conversion there (panic); message bytes are not comparable
            assert ge == oe, ctx
        elif op == 2:  # add_hash stream with repeats
            hs = rng.integers(0, 1 << 63, size=int(rng.integers(1, 400)), dtype=np.uint64)
            hs = np.concatenate([hs, hs[: len(hs) // 3]])
            if mx:
                hs = np.concatenate([hs, rng.integers(0, mx, size=50, dtype=np.uint64)])
            rng.shuffle(hs)
            g.add_many(hs.tolist()); o.add_many(hs)
        elif op == 3:  # second sketch gets data, then merge into the first
            s = _dirty(rng, int(rng.choice([500, 4000])))
            g2.add_sequence(s, True); o2.add_sequence(s, True)
            _same(g2, o2, ctx)
            g.merge(g2); o.merge(o2)
            merged = True
            assert g.track_abundance() == o.track_abundance()
        elif op == 4:  # pair operations
            assert g.count_common(g2) == o.count_common(o2), ctx
            assert g.compare(g2) == o.compare(o2), ctx
            assert g2.compare(g) == o2.compare(o), ctx
        elif op == 5:  # batch of ragged sequences
            lens = rng.integers(0, 400, size=int(rng.integers(1, 40)))
            offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
            buf = _dirty(rng, int(offs[-1])) if offs[-1] else b""
            if merged:
                continue
            if offs[-1]:
                g.add_sequences(buf, offs, force=True)
                for i in range(len(lens)):
                    o.add_sequence(buf[int(offs[i]):int(offs[i + 1])], True)
        else:  # reads of fixed length
            if merged:
                continue
            L = int(rng.choice([k, k + 3, 100, 150]))
            n = int(rng.integers(1, 60))
            buf = _dirty(rng, L * n)
            g.add_reads(buf, n, L, force=True); o.add_reads(buf, n, L, True)
        if merged and abund:
            # lib.rs:395-400: merge leaves abunds untruncated; only mins are comparable from here on
            assert np.array_equal(g.mins_np(), o.mins_np()), ctx
        else:
            _same(g, o, ctx)
    assert g.md5sum() == o.md5sum()
