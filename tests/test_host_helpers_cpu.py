"""Host-side helpers that need no GPU: the C caller's loop of host/feed_reads.c (driven here with recording callbacks
instead of the library's entry points) and the 2-bit packing helper used by the packed-input tests and bench leg."""
import ctypes as C
import os
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _feed_lib():
    from sourmash_rust_b200 import build
    F = C.CDLL(build.build_feed())
    F.feed_reads_mt.restype = C.c_uint64
    F.feed_reads_mt.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p, C.c_uint64,
                                C.c_uint64, C.c_bool, C.c_uint64]
    return F


def _run(n_reads, n_threads, n_mhs, warm, with_size=True):
    F = _feed_lib()
    stride = 12
    reads = np.zeros(n_reads * stride, dtype=np.uint8)
    for i in range(n_reads):
        s = b"R%06d" % i
        reads[i * stride:i * stride + len(s)] = np.frombuffer(s, dtype=np.uint8)
    seen, sized, lock = [], [], threading.Lock()

    @C.CFUNCTYPE(None, C.c_void_p, C.c_char_p, C.c_bool)
    def add_sequence(mh, seq, force):
        with lock:
            seen.append((mh, seq, force, threading.get_ident()))

    @C.CFUNCTYPE(C.c_size_t, C.c_void_p)
    def size(mh):
        with lock:
            sized.append((mh, len(seen)))
        return 0

    handles = (C.c_void_p * (n_threads * n_mhs))(*range(1000, 1000 + n_threads * n_mhs))
    ns = F.feed_reads_mt(C.cast(add_sequence, C.c_void_p), C.cast(size, C.c_void_p) if with_size else None, handles, n_mhs,
                         n_threads, reads.ctypes.data, n_reads, stride, True, warm)
    return ns, seen, sized


def test_feed_reads_every_read_once_per_sketch_in_contiguous_shares():
    n_reads, n_threads, n_mhs = 103, 4, 3
    ns, seen, sized = _run(n_reads, n_threads, n_mhs, warm=5)
    assert ns > 0
    assert all(force for _, _, force, _ in seen)
    per = (n_reads + n_threads - 1) // n_threads
    for t in range(n_threads):
        lo, hi = min(per * t, n_reads), min(per * t + per, n_reads)
        for j in range(n_mhs):
            h = 1000 + t * n_mhs + j
            got = [seq for mh, seq, _, _ in seen if mh == h]
            # warm part first, then the rest: together the thread's share, in order, each read once
            assert got == [b"R%06d" % i for i in range(lo, hi)], (t, j)
        # one host thread per share
        assert len({tid for mh, _, _, tid in seen if 1000 + t * n_mhs <= mh < 1000 + (t + 1) * n_mhs}) == 1
    assert len(seen) == n_reads * n_mhs
    # the flush call: every sketch after the warm part and again at the end
    assert sorted(mh for mh, _ in sized) == sorted(list(range(1000, 1000 + n_threads * n_mhs)) * 2)


def test_feed_reads_fewer_reads_than_threads_and_no_flush():
    ns, seen, sized = _run(3, 8, 1, warm=0, with_size=False)
    assert sorted(seq for _, seq, _, _ in seen) == [b"R%06d" % i for i in range(3)]
    assert sized == []


def test_pack_2bit_layout():
    import sourmash_rust_b200 as smb
    rng = np.random.default_rng(3)
    for read_len in (1, 4, 7, 150):
        reads = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(9, read_len))
        packed = smb.pack_2bit(reads, read_len)
        assert packed.shape == (9, (read_len + 3) // 4) and packed.dtype == np.uint8
        code = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3}
        for r in range(9):
            for b in range(read_len):   # base b: bits 2 * (b % 4) of byte b // 4; padding bits are zero
                assert (packed[r, b // 4] >> (2 * (b % 4))) & 3 == code[reads[r, b]]
            if read_len % 4:
                assert packed[r, -1] >> (2 * (read_len % 4)) == 0
