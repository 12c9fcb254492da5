"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in sourmash_rust_b200/sharding.py:
row-block partition + all-gather of the packed sketches, partial-sketch exchange + merge, and the
ordered concatenation of sharded search hits.  The arithmetic is checked with the oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as orc
        from sourmash_rust_b200 import sharding
        from util import MAX_HASH_1000, make_reads, random_dna

        # --- row-block all-gather of fixed-width sketches (ragged tail) -------------------------
        n_total, width = 7, 5
        rows = (np.arange(n_total * width, dtype=np.int64).reshape(n_total, width) * 1_000_003) % (1 << 40)
        lo, hi = sharding.shard_range(n_total, rank, world)
        full = sharding.allgather_rows(torch.from_numpy(rows[lo:hi].copy()), n_total)
        assert full.shape == (n_total, width) and np.array_equal(full.numpy(), rows)
        # every item is owned exactly once
        owned = sorted(i for r in range(world) for i in range(*sharding.shard_range(n_total, r, world)))
        assert owned == list(range(n_total))

        # --- all-gather of variable-length rows (scaled sketches), one rank with an empty row ------------
        per = 3
        rng = np.random.default_rng(7)
        all_rows = [np.sort(rng.integers(0, 1 << 50, size=n)).astype(np.int64) for n in (5, 0, 9, 2, 7, 4)]
        mine_rows = all_rows[rank * per:(rank + 1) * per]
        h = torch.from_numpy(np.concatenate(mine_rows)) if sum(len(r) for r in mine_rows) else torch.zeros(0, dtype=torch.int64)
        ln = torch.tensor([len(r) for r in mine_rows], dtype=torch.int64)
        full_h, full_o = sharding.allgather_csr(h, ln)
        assert full_o.tolist() == np.concatenate([[0], np.cumsum([len(r) for r in all_rows])]).tolist()
        assert np.array_equal(full_h.numpy(), np.concatenate(all_rows))

        # --- sharded sketching of one sample: reads split by rank, partial sketches merged --------
        genome = random_dna(60_000, 0x5EED0010)
        n_reads = 3001
        reads = make_reads(genome, n_reads, 150, 0x5EED0011)
        rlo, rhi = sharding.shard_range(n_reads, rank, world)
        for num, mx, ab in ((0, MAX_HASH_1000 * 20, True), (200, 0, False)):
            part = orc.KmerMinHash(num, 31, False, 42, mx, ab)
            part.add_reads(reads[rlo * 150:rhi * 150], rhi - rlo, 150)
            states = sharding.allgather_sketch_state(part.mins_np(), part.abunds_np() if ab else None)
            assert len(states) == world
            merged = sharding.combine_partial_sketches(lambda: orc.KmerMinHash(num, 31, False, 42, mx, ab), states)
            whole = orc.KmerMinHash(num, 31, False, 42, mx, ab)
            whole.add_reads(reads, n_reads, 150)
            assert np.array_equal(merged.mins_np(), whole.mins_np())
            if ab:
                assert np.array_equal(merged.abunds_np()[:merged.size()], whole.abunds_np())

        # --- sharded linear search keeps insertion order --------------------------------------------
        sk = []
        for i in range(9):
            m = orc.KmerMinHash(50, 21)
            m.add_sequence(genome[i * 700:(i * 700) + 4000])
            sk.append(m)
        ilo, ihi = sharding.shard_range(len(sk), rank, world)
        local = [orc.linear_find(sk[ilo:ihi], q, "containment", 0.05) for q in sk[:3]]
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
        merged_hits = sharding.merge_hit_lists(gathered, len(sk), world)
        for q in range(3):
            assert merged_hits[q] == orc.linear_find(sk, sk[q], "containment", 0.05)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    import socket
    with socket.socket() as s:  # a port nobody holds right now
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
