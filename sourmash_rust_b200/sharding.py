"""Host-side sharding logic of the multi-GPU path (one process per GPU, torch.distributed).

Nothing here computes sketches or comparisons: it partitions work and moves packed sketch arrays
between ranks.  It works on any backend (NCCL on the B200 box, gloo in the CPU tests).

  * sketching shards by read batch / sample: no collective until the partial sketches of one
    sample are combined (allgather_sketch_state + merge, the reference's KmerMinHash::merge rule,
    src/lib.rs:307-403: set union, abundances summed);
  * the all-vs-all matrix shards by row block; every rank needs all columns, so the packed CSR is
    all-gathered: fixed-width rows of `num` sketches (allgather_rows), variable-length rows of scaled
    sketches (allgather_csr);
  * linear search shards the index; per-rank hit lists are concatenated in rank order, which keeps
    LinearIndex::find's insertion order (src/index/linear.rs:34-44) (merge_hit_lists).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of n items owned by `rank`; blocks differ in size by at most one item
    short of `per` only on the tail ranks."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def allgather_rows(mine, n_total, group=None):
    """mine: (rows_of_this_rank, width) tensor, this rank's block under shard_range(n_total, ...).
    Returns the (n_total, width) tensor of all blocks in rank order (the tail block is padded for
    the collective and the padding is dropped)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return mine[:n_total]
    per = (n_total + world - 1) // world
    if mine.shape[0] < per:
        pad = torch.zeros((per - mine.shape[0],) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
        mine = torch.cat([mine, pad])
    full = torch.empty((per * world,) + tuple(mine.shape[1:]), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(full, mine.contiguous(), group=group)
    return full[:n_total]


def allgather_csr(hashes, lens, group=None):
    """Packed sketches with VARIABLE row lengths (scaled sketches): this rank's `hashes` (1-D int64, its rows back to
    back) and `lens` (1-D int64, one per row; every rank holds the same number of rows) -> (all hashes in rank order,
    offsets of all rows, n_rows + 1).  Lengths travel first; the hash arrays are padded to the longest rank for the
    collective and the padding is dropped."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = hashes.device
    if world == 1:
        all_lens, full = lens, hashes
    else:
        all_lens = torch.empty(world * lens.numel(), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_lens, lens.contiguous(), group=group)
        totals = all_lens.view(world, lens.numel()).sum(dim=1).cpu().tolist()
        width = max(1, max(totals))
        padded = torch.zeros(width, dtype=torch.int64, device=dev)
        padded[:hashes.numel()] = hashes
        gathered = torch.empty(world * width, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered, padded, group=group)
        full = torch.cat([gathered[r * width: r * width + totals[r]] for r in range(world)])
    offsets = torch.zeros(all_lens.numel() + 1, dtype=torch.int64, device=dev)
    offsets[1:] = torch.cumsum(all_lens, 0)
    return full, offsets


def allgather_sketch_state(mins, abunds=None, group=None, device=None):
    """Variable-length (mins, abunds) of one partial sketch per rank -> list over ranks.
    mins/abunds: 1-D uint64 numpy arrays (abunds may be None on every rank)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return [(mins, abunds)]
    device = device or torch.device("cpu")
    # the two lengths travel separately: after KmerMinHash::merge the abundance vector is deliberately NOT
    # truncated with mins (lib.rs:395-400), so len(abunds) may exceed len(mins)
    n_ab = 0 if abunds is None else len(abunds)
    n = torch.tensor([len(mins), n_ab], dtype=torch.int64, device=device)
    sizes = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [(int(s[0].item()), int(s[1].item())) for s in sizes]
    width = max(1, max(max(s) for s in sizes))
    cols = 2 if abunds is not None else 1
    buf = torch.zeros((cols, width), dtype=torch.int64, device=device)
    buf[0, :len(mins)] = torch.from_numpy(np.ascontiguousarray(mins, dtype=np.uint64).view(np.int64)).to(device)
    if abunds is not None:
        buf[1, :n_ab] = torch.from_numpy(np.ascontiguousarray(abunds, dtype=np.uint64).view(np.int64)).to(device)
    full = torch.empty((world, cols, width), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(full.view(world * cols, width), buf, group=group)
    out = []
    for r in range(world):
        m = full[r, 0, :sizes[r][0]].cpu().numpy().view(np.uint64).copy()
        a = full[r, 1, :sizes[r][1]].cpu().numpy().view(np.uint64).copy() if abunds is not None else None
        out.append((m, a))
    return out


def combine_partial_sketches(make_sketch, states):
    """Fold the per-rank partial states into one sketch with the sketch type's own merge.
    make_sketch() -> empty sketch object exposing set_mins(mins, abunds) (or mins_push/abunds_push)
    and merge(other)."""
    acc = None
    for mins, abunds in states:
        part = make_sketch()
        if hasattr(part, "set_mins"):
            part.set_mins(mins, abunds)
        else:
            for v in mins:
                part.mins_push(int(v))
            for v in (abunds if abunds is not None else []):
                part.abunds_push(int(v))
        if acc is None:
            acc = part
        else:
            acc.merge(part)
    return acc


def merge_hit_lists(per_rank_hits, n_index, world):
    """per_rank_hits[r][q] = local row ids (ascending) rank r found for query q in its index shard.
    Returns, per query, the global ids in index insertion order."""
    n_q = len(per_rank_hits[0]) if per_rank_hits else 0
    out = [[] for _ in range(n_q)]
    for r, hits in enumerate(per_rank_hits):
        lo, _ = shard_range(n_index, r, world)
        for q in range(n_q):
            out[q].extend(lo + int(h) for h in hits[q])
    return out
