"""BASELINE.json configs[3] end to end: containment search (search_minhashes_containment, threshold 0.1) of 1,000
scaled=1000 query sketches against a linear index of N sketches (~5,000 hashes each), the index sharded over the GPUs
of one node.  Launch with torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 \
        tests/manual/run_cfg4.py --index 1000000

Every rank holds N / world index sketches (sorted distinct random hashes <= max_hash) in HBM; each query is half the
hashes of one index sketch plus fresh ones; the query batch is all-gathered, then ONE library call per search
(smgpu_linear_find_sharded: each rank searches its shard, the hit lists are exchanged over NCCL inside the library and
concatenated in rank order = LinearIndex::find's insertion order).  Times are host wall clock around the search call,
max over ranks; --local times the rank-local search (smgpu_linear_find) alone as well."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import sourmash_rust_b200 as smb

def arg(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default

N, NQ, L = arg("--index", 100_000), arg("--queries", 1000), 5000
MAX_HASH = 18446744073709552
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
smb.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
assert N % world == 0 and NQ % world == 0
per, qper = N // world, NQ // world
g = torch.Generator(device=dev); g.manual_seed(0x5EED1000 + rank)
idx = torch.empty((per, L), dtype=torch.int64, device=dev)
for b0 in range(0, per, 25000):  # generate and sort in slabs
    b1 = min(per, b0 + 25000)
    x = torch.randint(0, MAX_HASH, (b1 - b0, L), generator=g, device=dev, dtype=torch.int64)
    idx[b0:b1], _ = torch.sort(x, dim=1)
    del x
assert not (idx[:, 1:] == idx[:, :-1]).any().item()
# this rank's share of the queries: half of the hashes of local index sketch src, half fresh
src = (torch.arange(qper, device=dev) * 37 + 11) % per
q = torch.cat([idx[src][:, ::2], torch.randint(0, MAX_HASH, (qper, L - L // 2), generator=g, device=dev, dtype=torch.int64)], dim=1)
q, _ = torch.sort(q, dim=1)
assert not (q[:, 1:] == q[:, :-1]).any().item()
if world > 1:
    allq = torch.empty((NQ, L), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allq, q.contiguous())
else:
    allq = q
offs_i = torch.arange(per + 1, device=dev, dtype=torch.int64) * L
offs_q = torch.arange(NQ + 1, device=dev, dtype=torch.int64) * L
index = smb.SketchCollection.from_csr(idx.data_ptr(), offs_i.data_ptr(), per, 0, 31, 42, MAX_HASH, on_device=True)
queries = smb.SketchCollection.from_csr(allq.data_ptr(), offs_q.data_ptr(), NQ, 0, 31, 42, MAX_HASH, on_device=True)
del idx
torch.cuda.empty_cache()
if world > 1:
    smb.comm_init_from_torch()
smb.profile_enable(True)
times, local_times = [], []
cap = 64 * NQ
h_offs, h_hits = np.zeros(NQ + 1, dtype=np.uint64), np.zeros(cap, dtype=np.uint64)   # the caller's result buffers (C ABI)
for rep in range(6):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    total = smb._call("smgpu_linear_find_sharded", index._p, queries._p, 1, 0.1, smb._vp(h_offs), smb._vp(h_hits), cap)   # global ids, every rank
    torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
    hits = [h_hits[int(h_offs[q]):int(h_offs[q + 1])].tolist() for q in range(NQ)]
    if rep == 0:
        smb.profile_read("find_stream", reset=True)
        smb.profile_read("join_sort", reset=True)
    if "--local" in sys.argv:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        smb._call("smgpu_linear_find", index._p, queries._p, 1, 0.1, smb._vp(h_offs), smb._vp(h_hits), cap)
        torch.cuda.synchronize(); local_times.append(time.perf_counter() - t0)
kms, kn = smb.profile_read("find_stream", reset=True)
jms, jn = smb.profile_read("join_sort", reset=True)
t = torch.tensor([min(times[1:]), min(local_times[1:]) if local_times else 0.0, kms / max(1, kn)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
# every query must be found at the GLOBAL row id of its source sketch (owner rank * per + local id)
own = all((rank * per + int(src[j])) in hits[rank * qper + j] for j in range(qper))
n_total = sum(len(h) for h in hits)
stat = torch.tensor([n_total, int(own)], dtype=torch.int64, device=dev)
mn = stat.clone()
if world > 1:
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
if rank == 0:
    sec, sec_local, k_ms = float(t[0].item()), float(t[1].item()), float(t[2].item())
    shard_bytes = per * L * 8
    print(json.dumps({"workload": "cfg4: %d scaled=1000 queries x %d-sketch linear index (~%d hashes each), containment > 0.1" % (NQ, N, L),
                      "n_gpus": world, "index_sketches_per_gpu": per, "index_bytes_per_gpu": shard_bytes,
                      "search_ms": sec * 1e3, "local_search_ms": sec_local * 1e3 or None, "pairs": N * NQ, "pairs_per_s": N * NQ / sec,
                      "index_sketches_per_s": N / sec,
                      "query_table_build_ms": (jms / jn) if jn else None, "stream_kernel_ms": k_ms or None, "stream_kernel_gbs": (shard_bytes / (k_ms * 1e-3) / 1e9) if k_ms else None,
                      "whole_search_index_gbs_per_gpu": shard_bytes / sec / 1e9,
                      "hits_total": int(stat[0].item()), "every_planted_source_found": bool(mn[1].item()),
                      "find_path": os.environ.get("SMB200_FIND_PATH", "0")}), flush=True)
if world > 1:
    dist.destroy_process_group()
