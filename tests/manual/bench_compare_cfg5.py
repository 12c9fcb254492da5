"""All-vs-all Jaccard at BASELINE config 5's size WITHOUT the genomes: N synthetic scaled sketches (~5 000 hashes each,
clusters of 100 sharing most of their hashes), made on the device, rows sharded over the ranks; each rank computes its
row block (smgpu_compare_matrix_allgather; the f64 matrix is produced in row blocks through one output buffer).
The same script on 1 GPU and on 8 gives the pair of numbers for the compare step alone.

    python tests/manual/bench_compare_cfg5.py --sketches 100000                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29546 \
        tests/manual/bench_compare_cfg5.py --sketches 100000
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import sourmash_rust_b200 as smb


def arg(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


N, CLUSTER, L = arg("--sketches", 20000), arg("--cluster", 100), arg("--hashes", 5000)
MAX_HASH = 18446744073709552
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
smb.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    smb.comm_init_from_torch()
assert N % world == 0 and (N // world) % CLUSTER == 0
per = N // world

# ---- this rank's sketches: cluster roots of 2 L hashes; a member keeps a share of the root (60-95 %) + fresh hashes
g = torch.Generator(device=dev); g.manual_seed(0x5EED3000 + rank)
rows = torch.empty((per, L), dtype=torch.int64, device=dev)
for c in range(per // CLUSTER):
    root = torch.randint(0, MAX_HASH, (2 * L,), generator=g, device=dev, dtype=torch.int64)
    for m in range(CLUSTER):
        share = int(L * (0.6, 0.75, 0.85, 0.95)[m % 4])
        pick = torch.randperm(2 * L, generator=g, device=dev)[:share]
        fresh = torch.randint(0, MAX_HASH, (L - share,), generator=g, device=dev, dtype=torch.int64)
        rows[c * CLUSTER + m], _ = torch.sort(torch.cat([root[pick], fresh]))
assert not (rows[:, 1:] == rows[:, :-1]).any().item()
offs = torch.arange(per + 1, device=dev, dtype=torch.int64) * L
mine = smb.SketchCollection.from_csr(rows.data_ptr(), offs.data_ptr(), per, 0, 31, 42, MAX_HASH, on_device=True)
del rows
lib_stream = torch.cuda.ExternalStream(smb._call("smgpu_stream"))

# ---- row blocks of this rank's share through one output buffer
blk = max(CLUSTER, min(per, (1 << 30) // N // CLUSTER * CLUSTER))
ratio = torch.empty((blk, N), dtype=torch.float64, device=dev)
allc = smb.collection_allgather(mine)


def matrix():
    related = 0
    for b0 in range(0, per, blk):
        bn = min(blk, per - b0)
        smb.compare_matrix_device(mine, allc, "compare", b0, bn, 0, N, None, None, ratio.data_ptr(), N)
        related += int((ratio[:bn] > 0.02).sum().item())
    return related


matrix()                                 # scratch grows out of fresh device memory the first time
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record(lib_stream)
allc = smb.collection_allgather(mine)
e1.record(lib_stream)
related = matrix()
e2.record(lib_stream)
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1), e1.elapsed_time(e2), float(related)], dtype=torch.float64, device=dev)
if world > 1:
    mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = t.clone(); dist.all_reduce(sm)
    t = torch.stack([mx[0], mx[1], sm[2]])
if rank == 0:
    ex, cm = float(t[0]), float(t[1])
    print(json.dumps({"workload": "compare step of cfg5 alone: %d scaled sketches x %d hashes, clusters of %d, all-vs-all Jaccard (f64), "
                                  "row blocks of %d through one output buffer" % (N, L, CLUSTER, blk),
                      "n_gpus": world, "sketches_per_gpu": per, "hashes_total": N * L, "exchange_ms": ex, "compare_ms": cm,
                      "pairs": N * N, "pairs_per_s": N * N / (cm * 1e-3), "related_pairs_ratio_gt_0.02": int(t[2])}), flush=True)
if world > 1:
    smb.comm_destroy()
    dist.destroy_process_group()
