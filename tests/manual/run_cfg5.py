"""BASELINE.json configs[4] end to end: sketch + all-vs-all Jaccard of N synthetic 5 Mbp genomes (scaled=1000, k=31),
sharded over the GPUs of one node.  Launch with torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 \
        tests/manual/run_cfg5.py --genomes 100000 --cluster 100

Each rank generates its own genomes on the device (clusters of related genomes: a random root, members with
0.1-5 % substitutions), sketches them in one pass (smgpu_sketch_collection); then ONE library call
(smgpu_compare_matrix_allgather) all-gathers the packed sketches over NCCL inside the library -- building the
join's hash table over the rank's own rows meanwhile -- and computes the rank's row block of the N x N matrix.
Times are device-side, max over ranks.  A few cells are checked against the per-object reference ABI."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import sourmash_rust_b200 as smb

def arg(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default

N, CLUSTER, L = arg("--genomes", 1024), arg("--cluster", 16), arg("--length", 5_000_000)
MAX_HASH = 18446744073709552
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
smb.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
assert N % world == 0 and (N // world) % CLUSTER == 0, "genomes per rank must be a whole number of clusters"
per = N // world
lo = rank * per

class _Ptr:  # zero-copy torch view of library-owned device memory
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}

def sync_max(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

# ---- synthetic genomes, generated on the device -----------------------------------------------------------
g = torch.Generator(device=dev); g.manual_seed(0x5EED2000 + rank)
acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
buf = torch.empty(per * L + 64, dtype=torch.uint8, device=dev)
t0 = time.time()
for c in range(per // CLUSTER):
    root = torch.randint(0, 4, (L,), generator=g, device=dev, dtype=torch.uint8)
    for m in range(CLUSTER):
        i = c * CLUSTER + m
        rate = (0.001, 0.005, 0.01, 0.02, 0.05)[m % 5]
        hit = torch.rand(L, generator=g, device=dev) < rate
        codes = torch.where(hit, (root + torch.randint(1, 4, (L,), generator=g, device=dev, dtype=torch.uint8)) % 4, root)
        buf[i * L:(i + 1) * L] = acgt[codes.long()]
offs = torch.arange(per + 1, device=dev, dtype=torch.int64) * L
torch.cuda.synchronize()
gen_s = time.time() - t0
lib_stream = torch.cuda.ExternalStream(smb._call("smgpu_stream"))
if world > 1:
    dist.barrier()

# ---- sketch: one pass over this rank's genomes ----------------------------------------------------------------
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
smb.SketchCollection.sketch_sequences(buf.data_ptr(), offs.data_ptr(), 0, 31, 42, MAX_HASH, on_device=True, n_seqs=min(per, 8))  # warm
e0.record(lib_stream)
mine = smb.SketchCollection.sketch_sequences(buf.data_ptr(), offs.data_ptr(), 0, 31, 42, MAX_HASH, on_device=True, n_seqs=per)
e1.record(lib_stream)
torch.cuda.synchronize()
sketch_ms = sync_max(e0.elapsed_time(e1))
# two sketches through the per-object ABI, for the spot check below
ref = []
for i in (0, 1):
    m = smb.KmerMinHash(0, 31, False, 42, MAX_HASH, False)
    m.add_reads(buf.data_ptr() + i * L, 1, L, on_device=True)
    ref.append(m)
ref_j, ref_mins = ref[0].compare(ref[1]), [r.mins_np() for r in ref]
del buf
torch.cuda.empty_cache()

# ---- exchange + compare: all-gather of the packed sketches inside the library, this rank's row block ------------------
if world > 1:
    smb.comm_init_from_torch()          # NCCL communicator of the library (set-up is not the exchange)
    dist.barrier()
    torch.cuda.synchronize()
common = torch.empty((per, N), dtype=torch.int32, device=dev)
size = torch.empty((per, N), dtype=torch.int32, device=dev)
ratio = torch.empty((per, N), dtype=torch.float64, device=dev)
t_x0 = torch.cuda.Event(enable_timing=True); t_x1 = torch.cuda.Event(enable_timing=True); t_c1 = torch.cuda.Event(enable_timing=True)
t_x0.record(lib_stream)
full = smb.collection_allgather(mine)    # exchange alone, for the record
t_x1.record(lib_stream)
torch.cuda.synchronize()
del full
if world > 1:
    dist.barrier()
t_x1b = torch.cuda.Event(enable_timing=True)
t_x1b.record(lib_stream)
coll = smb.compare_matrix_allgather_device(mine, "compare", common.data_ptr(), size.data_ptr(), ratio.data_ptr(), N)
t_c1.record(lib_stream)
torch.cuda.synchronize()
exch_ms, cmp_ms = sync_max(t_x0.elapsed_time(t_x1)), sync_max(t_x1b.elapsed_time(t_c1))
# the same once more: the first call also grew the library's scratch (GBs of hash table) out of fresh device memory
t_w0 = torch.cuda.Event(enable_timing=True); t_w1 = torch.cuda.Event(enable_timing=True)
if world > 1:
    dist.barrier()
t_w0.record(lib_stream)
coll = smb.compare_matrix_allgather_device(mine, "compare", common.data_ptr(), size.data_ptr(), ratio.data_ptr(), N)
t_w1.record(lib_stream)
torch.cuda.synchronize()
cmp_warm_ms = sync_max(t_w0.elapsed_time(t_w1))
all_rows01 = None

# ---- checks ------------------------------------------------------------------------------------------------------------
h_ptr, o_ptr, total_h = coll.csr_device()
full_o = torch.as_tensor(_Ptr(o_ptr, N + 1), device=dev)
full_h = torch.as_tensor(_Ptr(h_ptr, max(1, total_h)), device=dev)
rows01 = [full_h[int(full_o[lo + i].item()): int(full_o[lo + i + 1].item())].cpu().numpy().view(np.uint64) for i in (0, 1)]
assert all(np.array_equal(a, b) for a, b in zip(rows01, ref_mins)), "one-pass sketch differs from the per-object ABI"
assert float(ratio[0, lo + 1].item()) == ref_j, "matrix cell differs from kmerminhash_compare"
assert bool((ratio[torch.arange(per, device=dev), lo + torch.arange(per, device=dev)] == 1.0).all().item())
related = int((ratio > 0.02).sum().item())
rel = torch.tensor([related], dtype=torch.int64, device=dev)
if world > 1:
    dist.all_reduce(rel)
if rank == 0:
    print(json.dumps({
        "workload": "cfg5: %d genomes x %d bp, scaled=1000, k=31, clusters of %d; sketch + all-vs-all Jaccard" % (N, L, CLUSTER),
        "n_gpus": world, "genomes_per_gpu": per, "generate_s_per_rank": round(gen_s, 1),
        "sketch_ms": sketch_ms, "sketch_gbp_s": N * L / (sketch_ms * 1e-3) / 1e9,
        "exchange_alone_ms": exch_ms, "exchange_and_compare_ms": cmp_ms, "exchange_and_compare_again_ms": cmp_warm_ms, "pairs": N * N,
        "pairs_per_s": N * N / (cmp_ms * 1e-3), "pairs_per_s_again": N * N / (cmp_warm_ms * 1e-3),
        "end_to_end_ms": sketch_ms + cmp_ms, "hashes_total": int(full_o[-1].item()),
        "related_pairs_ratio_gt_0.02": int(rel.item()), "spot_checks": "rows 0,1 and cell (0,1) equal the per-object ABI; diagonal = 1.0"}),
        flush=True)
if world > 1:
    smb.comm_destroy()
    dist.destroy_process_group()
