"""A/B timing of the sketch kernels: `SMB200_LIB=<variant .so> python tests/manual/ab_sketch.py [md5]`.
Prints per-kernel average ms (CUDA events around each launch, kernels serialised) and the md5 of the
three sketches so that variants can be checked against each other."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import sourmash_rust_b200 as smb
from bench import torch_reads, MAX_HASH_1000
dev = torch.device("cuda", 0)
R = 1 << 21
g = torch.Generator(device=dev); g.manual_seed(7)
genome = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[torch.randint(0, 4, (100_000_000,), generator=g, device=dev)]
batches = [torch_reads(genome, R, 11 + b, dev) for b in range(3)]
def run(fuse):
    smb.fuse_multi_k(fuse)
    mhs = [smb.KmerMinHash(0, k, False, 42, MAX_HASH_1000, True) for k in (21, 31, 51)]
    for b in batches:
        smb.add_reads(mhs, b.data_ptr(), R, 150, force=False, on_device=True)
    [m.size() for m in mhs]
    smb.profile_enable(True)
    for kind in smb.PROFILE_KINDS:
        smb.profile_read(kind, reset=True)
    for it in range(4):
        for b in batches:
            smb.add_reads(mhs, b.data_ptr(), R, 150, force=False, on_device=True)
    torch.cuda.synchronize()
    out = {}
    for kind in ("sketch_k21", "sketch_k31", "sketch_k51", "sketch_multi"):
        ms, n = smb.profile_read(kind, reset=True)
        if n:
            out[kind] = round(ms / n, 4)
    smb.profile_enable(False)
    tot = sum(out.values())
    print(os.environ.get("SMB200_LIB", "default"), "fused" if fuse else "per-k", out,
          "sum %.4f ms -> %.2f Gbp/s multi-k" % (tot, R * 150 / tot / 1e6), [m.md5sum()[:8] for m in mhs])


run(False)
run(True)
