#!/usr/bin/env python
"""bench.py -- sketch-and-compare hot path on N B200s (one process per GPU), or the reference
algorithm's CPU path on the host cores (--impl reference).

Workload (BASELINE.json configs[1], "cfg2"): scaled=1000 (max_hash = 18446744073709552), multi-k
k=21/31/51, synthetic error-free 150 bp reads from a random-ACGT genome, track_abundance.  One
step = one batch of READS_PER_STEP reads added to the three sketches (a batch is 315 MB of ASCII,
larger than the 126 MB L2).  Metric: Gbp/s sketched = input bases per second (every base is
sketched at all three k).  A second section times the all-vs-all Jaccard matrix of configs[2]
("cfg3": 10,000 sketches, num=500, k=31) and reports comparisons/s.

Prints ONE JSON line on rank 0 (see the keys below).  Timing: CUDA events on the library's own
stream (wrapped as a torch ExternalStream), barrier + synchronize on both sides, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAX_HASH_1000 = 18446744073709552
KSIZES = (21, 31, 51)
READ_LEN = 150
GENOME_LEN = 100_000_000
SEED_GENOME = 0x5EED0010
SEED_READS = 0x5EED0011
# From the ncu --set full capture of the sketch kernels (profiles/sketch_r1_summary.md): executed SASS
# thread-instructions per window (smsp__inst_executed x 32 / windows), DRAM bytes per window
# (dram__bytes_read + dram__bytes_write) and pipe utilisations; used for the integer-issue roofline and
# roofline.traffic.  "multi" = the fused k=21/31/51 kernel (three hashes per window start).
INSTR_PER_WINDOW = {21: 141.0, 31: 151.9, 51: 233.9, "multi": 483.2}
DRAM_BYTES_PER_WINDOW = {21: 1.028, 31: 1.028, 51: 1.042, "multi": 1.035}
NCU_ALU_PIPE_PCT = {21: 66.6, 31: 64.6, 51: 68.3, "multi": 66.1}
NCU_FMAHEAVY_PIPE_PCT = {21: 65.0, 31: 68.0, 51: 67.1, "multi": 71.9}
NCU_ISSUE_PCT = {21: 74.6, 31: 74.4, 51: 72.6, "multi": 74.1}


# ----------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------
def np_genome(n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n, dtype=np.uint8)]


def np_reads(genome, n_reads, seed):
    """error-free reads, uniform start, strand flipped w.p. 0.5"""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = np.empty((n_reads, READ_LEN), dtype=np.uint8)
    comp = np.zeros(256, dtype=np.uint8)
    comp[list(b"ACGT")] = list(b"TGCA")
    ar = np.arange(READ_LEN, dtype=np.int64)
    for lo in range(0, n_reads, 1 << 18):
        hi = min(n_reads, lo + (1 << 18))
        starts = rng.integers(0, len(genome) - READ_LEN, size=hi - lo, dtype=np.int64)
        flips = rng.integers(0, 2, size=hi - lo, dtype=np.uint8).astype(bool)
        r = genome[starts[:, None] + ar[None, :]]
        r[flips] = comp[r[flips][:, ::-1]]
        out[lo:hi] = r
    return out.reshape(-1)


def torch_reads(genome_t, n_reads, seed, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty((n_reads, READ_LEN), dtype=torch.uint8, device=device)
    comp = torch.zeros(256, dtype=torch.uint8, device=device)
    comp[torch.tensor(list(b"ACGT"), device=device, dtype=torch.long)] = torch.tensor(list(b"TGCA"), dtype=torch.uint8, device=device)
    ar = torch.arange(READ_LEN, device=device)
    for lo in range(0, n_reads, 1 << 19):
        hi = min(n_reads, lo + (1 << 19))
        starts = torch.randint(0, genome_t.numel() - READ_LEN, (hi - lo,), generator=g, device=device)
        flips = torch.randint(0, 2, (hi - lo,), generator=g, device=device).bool()
        r = genome_t[starts[:, None] + ar[None, :]]
        rc = comp[r.flip(1).long()]
        out[lo:hi] = torch.where(flips[:, None], rc, r)
    return out.reshape(-1)


def planted_sketches(n_rows, num, seed):
    """cfg3 compare-only input: 100-member clusters; a member keeps each hash of its
    cluster root with the k-mer survival probability of its substitution rate and fills up with
    fresh hashes; rows are the `num` smallest, sorted, distinct."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rates = (0.001, 0.005, 0.01, 0.02, 0.05)
    rows = np.empty((n_rows, num), dtype=np.uint64)
    i = 0
    while i < n_rows:
        root = np.unique(rng.integers(0, 1 << 52, size=4 * num, dtype=np.uint64))
        for m in range(min(100, n_rows - i)):
            keep_p = (1.0 - rates[m % 5]) ** 31
            kept = root[rng.random(root.size) < keep_p]
            fresh = rng.integers(0, 1 << 52, size=4 * num - kept.size + 8, dtype=np.uint64)
            row = np.unique(np.concatenate([kept, fresh]))[:num]
            assert row.size == num
            rows[i] = row
            i += 1
    return rows


# ----------------------------------------------------------------------------------------------
# clocks sampled during the timed regions
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.samples = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, windows):
        sm, mx, reasons = [], 0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for t, line in self.samples:
            if not any(a <= t <= b for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except (ValueError, IndexError):
                continue
            for nm, val in zip(names, f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm restated in C (oracle/), all host threads
# ----------------------------------------------------------------------------------------------
def cpu_sketch_rate(reads, n_reads, threads):
    from oracle import oracle as orc
    t0 = time.perf_counter()
    sk = orc.mt_sketch_reads(reads.tobytes() if isinstance(reads, np.ndarray) else reads, n_reads, READ_LEN,
                             list(KSIZES), 0, MAX_HASH_1000, True, threads)
    dt = time.perf_counter() - t0
    return n_reads * READ_LEN / dt / 1e9, dt, sk


def run_reference(args):
    """--impl reference: same metric/config on the host cores.  The Rust crate cannot be built in
    this image (no cargo/rustc), so this is the C restatement of its algorithm (oracle/oracle.c),
    structured like the reference and pinned to its known-answer tests."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    threads = os.cpu_count() or 1
    genome = np_genome(4_000_000, SEED_GENOME)
    calib = np_reads(genome, 1 << 13, SEED_READS)
    _, dt, _ = cpu_sketch_rate(calib, 1 << 13, threads)
    per_step = int(min(1 << 20, max(1 << 13, (1 << 13) * 2.0 / max(dt, 1e-3))))  # about 2 s per step
    batches = [np_reads(genome, per_step, SEED_READS + 1 + b) for b in range(2)]
    for w in range(args.warmup):
        cpu_sketch_rate(batches[w % 2], per_step, threads)
    t0 = time.perf_counter()
    for s in range(args.steps):
        cpu_sketch_rate(batches[s % 2], per_step, threads)
    dt = time.perf_counter() - t0
    value = args.steps * per_step * READ_LEN / dt / 1e9
    line = {
        "impl": "reference", "metric": "Gbp/s sketched", "value": value, "unit": "Gbp/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "cfg2: scaled=1000 (max_hash=18446744073709552), k=21/31/51 multi-k, 150 bp reads, "
                               "track_abundance; each step a bounded sample of %d reads" % per_step,
                   "reads_per_step": per_step, "read_len": READ_LEN},
        "cpu_baseline": {"value": value, "unit": "Gbp/s", "cores": threads, "kind": "port",
                         "sample": "%d steps x %d reads x 150 bp, 3 k-sizes per base, reads spread over %d threads "
                                   "(oracle/baseline_mt.c)" % (args.steps, per_step, threads)},
        "e2e": {"value": value, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(torch, local_rank):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is allocated:
    first-touch then places the staging buffers on that node, so that eight ranks feeding eight GPUs do not all pull
    their batches across the socket interconnect.  Best effort: returns the node or None."""
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    import sourmash_rust_b200 as smb
    from sourmash_rust_b200 import build
    build.build_library()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: this build has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank)
    smb.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _, sm_count = smb.device_info()
    lib_stream = torch.cuda.ExternalStream(smb.stream_handle(), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    R = args.reads_per_step
    n_bytes = R * READ_LEN
    # ---- inputs: same genome on every rank, rank-private read batches (sharded by read batch) ------
    gg = torch.Generator(device=dev)
    gg.manual_seed(SEED_GENOME)
    genome_t = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[
        torch.randint(0, 4, (GENOME_LEN,), generator=gg, device=dev)]
    n_batches = min(3, args.steps + args.warmup)
    dev_batches = [torch_reads(genome_t, R, SEED_READS + 1000 * rank + b, dev) for b in range(n_batches)]
    host_batches = []
    for b in dev_batches:
        h = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
        h.copy_(b)
        host_batches.append(h)
    torch.cuda.synchronize()

    def new_sketches():
        return [smb.KmerMinHash(0, k, False, 42, MAX_HASH_1000, True) for k in KSIZES]

    clocks = ClockSampler(local_rank)
    clocks.start()
    windows = []

    # ---- device-resident path: `value` --------------------------------------------------------------
    mhs = new_sketches()
    for w in range(args.warmup):
        smb.add_reads(mhs, dev_batches[w % n_batches].data_ptr(), R, READ_LEN, force=False, on_device=True)
    for m in mhs:
        m.size()
    mhs = new_sketches()
    barrier()
    launches0 = smb.launch_count()
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(lib_stream)
    for s in range(args.steps):
        smb.add_reads(mhs, dev_batches[s % n_batches].data_ptr(), R, READ_LEN, force=False, on_device=True)
    sizes = [m.size() for m in mhs]  # folds the last candidates into the sorted sketches
    ev1.record(lib_stream)
    barrier()
    windows.append((t_wall0, time.time()))
    ms_dev = max_over_ranks(ev0.elapsed_time(ev1))
    launches = smb.launch_count() - launches0
    # the same K steps once more with per-kernel CUDA events: the three kernels of a step then run
    # one after the other on the library's stream (in the pass above they overlap at their edges on
    # three streams), so that each duration is the kernel's own -- the roofline's denominator
    # (the timed pass uses the fused k=21/31/51 launch; here first the fused kernel alone, then -- with
    # fusion switched off -- one kernel per k-size, so that the k=31 kernel the metric names is measured too)
    smb.profile_enable(True)
    kern = {}
    for fuse, kinds in ((True, ("sketch_multi",)), (False, ("sketch_k21", "sketch_k31", "sketch_k51"))):
        mhs_k = new_sketches()
        smb.fuse_multi_k(fuse)
        for kind in smb.PROFILE_KINDS:
            smb.profile_read(kind, reset=True)
        barrier()
        t_wall0 = time.time()
        for s in range(args.steps):
            smb.add_reads(mhs_k, dev_batches[s % n_batches].data_ptr(), R, READ_LEN, force=False, on_device=True)
        barrier()
        windows.append((t_wall0, time.time()))
        for kind in kinds:
            kern[kind] = smb.profile_read(kind, reset=True)
        del mhs_k
    smb.fuse_multi_k(True)
    smb.profile_enable(False)
    total_bases = sum_over_ranks(float(args.steps * n_bytes))
    value = total_bases / (ms_dev * 1e-3) / 1e9
    md5_dev = [m.md5sum() for m in mhs]

    # ---- end-to-end path: pinned host buffers in, sketches read back every step ------------------------
    mhs2 = new_sketches()
    # pinned host buffers for the sketches read back every step
    out_m = [torch.zeros(1 << 22, dtype=torch.int64, pin_memory=True) for _ in KSIZES]
    out_a = [torch.zeros(1 << 22, dtype=torch.int64, pin_memory=True) for _ in KSIZES]

    def e2e_step(s):
        smb.add_reads(mhs2, host_batches[s % n_batches].data_ptr(), R, READ_LEN, force=False, on_device=False)
        d2h = 0
        for i, m in enumerate(mhs2):
            n = smb._call("kmerminhash_copy_mins", m._p, smb._vp(out_m[i].data_ptr()), smb._vp(out_a[i].data_ptr()), False)
            d2h += 16 * n
        return d2h

    warm = new_sketches()
    mhs2, keep = warm, mhs2
    for w in range(min(args.warmup, 2)):
        e2e_step(w)
    mhs2 = keep
    barrier()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(lib_stream)
    d2h_total = 0
    for s in range(args.steps):
        d2h_total += e2e_step(s)
    e1.record(lib_stream)
    barrier()
    windows.append((t_wall0, time.time()))
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = total_bases / (ms_e2e * 1e-3) / 1e9
    assert [m.md5sum() for m in mhs2] == md5_dev, "device-resident and host-fed paths disagree"

    # ---- roofline of the dominant kernel (k=51 is the heaviest of the three; k=31 is the metric's k) ---
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    k31_ms, k31_n = kern["sketch_k31"]
    km_ms, km_n = kern["sketch_multi"]
    roof = None
    per_k = {}
    for k, kind in zip(KSIZES + ("multi",), ("sketch_k21", "sketch_k31", "sketch_k51", "sketch_multi")):
        ms, n = kern[kind]
        if n:
            per_k["k%s" % k] = {"launches": n, "avg_ms": ms / n, "gbp_s": args.steps * n_bytes / (ms * 1e-3) / 1e9}
    if km_n:
        # the dominant kernel of the timed step is the fused launch: it reads every base once
        bytes_per_launch = args.steps * n_bytes / km_n  # 1 B (one ASCII base) per window start, SURVEY 8(d)
        achieved = bytes_per_launch / (km_ms / km_n * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "sketch_kernel<21,31,51>", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": DRAM_BYTES_PER_WINDOW["multi"] * bytes_per_launch,
                "traffic_source": "ncu dram__bytes_read+write per window (profiles/sketch_r1_summary.md) x windows per launch",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": km_ms / km_n,
                "duration_source": "CUDA events around each launch on the library's stream, separate pass of the same K steps",
                "note": "HBM is not the binding resource of this kernel (three MurmurHash3 per byte read): see int_pipe"}
    int_pipe = None
    # ---- all-vs-all compare (cfg3), rows sharded by rank, CSR all-gathered over NCCL -------------------
    compare = None
    if not args.no_compare:
        compare = bench_compare(args, smb, torch, dist, dev, rank, world, barrier, max_over_ranks, lib_stream, windows)

    clocks.stop()

    # ---- CPU baseline next to it (rank 0, N=1 only): bounded sample of the same workload --------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as orc
        orc.build()
        threads = os.cpu_count() or 1
        sample_reads = host_batches[0].numpy()
        _, dt, _ = cpu_sketch_rate(sample_reads[: (1 << 13) * READ_LEN], 1 << 13, threads)
        n_s = int(min(R, max(1 << 13, (1 << 13) * 12.0 / max(dt, 1e-3))))
        rate, dt, osk = cpu_sketch_rate(sample_reads[: n_s * READ_LEN], n_s, threads)
        cpu = {"value": rate, "unit": "Gbp/s", "cores": threads, "kind": "port",
               "sample": "first %d reads of batch 0 (%.1f s on %d threads), 3 k-sizes per base" % (n_s, dt, threads)}
        # the same sample through the GPU path must give the same sketches
        chk = new_sketches()
        smb.add_reads(chk, host_batches[0].data_ptr(), n_s, READ_LEN, force=False, on_device=False)
        for g, o in zip(chk, osk):
            assert g.md5sum() == o.md5sum() and np.array_equal(g.abunds_np(), o.abunds_np()), "GPU/CPU sketches differ"
        cpu["parity_checked"] = True

    clk = clocks.summary(windows)
    if rank == 0 and km_n:
        # integer roofline of the fused kernel, two ways: (1) issue slots -- 4 warp instructions per clock
        # per SM at the SM clock seen during the run; (2) the hash-only ceiling -- MurmurHash3 of
        # register-resident k-mers and nothing else, measured live (smgpu_int_peak modes 10-12)
        mhz = clk.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
        issue_peak = 4 * 32 * sm_count * mhz * 1e6
        inst_rate = INSTR_PER_WINDOW["multi"] * (args.steps * n_bytes) / (km_ms * 1e-3)
        hash_rate = {k: smb.int_peak(mode, 256, sm_count * 8) for mode, k in ((10, 21), (11, 31), (12, 51))}
        ceiling = 1.0 / sum(1.0 / hash_rate[k] for k in KSIZES)  # windows/s if the three hashes were all there is
        achieved_w = (args.steps * n_bytes) / (km_ms * 1e-3)
        int_pipe = {"kernel": "sketch_kernel<21,31,51>", "achieved_tinstr_s": inst_rate / 1e12, "peak_tinstr_s": issue_peak / 1e12,
                    "frac": inst_rate / issue_peak, "instr_per_window": INSTR_PER_WINDOW["multi"],
                    "peak_source": "4 warp-instr/clk/SM x 32 lanes x %d SMs x %.0f MHz (sampled during the run)" % (sm_count, mhz),
                    "ncu_issue_slots_pct": NCU_ISSUE_PCT["multi"], "ncu_alu_pipe_pct": NCU_ALU_PIPE_PCT["multi"],
                    "ncu_fmaheavy_pipe_pct": NCU_FMAHEAVY_PIPE_PCT["multi"],
                    "hash_only_ceiling": {"g_hashes_s": {"k%d" % k: hash_rate[k] / 1e9 for k in KSIZES},
                                          "fused_gbp_s": ceiling / 1e9, "frac": achieved_w / ceiling,
                                          "how": "MurmurHash3 x64_128 of register-resident k-mers, no staging / strand "
                                                 "choice / shared memory (hash_peak_kernel), measured in this run"},
                    "k31_kernel": {"gbp_s": args.steps * n_bytes / (k31_ms * 1e-3) / 1e9 if k31_n else None,
                                   "instr_per_window": INSTR_PER_WINDOW[31],
                                   "issue_frac": (INSTR_PER_WINDOW[31] * (args.steps * n_bytes) / (k31_ms * 1e-3) / issue_peak) if k31_n else None,
                                   "hash_only_frac": (args.steps * n_bytes / (k31_ms * 1e-3) / hash_rate[31]) if k31_n else None},
                    "note": "binding resource of the sketch kernels: ALU + FMA-heavy (IMAD) pipes; instruction counts and "
                            "pipe utilisations from the ncu capture in profiles/"}
    if rank == 0:
        line = {
            "metric": "Gbp/s sketched", "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "cfg2: scaled=1000 (max_hash=18446744073709552), k=21/31/51 multi-k, 150 bp reads, "
                                   "track_abundance, %d reads (%d MB ASCII) per step per GPU" % (R, n_bytes >> 20),
                       "reads_per_step_per_gpu": R, "read_len": READ_LEN, "sharding": "read batches per rank, no collective",
                       "l2": "inputs (%d MB per step) larger than the 126 MB L2; %d distinct batches cycled" % (n_bytes >> 20, n_batches),
                       "sketch_sizes": sizes, "host_numa_node": numa},
            "e2e": {"value": e2e_value, "unit": "Gbp/s", "h2d_bytes_per_step": n_bytes,
                    "d2h_bytes_per_step": d2h_total // max(1, args.steps), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": roof, "int_pipe": int_pipe, "sketch_kernels": per_k,
            "cpu_baseline": cpu,
            "clocks": clk,
            "compare": compare,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_compare(args, smb, torch, dist, dev, rank, world, barrier, max_over_ranks, lib_stream, windows):
    N, NUM = args.compare_sketches, 500
    rows = planted_sketches(N, NUM, 0x5EED0100)
    from sourmash_rust_b200 import sharding
    r0, r1 = sharding.shard_range(N, rank, world)
    offsets = (np.arange(N + 1, dtype=np.uint64) * np.uint64(NUM))
    offs_t = torch.from_numpy(offsets.view(np.int64)).to(dev)
    mine = torch.from_numpy(rows[r0:r1].view(np.int64).copy()).to(dev)
    common = torch.empty((r1 - r0, N), dtype=torch.int32, device=dev)
    size = torch.empty((r1 - r0, N), dtype=torch.int32, device=dev)
    ratio = torch.empty((r1 - r0, N), dtype=torch.float64, device=dev)
    steps = max(1, min(args.steps, 5))

    def step():
        full = sharding.allgather_rows(mine, N)  # NCCL all-gather of the packed sketches (no-op at N=1)
        if world > 1:
            torch.cuda.current_stream().synchronize()
        coll = smb.SketchCollection.from_csr(full.data_ptr(), offs_t.data_ptr(), N, NUM, 31, 42, 0, on_device=True)
        smb.compare_matrix_device(coll, coll, "compare", r0, r1 - r0, 0, N, common.data_ptr(), size.data_ptr(),
                                  ratio.data_ptr(), N)
        return coll

    for _ in range(2):
        step()
    smb.profile_enable(True)
    smb.profile_read("compare", reset=True)
    barrier()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(lib_stream)
    for _ in range(steps):
        step()
    e1.record(lib_stream)
    barrier()
    windows.append((t_wall0, time.time()))
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    kms, kn = smb.profile_read("compare", reset=True)
    smb.profile_enable(False)
    # the same matrix with the data-driven path choice overridden: every pair walked (dense kernel)
    smb.compare_path("dense")
    step()
    barrier()
    e0.record(lib_stream)
    for _ in range(2):
        step()
    e1.record(lib_stream)
    barrier()
    ms_dense = max_over_ranks(e0.elapsed_time(e1)) / 2
    smb.compare_path("auto")
    # end to end: host CSR in, f64 Jaccard matrix out to pinned host memory
    out = torch.empty((r1 - r0, N), dtype=torch.float64, pin_memory=True)
    # inputs in pinned host memory (numpy views of pinned torch buffers), as the sketch arm's are
    rows_pin = torch.from_numpy(np.ascontiguousarray(rows).view(np.int64).reshape(-1)).pin_memory()
    offs_pin = torch.from_numpy(np.ascontiguousarray(offsets).view(np.int64)).pin_memory()
    rows_c, offs_c = rows_pin.numpy().view(np.uint64), offs_pin.numpy().view(np.uint64)

    def e2e():
        coll = smb.SketchCollection.from_csr(rows_c, offs_c, N, NUM, 31, 42, 0, on_device=False)
        smb._call("smgpu_compare_matrix", coll._p, r0, r1 - r0, coll._p, 0, N, 0, None, None, smb._vp(out.data_ptr()), N, False)

    e2e()
    barrier()
    e0.record(lib_stream)
    for _ in range(steps):
        e2e()
    e1.record(lib_stream)
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / steps
    # spot parity: 64 x 64 block against the oracle
    ok = None
    if rank == 0:
        from oracle import oracle as orc
        osk = []
        for i in range(64):
            o = orc.KmerMinHash(NUM, 31)
            o.add_many(rows[i])
            osk.append(o)
        oc, osz = orc.compare_matrix(osk, osk)
        ok = bool(np.array_equal(common[:64, :64].cpu().numpy(), oc.astype(np.int32)) and
                  np.array_equal(out[:64, :64].numpy(), oc / np.maximum(1, osz)))
        assert ok, "compare matrix differs from the oracle"
    pairs = float(N) * float(N)
    return {"metric": "Jaccard comparisons/s all-vs-all", "value": pairs / (ms * 1e-3), "unit": "pairs/s",
            "config": "cfg3: %d sketches, num=500, k=31, full ordered matrix (common,size u32 + Jaccard f64), rows "
                      "sharded over %d rank(s), CSR all-gathered over NCCL inside the step" % (N, world),
            "ms_per_step": ms, "steps": steps, "scaling": "strong",
            "kernel_ms_per_step": (kms / kn * (kn / steps)) if kn else None,
            "e2e": {"value": pairs / (ms_e2e * 1e-3), "unit": "pairs/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(rows_c.nbytes + offs_c.nbytes),
                    "d2h_bytes_per_step": int((r1 - r0) * N * 8)},
            "path": "auto: hash-grouped inverted-index join (row hashes in a hash table, column hashes probe it) finds the related pairs (here 1% of all), only those are walked",
            "dense_path": {"value": pairs / (ms_dense * 1e-3), "unit": "pairs/s", "ms_per_step": ms_dense,
                           "note": "every pair walked: rank-compressed fixed-length walk, smgpu_compare_path(1)"},
            "operand_bytes_per_pair": 2 * NUM * 8, "parity_checked_64x64": ok}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads-per-step", type=int, default=1 << 21)
    ap.add_argument("--compare-sketches", type=int, default=10000)
    ap.add_argument("--no-compare", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
