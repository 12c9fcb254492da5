"""Summarises an ncu report into a markdown table for profiles/:
   python tests/manual/ncu_summary.py REPORT.ncu-rep UNITS_PER_LAUNCH [title] [--json profiles/kernel_constants.json --source profiles/<name>.md] > profiles/<name>.md
Reads the report with `ncu -i ... --page raw --csv` (no GPU needed).  With --json, the per-kernel constants bench.py
needs for its roofline objects (instructions / DRAM bytes per unit, pipe utilisations) are merged into that file under
the kernel's name, so that they come from the same capture as the committed table."""
import csv, io, json, os, subprocess, sys

args = sys.argv[1:]
json_path = source = None
if "--json" in args:
    i = args.index("--json"); json_path = args[i + 1]; del args[i:i + 2]
if "--source" in args:
    i = args.index("--source"); source = args[i + 1]; del args[i:i + 2]
rep, units_per_launch = args[0], float(args[1])
title = args[2] if len(args) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
ci = {h: i for i, h in enumerate(hdr)}

METRICS = [
    ("duration", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("regs/thread", "launch__registers_per_thread"),
    ("dyn smem/block", "launch__shared_mem_per_block_dynamic"),
    ("warps active % of 64/SM (achieved occupancy)", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("warp instructions executed", "smsp__inst_executed.sum"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("eligible warps / scheduler / cycle", "smsp__warps_eligible.avg.per_cycle_active"),
    ("ALU pipe cycles active %", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("FMA-heavy pipe (IMAD) cycles active %", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("FMA pipe instr % (both halves)", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("LSU pipe instr %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("SM throughput %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("DRAM throughput %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 hit rate %", "lts__t_sector_hit_rate.pct"),
    ("smem wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("smem bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("SM cycles elapsed", "sm__cycles_elapsed.avg"),
    ("stall: math pipe throttle (warps/issue)", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall: not selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ("stall: wait (fixed latency)", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall: barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("stall: dispatch", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"),
    ("stall: short scoreboard (smem)", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("stall: long scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
]


def val(r, name):
    if name not in ci:
        return None
    x = r[ci[name]].replace(",", "")
    try:
        return float(x)
    except ValueError:
        return None


def fmt(x):
    if x is None:
        return "n/a"
    if abs(x) >= 1e6:
        return "%.4g" % x
    return ("%.3f" % x).rstrip("0").rstrip(".")


names = []
for r in body:
    n = r[ci["Kernel Name"]]
    n = n.replace("smb200::", "").split("(")[0].replace("void ", "").replace("(int)", "")
    names.append(n)
print("# %s\n" % title)
print("Values are per launch; `--clock-control none`.  Units per launch (windows / pairs): %g.\n" % units_per_launch)
print("| metric | " + " | ".join("`%s`" % n for n in names) + " |")
print("|---|" + "---:|" * len(names))
for label, m in METRICS:
    if m not in ci:
        continue
    u = units[ci[m]]
    print("| %s%s | " % (label, (" (%s)" % u) if u else "") + " | ".join(fmt(val(r, m)) for r in body) + " |")
# derived
def per_unit(r, m, scale=1.0):
    v = val(r, m)
    return None if v is None else v * scale / units_per_launch
def dur_s(r):
    v, u = val(r, "gpu__time_duration.sum"), units[ci["gpu__time_duration.sum"]]
    return v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1e-9)
def to_bytes(r, m):
    v, u = val(r, m), units[ci[m]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
print("| **thread instructions per unit** | " + " | ".join(fmt(per_unit(r, "smsp__inst_executed.sum", 32.0)) for r in body) + " |")
print("| **G units/s (this launch, under ncu)** | " + " | ".join(fmt(units_per_launch / dur_s(r) / 1e9) for r in body) + " |")
print("| **DRAM bytes per unit (read+write)** | " + " | ".join(
    fmt((to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")) / units_per_launch) for r in body) + " |")
print("| **achieved DRAM GB/s** | " + " | ".join(
    fmt((to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")) / dur_s(r) / 1e9) for r in body) + " |")

if json_path:
    try:
        consts = json.load(open(json_path))
    except (OSError, ValueError):
        consts = {}
    for n, r in zip(names, body):
        key = n.replace(" ", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        consts[key] = {
            "instr_per_unit": per_unit(r, "smsp__inst_executed.sum", 32.0),
            "dram_bytes_per_unit": (to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum")) / units_per_launch,
            "issue_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "alu_pct": val(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "fmaheavy_pct": val(r, "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "dram_pct": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "units_per_launch": units_per_launch,
            "source": source or os.path.basename(rep),
        }
    json.dump(consts, open(json_path, "w"), indent=1, sort_keys=True)
    open(json_path, "a").write("\n")
