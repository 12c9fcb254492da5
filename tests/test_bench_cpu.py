"""CPU check of bench.py's reference arm (the contract's `--impl reference` leg): it must run without a GPU and print
ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Gbp/s sketched" and d["unit"] == "Gbp/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]
    # the compare half of the metric has a CPU arm too (a bounded block of cfg3)
    c = d["compare"]
    assert c["impl"] == "reference" and c["unit"] == "pairs/s" and c["value"] > 0 and c["cpu_baseline"]["kind"] == "port"


def test_kernel_constants_file_names_the_kernels_bench_reads():
    sys.path.insert(0, ROOT)
    import bench
    consts = bench.kernel_constants()
    for key in bench.KERNEL_OF.values():
        assert key in consts and consts[key]["instr_per_unit"] > 0 and consts[key]["dram_bytes_per_unit"] > 0, key
