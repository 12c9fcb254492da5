"""GPU unit tests of individual kernels, driven by small CUDA programs under tests/host/."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sort_scan_reduce_kernels(tmp_path):
    from sourmash_rust_b200 import build
    build.build_library()
    exe = str(tmp_path / "sortops_gpu_test")
    objs = [os.path.join(build.OBJ, o) for o in ("sortops.o", "device.o")]
    subprocess.check_call([build.NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "sortops_gpu_test.cu")] + objs)
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0, out.stdout
