"""Builds sourmash_rust_b200/libsourmash.so from csrc/ with nvcc for sm_100a (in-tree, so the
binary travels with the repository snapshot).  `python -m sourmash_rust_b200.build [--force]`."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libsourmash.so")
FEED_SRC = os.path.join(HERE, "host", "feed_reads.c")   # a C caller's loop over the reference ABI (bench / percall timing)
FEED_LIB = os.path.join(HERE, "libfeedreads.so")
SOURCES = ["device.cu", "sketch.cu", "sortops.cu", "compare.cu", "join.cu", "find_stream.cu", "protein.cu", "minhash.cu", "collection.cu", "sketch_many.cu", "nodegraph.cu", "comm.cu", "signature.cpp", "ffi.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-Wall",
         "-x", "cu"]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(os.path.dirname(HERE), "include", f) for f in ("sourmash.h", "sourmash_b200.h")]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def _compile(src):
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout))
    return obj


def build_feed(force=False):
    if force or not os.path.exists(FEED_LIB) or os.path.getmtime(FEED_SRC) > os.path.getmtime(FEED_LIB):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-Wall", "-o", FEED_LIB, FEED_SRC, "-lpthread"])
    return FEED_LIB


def build_library(force=False, verbose=False):
    build_feed(force)
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(_compile, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt", "-lz"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    if verbose:
        print("built", LIB)
    return LIB


def build_tsan(verbose=False):
    """A ThreadSanitizer build of the same sources (host code instrumented; device code as usual) under build/tsan/, for
    tests/test_tsan_cpu.py.  Never the library the package loads."""
    out_dir = os.path.join(OBJ, "tsan")
    lib = os.path.join(out_dir, "libsourmash_tsan.so")
    if os.path.exists(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in _deps()):
        return lib
    os.makedirs(out_dir, exist_ok=True)
    san = ["-Xcompiler", "-fsanitize=thread,-g,-fno-omit-frame-pointer"]

    def compile_one(src):
        obj = os.path.join(out_dir, os.path.splitext(src)[0] + ".o")
        r = subprocess.run([NVCC] + FLAGS + san + ["-c", os.path.join(CSRC, src), "-o", obj],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc (tsan) failed for %s:\n%s" % (src, r.stdout))
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([NVCC, "-shared", "-o", lib] + objs + san + ["-lcudart_static", "-lpthread", "-ldl", "-lrt", "-lz"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link (tsan) failed:\n" + r.stdout)
    if verbose:
        print("built", lib)
    return lib


if __name__ == "__main__":
    if "--tsan" in sys.argv:
        build_tsan(verbose=True)
    else:
        build_library(force="--force" in sys.argv, verbose=True)
