"""Builds a kernel-experiment variant of the library: `python tests/manual/build_variant.py NAME -DX=1 ...`
recompiles csrc/sketch.cu with the extra flags and links it with the regular objects into
sourmash_rust_b200/build/libsourmash_NAME.so (load it with SMB200_LIB=...)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sourmash_rust_b200 import build as b
name, extra = sys.argv[1], sys.argv[2:]
src = "sketch.cu"
if "--src" in extra:  # which translation unit takes the extra flags (default: sketch.cu)
    i = extra.index("--src")
    src = extra[i + 1]
    del extra[i:i + 2]
b.build_library()
obj = os.path.join(b.OBJ, "%s_%s.o" % (os.path.splitext(src)[0], name))
subprocess.run([b.NVCC] + b.FLAGS + extra + ["-c", os.path.join(b.CSRC, src), "-o", obj], check=True)
objs = [os.path.join(b.OBJ, os.path.splitext(s)[0] + ".o") for s in b.SOURCES if s != src] + [obj]
out = os.path.join(b.OBJ, "libsourmash_%s.so" % name)
subprocess.run([b.NVCC, "-shared", "-o", out] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt", "-lz"], check=True)
print(out)
