"""CPU-side checks of the product library (no GPU needed): it builds, loads, exports every
symbol declared in include/*.h, fails loudly without a device, and its host+device bit helpers
agree with a byte-wise restatement of the reference loop body."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def smb():
    from sourmash_rust_b200 import build
    build.build_library()
    import sourmash_rust_b200 as s
    s.lib()
    return s


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return set(re.findall(r"\b([a-z][a-z0-9_]*)\s*\(", txt)) - {"defined", "extern"}


def test_exports_every_declared_symbol(smb):
    L = C.CDLL(smb.LIB_PATH)
    ref = _declared("sourmash.h")
    ext = _declared("sourmash_b200.h")
    assert len(ref) == 48, sorted(ref)  # 48 extern "C" symbols in the reference
    for name in sorted(ref | ext):
        assert hasattr(L, name), name
    assert ref == set(smb.REFERENCE_ABI)
    assert ext == set(smb.EXTENSION_ABI)


def test_scalar_calls_need_no_gpu(smb):
    assert smb.hash_murmur(b"ACG", 42) == 1731421407650554201  # tests/test.rs:3-6
    mh = smb.KmerMinHash(500, 31, False, 42, 0, True)
    assert (mh.num, mh.ksize, mh.seed, mh.max_hash, mh.is_protein) == (500, 31, 42, 0, False)
    assert mh.track_abundance()


def test_error_slot_protocol(smb):
    L = smb.lib()
    L.sourmash_err_clear()
    assert L.sourmash_err_get_last_code() == 0
    s = L.sourmash_err_get_last_message()
    assert s.len == 0
    # NULL handle: the reference asserts (panic); here it is a recorded Panic error and a zero return
    assert L.kmerminhash_get_mins_size(None) == 0
    assert L.sourmash_err_get_last_code() == 1
    L.sourmash_err_clear()
    s = L.sourmash_str_from_cstr(b"hello")
    assert C.string_at(s.data, s.len) == b"hello" and s.owned
    L.sourmash_str_free(C.byref(s))
    assert not s.owned and s.len == 0


def test_no_cpu_fallback(smb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    mh = smb.KmerMinHash(20, 10)
    with pytest.raises(smb.SourmashError) as e:
        mh.add_sequence(b"TGCCGCCCAGCA")
    assert e.value.code == 2 and "no CPU path" in e.value.message


def test_host_bit_helpers(tmp_path):
    exe = str(tmp_path / "tile_views_test")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "tile_views_test.cpp")])
    out = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0, out.stdout


def _signature_text():
    import json
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import golden
    g = golden("genome_s10_s11.json")
    doc = [{"class": g["class"], "email": g["email"], "hash_function": g["hash_function"], "filename": g["filename"],
            "name": g["name"], "license": "CC0", "signatures": g["sketches"], "version": 0.4}]
    return json.dumps(doc).encode()


def test_load_path_sniffs_compression(smb, tmp_path):
    """signatures_load_path goes through get_input (reference src/ffi.rs:557, src/file.rs:47-77): gzip is inflated
    (first member only, as flate2's GzDecoder does), a plain file is read as is, and both parse to the same thing."""
    import gzip
    txt = _signature_text()
    plain, packed, twice, cut = (str(tmp_path / n) for n in ("a.sig", "a.sig.gz", "two.sig.gz", "cut.sig.gz"))
    open(plain, "wb").write(txt)
    open(packed, "wb").write(gzip.compress(txt, 9))
    open(twice, "wb").write(gzip.compress(txt, 1) + gzip.compress(b"trailing member, never parsed"))
    open(cut, "wb").write(gzip.compress(txt, 9)[:-40])
    want = [s.save_json() for s in smb.signatures_load_buffer(txt)]
    assert len(want) == 4
    for path in (plain, packed, twice):
        got = smb.signatures_load_path(path)
        assert [s.save_json() for s in got] == want, path
    # the filters still apply after inflation
    assert len(smb.signatures_load_path(packed, ksize=21)) == sum(1 for s in smb.signatures_load_buffer(txt, ksize=21))
    with pytest.raises(smb.SourmashError) as e:
        smb.signatures_load_path(cut)
    assert e.value.code == 4               # Unknown: an io error under serde_json::from_reader (errors.rs:54-77)
    with pytest.raises(smb.SourmashError) as e:
        smb.signatures_load_path(str(tmp_path / "missing.sig"))
    assert e.value.code == 1               # Panic: file.rs:83 expect()
    for magic in (b"BZh91AY&SY", b"\xfd7zXZ\x00\x00"):
        p = str(tmp_path / "other.bin")
        open(p, "wb").write(magic + b"\0" * 32)
        with pytest.raises(smb.SourmashError) as e:
            smb.signatures_load_path(p)
        assert e.value.code == 2 and "not supported" in e.value.message


def test_load_path_reads_stdin(smb, tmp_path):
    """`-` names standard input (file.rs:49-51)."""
    txt = _signature_text()
    code = ("import sys; sys.path.insert(0, %r); import sourmash_rust_b200 as s; "
            "print(len(s.signatures_load_path('-', ksize=31)))" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], input=txt, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert r.returncode == 0, r.stderr.decode()
    assert int(r.stdout.split()[-1]) == len(smb.signatures_load_buffer(txt, ksize=31))


def test_collection_push_signatures_host_side(smb):
    """Rows are staged on the host until first use, so the row count and the compatibility checks of
    smgpu_collection_push_signatures need no GPU (check order of lib.rs:176-190; empty signature = the reference's panic)."""
    import json
    sk = {"num": 3, "ksize": 21, "seed": 42, "max_hash": 0, "mins": [1, 5, 9], "md5sum": "", "molecule": "DNA"}
    def sigs(*sketch_lists):
        return smb.signatures_load_buffer(json.dumps([{"hash_function": "h", "signatures": [dict(sk, **kw) for kw in sl]} for sl in sketch_lists]).encode())
    ok = sigs([{}], [{"mins": [2, 3]}], [{"mins": []}])
    assert len(smb.SketchCollection.from_signatures(ok)) == 3
    assert len(smb.SketchCollection.from_signatures([])) == 0
    for kw, code in (({"ksize": 31}, 101), ({"molecule": "protein"}, 102), ({"max_hash": 7, "num": 0}, 103), ({"seed": 1}, 104)):
        with pytest.raises(smb.SourmashError) as e:
            smb.SketchCollection.from_signatures(sigs([{}], [kw]))
        assert e.value.code == code
    with pytest.raises(smb.SourmashError) as e:      # unsorted mins cannot be a row
        smb.SketchCollection.from_signatures(sigs([{"mins": [5, 1]}]))
    assert e.value.code == 2
    empty = smb.Signature()
    with pytest.raises(smb.SourmashError) as e:
        smb.SketchCollection.from_signatures([empty])
    assert e.value.code == 1


def test_host_hash_matches_oracle_over_lengths_and_seeds(smb):
    """hash_murmur is a scalar call answered on the host (lib.rs:33-35): every tail length 0..15 with 0..4 body
    blocks, several seeds (NUL-free bytes: the ABI takes a C string, ffi.rs:16-24)."""
    import random
    sys.path.insert(0, ROOT)
    from oracle import oracle as orc
    rng = random.Random(99)
    for n in list(range(0, 81)) + [127, 128, 129, 1000]:
        for seed in (42, 0, 1, 0xFFFFFFFF, 123456789):
            word = bytes(rng.randrange(1, 256) for _ in range(n))
            assert smb.hash_murmur(word, seed) == orc.hash_murmur(word, seed), (n, seed)
