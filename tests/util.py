"""Shared helpers for the parity tests (deterministic synthetic inputs, golden loaders)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAX_HASH_1000 = 18446744073709552  # round(2**64 / 1000), SURVEY 8(d)


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def splitmix64(seed, n):
    """n SplitMix64 outputs (vectorised), the generator named in SURVEY 8(d)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def random_dna(n, seed):
    """n random bases, 2 bits per base from SplitMix64 -> b'ACGT'[bits]."""
    words = splitmix64(seed, (n + 31) // 32)
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]
    codes = ((words[:, None] >> shifts) & np.uint64(3)).astype(np.uint8).reshape(-1)[:n]
    return np.frombuffer(b"ACGT", dtype=np.uint8)[codes].tobytes()


def mutate(seq: bytes, rate, seed):
    """Independent substitutions (uniform over the 3 other bases) with probability rate."""
    a = np.frombuffer(seq, dtype=np.uint8).copy()
    r = splitmix64(seed, len(a))
    hit = (r >> np.uint64(11)).astype(np.float64) / float(1 << 53) < rate
    lut = np.zeros(256, dtype=np.uint8)
    lut[list(b"ACGT")] = [0, 1, 2, 3]
    codes = lut[a]
    shift = ((r & np.uint64(0xFFFF)) % np.uint64(3)).astype(np.uint8) + 1
    codes = np.where(hit, (codes + shift) % 4, codes)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[codes].tobytes()


def revcomp(seq: bytes) -> bytes:
    return seq.translate(bytes.maketrans(b"ACGTacgt", b"TGCAtgca"))[::-1]


def make_reads(genome: bytes, nreads, readlen, seed):
    """Error-free reads: uniform start, strand flipped w.p. 0.5."""
    r = splitmix64(seed, 2 * nreads)
    starts = (r[:nreads] % np.uint64(len(genome) - readlen)).astype(np.int64)
    flips = (r[nreads:] & np.uint64(1)).astype(bool)
    out = bytearray()
    for s, f in zip(starts, flips):
        rd = genome[s:s + readlen]
        out += revcomp(rd) if f else rd
    return bytes(out)


def sbt_v5_tree():
    """The reference's tests/data/v5.sbt.json tree (golden copy): (d, {pos: (oxli_bytes, min_n_below)},
    leaf positions, the raw golden dict)."""
    import base64
    import zlib
    t = golden("sbt_v5_tree.json")
    nodes = {int(p): (zlib.decompress(base64.b64decode(n["oxli_zlib_b64"])), n["min_n_below"]) for p, n in t["nodes"].items()}
    return t["d"], nodes, t["leaf_positions"], t
