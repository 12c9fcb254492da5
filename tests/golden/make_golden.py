"""Regenerates tests/golden/*.json from the reference's own test fixtures.

Run in the build container only (reads /root/reference/tests/data, which does
not exist on the GPU box):  python tests/golden/make_golden.py

Outputs (committed):
  sbt_v5_leaves.json   the 7 leaf sketches of tests/data/v5.sbt.json (num=500,k=31),
                       keyed by tree position 6..12, + the hit counts asserted by
                       src/index/sbt.rs:543-588 (load_sbt test)
  genome_s10_s11.json  the 4 sketches of tests/data/genome-s10+s11.sig with their
                       stored md5sum (tests/signature.rs:10-32) and signature metadata
  subset_scaled.json   10 of the 100 scaled (max_hash=9223372036854776, k=21, with
                       abundances) leaves of tests/data/.sbt.subset, as stored (UNSORTED)
  sbt_v5_tree.json     the tree of tests/data/v5.sbt.json: d, factory, internal nodes with their
                       khmer Nodegraph files tests/data/.sbt.v5/internal.N (zlib + base64 of the
                       OXLI v4 bytes), leaf positions, the table sizes and the hash list that
                       src/index/nodegraph.rs:292-821 (load_nodegraph) asserts present / absent, and
                       the hit counts src/index/sbt.rs:543-551 asserts for sbt.find
"""
import base64
import re
import zlib
import json
import os

REF = "/root/reference/tests/data"
OUT = os.path.dirname(os.path.abspath(__file__))


def sketch(s):
    d = {k: s[k] for k in ("num", "ksize", "seed", "max_hash", "md5sum", "molecule", "mins")}
    if "abundances" in s:
        d["abundances"] = s["abundances"]
    return d


def main():
    sbt = json.load(open(os.path.join(REF, "v5.sbt.json")))
    leaves = {}
    for pos, leaf in sorted(sbt["leaves"].items(), key=lambda kv: int(kv[0])):
        sig = json.load(open(os.path.join(REF, ".sbt.v5", leaf["filename"])))[0]
        leaves[pos] = {"filename": leaf["filename"], "sig_name": sig.get("name"),
                       "sketch": sketch(sig["signatures"][0])}
    golden = {
        "source": "tests/data/v5.sbt.json + tests/data/.sbt.v5/*",
        "query_position": "7",
        # src/index/sbt.rs:543-588: (search fn, threshold) -> number of hits
        "asserted_hits": {"similarity@0.5": 1, "similarity@0.1": 2, "containment@0.5": 2, "containment@0.1": 4},
        "leaves": leaves,
    }
    json.dump(golden, open(os.path.join(OUT, "sbt_v5_leaves.json"), "w"), separators=(",", ":"))

    sigs = json.load(open(os.path.join(REF, "genome-s10+s11.sig")))
    s = sigs[0]
    json.dump({
        "source": "tests/data/genome-s10+s11.sig",
        "n_signatures": len(sigs),
        "class": s["class"], "email": s["email"], "filename": s.get("filename"), "name": s.get("name"),
        "hash_function": s["hash_function"],
        "sketches": [sketch(x) for x in s["signatures"]],
    }, open(os.path.join(OUT, "genome_s10_s11.json"), "w"), separators=(",", ":"))

    sub = json.load(open(os.path.join(REF, "subset.sbt.json")))
    picks = sorted(sub["leaves"].items(), key=lambda kv: int(kv[0]))[:10]
    out = []
    for pos, leaf in picks:
        sig = json.load(open(os.path.join(REF, ".sbt.subset", leaf["filename"])))[0]
        out.append({"position": pos, "filename": leaf["filename"], "sketch": sketch(sig["signatures"][0])})
    json.dump({"source": "tests/data/subset.sbt.json + .sbt.subset/* (first 10 leaves, stored order)",
               "leaves": out}, open(os.path.join(OUT, "subset_scaled.json"), "w"), separators=(",", ":"))


def tree():
    sbt = json.load(open(os.path.join(REF, "v5.sbt.json")))
    nodes = {}
    for pos, nd in sorted(sbt["nodes"].items(), key=lambda kv: int(kv[0])):
        raw = open(os.path.join(REF, ".sbt.v5", nd["filename"]), "rb").read()
        nodes[pos] = {"filename": nd["filename"], "min_n_below": nd["metadata"]["min_n_below"], "n_bytes": len(raw),
                      "oxli_zlib_b64": base64.b64encode(zlib.compress(raw, 9)).decode()}
    # the hashes asserted by load_nodegraph (src/index/nodegraph.rs:292-821) on tests/data/internal.0
    src = open("/root/reference/src/index/nodegraph.rs").read()
    body = src[src.index("fn load_nodegraph()"):]
    absent = int(re.search(r"ng\.get\((\d+)\), 0\)", body).group(1))
    lst = body[body.index("for h in ["):body.index(".iter()")]
    present = [int(x) for x in re.findall(r"\d+", lst)]
    json.dump({
        "source": "tests/data/v5.sbt.json + tests/data/.sbt.v5/internal.* + src/index/nodegraph.rs:292-821",
        "d": sbt["d"], "factory": sbt["factory"],
        "nodes": nodes,
        "leaf_positions": sorted(int(p) for p in sbt["leaves"]),
        "load_nodegraph": {"file": "internal.0", "tablesizes": [99991, 99989, 99971, 99961], "absent": [absent],
                           "present": present},
        # update_nodegraph (nodegraph.rs:271-290): new([99991,99989,99971,99961]) | internal.1 | internal.2 == internal.0
        "update_nodegraph": {"parent": "0", "children": ["1", "2"]},
        # load_sbt (sbt.rs:543-551): sbt.find(search_minhashes, leaf 7, threshold) -> number of hits
        "asserted_sbt_find": {"query_position": 7, "similarity@0.5": 1, "similarity@0.1": 2},
    }, open(os.path.join(OUT, "sbt_v5_tree.json"), "w"), separators=(",", ":"))


if __name__ == "__main__":
    tree()
    main()
