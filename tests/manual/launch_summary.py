"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel share table (markdown)."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ci = {n: i for i, n in enumerate(h)}
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ci["Kernel Name"]].replace("smb200::", "").split("(")[0].replace("void ", "").replace("(int)", "")
    if "<" in r[ci["Kernel Name"]] and "<" not in name:
        name = r[ci["Kernel Name"]].replace("smb200::", "").replace("void ", "").replace("(int)", "")
        name = name[:name.index(">") + 1]
    v = float(r[ci["Metric Value"]].replace(",", ""))
    u = r[ci["Metric Unit"]]
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f%% |" % (name, n, ms, 100 * ms / tot))
print("| **all** | %d | %.3f | 100%% |" % (sum(a[0] for a in agg.values()), tot))
