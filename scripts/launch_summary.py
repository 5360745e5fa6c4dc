"""summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
per-kernel time (and DRAM traffic) of ONE training step (between the last two image-upload kernels)"""
import collections, csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]; data = rows[hdr + 1:]
ki, vi, mi, ii = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name"), H.index("ID")
launch = collections.OrderedDict()
for r in data:
    d = launch.setdefault(r[ii], {"name": r[ki]})
    d[r[mi]] = float(r[vi].replace(",", ""))
L = list(launch.values())
marks = [i for i, d in enumerate(L) if "nchw_to_padded" in d["name"] or "nchw_to_nhwc" in d["name"]]
a, b = marks[-2], marks[-1]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in L[a:b]:
    n = re.sub(r"\(.*", "", d["name"]); n = re.sub(r"^void ", "", n)
    agg[n][0] += 1
    agg[n][1] += d.get("gpu__time_duration.sum", 0.0)
    agg[n][2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(v[1] for v in agg.values())
print("launches in one step: %d, summed kernel time %.2f ms (cold-cache, serialised)" % (b - a, tot / 1e6))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for n, (c, t, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%9.1f us %5.1f%% x%-4d dram %8.1f MB  %s" % (t / 1e3, 100 * t / tot, c, by / 1e6, n[:90]))
if len(sys.argv) > 3:
    conv = {n: {"launches": c, "time_us": t / 1e3, "dram_bytes": by} for n, (c, t, by) in agg.items() if "conv_" in n}
    json.dump({"source": sys.argv[1], "per_step": conv,
               "conv_dram_bytes_per_launch": sum(v["dram_bytes"] for v in conv.values()) /
               max(1, sum(v["launches"] for v in conv.values()))}, open(sys.argv[3], "w"), indent=1)
