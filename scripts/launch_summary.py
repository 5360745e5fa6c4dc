"""summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of ONE training step"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]; data = rows[hdr + 1:]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
names = [r[ki] for r in data]
marks = [i for i, n in enumerate(names) if "nchw_to_padded" in n or "nchw_to_nhwc" in n]
a, b = marks[-2], marks[-1]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data[a:b]:
    n = re.sub(r"\(.*", "", r[ki]); n = re.sub(r"^void ", "", n)
    agg[n][0] += 1; agg[n][1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print("launches in one step: %d, summed kernel time %.2f ms (cold-cache, serialised)" % (b - a, tot / 1e6))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%9.1f us %5.1f%% x%-4d %s" % (t / 1e3, 100 * t / tot, c, n[:100]))
