"""summarise an `ncu --set full` report: one line per profiled launch with duration, DRAM bytes, achieved DRAM GB/s vs the
measured HBM peak, tensor-pipe activity, L2 / SM throughput, registers, grid.   usage: ncu_summary.py file.ncu-rep [peak_gbs]"""
import csv, io, json, os, re, subprocess, sys
rep = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
if peak is None:
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6553.6
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
def col(name):
    for i, h in enumerate(hdr):
        if h == name:
            return i
    return None
want = {"t": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "tensor": "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "tensor2": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l2": "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "regs": "launch__registers_per_thread", "warps": "sm__warps_active.avg.pct_of_peak_sustained_active"}
idx = {k: col(v) for k, v in want.items()}
kn, grid = col("Kernel Name"), col("Grid Size")
units = rows[1]
def val(r, k):
    i = idx[k]
    if i is None or r[i] in ("", "n/a"):
        return None
    v = float(r[i].replace(",", ""))
    u = units[i]
    if k == "t":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)      # -> us
    if k in ("rd", "wr"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    return v
print("ncu --set full --clock-control none; HBM peak %.1f GB/s (MEASURED_PEAKS.json); per launch" % peak)
print("%-46s %9s %9s %9s %8s %7s %8s %6s %6s %5s %s" % ("kernel", "time_us", "rd_MB", "wr_MB", "GB/s", "ofpeak", "tensor%", "L2%", "SM%", "regs", "grid"))
for r in rows[2:]:
    if len(r) <= kn:
        continue
    name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("dn::", "")
    t, rd, wr = val(r, "t"), val(r, "rd") or 0.0, val(r, "wr") or 0.0
    gbs = (rd + wr) / (t * 1e-6) / 1e9 if t else 0.0
    tp = val(r, "tensor")
    if tp is None:
        tp = val(r, "tensor2")
    print("%-46s %9.1f %9.1f %9.1f %8.0f %7.2f %8s %6s %6s %5s %s" % (
        name[:46], t or 0, rd / 1e6, wr / 1e6, gbs, gbs / peak, "%.1f" % tp if tp is not None else "-",
        "%.0f" % val(r, "l2") if val(r, "l2") is not None else "-", "%.0f" % val(r, "sm") if val(r, "sm") is not None else "-",
        "%d" % val(r, "regs") if val(r, "regs") is not None else "-", r[grid] if grid is not None else ""))
