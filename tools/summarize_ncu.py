"""Summarises an .ncu-rep (raw page) and a launch-list csv into small text files for profiles/."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "l1tex__t_bytes.sum", "lts__t_bytes.sum"]


def summarize_rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(f"kernel: {name}\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"  {k} = {r[i]} {units[i]}\n")
            for i, h in enumerate(hdr):
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                    f.write(f"  {h} = {r[i]}\n")
            f.write("\n")


def summarize_launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# per-kernel device time over the captured launches (ncu, cold cache, serialised); total {tot / 1e3:.1f} us\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t / 1e3:10.1f} us  {100 * t / tot:5.1f}%  n={n:4d}  avg {t / n / 1e3:8.1f} us  {name}\n")


if __name__ == "__main__":
    kind, src, dst = sys.argv[1:4]
    (summarize_rep if kind == "rep" else summarize_launches)(src, dst)
