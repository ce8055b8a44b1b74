"""Turn ncu output into the text summaries kept under profiles/.
  python ncu_summary.py launches gpurun_out/launches.csv           per-kernel launch count / total / max / share
  python ncu_summary.py full gpurun_out/full.ncu-rep [kernel-regex] selected --set full metrics of every captured launch
Runs here (the container has ncu for reading reports; it needs no GPU)."""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def launches(path):
    rows = []
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    for r in csv.DictReader(io.StringIO(text[start:])):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(unit, 1e-6)
            rows.append((re.sub(r"\(.*", "", r["Kernel Name"]), v))
    agg = collections.OrderedDict()
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    total = sum(a[1] for a in agg.values())
    print(f"# {len(rows)} launches, {total:.3f} ms of kernel time (ncu serialises launches and runs them cache-cold: shares, not absolutes)")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:72]:72s} n={a[0]:6d} total={a[1]:9.3f} ms max={a[2]:8.3f} ms share={100 * a[1] / total:5.1f}%")


def full(path, pattern="."):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    start = out.find('"ID"')
    rd = list(csv.reader(io.StringIO(out[start:])))
    head, units = rd[0], rd[1]
    for row in rd[2:]:
        rec = dict(zip(head, row))
        name = rec.get("Kernel Name", "")
        if not re.search(pattern, name):
            continue
        print(f"== {name[:110]}  grid {rec.get('Grid Size')} block {rec.get('Block Size')}")
        for mname in METRICS:
            if mname in rec:
                print(f"   {mname:72s} {rec[mname]:>18s} {units[head.index(mname)]}")
        try:
            rd_b = float(rec["dram__bytes_read.sum"].replace(",", ""))
            wr_b = float(rec["dram__bytes_write.sum"].replace(",", ""))
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            u_r = mult.get(units[head.index("dram__bytes_read.sum")], 1.0)
            u_w = mult.get(units[head.index("dram__bytes_write.sum")], 1.0)
            dur = float(rec["gpu__time_duration.sum"].replace(",", ""))
            u_t = {"nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3}.get(
                units[head.index("gpu__time_duration.sum")], 1e-9)
            tot = rd_b * u_r + wr_b * u_w
            print(f"   -> DRAM traffic {tot / 1e9:.3f} GB in {dur * u_t * 1e3:.3f} ms = {tot / (dur * u_t) / 1e9:.0f} GB/s")
        except (KeyError, ValueError):
            pass


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ".")
