"""Turns ncu output into the small text/JSON summaries kept under profiles/.
  launch list:  python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rN_launch_list_bench.txt
  one kernel:   python tools/ncu_summary.py kernel gpurun_out/x.ncu-rep "<source command>" > profiles/rN_x_ncu.json
"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max"]


def short(name):
    name = name.replace("void ", "").replace("orbit::", "")
    return name.split("(")[0] if "<" not in name else name[:name.index(">") + 1].replace("(int)", "").replace("(bool)", "")


def launches(path, last=0):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
    hdr = rows[hdr_i]
    ki, mi, vi, ui, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    by_id = {}
    for r in rows[hdr_i + 1:]:
        if len(r) != len(hdr):
            continue
        d = by_id.setdefault(int(r[ii]), {"kernel": short(r[ki])})
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        if r[mi] == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if unit in ("ns", "nsecond") else v * (1e3 if unit in ("ms", "msecond") else 1.0)
        else:
            scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
            d[r[mi]] = v * scale
    print("# id  kernel  time_us  dram_read_MB  dram_write_MB")
    for i in sorted(by_id):
        d = by_id[i]
        print(i, d["kernel"], round(d.get("us", 0.0), 2), round(d.get("dram__bytes_read.sum", 0.0), 3), round(d.get("dram__bytes_write.sum", 0.0), 3))
    tot = {}
    ids = sorted(by_id)[-last:] if last else sorted(by_id)
    if last:
        print("# the last %d launches = the timed region of `bench.py --step-only` (%d steps x 7 kernels)" % (last, last // 7))
    for d in (by_id[i] for i in ids):
        t = tot.setdefault(d["kernel"], [0, 0.0]); t[0] += 1; t[1] += d.get("us", 0.0)
    all_us = sum(t[1] for t in tot.values())
    print("# share of profiled GPU time per kernel (cold-cache, serialised launches)")
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("# %-50s launches %4d  avg %8.2f us  share %5.1f %%" % (k, n, us / n, 100 * us / all_us))


def kernel(path, source):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = {"source": source, "launches": []}
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEEP:
            if k in hdr:
                d[k] = [float(r[hdr.index(k)].replace(",", "")), units[hdr.index(k)]]
        out["launches"].append(d)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    first = out["launches"][0]
    out["dram_traffic_bytes_per_launch"] = int(sum(first[k][0] * scale[first[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    else:
        kernel(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
