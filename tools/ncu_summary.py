"""Turns ncu output into the small JSON summaries committed under profiles/.

  python tools/ncu_summary.py launches <launches.csv> <out.json> [note]
      per-kernel launch counts / total device time / share of the timed region from the
      `ncu --metrics gpu__time_duration.sum --csv` launch list
  python tools/ncu_summary.py full <out.json> <report.ncu-rep> [...]
      selected `--set full` metrics of every profiled launch (needs ncu on PATH to read the reports)
"""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = [
    "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpc__cycles_elapsed.max",
    "gpc__cycles_elapsed.avg.per_second",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").strip()


def launches(path, out, note=""):
    lines = [l for l in open(path, errors="replace") if l.startswith('"')]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        if len(r) <= vi or not r[vi]:
            continue
        t = float(r[vi].replace(",", ""))
        unit = r[ui]
        ms = t / 1e6 if unit in ("ns", "nsecond") else (t / 1e3 if unit in ("us", "usecond") else t)
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(v[1] for v in agg.values())
    ks = [dict(kernel=k, launches=v[0], total_ms=round(v[1], 4), share_pct=round(100 * v[1] / total, 2))
          for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    json.dump(dict(note=note, launches=sum(v[0] for v in agg.values()), total_ms=round(total, 3), kernels=ks),
              open(out, "w"), indent=1)
    print(f"{out}: {len(ks)} kernels, {total:.2f} ms")


def full(out, reports):
    res = []
    for rep in reports:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = {}
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    u = units[i]
                    d[k + (f" [{u}]" if u else "")] = r[i]
            d["Kernel Name"] = short(d.get("Kernel Name", ""))
            d["report"] = rep.split("/")[-1]
            res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    print(f"{out}: {len(res)} launches")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3:])
