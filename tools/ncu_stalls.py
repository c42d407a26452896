"""Per-instruction warp-stall summary of one kernel from an `ncu --set full --import-source on` report.

  python tools/ncu_stalls.py <report.ncu-rep> [--exec-count N] [--top K]

Reads `ncu -i <rep> --page source --csv --print-source sass` (needs ncu on PATH; works without a GPU) and prints
  * the stall-reason totals of the whole kernel,
  * the K instructions with the most samples (address, SASS, samples, executions, top stall reasons),
  * for the hot loop (the contiguous SASS range whose instructions were executed exactly N times; default: the execution
    count shared by most MUFU/FFMA instructions): instructions per iteration, stall shares and the opcode mix.
This is how profiles/r01e_attention_stalls.md was produced."""
import argparse
import csv
import io
import subprocess
from collections import Counter


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--exec-count", type=float, default=None)
    ap.add_argument("--top", type=int, default=25)
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[start - 1][1] if start else "")
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[start + 1:] if len(r) == len(hdr)]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(f(r[ix["# Samples"]]) for r in data)
    agg = {s: sum(f(r[ix[s]]) for r in data) for s in stalls}
    print(f"{len(data)} SASS instructions, {int(tot)} samples")
    print("kernel stall shares %:", {k[6:]: round(v / tot * 100, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])
                                     if v / tot > 0.01})
    print(f"\ntop {a.top} instructions by samples:")
    for r in sorted(data, key=lambda r: -f(r[ix["# Samples"]]))[:a.top]:
        st = sorted(((s[6:], int(f(r[ix[s]]))) for s in stalls if f(r[ix[s]]) > 0), key=lambda kv: -kv[1])[:3]
        print(f"  {r[ix['Address']][-5:]} {r[ix['Source']][:60]:60s} {int(f(r[ix['# Samples']])):6d} "
              f"{int(f(r[ix['Instructions Executed']])):9d} {dict(st)}")
    ex = [f(r[ix["Instructions Executed"]]) for r in data]
    if a.exec_count is None:
        hot = [ex[i] for i, r in enumerate(data) if any(op in r[ix["Source"]] for op in ("MUFU", "FFMA", "HMMA", "UTCHMMA"))]
        a.exec_count = Counter(hot).most_common(1)[0][0] if hot else max(ex)
    idx = [i for i in range(len(data)) if ex[i] == a.exec_count]
    lo, hi = min(idx), max(idx)
    reg = data[lo:hi + 1]
    rt = sum(f(r[ix["# Samples"]]) for r in reg)
    ragg = {s: sum(f(r[ix[s]]) for r in reg) for s in stalls}
    ops = Counter()
    for r in reg:
        if f(r[ix["Instructions Executed"]]) >= 0.9 * a.exec_count:
            ops[[o for o in r[ix["Source"]].split() if not o.startswith("@")][0].split(".")[0]] += 1
    print(f"\nhot loop (executed {int(a.exec_count)} times per instruction): {hi - lo + 1} static instructions, "
          f"{sum(f(r[ix['Instructions Executed']]) for r in reg) / a.exec_count:.0f} executed per iteration, "
          f"{int(rt)} samples ({rt / tot * 100:.0f} % of the kernel's)")
    print("  stall shares %:", {k[6:]: round(v / rt * 100, 1) for k, v in sorted(ragg.items(), key=lambda kv: -kv[1])
                                if rt and v / rt > 0.01})
    print("  opcode mix:", dict(ops.most_common(16)))


if __name__ == "__main__":
    main()
