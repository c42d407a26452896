"""Static issue-time estimate of a SASS region from the scheduling control bits ptxas encodes in every instruction.

  python tools/sass_stalls.py <object-or-library> <mangled-kernel-name-substring> [--from PATTERN --to PATTERN]

sm_70+ instructions are 128 bits; the upper word carries the control field: bits 105-108 = stall count (cycles the warp
waits before ITS next instruction may issue), bit 109 = yield, bits 110-112 / 113-115 = scoreboard set by this
instruction (write / read), bits 116-121 = scoreboards waited on.  Summing the stall counts over a loop body gives the
minimum number of cycles ONE warp needs per iteration even if every variable-latency result (MUFU, TMEM, shared memory)
were ready in time - i.e. `selected` + `wait` of ncu's warp-state sampling - and can be evaluated here, without a GPU,
to compare code-generation variants (register budgets, unrolling, instruction mixes).  Works on cuobjdump -sass text.
The region is the first match of --from up to the first later match of --to (regular expressions on the SASS text)."""
import argparse
import re
import subprocess
from collections import Counter


def parse(obj, kernel):
    names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out, cur, keep = [], None, False
    lines = names.splitlines()
    insts = []
    for i, l in enumerate(lines):
        m = re.match(r"\s*Function : (\S+)", l)
        if m:
            keep = kernel in m.group(1)
            continue
        if not keep:
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", l)
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s*/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ctrl = hi >> 41
                insts.append(dict(addr=int(m.group(1), 16), text=m.group(2).strip(), stall=ctrl & 0xF,
                                  yield_=(ctrl >> 4) & 1, wr=(ctrl >> 5) & 7, rd=(ctrl >> 8) & 7, wait=(ctrl >> 11) & 0x3F))
    return insts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("kernel")
    ap.add_argument("--from", dest="lo", default=None)
    ap.add_argument("--to", dest="hi", default=None)
    a = ap.parse_args()
    insts = parse(a.obj, a.kernel)
    if not insts:
        raise SystemExit("kernel not found")
    lo = 0
    if a.lo:
        lo = next(i for i, x in enumerate(insts) if re.search(a.lo, x["text"]))
    hi = len(insts) - 1
    if a.hi:
        hi = next(i for i in range(lo + 1, len(insts)) if re.search(a.hi, insts[i]["text"]))
    reg = insts[lo:hi + 1]
    ops = Counter()
    by_op = Counter()
    for x in reg:
        op = [o for o in x["text"].split() if not o.startswith("@")][0].split(".")[0]
        ops[op] += 1
        by_op[op] += x["stall"]
    total = sum(x["stall"] for x in reg)
    print(f"region {insts[lo]['addr']:#x}..{insts[hi]['addr']:#x}: {len(reg)} instructions, sum of stall counts {total} cycles "
          f"({total / len(reg):.2f} per instruction), {sum(1 for x in reg if x['wait'])} instructions wait on a scoreboard")
    print("  stall cycles by opcode:", dict(by_op.most_common(12)))
    print("  opcode mix:", dict(ops.most_common(14)))


if __name__ == "__main__":
    main()
