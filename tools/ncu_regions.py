#!/usr/bin/env python
"""Where do a kernel's warp-level instructions go?  Reads `ncu -i REP --page source --csv` and groups
consecutive SASS lines with similar execution counts into regions (loops), printing each region's share
of executed instructions, its packed-FP32x2 share and its stall samples.
    python tools/ncu_regions.py gpurun_out/<name>.ncu-rep [top_n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[start - 1][1] if start else "")
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    body = rows[start + 1:]

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, IndexError):
            return 0.0

    ex = [f(r, "Instructions Executed") for r in body]
    smp = [f(r, "# Samples") for r in body]
    ops = []
    for r in body:
        src = r[ix["Source"]].split()
        op = src[0] if src else ""
        if op.startswith("@") and len(src) > 1:
            op = src[1]
        ops.append(op.rstrip(";"))
    tot, tots = sum(ex), sum(smp)
    packed = ("FFMA2", "FMUL2", "FADD2")
    regions, a = [], 0
    for i in range(1, len(body) + 1):
        if i == len(body) or not (0.7 <= (ex[i] + 1) / (ex[a] + 1) <= 1.4):
            regions.append((a, i))
            a = i
    regions.sort(key=lambda t: -sum(ex[t[0]:t[1]]))
    print(f"total warp instructions {tot:.4g}, samples {tots:.0f}")
    for a, b in regions[:top]:
        sm = sum(ex[a:b])
        pk = sum(ex[i] for i in range(a, b) if ops[i] in packed)
        mov = sum(ex[i] for i in range(a, b) if ops[i] == "MOV")
        print(f"  SASS lines {a:5d}-{b:5d} ({b - a:4d} instr) exec/line {ex[a]:9.3g}  share {100 * sm / tot:5.2f}%  "
              f"samples {100 * sum(smp[a:b]) / max(tots, 1):5.2f}%  packed {100 * pk / max(sm, 1):3.0f}%  MOV {100 * mov / max(sm, 1):3.0f}%")
    pk = sum(e for e, o in zip(ex, ops) if o in packed)
    mu = sum(e for e, o in zip(ex, ops) if o.startswith("MUFU"))
    print(f"packed {pk:.4g}  MUFU {mu:.4g}  other {tot - pk - mu:.4g}  -> packed share of issue cycles (2 per packed) "
          f"{2 * pk / (2 * pk + tot - pk):.3f}")
    by = {}
    for e, o in zip(ex, ops):
        by[o] = by.get(o, 0) + e
    print("  " + "  ".join(f"{o} {100 * e / tot:.1f}%" for o, e in sorted(by.items(), key=lambda kv: -kv[1])[:14]))


if __name__ == "__main__":
    main()
