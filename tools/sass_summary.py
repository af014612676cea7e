#!/usr/bin/env python
"""Static evidence from the built library (no GPU needed): per hot kernel, the ptxas resource line
(registers, spills, shared memory) and the SASS opcode mix of its hottest loop.

    python tools/sass_summary.py > profiles/round1_sass_static.md

Reads pycpet_b200/libcpetb200.so with `cuobjdump -sass` / `cuobjdump -res-usage`.  The "inner loop"
of a kernel is taken as the backward branch whose body has the highest share of packed-FP32
instructions (the innermost unrolled charge loop); the counts are per trip through that loop body."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pycpet_b200", "libcpetb200.so")

HOT = [
    ("k2p_topo_kernel<false, 6>", "K2 streamline integrator, points-packed hybrid form (default for long queues: 3A / MD frames)"),
    ("k2p_topo_kernel<true, 6>", "K2 points-packed, second-difference curvature instantiation"),
    ("k2x_topo_kernel<false, 6, 4>", "K2 hybrid form with charge pairs packed (short queues)"),
    ("k2w_topo_kernel<false, 4, 4>", "K2 round-1 direct form (k2_form=1, kept for A/B)"),
    ("k1x_grid_kernel<false>", "K1 general, hybrid near/far form (field sums over >= 2,048 listed points), raw"),
    ("k1x_grid_kernel<true>", "K1 general, hybrid near/far form, softened (`volume` semantics on the near class)"),
    ("k1_grid_kernel<1, 4, 1>", "K1 general, raw field, 4 points per thread"),
    ("k1_grid_kernel<0, 4, 1>", "K1 general, softened field (`volume` on non-mesh point lists), 4 points per thread"),
    ("k1_grid_kernel<2, 2, 1>", "K1 general, ESP, 2 points per thread (esp101 default)"),
    ("k1_grid_kernel<0, 1, 32>", "K1 general, 32 lanes per point (11^3 meshes, single points)"),
    ("k1_lattice_kernel<0, 5, 4, 0>", "K1 lattice, softened, 5 z-nodes per thread"),
    ("k1_lattice_kernel<1, 5, 4, 0>", "K1 lattice, unsoftened instantiation picked by the softening scan (`volume` 100^3)"),
    ("k1_lattice_kernel<2, 5, 4, 0>", "K1 lattice, ESP, every rsqrt on the MUFU"),
    ("k1_lattice_kernel<2, 6, 4, 1>", "K1 lattice, ESP, 6 z-nodes per thread, one rsqrt in six on the FMA pipe (esp101 default)"),
    ("k3_hist2d_kernel<float, true>", "K3 2-D histogram, shared-memory bins"),
]
PACKED = ("FFMA2", "FADD2", "FMUL2")
SHOW = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "FMNMX", "MUFU.RSQ", "MUFU.RCP", "MUFU.SQRT",
        "LDS.128", "LDS.64", "LDS", "LDG", "STG", "DADD", "DFMA", "DMUL", "SHFL", "ATOMS", "ATOMG", "RED",
        "MATCH", "UBLKCP", "SYNCS", "BAR", "MOV", "IADD3", "BRA"]


def sh(*cmd):
    return subprocess.run(cmd, capture_output=True, text=True, check=True).stdout


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def parse_sass(text):
    """-> {mangled: [(addr, opcode-with-modifiers, operands)]}"""
    funcs, cur = {}, None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)\s*(.*?);", line)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), m.group(2), m.group(3)))
    return funcs


def key(op):
    for k in ("MUFU.RSQ", "MUFU.RCP", "MUFU.SQRT", "LDS.128", "LDS.64"):
        if op.startswith(k):
            return k
    return op.split(".")[0]


def hottest_loop(ins):
    best = None
    addrs = [a for a, _, _ in ins]
    for i, (a, op, args) in enumerate(ins):
        if not op.startswith("BRA"):
            continue
        m = re.search(r"0x([0-9a-f]+)", args)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a or tgt not in addrs:
            continue
        body = ins[addrs.index(tgt):i + 1]
        packed = sum(1 for _, o, _ in body if o.split(".")[0] in PACKED)
        # the innermost unrolled charge loop: dense in packed instructions and, among the dense ones (the kernels carry
        # 1-, 2- and 4-point variants of the same loop), the longest
        dense = packed / len(body) >= 0.7
        score = (dense, packed if dense else packed / len(body), packed)
        if packed and (best is None or score > best[0]):
            best = (score, body)
    return best[1] if best else []


def main():
    sass = sh("cuobjdump", "-sass", LIB)
    res = sh("cuobjdump", "-res-usage", LIB)
    funcs = parse_sass(sass)
    names = demangle(list(funcs))
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = line.strip()
            cur = None
    print("# Static SASS evidence, round 2 (`tools/sass_summary.py`, no GPU needed)\n")
    print("Built by `pycpet_b200/build.py`: `nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo`. Resource lines are")
    print("`cuobjdump -res-usage`; opcode counts are per trip through the hottest loop body (the backward branch whose")
    print("body has the highest share of packed-FP32 instructions: the innermost unrolled charge loop) and over the")
    print("whole kernel. FieldMode template argument: 0 = softened field, 1 = raw field, 2 = ESP.\n")
    whole_rows = []
    for pat, what in HOT:
        hit = [m for m, d in names.items() if pat in d]
        if not hit:
            print(f"## `{pat}` — not found in the library\n")
            continue
        mangled = hit[0]
        ins = funcs[mangled]
        body = hottest_loop(ins)
        cw = collections.Counter(key(o) for _, o, _ in ins)
        cb = collections.Counter(key(o) for _, o, _ in body)
        print(f"## `{pat}` — {what}\n")
        print(f"* resources: `{usage.get(mangled, 'n/a')}`")
        print(f"* {len(ins)} SASS instructions; hottest loop body {len(body)} instructions")
        if body:
            packed = sum(cb[k] for k in PACKED)
            mufu = sum(v for k, v in cb.items() if k.startswith("MUFU"))
            other = len(body) - packed - mufu
            print(f"* loop body: {packed} packed FP32x2 + {mufu} MUFU + {other} other "
                  f"-> FMA-pipe issue share 2P/(2P+rest) = {2 * packed / (2 * packed + mufu + other):.3f}")
            print("* loop opcodes: " + ", ".join(f"{k} {cb[k]}" for k in SHOW if cb.get(k)) +
                  "; rest " + str(sum(v for k, v in cb.items() if k not in SHOW)))
        print("* whole kernel: " + ", ".join(f"{k} {cw[k]}" for k in SHOW if cw.get(k)) + "\n")
        whole_rows.append((pat, cw))
    print("## TMA / mbarrier evidence\n")
    print("| kernel | UBLKCP (1-D TMA bulk copy) | SYNCS (mbarrier arrive/try_wait) |")
    print("|---|---|---|")
    for pat, cw in whole_rows:
        print(f"| `{pat}` | {cw.get('UBLKCP', 0)} | {cw.get('SYNCS', 0)} |")
    spills = [ln for ln in res.splitlines() if "REG:" in ln and not re.search(r"STACK:0\b", ln)]
    print(f"\nKernels with a non-zero stack frame in the whole library: {len(spills)}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
