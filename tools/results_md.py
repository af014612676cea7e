#!/usr/bin/env python
"""Collect the bench.py JSON lines under gpurun_out/ into profiles/<round>_results.md.

    python tools/results_md.py [round1]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "gpurun_out")
rnd = sys.argv[1] if len(sys.argv) > 1 else "round1"


def last_json(name):
    p = os.path.join(GO, name)
    if not os.path.exists(p):
        return None
    for line in reversed(open(p).read().strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    return None


single = [("topo3a", "bench_topo3a_1gpu.log"), ("md1m", "bench_md1m.log"), ("topo_fine", "bench_topo_fine.log"),
          ("volume", "bench_volume.log"), ("esp101", "bench_esp101.log"), ("volume2a", "bench_volume2a.log")]
multi = [("topo3a", 2, "bench_topo3a_2gpu.log"), ("topo3a", 4, "bench_topo3a_4gpu.log"),
         ("topo3a", 8, "bench_topo3a_8gpu.log"), ("md1m", 8, "bench_md1m_8gpu.log"),
         ("volume", 8, "bench_volume_8gpu.log"), ("esp101", 8, "bench_esp101_8gpu.log")]
md = [f"# {rnd} — bench.py lines (one B200 per rank)\n",
      "Each block is the single JSON line `python bench.py --workload <name>` printed on the GPU box (gpurun), reformatted.",
      "`value` = device-resident inputs, CUDA-event timed per step, L2 flushed between steps; `e2e` = host-pointer C-ABI with "
      "pinned host buffers (topo workloads: the MD-frame batch call, one frame per step);",
      "`roofline` = dominant kernel vs the live FFMA2 probe; `cpu_baseline` = the reference's own C (oracle/_ref) on the box's "
      "host cores (forked workers).\n",
      "| workload | value pair-evals/s | ms/step | e2e pair-evals/s | roofline frac (measured FP32 peak) | CPU reference pair-evals/s | e2e speed-up |",
      "|---|---|---|---|---|---|---|"]
lines = {}
for name, f in single:
    l = last_json(f)
    if l is None:
        continue
    lines[name] = l
    cb = l.get("cpu_baseline", {}).get("value")
    md.append(f"| `{name}` | {l['value']:.3e} | {l['ms_per_step']:.3f} | {l['e2e']['value']:.3e} | "
              f"{l['roofline']['frac']:.3f} | {cb:.3e} | {l['e2e']['value'] / cb:.0f}x |")
md.append("")
sc = []
for name, n, f in multi:
    l = last_json(f)
    if l is None or name not in lines:
        continue
    sc.append(f"{name} x{n}: {l['value']:.3e} ({100 * l['value'] / (n * lines[name]['value']):.1f} % of {n} x the 1-GPU value, "
              f"{l['ms_per_step']:.3f} ms/step)")
if sc:
    md.append("Scaling (weak, one frame per GPU per step, a different jittered frame on every rank, max-over-ranks timing): "
              + "; ".join(sc) + ". The loss is load imbalance between the ranks' different frames, not communication.\n")
for name, f in single:
    if name in lines:
        md += [f"## `{name}`\n", "```json", json.dumps(lines[name], indent=1), "```\n"]
ref = last_json("bench_reference.log")
if ref:
    md += ["## `--impl reference` (CPU arm alone, default workload)\n", "```json", json.dumps(ref, indent=1), "```\n"]
for name, n, f in multi:
    l = last_json(f)
    if l:
        md += [f"## `{name}`, {n} GPUs (torchrun; per-frame histograms all-gathered with NCCL)\n", "```json",
               json.dumps(l, indent=1), "```\n"]
out = os.path.join(ROOT, "profiles", f"{rnd}_results.md")
open(out, "w").write("\n".join(md))
print(out)
