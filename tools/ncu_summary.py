#!/usr/bin/env python
"""Turn an ncu report (+ optional launch-list csv) into the markdown summaries kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_r1b.ncu-rep profiles/round1_ncu_summary.md \
        [gpurun_out/launches_r1.csv profiles/round1_launches.md]
"""
import collections
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "sm__cycles_elapsed.avg",
    "smsp__warps_eligible.avg.per_cycle_active",
]


def raw_rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(txt)))


def ncu_md(rep, out, title):
    rows = raw_rows(rep)
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    md = [f"# {title}\n",
          "Captured with `ncu --set full --clock-control none --import-source on -k regex:\"k1_grid|k2w_topo|k1_lattice\" "
          "-c 8 python tools/prof_target.py all 1` (M = 7,890 charges; K1 on a 101^3 grid, K2 on the 3A line set 47^3).",
          "Numbers under ncu are for pipe utilisation, stall reasons and DRAM traffic only; throughput is quoted from "
          "bench.py / tools/sweep.py (CUDA events, no profiler).\n"]
    for r in rows[2:]:
        md.append(f"## `{r[idx['Kernel Name']]}`\n")
        md.append("| metric | value | unit |\n|---|---|---|")
        for k in KEEP:
            if k in idx:
                md.append(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |")
        vals = sorted([(float(r[idx[k]]), k) for k in stall], reverse=True)[:6]
        md.append("| top stall reasons (warps per issue) | " + ", ".join(
            f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}"
            for v, k in vals) + " | |")
        md.append("")
    open(out, "w").write("\n".join(md) + "\n")


def launches_md(csv_path, out, title):
    rows = list(csv.reader(l for l in open(csv_path) if l.startswith('"')))
    h = rows[0]
    i_name, i_val = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[i_name], [0, 0.0])
        a[0] += 1
        a[1] += float(r[i_val].replace(",", ""))
    tot = sum(a[1] for n, a in agg.items() if "fp32_probe" not in n)
    md = [f"# {title}\n",
          "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3`.",
          "Per-launch times are cold-cache and serialised: compare SHARES. The FP32 probe kernels run once after the "
          "timed region and are excluded from the share; `FillFunctor` is the L2 flush between steps.\n",
          "| kernel | launches | total ms | share of step kernels |", "|---|---|---|---|"]
    for n, a in agg.items():
        share = "" if "fp32_probe" in n else f"{100 * a[1] / tot:.2f} %"
        md.append(f"| `{n[:100]}` | {a[0]} | {a[1] / 1e6:.3f} | {share} |")
    open(out, "w").write("\n".join(md) + "\n")


if __name__ == "__main__":
    ncu_md(sys.argv[1], sys.argv[2], "Round 1 — ncu full-set summary of the hot kernels (one B200)")
    if len(sys.argv) > 4:
        launches_md(sys.argv[3], sys.argv[4], "Round 1 — launch list of the default bench step")
