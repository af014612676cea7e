#!/usr/bin/env python
"""One launch of the dominant kernel of every bench.py workload (same synthetic inputs), for
`ncu --set full`: the per-launch DRAM traffic that bench.py reports as roofline.traffic.

    ncu --set full --clock-control none -k regex:"k1_grid_kernel|k2p_topo_kernel|k1_lattice" -o gpurun_out/round2_workloads \
        python tools/prof_workloads.py
    python tools/prof_workloads.py --collect gpurun_out/round2_workloads.ncu-rep [round2]   # -> profiles/round2_traffic.json
"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
ORDER = ["topo3a", "md1m", "topo_fine", "volume", "esp101", "volume2a"]

if len(sys.argv) > 2 and sys.argv[1] == "--collect":
    txt = subprocess.run(["ncu", "-i", sys.argv[2], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    idx = {h: i for i, h in enumerate(rows[0])}
    units = rows[1]
    def to_bytes(r, k):
        v = float(r[idx[k]].replace(",", "")); u = units[idx[k]].lower()
        return int(round(v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]))
    launches = [r for r in rows[2:] if "finalize" not in r[idx["Kernel Name"]] and
                (float(r[idx["gpu__time_duration.sum"]].replace(",", "")) > 0.008 or "_topo_kernel" in r[idx["Kernel Name"]])]
    # the softened lattice instantiation that exits at once is not a workload kernel
    launches = [r for r in launches if not ("k1_lattice" in r[idx["Kernel Name"]] and to_bytes(r, "dram__bytes_read.sum") < 40000
                                            and float(r[idx["gpu__time_duration.sum"]].replace(",", "")) < 0.01)]
    assert len(launches) == len(ORDER), [r[idx["Kernel Name"]] for r in launches]
    out = {}
    for name, r in zip(ORDER, launches):
        rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
        out[name] = {"kernel": r[idx["Kernel Name"]], "dram_read": rd, "dram_write": wr, "traffic": rd + wr,
                     "duration_ms_under_ncu": float(r[idx["gpu__time_duration.sum"]].replace(",", "")) *
                     {"ms": 1.0, "msecond": 1.0, "us": 1e-3, "usecond": 1e-3, "s": 1e3, "second": 1e3, "ns": 1e-6, "nsecond": 1e-6}[units[idx["gpu__time_duration.sum"]].lower()],
                     "source": f"ncu --set full --clock-control none, one launch, {os.path.basename(sys.argv[2])} (tools/prof_workloads.py)"}
    rnd = sys.argv[3] if len(sys.argv) > 3 else "round2"
    json.dump(out, open(os.path.join(ROOT, "profiles", f"{rnd}_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
    sys.exit(0)

import numpy as np, torch
import bench
from pycpet_b200.device import Engine
eng = Engine(0)
for name in ORDER:
    kind, desc, prm = bench.WORKLOADS[name]
    inp = bench.make_inputs(kind, prm, 0)
    eng.set_charges(torch.from_numpy(inp["x"]).cuda(), torch.from_numpy(inp["Q"]).cuda())
    if kind == "topo":
        eng.topo_batch(torch.from_numpy(inp["seeds"]).cuda(), torch.from_numpy(inp["n_iter"]).cuda(), inp["h"], inp["dims"])
    elif kind == "field":
        pts = torch.from_numpy(inp["points"]).cuda()
        if len(pts) >= 4096:
            ax = torch.from_numpy(inp["axis"]).cuda()
            eng.field_lattice(ax, ax, ax, soften=True, concat=True)
        else:
            eng.field_grid(pts, soften=True, concat=True)
    else:
        ax = torch.from_numpy(inp["axis"]).cuda()
        eng.esp_lattice(ax, ax, ax, concat_half=True)
    torch.cuda.synchronize()
print("done")
