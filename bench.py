#!/usr/bin/env python
"""bench.py -- the measurement contract for the PyCPET hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A *step* = one pass of the hot path over one frame of synthetic input:
upload/pack the frame's charges, run the kernel(s) of the workload, bin the result.  Default
workload = BASELINE.json configs[1] ("3A_field-topology": 47^3 = 103,823 streamlines in a 0.5 A
box, step 0.1, on a 7,890-charge protein-like frame, then the distance x curvature histogram).

One JSON line on stdout (rank 0):
  value      : whole-job pair-evaluations/s with inputs already resident in HBM (device-pointer
               C-ABI), CUDA-event timed per step on the launching stream, L2 flushed between steps
  e2e        : the same metric through the host-pointer C-ABI the reference-facing plugin calls
               (pinned host buffers in, host buffers out; copies inside the timed region)
  roofline   : the dominant kernel against the non-tensor FP32 issue rate (20 flop / pair-eval),
               its duration measured live with CUDA events around that kernel alone
  cpu_baseline: the reference's own C (oracle/_ref, compiled from /root/reference in place) on
               this box's host cores, on a bounded sample of the same workload
  sustained  : (N = 1) the same step loop continued for >= 5 s: pair-evals/s, SM clock, power, throttle reasons
  others     : (N = 1, default workload) short records of the other BASELINE configurations (md1m, volume, esp101, volume2a)
  parity_checked: every gathered histogram of the timed steps (N > 1) and the end-to-end call's histograms against
               the device arm's, compared outside the timed region
`--impl reference` times that CPU reference alone (rank 0 only) and prints the same line shape.
`--split seeds|slab` (with --gpus N) are the strong-scaling modes: ONE 1M-line frame dealt by streamlines, ONE 464^3 mesh
by slabs of x-planes, the final NCCL gather inside the timed bracket ("scaling": "strong").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

# torchrun sets this for N > 1; do the same at N = 1 so that idle OpenMP workers of the host math
# libraries do not spin next to the thread that enqueues the end-to-end arm (the CPU reference arm
# uses forked worker processes, not OpenMP)
os.environ.setdefault("OMP_NUM_THREADS", "1")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FRAME_POOL = 8                   # distinct MD frames cycled by the streamline workloads
GATHER_WINDOW = 4                # per-frame histogram all-gathers allowed in flight (N > 1)
FLOPS_PER_PAIR = 20.0            # BASELINE.json north_star: "counted at 20 flops/pair"
ESP_FLOPS_PER_PAIR = 11.0
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # 74.5: 148 SM x 128 lanes x 2 x 1.965 GHz

# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from one
# `ncu --set full` capture of that kernel on the same synthetic inputs (tools/prof_workloads.py;
# summaries in profiles/round2_k2x_ncu.md, profiles/round1_k2w_ncu.md and profiles/round1_ncu_summary.md)
def _load_traffic():
    out = {}
    for name in ("round1_traffic.json", "round2_traffic.json"):       # later rounds override
        try:
            with open(os.path.join(ROOT, "profiles", name)) as fh:
                data = json.load(fh)
        except (OSError, ValueError):
            continue
        out.update({k: (int(v["traffic"]), f"profiles/{name}: {v['kernel']}, {v['dram_read']} B read + "
                                           f"{v['dram_write']} B written back to DRAM during the launch ({v['source']})")
                    for k, v in data.items()})
    return out


NCU_TRAFFIC = _load_traffic()

WORKLOADS = {
    # name: (kind, description, params)
    "topo3a": ("topo", "3A_field-topology: topo, 47^3=103,823 streamlines, 7,890 synthetic protein-like "
               "charges, box half-width 0.5 A, step 0.1 A (max_steps 17), + 2-D histogram",
               dict(m=7890, n_axis=47, half=0.5, h=0.1)),
    "md1m": ("topo", "MD-frame topology: 100^3=1,000,000 streamlines per frame, 7,890 charges, "
             "box 0.5 A, step 0.1 A, + 2-D histogram", dict(m=7890, n_axis=100, half=0.5, h=0.1)),
    "topo_fine": ("topo", "topo, 47^3 streamlines, 7,890 charges, box 0.5 A, step 0.01 A (max_steps 173)",
                  dict(m=7890, n_axis=47, half=0.5, h=0.01)),
    "volume2a": ("field", "2A_3D-field: volume E-field, 11^3 grid, 7,890 charges", dict(m=7890, n_axis=11, half=0.5)),
    "volume": ("field", "synthetic sweep cell: volume E-field, 100^3 grid points x 100,000 charges",
               dict(m=100_000, n_axis=100, half=1.5)),
    "esp101": ("esp", "volume_ESP: 101^3 grid x ~100k-charge solvated system", dict(m=100_000, n_axis=101, half=5.0)),
    # sweep corner N = 1e8 (BASELINE configs[4]); only with --split slab (the mesh is never expanded on the host)
    "esp464": ("esp", "volume_ESP: 464^3 = 99,897,344 grid points x 100,000 charges", dict(m=100_000, n_axis=464, half=1.5)),
    "volume464": ("field", "volume E-field: 464^3 = 99,897,344 grid points x 100,000 charges",
                  dict(m=100_000, n_axis=464, half=1.5)),
}
MUFU_PEAK = 148 * 16 * 1.965e9   # rsqrt/s: 16 MUFU lanes per SM per clock, binds ESP


# ------------------------------------------------------------------------------------------------
# clocks: sample NVML during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4, "hw_power_brake": 0x80}

    def __init__(self, local_rank: int):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        # NVML numbers the physical devices; CUDA_VISIBLE_DEVICES (indices or GPU-/MIG- UUIDs) maps
        # this process's local rank onto them
        entry = None
        vis = [v.strip() for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
        if local_rank < len(vis):
            entry = vis[local_rank]
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            if entry is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            elif entry.isdigit():
                self.h = pynvml.nvmlDeviceGetHandleByIndex(int(entry))
            else:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(entry.encode())
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:
                    pass
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        self.samples, self.reasons, self.power = [], set(), []
        self._stop.clear()
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
            self._thr = None

    def summary(self):
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None),
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "sm_mhz_min": (min(self.samples) if self.samples else None),
                "power_w_max": (max(self.power) if self.power else None)}

    def rejected(self):
        if self.reasons & set(self.BAD):
            return True
        s = self.summary()
        if s["sm_mhz"] and s["sm_max_mhz"] and s["sm_mhz"] < 0.6 * s["sm_max_mhz"] and not self.reasons:
            return True       # stuck well below max with no reason: a leftover clock lock
        return False


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
def make_inputs(kind, prm, frame_id):
    """Synthetic frame `frame_id`: base charge set + per-frame Gaussian jitter sigma = 0.3 A
    (seed = frame id), as SURVEY.md section 8(d) specifies for MD batches."""
    import synth

    x, Q = synth.charges(prm["m"], seed=1, box=prm["half"])
    if frame_id > 0:
        rng = np.random.default_rng(1000 + frame_id)
        x = (x + rng.normal(0.0, 0.3, x.shape)).astype(np.float32)
        inside = np.all(np.abs(x) < prm["half"] * 1.05, axis=1)
        x[inside] *= np.float32(3.0)          # keep jittered charges out of the sampling box
    d = dict(x=np.ascontiguousarray(x), Q=np.ascontiguousarray(Q))
    if kind == "topo":
        seeds, n_iter, dims, max_steps = synth.seeds(prm["n_axis"], prm["half"], prm["h"])
        d.update(seeds=seeds, n_iter=n_iter.astype(np.int32), dims=dims, h=prm["h"], max_steps=max_steps)
    else:
        d.update(points=synth.grid(prm["n_axis"], prm["half"]),
                 axis=np.linspace(-prm["half"], prm["half"], prm["n_axis"]).astype(np.float32))
    return d


def hist_edges(kind, prm):
    # fixed-range variant (scripts/residue_breakdown_analysis.py:28-37 style): 50 x 50 bins
    h = prm.get("h", 0.1)
    n_max = round(2 * np.sqrt(3) * prm["half"] / h)
    return np.linspace(0.0, n_max * h, 51), np.linspace(0.0, 5.0, 51)


def pinned(a):
    import torch

    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


# the three places run_gpu touches torch.cuda directly (tests/test_bench_dryrun.py swaps them, the
# Engine and Math_ops for host stand-ins to exercise the accounting and the N > 1 flow over gloo)
def _device(local_rank):
    import torch

    torch.cuda.set_device(local_rank)
    return torch.device("cuda", local_rank)


def _event():
    import torch

    return torch.cuda.Event(enable_timing=True)


def _sync():
    import torch

    torch.cuda.synchronize()


def run_gpu(args, rank, world, local_rank, sub=False):
    """One workload, frames sharded over the ranks (weak scaling).  sub=True: a short side record for
    the N = 1 line's "others" (no CPU baseline, no sustained run, no nesting)."""
    import torch
    import torch.distributed as dist

    from pycpet_b200 import Math_ops
    from pycpet_b200.device import Engine

    kind, desc, prm = WORKLOADS[args.workload]
    if prm["n_axis"] > 300:
        raise SystemExit(f"--workload {args.workload} expands to 1e8 points: run it with --split slab")
    dev = _device(local_rank)
    # Streamline workloads run a trajectory: POOL frames (base charge set + per-frame jitter),
    # rank r integrates frame (r + s) mod POOL at step s -- a different frame on every GPU at every step, and
    # the same mix of frames on every rank whatever N is.  Grid workloads: one frame per rank (equal work).
    if kind == "topo":
        pool = [make_inputs(kind, prm, frame_id=j) for j in range(FRAME_POOL)]
        inp = pool[0]
    else:
        inp = make_inputs(kind, prm, frame_id=rank)
        pool = [inp]
    de, ce = hist_edges(kind, prm)

    # ---- device-resident arm ---------------------------------------------------------------
    eng = Engine(local_rank)
    eng.set_tuning(timing=1)
    dcharges = [(torch.from_numpy(f["x"]).to(dev), torch.from_numpy(f["Q"]).to(dev)) for f in pool]
    pool_pairs = [0] * len(pool)        # pair-evaluations of every pool frame (probe pass, untimed)
    seq = {"s": 0}                      # steps issued so far on this rank (warm-up included)
    if kind == "topo":
        dseeds, dnit = torch.from_numpy(inp["seeds"]).to(dev), torch.from_numpy(inp["n_iter"]).to(dev)
        dout = torch.empty((len(inp["seeds"]), 2), dtype=torch.float32, device=dev)
        dcounts = torch.empty((1, 50, 50), dtype=torch.int64, device=dev)
    else:
        dpts = torch.from_numpy(inp["points"]).to(dev)
        daxis = torch.from_numpy(inp["axis"]).to(dev)       # the box mesh is axis x axis x axis, z fastest
        n = len(inp["points"])
        dout = (torch.empty((n, 6), dtype=torch.float32, device=dev) if kind == "field"
                else torch.empty((n, 4), dtype=torch.float16, device=dev))
    pending = []        # in-flight NCCL gathers of per-frame histograms (world > 1)
    pool_counts = [None] * len(pool)    # this rank's own histogram of every pool frame (probe pass)
    # histograms gathered per NCCL call: 1 = one all-gather per frame (default); W > 1 batches W frames per call;
    # 0 = no gathers at all (diagnostic only: how much of the N-GPU step is the exchange and its interference)
    GW = int(getattr(args, "gather_every", 1))
    if kind == "topo" and world > 1 and GW > 0:
        # preallocated gather ring, one slot per batch of GW steps (enough slots for a timed region + the gathers
        # still in flight from the warm-up): all_gather_into_tensor writes slot i, nothing is allocated inside the
        # timed region and every gathered histogram of the timed steps is still there for the parity check afterwards
        n_ring = (args.steps + GATHER_WINDOW * GW) // GW + 4
        ring = torch.empty((n_ring, world, GW, 50, 50), dtype=torch.int64, device=dev)
        snaps = torch.empty((n_ring, GW, 50, 50), dtype=torch.int64, device=dev)
        ring_step = [[-1] * GW for _ in range(n_ring)]      # global step number whose histograms slot i, entry j holds

    def issue_gather(slot):
        pending.append(dist.all_gather_into_tensor(ring[slot].view(world * GW, 50, 50), snaps[slot], async_op=True))

    def drain(keep=0):
        """Wait for the oldest histogram gathers until at most `keep` are in flight."""
        while len(pending) > keep:
            pending.pop(0).wait()
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    launches = {"n": 0}
    kern_ms = []
    work = {"pairs": 0, "units": 0, "k_launch": 1}

    def step_device(probe=False):
        """One step, fully asynchronous (no host synchronisation inside the timed region).
        probe=True (warm-up only) additionally reads the work counters back."""
        s_now = seq["s"]
        j = (rank + s_now) % len(pool)
        seq["s"] += 1
        eng.set_charges(*dcharges[j])                                    # pack kernel
        n_launch = 1
        if kind == "topo":
            eng.topo_batch(dseeds, dnit, inp["h"], inp["dims"], out=dout)
            if probe:
                c = eng.last_counters()
                pool_pairs[j] = c["pair_evals"]
                work["units"], work["k_launch"] = len(inp["seeds"]), c["launches"]
                work["path"] = eng.last_path() if hasattr(eng, "last_path") else "k2p"
            eng.hist2d(dout, de, ce, out=dcounts)
            if probe:
                pool_counts[j] = dcounts[0].clone()
            n_launch += work["k_launch"] + 1
            if world > 1 and GW > 0:
                # the path's one exchange: per-frame histograms to every rank.  Issued asynchronously
                # on NCCL's stream from a snapshot of the counts, so the next frame's kernels never
                # wait on communication; all handles are waited for before the timed region closes.
                slot, sub = (s_now // GW) % n_ring, s_now % GW
                snaps[slot][sub].copy_(dcounts[0])
                ring_step[slot][sub] = s_now
                if sub == GW - 1:
                    issue_gather(slot)
        elif kind == "field":
            # device arm: the mesh is described by its axes (what the host entry point derives from the
            # flat list by itself, see cpet_field_grid); the e2e arm below hands over the flat list
            if n >= 4096:
                eng.field_lattice(daxis, daxis, daxis, soften=True, concat=True, out=dout)
            else:            # below the library's own mesh-detection threshold: general kernel
                eng.field_grid(dpts, soften=True, concat=True, out=dout)
            if probe:
                c = eng.last_counters()
                pool_pairs[j] = c["pair_evals"]
                work["units"], work["k_launch"] = len(inp["points"]), c["launches"]
            n_launch += work["k_launch"]
        else:
            if n >= 4096:    # same arrangement as the field arm: the mesh by its axes (lattice kernel)
                eng.esp_lattice(daxis, daxis, daxis, concat_half=True, out=dout)
            else:
                eng.esp_grid(dpts, concat_half=True, out=dout)
            if probe:
                c = eng.last_counters()
                pool_pairs[j] = c["pair_evals"]
                work["units"], work["k_launch"] = len(inp["points"]), c["launches"]
            n_launch += work["k_launch"]
        launches["n"] += n_launch

    def barrier():
        _sync()
        if world > 1:
            dist.barrier()
        _sync()

    sampler = ClockSampler(local_rank)

    def timed_device():
        for _ in range(len(pool)):       # probe pass: one untimed step per pool frame, counters read back
            drain(keep=GATHER_WINDOW)
            step_device(probe=True)
        for _ in range(args.warmup):
            flush_buf.zero_()
            drain(keep=GATHER_WINDOW)
            step_device()
        drain()
        evs = [(_event(), _event()) for _ in range(args.steps + 1)]
        barrier()
        eng.kernel_times()               # reset the library's per-launch event record
        launches["n"] = 0
        work["pairs_total"] = sum(pool_pairs[(rank + seq["s"] + k) % len(pool)] for k in range(args.steps))
        work["pairs"] = work["pairs_total"] / args.steps
        work["s_start"] = seq["s"]
        sampler.start()
        for a, b in evs[:-1]:
            flush_buf.zero_()            # L2 flush, outside the per-step event bracket
            a.record()
            drain(keep=GATHER_WINDOW)    # results are collected a few frames behind the integrator, so one
            step_device()                # heavy frame on one rank does not stall the others at every step
            b.record()
        evs[-1][0].record()
        if kind == "topo" and world > 1 and GW > 1 and seq["s"] % GW != 0:
            issue_gather(((seq["s"] - 1) // GW) % n_ring)      # the last, partly filled batch
        drain()                          # the last frame's gather, inside its own timed bracket
        evs[-1][1].record()
        barrier()
        sampler.stop()
        kern_ms[:] = eng.kernel_times()  # dominant kernel alone, one entry per timed step
        work["launches"] = launches["n"]  # kernels launched inside the timed region
        work["step_ms"] = [a.elapsed_time(b) for a, b in evs[:-1]]
        return sum(a.elapsed_time(b) for a, b in evs) * 1e-3

    t_dev = timed_device()
    remeasured = False
    if sampler.rejected():
        remeasured = True
        time.sleep(2.0)
        t_dev = timed_device()
    clocks = sampler.summary()
    clocks["remeasured"] = remeasured

    # ---- parity of what was just timed (outside the timed region) -------------------------------------
    # N > 1: every gathered per-frame histogram of the timed steps (steps x ranks) against THIS rank's own
    # histogram of that pool frame (probe pass): rank q computed frame (q + s) mod POOL at step s.
    parity = {"checked": True, "what": []}
    if kind == "topo" and world > 1 and GW > 0:
        n_cmp = 0
        for sg in range(work["s_start"], work["s_start"] + args.steps):
            slot, sub = (sg // GW) % n_ring, sg % GW
            ok = ring_step[slot][sub] == sg
            for q in range(world):
                ok = ok and bool(torch.equal(ring[slot][q][sub], pool_counts[(q + sg) % len(pool)]))
                n_cmp += 1
            parity["checked"] = parity["checked"] and ok
        parity["what"].append(f"{n_cmp} gathered per-frame histograms of the timed steps == this rank's own "
                              "histogram of the same pool frame, bit for bit")

    # ---- sustained run (N = 1): the same step loop for several seconds ------------------------------
    sustained = None
    sus_s = float(getattr(args, "sustained", 0.0) or 0.0)
    if world == 1 and not sub and sus_s > 0:
        sampler.start()
        t_busy, pairs_sus, n_sus = 0.0, 0.0, 0
        t_wall0 = time.perf_counter()
        while t_busy < sus_s and time.perf_counter() - t_wall0 < 4 * sus_s + 10:
            evs = [(_event(), _event()) for _ in range(100)]
            for a, b in evs:
                flush_buf.zero_()
                a.record()
                pairs_sus += pool_pairs[(rank + seq["s"]) % len(pool)]
                step_device()
                b.record()
            _sync()
            t_busy += sum(a.elapsed_time(b) for a, b in evs) * 1e-3
            n_sus += len(evs)
        wall = time.perf_counter() - t_wall0
        sampler.stop()
        cs = sampler.summary()
        sustained = {"value": pairs_sus / t_busy, "unit": "pair-evals/s", "steps": n_sus, "seconds": t_busy,
                     "wall_s": wall, "ms_per_step": t_busy / n_sus * 1e3,
                     "fp32_frac_of_nominal": pairs_sus / t_busy * (ESP_FLOPS_PER_PAIR if kind == "esp" else FLOPS_PER_PAIR)
                                             / 1e12 / NOMINAL_FP32_TFLOPS,
                     "sm_mhz": cs["sm_mhz"], "sm_mhz_min": cs["sm_mhz_min"], "power_w_max": cs["power_w_max"],
                     "reasons": cs["reasons"], "clock_samples": cs["samples"],
                     "how": "the timed region's step loop (L2 flush, one event pair per step) continued back to back"}

    # ---- end-to-end arm: host-pointer C-ABI, pinned host buffers ---------------------------------
    m = Math_ops(device=local_rank)
    keep = []
    def pin(a):
        t, v = pinned(a)
        keep.append(t)
        return v
    hcharges = [(pin(f["x"]), pin(f["Q"])) for f in pool]
    hx, hq = hcharges[0]
    if kind == "topo":
        hseeds, hnit = pin(inp["seeds"]), pin(inp["n_iter"])
        h2d = hx.nbytes + hq.nbytes + hseeds.nbytes + hnit.nbytes + de.nbytes + ce.nbytes
        d2h = len(hseeds) * 8 + 50 * 50 * 8
    else:
        hpts = pin(inp["points"])
        h2d = hx.nbytes + hq.nbytes + hpts.nbytes
        d2h = len(hpts) * (24 if kind == "field" else 8)

    # outputs land in pinned host buffers too (the caller-allocated arrays of the reference's API)
    if kind == "topo":
        # MD-frame batch call: K frames per call, every frame with its own charges and n_iter row
        # (copied host->device per frame), rows and counts copied back per frame; seeds and bin edges
        # are shared by the frames of a call and uploaded once per call
        K = args.steps
        o_rows = pin(np.zeros((K, len(hseeds), 2), np.float32))
        o_counts = pin(np.zeros((K, 50, 50), np.int64))
        hnit_frames = pin(np.broadcast_to(inp["n_iter"], (K, len(hseeds))).copy())
        e2e_ids = [(rank + k) % len(pool) for k in range(K)]            # the same rotation as the device arm
        frames = [hcharges[j] for j in e2e_ids]
        work["pairs_e2e_total"] = sum(pool_pairs[j] for j in e2e_ids)
        h2d = hx.nbytes + hq.nbytes + hnit.nbytes + (hseeds.nbytes + de.nbytes + ce.nbytes) // K
    elif kind == "field":
        o_field = pin(np.zeros((len(hpts), 6), np.float32))
    else:
        o_esp = pin(np.zeros((len(hpts), 4), np.float16))

    def e2e_frames(k):
        return m.topo_hist_frames(frames[:k], hseeds, hnit_frames[:k], de, ce, step_size=inp["h"],
                                  dimensions=inp["dims"], want_rows=True, rows_out=o_rows[:k],
                                  counts_out=o_counts[:k])

    if kind != "topo":
        work["pairs_e2e_total"] = work["pairs"] * args.steps

    def step_e2e():
        m.set_charges(hx, hq)
        if kind == "field":
            return m.field_grid(hpts, soften=True, concat=True, out=o_field)
        return m.esp_grid(hpts, concat_half=True, out=o_esp)

    if kind == "topo":
        e2e_frames(min(K, max(3, args.warmup)))
    else:
        for _ in range(max(3, args.warmup)):
            step_e2e()
    barrier()
    t0 = time.perf_counter()
    if kind == "topo":
        res = e2e_frames(K)              # one call, K frames (steps), copies overlapped with kernels
    else:
        for _ in range(args.steps):
            res = step_e2e()
    _sync()
    t_e2e = time.perf_counter() - t0
    barrier()
    if kind == "topo":
        # the host-pointer arm against the device-pointer arm: same frames, same counts
        ok = all(np.array_equal(np.asarray(o_counts[k]), pool_counts[j].cpu().numpy()) for k, j in enumerate(e2e_ids))
        parity["checked"] = parity["checked"] and ok
        parity["what"].append(f"{K} histograms returned by the end-to-end call == the device arm's histograms of "
                              "the same pool frames")

    # ---- reduce over ranks ---------------------------------------------------------------------------
    step_ms = work["step_ms"]
    # topo steps record two timed launches each (integrator, then histogram): keep the integrator.
    # The library keeps the last 256 launches; 2 x steps and 256 are both even, so a truncated record
    # still starts on an integrator launch.
    kt = kern_ms[0::2] if (kind == "topo" and len(kern_ms) % 2 == 0) else kern_ms
    per_rank = torch.tensor([[statistics.median(step_ms), max(step_ms), float(np.mean(kt)) if len(kt) else 0.0,
                              1.0 if parity["checked"] else 0.0]], dtype=torch.float64, device=dev)
    if world > 1:
        allr = torch.empty((world, 4), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, per_rank)
        per_rank = allr
    per_rank = per_rank.cpu().numpy()
    tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
    ww = torch.tensor([float(work["pairs_total"]), float(work["units"]), float(work["pairs_e2e_total"])],
                      dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ww, op=dist.ReduceOp.SUM)
    t_dev, t_e2e = float(tt[0]), float(tt[1])
    pairs_all, units_all, pairs_e2e_all = float(ww[0]), float(ww[1]), float(ww[2])   # pairs: whole timed region

    if rank != 0:
        return None
    flops = ESP_FLOPS_PER_PAIR if kind == "esp" else FLOPS_PER_PAIR
    peak_ffma2 = eng.fp32_peak_tflops(True)
    peak_ffma = eng.fp32_peak_tflops(False)
    peak = max(peak_ffma2, peak_ffma)
    k_ms = float(np.mean(kt))
    achieved = work["pairs"] * flops / (k_ms * 1e-3) / 1e12
    value = pairs_all / t_dev
    line = {
        "metric": "pair-evals/s", "value": value, "unit": "pair-evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": desc, "name": args.workload, "charges": int(len(inp["Q"])),
                   "units_per_step_per_gpu": int(work["units"]),
                   "pair_evals_per_step_per_gpu": int(work["pairs"]),
                   "l2": "flushed between timed steps (256 MiB write)",
                   "frames": (f"{len(pool)} jittered MD frames; rank r integrates frame (r + step) mod {len(pool)}; "
                              "pair_evals_per_step_per_gpu is rank 0's mean over the timed steps"
                              if kind == "topo" else "one frame per rank"),
                   "parallelism": f"frames sharded, 1 frame per GPU per step, x{world}" + (
                       "" if world == 1 or kind != "topo" else
                       (", per-frame histograms all-gathered" + (f" in batches of {GW} frames" if GW > 1 else "")
                        if GW > 0 else ", DIAGNOSTIC: histogram gathers disabled"))},
        "units_per_s": units_all * args.steps / t_dev,
        "units": "streamlines" if kind == "topo" else "grid points",
        "fp32_frac_of_nominal": value / world * flops / 1e12 / NOMINAL_FP32_TFLOPS,
        "clocks": clocks,
        "e2e": {"value": pairs_e2e_all / t_e2e, "unit": "pair-evals/s",
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": t_e2e / args.steps * 1e3,
                "api": ("Math_ops.topo_hist_frames -> cpet_topo_hist_frames: one call, one frame per step, pinned "
                        "host buffers, charges + n_iter in and rows + counts out per frame, seeds/edges once per call"
                        if kind == "topo" else
                        "Math_ops.field_grid / esp_grid -> cpet_field_grid / cpet_esp_grid, one call per step, "
                        "pinned host buffers")},
        "gpu_launches": int(work["launches"]),
        "parity_checked": bool(per_rank[:, 3].min() == 1.0) if kind == "topo" else None,
        "parity": ("; ".join(parity["what"]) if kind == "topo" else
                   "grid workloads: see tests/test_gpu_parity.py (oracle parity at this size)"),
        "step_ms_by_rank": {"median": [round(float(v), 4) for v in per_rank[:, 0]],
                            "max": [round(float(v), 4) for v in per_rank[:, 1]],
                            "kernel_mean": [round(float(v), 4) for v in per_rank[:, 2]]},
        "roofline": {"bound": "fp32-non-tensor" if kind != "esp" else "mufu (fp32 fraction quoted at 11 flop/pair)",
                     "kernel": (work.get("path", "k2p") + "_topo_kernel" if kind == "topo" else
                                (("k1_lattice_nodes_kernel" if kind == "field" and len(inp["points"]) >= 100000
                                  else "k1_lattice_kernel") if len(inp["points"]) >= 4096 else "k1_grid_kernel")),
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "peak_source": "measured live: register-resident FMA loop (cpet_fp32_peak_probe, "
                                    f"FFMA2 {peak_ffma2:.1f} / FFMA {peak_ffma:.1f} TFLOP/s); "
                                    "MEASURED_PEAKS.json has no FP32 entry",
                     "peak_nominal": NOMINAL_FP32_TFLOPS, "frac_nominal": achieved / NOMINAL_FP32_TFLOPS,
                     "flops_per_pair": flops, "kernel_ms": k_ms,
                     "traffic": NCU_TRAFFIC.get(args.workload, (None, None))[0],
                     "traffic_source": NCU_TRAFFIC.get(args.workload, (None, None))[1],
                     "algorithmic_bytes": int(work["units"]) * (28 if kind == "topo" else (36 if kind == "field" else 20))},
    }
    if kind == "esp":
        line["roofline"]["mufu_peak_pairs_per_s"] = MUFU_PEAK
        line["roofline"]["mufu_frac"] = work["pairs"] / (k_ms * 1e-3) / MUFU_PEAK
    if sub:
        return line
    if sustained is not None:
        line["sustained"] = sustained
    if world == 1:
        line["cpu_baseline"] = cpu_baseline(kind, prm, inp,
                                            budget_s=12.0 if args.cpu_seconds is None else args.cpu_seconds)
        others = getattr(args, "others", None)
        if others:
            # short, clock-sampled records of the other BASELINE configurations on the same box (the driver only
            # runs the default workload): same code path as their own `--workload` lines, fewer steps
            line["others"] = {}
            for name in others:
                if name == args.workload:
                    continue
                sa = argparse.Namespace(**vars(args))
                sa.workload, sa.steps, sa.warmup = name, (20 if name == "volume2a" else 5), 3
                o = run_gpu(sa, rank, world, local_rank, sub=True)
                line["others"][name] = {k: o[k] for k in ("value", "unit", "steps", "ms_per_step", "units_per_s", "units",
                                                         "fp32_frac_of_nominal", "clocks", "e2e", "gpu_launches",
                                                         "parity_checked", "roofline")}
                line["others"][name]["workload"] = o["config"]["workload"]
    return line


# ------------------------------------------------------------------------------------------------
# strong scaling: ONE frame / ONE grid split over the ranks, the final gather inside the timed bracket
# ------------------------------------------------------------------------------------------------
def run_split(args, rank, world, local_rank):
    """--split seeds: one topology frame, its streamlines dealt over the ranks (LPT-sorted serpentine
    deal of 32-line blocks, pycpet_b200.sharding.deal_lines_all), rows all-gathered and restored to
    seed order on every rank, histogram on every rank.
    --split slab: one box mesh split by slabs of x-planes (sharding.lattice_sharded), the reference's
    return rows ((N,6) f32 [x|E] or (N,4) f16 [x|phi]) all-gathered on every rank.
    A step = upload-free device pass over the whole frame: pack charges -> shard kernel -> NCCL gather
    (-> scatter + histogram).  `value` = the frame's pair-evaluations x steps / max-over-ranks time."""
    import torch
    import torch.distributed as dist

    import synth
    from pycpet_b200 import sharding
    from pycpet_b200.device import Engine

    kind, desc, prm = WORKLOADS[args.workload]
    if (args.split == "seeds") != (kind == "topo"):
        raise SystemExit("--split seeds needs a streamline workload, --split slab a grid workload")
    dev = _device(local_rank)
    eng = Engine(local_rank)
    eng.set_tuning(timing=1)
    x, Q = synth.charges(prm["m"], seed=1, box=prm["half"])
    dx, dq = torch.from_numpy(x).to(dev), torch.from_numpy(Q).to(dev)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    gather_ev = []

    if kind == "topo":
        seeds, n_iter, dims, _ = synth.seeds(prm["n_axis"], prm["half"], prm["h"])
        n_iter = n_iter.astype(np.int32)
        de, ce = hist_edges(kind, prm)
        deal = sharding.deal_lines_all(n_iter, world)
        counts = [len(d) for d in deal]
        ids = deal[rank]
        dseeds = torch.from_numpy(np.ascontiguousarray(seeds[ids])).to(dev)
        dnit = torch.from_numpy(np.ascontiguousarray(n_iter[ids])).to(dev)
        perm = torch.from_numpy(np.concatenate(deal).astype(np.int64)).to(dev)
        local = torch.empty((len(ids), 2), dtype=torch.float32, device=dev)
        gathered = torch.empty((len(seeds), 2), dtype=torch.float32, device=dev)
        full = torch.empty_like(gathered)
        dcounts = torch.empty((1, 50, 50), dtype=torch.int64, device=dev)
        units = len(seeds)

        def step():
            eng.set_charges(dx, dq)
            eng.topo_batch(dseeds, dnit, prm["h"], dims, out=local)
            a, b = _event(), _event()
            a.record()
            sharding.all_gather_blocks(local, counts, out=gathered)
            full[perm] = gathered
            b.record()
            gather_ev.append((a, b))
            eng.hist2d(full, de, ce, out=dcounts)
    else:
        n = prm["n_axis"]
        axis = np.linspace(-prm["half"], prm["half"], n).astype(np.float32)
        dax = torch.from_numpy(axis).to(dev)
        lo, hi = sharding.slab(n, rank, world)
        counts = [(sharding.slab(n, r, world)[1] - sharding.slab(n, r, world)[0]) * n * n for r in range(world)]
        cols, dt = (6, torch.float32) if kind == "field" else (4, torch.float16)
        local = torch.empty((counts[rank], cols), dtype=dt, device=dev)
        full = torch.empty((n ** 3, cols), dtype=dt, device=dev) if world > 1 else local
        units = n ** 3

        def step():
            eng.set_charges(dx, dq)
            if kind == "field":
                eng.field_lattice(dax[lo:hi], dax, dax, soften=True, concat=True, out=local)
            else:
                eng.esp_lattice(dax[lo:hi], dax, dax, concat_half=True, out=local)
            if world > 1:
                a, b = _event(), _event()
                a.record()
                sharding.all_gather_blocks(local, counts, out=full)
                b.record()
                gather_ev.append((a, b))

    def barrier():
        _sync()
        if world > 1:
            dist.barrier()
        _sync()

    step()                                                  # probe: this rank's share of the work
    _sync()
    c = eng.last_counters() if kind != "topo" else None
    if kind == "topo":
        eng.topo_batch(dseeds, dnit, prm["h"], dims, out=local)
        _sync()
        c = eng.last_counters()
    my_pairs = float(c["pair_evals"])
    for _ in range(args.warmup):
        flush_buf.zero_()
        step()
    sampler = ClockSampler(local_rank)
    evs = [(_event(), _event()) for _ in range(args.steps)]
    barrier()
    eng.kernel_times()
    gather_ev.clear()
    sampler.start()
    for a, b in evs:
        flush_buf.zero_()
        a.record()
        step()
        b.record()
    barrier()
    sampler.stop()
    kern_ms = eng.kernel_times()
    kt = kern_ms[0::2] if (kind == "topo" and len(kern_ms) % 2 == 0) else kern_ms
    t_dev = sum(a.elapsed_time(b) for a, b in evs) * 1e-3
    g_ms = float(np.mean([a.elapsed_time(b) for a, b in gather_ev])) if gather_ev else 0.0

    # ---- end to end: inputs from pinned host memory every step, the full result back on rank 0 -------
    hx, hq = pinned(x)[0], pinned(Q)[0]
    if kind == "topo":
        hs, hn = pinned(seeds[ids])[0], pinned(n_iter[ids])[0]
        h_rows = pinned(np.empty((len(seeds), 2), np.float32))[0] if rank == 0 else None
        h_cnt = pinned(np.empty((1, 50, 50), np.int64))[0] if rank == 0 else None
        h2d = x.nbytes + Q.nbytes + hs.numel() * 4 + hn.numel() * 4
        d2h = len(seeds) * 8 + 50 * 50 * 8
    else:
        hax = pinned(axis)[0]
        h_full = (pinned(np.empty(tuple(full.shape), np.float32 if kind == "field" else np.float16))[0]
                  if rank == 0 else None)
        h2d = x.nbytes + Q.nbytes + 3 * axis.nbytes
        d2h = full.numel() * full.element_size()

    def step_e2e():
        dx.copy_(hx, non_blocking=True)
        dq.copy_(hq, non_blocking=True)
        if kind == "topo":
            dseeds.copy_(hs, non_blocking=True)
            dnit.copy_(hn, non_blocking=True)
        else:
            dax.copy_(hax, non_blocking=True)
        step()
        if rank == 0:
            if kind == "topo":
                h_rows.copy_(full, non_blocking=True)
                h_cnt.copy_(dcounts, non_blocking=True)
            else:
                h_full.copy_(full, non_blocking=True)
        _sync()

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    t_e2e = time.perf_counter() - t0

    # ---- parity: rank 0 recomputes the whole frame on its own GPU -------------------------------------
    ok, what = True, ""
    if rank == 0 and world > 1:
        eng.set_charges(dx, dq)
        if kind == "topo":
            one = eng.topo_batch(torch.from_numpy(seeds).to(dev), torch.from_numpy(n_iter).to(dev), prm["h"], dims)
            ok = bool(torch.equal(one, full)) and bool(torch.equal(eng.hist2d(one, de, ce), dcounts[0]))
            what = "gathered rows (seed order) and histogram == the unsharded single-GPU run of the same frame, bit for bit"
        else:
            # the launcher's charge-range splits depend on the slab size, so the FP64 partial sums meet in another
            # order: compare within the field budget instead of bit for bit
            one = (eng.field_lattice(dax, dax, dax, soften=True, concat=True) if kind == "field"
                   else eng.esp_lattice(dax, dax, dax, concat_half=True))
            k0 = 3
            a, b = one[:, k0:].float(), full[:, k0:].float()
            err = float((a - b).abs().max() / a.abs().max())
            same_xyz = bool(torch.equal(one[:, :k0], full[:, :k0]))
            tol = 2e-6 if kind == "field" else 1.5e-3          # ESP rows are float16 (1 ulp = 9.8e-4 relative)
            ok = same_xyz and err <= tol
            what = (f"gathered rows vs the unsharded single-GPU run: coordinates identical, max-norm relative "
                    f"difference {err:.2e} (tolerance {tol:g})")
            del one, a, b

    # ---- reduce over ranks --------------------------------------------------------------------------------
    mine = torch.tensor([[t_dev, t_e2e, my_pairs, float(np.mean(kt)) if len(kt) else 0.0, g_ms]],
                        dtype=torch.float64, device=dev)
    allr = mine
    if world > 1:
        allr = torch.empty((world, 5), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
    allr = allr.cpu().numpy()
    if rank != 0:
        return None
    t_dev, t_e2e = float(allr[:, 0].max()), float(allr[:, 1].max())
    pairs_frame = float(allr[:, 2].sum())
    flops = ESP_FLOPS_PER_PAIR if kind == "esp" else FLOPS_PER_PAIR
    k_ms = float(allr[:, 3].max())
    value = pairs_frame * args.steps / t_dev
    gather_bytes = (len(seeds) * 8 if kind == "topo" else full.numel() * full.element_size())
    clocks = sampler.summary()
    clocks["remeasured"] = False
    peak = max(eng.fp32_peak_tflops(True), eng.fp32_peak_tflops(False))
    achieved = float(allr[0, 2]) * flops / (float(allr[0, 3]) * 1e-3) / 1e12
    return {
        "metric": "pair-evals/s", "value": value, "unit": "pair-evals/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "name": args.workload, "charges": int(len(Q)), "split": args.split,
                   "units_per_step": int(units), "pair_evals_per_step": int(pairs_frame),
                   "l2": "flushed between timed steps (256 MiB write)",
                   "parallelism": (f"one frame, streamlines dealt over {world} GPUs (LPT serpentine deal, plan made once "
                                   "on the host from n_iter), rows all-gathered + restored to seed order + histogram "
                                   "on every rank" if kind == "topo" else
                                   f"one mesh, {world} slabs of x-planes, rows all-gathered on every rank")},
        "units_per_s": units * args.steps / t_dev, "units": "streamlines" if kind == "topo" else "grid points",
        "fp32_frac_of_nominal": value / world * flops / 1e12 / NOMINAL_FP32_TFLOPS,
        "clocks": clocks,
        "e2e": {"value": pairs_frame * args.steps / t_e2e, "unit": "pair-evals/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": t_e2e / args.steps * 1e3,
                "api": "Engine (device pointers) under pycpet_b200.sharding: per step every rank uploads the frame's "
                       "charges and its shard's inputs from pinned host memory, rank 0 copies the gathered result "
                       "back to pinned host memory"},
        "gpu_launches": int(args.steps * (c["launches"] + 1 + (2 if kind == "topo" else 0))),
        "parity_checked": ok if world > 1 else None, "parity": what,
        "limiter": {"rank_kernel_ms": [round(float(v), 4) for v in allr[:, 3]],
                    "gather_ms": [round(float(v), 4) for v in allr[:, 4]], "gather_bytes": int(gather_bytes),
                    "gather_gb_per_s": (gather_bytes * (world - 1) / world / (g_ms * 1e-3) / 1e9 if g_ms > 0 else None),
                    "note": ("kernel_ms differs between ranks by the tail of the dealt queue; gather_ms includes the "
                             "wait for the slowest rank" if kind == "topo" else
                             "gather_ms = all_gather_into_tensor of the slabs, including the wait for the slowest rank")},
        "roofline": {"bound": "fp32-non-tensor" if kind != "esp" else "mufu (fp32 fraction quoted at 11 flop/pair)",
                     "kernel": ((eng.last_path() + "_topo_kernel") if kind == "topo" else
                                ("k1_lattice_nodes_kernel" if kind == "field" else "k1_lattice_kernel")),
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "peak_nominal": NOMINAL_FP32_TFLOPS, "frac_nominal": achieved / NOMINAL_FP32_TFLOPS,
                     "flops_per_pair": flops, "kernel_ms": float(allr[0, 3]), "traffic": None},
    }


# ------------------------------------------------------------------------------------------------
# CPU reference arm (oracle/_ref = the reference's own math_module.c; oracle port otherwise)
# ------------------------------------------------------------------------------------------------
_CREDIT = {}     # (frame, sample size) -> pair-evaluations credited; the accounting pass runs once per sample


def cpu_run_sample(kind, inp, n_units, threads, use_ref):
    from oracle import f64, ref

    x, Q = inp["x"], inp["Q"]
    if kind == "topo":
        sel = np.linspace(0, len(inp["seeds"]) - 1, n_units).astype(np.int64)
        seeds, n_iter = inp["seeds"][sel], inp["n_iter"][sel]
        key = (id(inp["x"]), int(n_units))
        if key not in _CREDIT:      # steps taken per line (float64 port, all cores), work accounting only
            _, steps = f64.topo_batch(seeds, n_iter, x, Q, inp["h"], inp["dims"])
            _CREDIT[key] = float((steps.astype(np.int64) + 2).sum()) * len(Q)
        credited = _CREDIT[key]
        t0 = time.perf_counter()
        if use_ref:
            ref.topo(seeds, n_iter, x, Q, inp["h"], inp["dims"], threads=threads)
        else:
            f64.topo_batch(seeds, n_iter, x, Q, inp["h"], inp["dims"])
        return time.perf_counter() - t0, credited
    sel = np.linspace(0, len(inp["points"]) - 1, n_units).astype(np.int64)
    pts = inp["points"][sel]
    t0 = time.perf_counter()
    if kind == "field":
        ref.field_grid(pts, x, Q, threads=threads) if use_ref else f64.field_grid(pts, x, Q, True)
    else:
        ref.esp_grid(pts, x, Q, threads=threads) if use_ref else f64.esp_grid(pts, x, Q)
    return time.perf_counter() - t0, float(len(pts)) * len(Q)


def host_cores() -> int:
    """Cores this process may actually use: the affinity mask, capped by a cgroup v2 CPU quota."""
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as fh:
            quota, period = fh.read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except (OSError, ValueError):
        pass
    return max(1, n)


def cpu_baseline(kind, prm, inp, budget_s=12.0, steps=1, warmup=0):
    from oracle import f64, ref

    use_ref = ref.available()
    threads = host_cores()
    # the float64 port (work accounting; the timed arm itself when oracle/_ref is absent) runs on all
    # cores whatever OMP_NUM_THREADS says -- torchrun exports 1, and so does this file at N = 1
    f64.set_num_threads(threads)
    total = len(inp["seeds"]) if kind == "topo" else len(inp["points"])
    probe = min(total, 64 * threads)
    # calibration, repeated until two passes agree: the first parallel burst of a process runs several
    # times slower than the steady state (cores idle until then), which would both shrink the sample
    # and under-report the reference
    t, w = cpu_run_sample(kind, inp, probe, threads, use_ref)
    for _ in range(8):
        t2, w = cpu_run_sample(kind, inp, probe, threads, use_ref)
        settled = t2 > 0.8 * t
        t = min(t, t2)
        if settled:
            break
    # second stage: about one second of work, whose rate sizes the sample (the probe is too short to
    # be a steady-state rate on a many-core box)
    n1 = int(min(total, max(probe, probe * min(1.0, budget_s) / max(t, 1e-6))))
    t1, _ = cpu_run_sample(kind, inp, n1, threads, use_ref)
    n = int(min(total, max(probe, n1 * budget_s / max(t1, 1e-6))))
    for _ in range(warmup):
        cpu_run_sample(kind, inp, n, threads, use_ref)
    tt, ww = 0.0, 0.0
    for _ in range(steps):
        t, w = cpu_run_sample(kind, inp, n, threads, use_ref)
        tt += t
        ww += w
    return {"value": ww / tt, "unit": "pair-evals/s", "cores": threads,
            "kind": "reference" if use_ref else "port",
            "sample": f"{n} of {total} {'streamlines' if kind == 'topo' else 'grid points'} of the same frame "
                      f"(evenly strided), {tt:.1f} s; work credited as sum(K+2) x M like the GPU arm"
                      if kind == "topo" else
                      f"{n} of {total} grid points of the same frame (evenly strided), {tt:.1f} s",
            "impl": ("oracle/_ref: reference CPET/utils/math_module.c compiled in place "
                     f"({os.path.basename(ref.path())}), one call per streamline/slab across host threads"
                     if use_ref else "oracle port (cpet_oracle.c, float64, OpenMP)"),
            "ms_per_step": tt / steps * 1e3}


def run_reference(args):
    kind, desc, prm = WORKLOADS[args.workload]
    inp = make_inputs(kind, prm, frame_id=0)
    budget = max(2.0, min(30.0, 150.0 / max(1, args.steps + args.warmup)))
    if args.cpu_seconds is not None:
        budget = args.cpu_seconds
    cb = cpu_baseline(kind, prm, inp, budget_s=budget, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "pair-evals/s", "value": cb["value"], "unit": "pair-evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "name": args.workload, "charges": int(len(inp["Q"]))},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "pair-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="topo3a", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sustained", type=float, default=5.0,
                    help="seconds of back-to-back steps after the timed region (N = 1; 0 = skip)")
    ap.add_argument("--others", default="md1m,volume,esp101,volume2a",
                    help="comma-separated workloads recorded briefly in the N = 1 line's \"others\" ('' = none)")
    ap.add_argument("--split", default="frames", choices=["frames", "seeds", "slab"],
                    help="frames: one frame per GPU per step (weak scaling, default); seeds: ONE topology frame dealt "
                         "over the GPUs by streamlines; slab: ONE box grid split by slabs of x-planes (strong scaling)")
    ap.add_argument("--gather-every", type=int, default=1,
                    help="N > 1, streamline workloads: frames per histogram all-gather (0 = no gathers, diagnostic)")
    ap.add_argument("--cpu-seconds", type=float, default=None,
                    help="CPU work per cpu_baseline sample (default 12 s; reference arm: sized from steps)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    args.others = [w for w in args.others.split(",") if w]
    explicit_workload = any(a.startswith("--workload") for a in sys.argv[1:])
    if explicit_workload or args.split != "frames":
        args.others = []          # side records ride on the default line only

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(run_reference(args)), flush=True)
        return 0

    args.warmup = max(args.warmup, 3)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # launched by hand as `python bench.py --gpus N`: start one rank per GPU ourselves, the way
        # the driver does
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000),
               os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.split == "frames":
            line = run_gpu(args, rank, world, local_rank)
        else:
            line = run_split(args, rank, world, local_rank)
        if rank == 0:
            print(json.dumps(line), flush=True)
    finally:
        if world > 1:
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
