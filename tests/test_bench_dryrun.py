"""CPU: bench.run_gpu's control flow and accounting with the device swapped for host stand-ins.

Nothing here computes a field: the Engine / Math_ops stand-ins only count calls and report a fixed,
frame-dependent amount of work, and CUDA events are replaced by a fake clock.  What is checked is
bench.py's own logic -- the frame rotation, that `value` is (work of the timed steps) / (timed
seconds), the JSON contract keys, and (world_size 2 over gloo) the asynchronous histogram gathers,
the max-over-ranks time and the sum-over-ranks work."""
import argparse
import json
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STEP_MS = 2.0            # fake duration of every timed event bracket
K_MS = 1.5               # fake duration of the dominant kernel


class FakeClock:
    now = 0.0


class FakeEvent:
    def __init__(self):
        self.t = None

    def record(self):
        FakeClock.now += STEP_MS / 2.0
        self.t = FakeClock.now

    def elapsed_time(self, other):
        return other.t - self.t


def _frame_code(tag):
    return int(abs(tag) * 1e6) % 100003


class FakeEngine:
    """Stand-in for pycpet_b200.device.Engine: same method names, no device."""
    instances = []

    def __init__(self, device=None, stream=None):
        self.frame_m = None
        self.frame_tag = 0.0
        self.times = []
        self.calls = {"set_charges": 0, "topo_batch": 0, "hist2d": 0, "grid": 0}
        self.frames_seen = []
        FakeEngine.instances.append(self)

    def set_tuning(self, **kv):
        pass

    def set_charges(self, x, Q):
        self.calls["set_charges"] += 1
        self.frame_m = int(Q.numel())
        self.frame_tag = float(x.reshape(-1)[0])       # distinguishes the jittered frames
        self.frames_seen.append(self.frame_tag)

    def _pairs(self, units):
        # a frame-dependent amount of work, as box exits make it on the device
        return int(units * self.frame_m * (10 + int(abs(self.frame_tag) * 1000) % 7))

    def topo_batch(self, seeds, n_iter, step_size, dimensions, second_diff=False, out=None, steps=None):
        self.calls["topo_batch"] += 1
        self.units = int(seeds.shape[0])
        self.times.append(K_MS)
        if out is None:
            out = torch.zeros((self.units, 2), dtype=torch.float32)
        out.zero_()
        return out

    def hist2d(self, values, d_edges, c_edges, out=None):
        self.calls["hist2d"] += 1
        self.times.append(0.01)
        if out is None:
            out = torch.zeros((1, len(d_edges) - 1, len(c_edges) - 1), dtype=torch.int64)
        out.fill_(_frame_code(self.frame_tag))          # a frame's histogram depends on the frame only
        return out[0] if values.dim() == 2 else out     # like Engine.hist2d: (n, 2) values -> (nd, nc)

    def field_grid(self, x0, soften=True, concat=False, out=None):
        self.calls["grid"] += 1
        self.units = int(x0.shape[0])
        self.times.append(K_MS)
        return out

    def field_lattice(self, xs, ys, zs, soften=True, concat=False, out=None):
        self.calls["grid"] += 1
        self.units = int(xs.numel() * ys.numel() * zs.numel())
        self.times.append(K_MS)
        if out is None:
            out = torch.ones((self.units, 6), dtype=torch.float32)
        else:
            out.fill_(1.0)
        return out

    def esp_grid(self, x0, concat_half=False, out=None):
        self.calls["grid"] += 1
        self.units = int(x0.shape[0])
        self.times.append(K_MS)
        return out

    def esp_lattice(self, xs, ys, zs, concat_half=False, out=None):
        self.calls["grid"] += 1
        self.units = int(xs.numel() * ys.numel() * zs.numel())
        self.times.append(K_MS)
        if out is None:
            out = torch.ones((self.units, 4), dtype=torch.float16)
        else:
            out.fill_(1.0)
        return out

    def last_path(self):
        return "k2p"

    def last_counters(self):
        return {"launches": 4, "pair_evals": self._pairs(self.units), "field_evals": 0}

    def kernel_times(self):
        t, self.times = self.times, []
        return t[-256:]                  # the library's event ring keeps the last 256 launches

    def fp32_peak_tflops(self, packed=True, iters=4096):
        return 73.4 if packed else 72.4


class FakeMath:
    """Stand-in for pycpet_b200.Math_ops (host-pointer arm)."""
    frames_calls = []

    def __init__(self, shared_loc=None, device=None):
        pass

    def set_charges(self, x, Q):
        pass

    def topo_hist_frames(self, frames, seeds, n_iter, d_edges, c_edges, step_size=0.1, dimensions=(1, 1, 1),
                         second_diff=False, want_rows=False, rows_out=None, counts_out=None):
        assert n_iter.shape == (len(frames), len(seeds))
        assert rows_out.shape == (len(frames), len(seeds), 2) and counts_out.shape[0] == len(frames)
        FakeMath.frames_calls.append([float(np.asarray(fx).reshape(-1)[0]) for fx, _ in frames])
        for k, (fx, _) in enumerate(frames):
            counts_out[k] = _frame_code(float(np.asarray(fx).reshape(-1)[0]))
        return rows_out, counts_out

    def field_grid(self, x_0, x=None, Q=None, soften=True, concat=False, out=None):
        return out

    def esp_grid(self, x_0, x=None, Q=None, concat_half=False, out=None):
        return out


def _install(monkey_set):
    """Swap the device-touching pieces of bench.py and the package for the stand-ins."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    import pycpet_b200
    import pycpet_b200.device as pdev

    def host_pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t, t.numpy()

    monkey_set(bench, "_device", lambda local_rank: torch.device("cpu"))
    monkey_set(bench, "_event", FakeEvent)
    monkey_set(bench, "_sync", lambda: None)
    monkey_set(bench, "pinned", host_pinned)
    monkey_set(bench, "cpu_baseline", lambda *a, **k: {"value": 1.0, "unit": "pair-evals/s", "cores": 1,
                                                        "kind": "port", "sample": "stub"})
    monkey_set(pdev, "Engine", FakeEngine)
    monkey_set(pycpet_b200, "Math_ops", FakeMath)
    FakeEngine.instances.clear()
    FakeMath.frames_calls.clear()
    FakeClock.now = 0.0
    # small shapes: 5^3 lines / points on a 300-charge frame
    for name, (kind, desc, prm) in list(bench.WORKLOADS.items()):
        small = dict(prm, m=300, n_axis=5)
        monkey_set(bench.WORKLOADS, name, (kind, desc, small), item=True)
    return bench


def _args(workload, steps=5, warmup=3, **kw):
    return argparse.Namespace(gpus=1, steps=steps, warmup=warmup, workload=workload, impl="b200", cpu_seconds=None, **kw)


CONTRACT_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                 "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline")


@pytest.fixture
def bench_mod(monkeypatch):
    def monkey_set(obj, name, value, item=False):
        if item:
            monkeypatch.setitem(obj, name, value)
        else:
            monkeypatch.setattr(obj, name, value)
    return _install(monkey_set)


def test_topo_accounting_single_rank(bench_mod):
    bench = bench_mod
    steps, warmup = 5, 3
    line = bench.run_gpu(_args("topo3a", steps, warmup), rank=0, world=1, local_rank=0)
    json.dumps(line)                                         # serialisable
    for k in CONTRACT_KEYS + ("cpu_baseline",):
        assert k in line, k
    assert line["n_gpus"] == 1 and line["steps"] == steps and line["vs_baseline"] is None
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["dtype"] == "f32"
    assert "workload" in line["config"] and "model" not in line["config"]
    eng = FakeEngine.instances[0]
    pool = bench.FRAME_POOL
    # probe pass (one step per pool frame) + warm-up + timed steps, each one set_charges -> topo -> hist
    assert eng.calls["set_charges"] == eng.calls["topo_batch"] == eng.calls["hist2d"] == pool + warmup + steps
    assert len(set(eng.frames_seen[:pool])) == pool          # the probe pass visits every pool frame once
    # rotation continues across warm-up into the timed steps: frame (rank + s) mod POOL at step s
    assert eng.frames_seen[pool:2 * pool][:warmup + steps] == (eng.frames_seen[:pool] * 2)[:warmup + steps]
    # value = pair-evals of exactly the timed steps / the fake clock's timed seconds
    tags = eng.frames_seen[-steps:]
    by_tag = {}
    for t in eng.frames_seen[:pool]:
        eng.frame_tag = t
        by_tag[t] = eng._pairs(125)
    want_pairs = sum(by_tag[t] for t in tags)
    timed_s = (steps + 1) * (STEP_MS / 2.0) * 1e-3           # one a->b bracket per step + the drain bracket
    assert line["value"] == pytest.approx(want_pairs / timed_s, rel=1e-12)
    assert line["ms_per_step"] == pytest.approx(timed_s / steps * 1e3)
    assert line["config"]["units_per_step_per_gpu"] == 125
    assert line["units_per_s"] == pytest.approx(125 * steps / timed_s)
    # launches claimed for the timed region: pack + K2 launches + histogram per step
    assert line["gpu_launches"] == steps * (1 + 4 + 1)
    # roofline: 20 flop x mean pair-evals per timed step / the integrator's own duration
    r = line["roofline"]
    assert r["kernel"] == "k2p_topo_kernel" and r["kernel_ms"] == pytest.approx(K_MS)
    assert r["achieved"] == pytest.approx(want_pairs / steps * 20.0 / (K_MS * 1e-3) / 1e12)
    assert r["peak"] == 73.4 and r["frac"] == pytest.approx(r["achieved"] / 73.4)
    # end-to-end arm: one warm-up call and ONE timed call of `steps` frames, in the rotation of the device arm
    assert [len(c) for c in FakeMath.frames_calls] == [3, steps]
    assert FakeMath.frames_calls[1] == (eng.frames_seen[:pool] * 2)[:steps]
    e = line["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] == 125 * 8 + 50 * 50 * 8
    assert e["value"] > 0 and e["unit"] == line["unit"]
    # the end-to-end call's histograms were compared with the device arm's
    assert line["parity_checked"] is True and "end-to-end" in line["parity"]
    assert line["step_ms_by_rank"]["median"] == [pytest.approx(STEP_MS / 2.0)]
    assert "sustained" not in line and "others" not in line


def test_sustained_and_other_workloads_ride_on_the_default_line(bench_mod):
    """N = 1: a sustained run of the same step loop after the timed region and one short record per
    other BASELINE configuration."""
    bench = bench_mod
    steps = 4
    line = bench.run_gpu(_args("topo3a", steps, 3, sustained=0.25, others=["md1m", "volume", "esp101", "topo3a"]),
                         rank=0, world=1, local_rank=0)
    json.dumps(line)
    s = line["sustained"]
    # fake clock: 1 ms per step bracket, batches of 100 steps until 0.25 s of busy time
    assert s["steps"] == 300 and s["seconds"] == pytest.approx(0.3) and s["ms_per_step"] == pytest.approx(1.0)
    eng = FakeEngine.instances[0]
    by_tag = {}
    for t in eng.frames_seen[:bench.FRAME_POOL]:
        eng.frame_tag = t
        by_tag[t] = eng._pairs(125)
    first = bench.FRAME_POOL + 3 + steps                      # steps issued before the sustained loop
    want = sum(by_tag[t] for t in eng.frames_seen[first:first + 300])
    assert s["value"] == pytest.approx(want / 0.3, rel=1e-12)
    assert sorted(line["others"]) == ["esp101", "md1m", "volume"]          # the line's own workload is not repeated
    for name, o in line["others"].items():
        assert o["value"] > 0 and o["e2e"]["value"] > 0 and o["roofline"]["kernel_ms"] == pytest.approx(K_MS)
        assert "cpu_baseline" not in o and "sustained" not in o
    assert line["others"]["esp101"]["roofline"]["mufu_frac"] > 0
    assert line["others"]["volume"]["roofline"]["kernel"] == "k1_grid_kernel"      # 5^3 points: below the mesh threshold


def test_long_runs_keep_the_integrator_times(bench_mod):
    """More timed launches than the library's ring holds: the histogram launches must still be
    left out of the dominant kernel's mean duration."""
    line = bench_mod.run_gpu(_args("topo3a", 140, 3), rank=0, world=1, local_rank=0)
    assert line["roofline"]["kernel_ms"] == pytest.approx(K_MS)


@pytest.mark.parametrize("workload,kernel", [("volume", "k1_grid_kernel"), ("esp101", "k1_grid_kernel"),
                                             ("volume2a", "k1_grid_kernel")])   # 5^3 points: below the mesh threshold
def test_grid_accounting_single_rank(bench_mod, workload, kernel):
    bench = bench_mod
    steps = 4
    line = bench.run_gpu(_args(workload, steps, 3), rank=0, world=1, local_rank=0)
    json.dumps(line)
    for k in CONTRACT_KEYS:
        assert k in line, k
    eng = FakeEngine.instances[0]
    assert eng.calls["grid"] == 1 + 3 + steps and eng.calls["topo_batch"] == 0      # one frame per rank: one probe
    pairs = eng._pairs(125)
    timed_s = (steps + 1) * (STEP_MS / 2.0) * 1e-3
    assert line["value"] == pytest.approx(pairs * steps / timed_s)
    assert line["roofline"]["kernel"] == kernel
    flops = 11.0 if workload == "esp101" else 20.0
    assert line["roofline"]["flops_per_pair"] == flops
    assert line["roofline"]["achieved"] == pytest.approx(pairs * flops / (K_MS * 1e-3) / 1e12)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, size, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)

    def monkey_set(obj, name, value, item=False):
        if item:
            obj[name] = value
        else:
            setattr(obj, name, value)

    bench = _install(monkey_set)
    steps = 7                                     # more than GATHER_WINDOW: the windowed drain is exercised
    line = bench.run_gpu(_args("topo3a", steps, 3), rank=rank, world=size, local_rank=0)
    eng = FakeEngine.instances[0]
    by_tag = {}
    for t in eng.frames_seen[:bench.FRAME_POOL]:
        eng.frame_tag = t
        by_tag[t] = eng._pairs(125)
    mine = sum(by_tag[t] for t in eng.frames_seen[-steps:])
    both = torch.tensor([float(mine)], dtype=torch.float64)
    dist.all_reduce(both)
    # the same with the histograms gathered in batches of 3 frames (7 steps: the last batch is partly filled) and
    # with the gathers disabled (diagnostic)
    FakeEngine.instances.clear()
    line3 = bench.run_gpu(_args("topo3a", steps, 3, gather_every=3), rank=rank, world=size, local_rank=0)
    FakeEngine.instances.clear()
    line0 = bench.run_gpu(_args("topo3a", steps, 3, gather_every=0), rank=rank, world=size, local_rank=0)
    if rank == 0:
        assert line3["parity_checked"] is True and "14 gathered" in line3["parity"]
        assert "batches of 3" in line3["config"]["parallelism"]
        assert line0["parity_checked"] is True and "gathered" not in line0["parity"]
        assert "DIAGNOSTIC" in line0["config"]["parallelism"]
        with open(os.path.join(tmp, "line.json"), "w") as fh:
            json.dump({"line": line, "pairs_all": float(both[0]), "first": eng.frames_seen[0]}, fh)
    else:
        assert line is None                       # only rank 0 reports
        with open(os.path.join(tmp, "rank1.json"), "w") as fh:
            json.dump({"first": eng.frames_seen[0]}, fh)
    dist.destroy_process_group()


def _split_worker(rank, size, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)

    def monkey_set(obj, name, value, item=False):
        if item:
            obj[name] = value
        else:
            setattr(obj, name, value)

    bench = _install(monkey_set)
    out = {}
    for workload, split in (("md1m", "seeds"), ("volume464", "slab"), ("esp464", "slab")):
        FakeEngine.instances.clear()
        line = bench.run_split(_args(workload, 4, 3, split=split), rank=rank, world=size, local_rank=0)
        eng = FakeEngine.instances[0]
        out[workload] = {"line": line, "units": eng.units}
    with open(os.path.join(tmp, f"split{rank}.json"), "w") as fh:
        json.dump(out, fh)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_strong_scaling_splits_over_gloo(tmp_path):
    """--split seeds / slab with world_size 2: the shards partition the frame, `value` credits the
    whole frame once per step, the line says "strong"."""
    mp.spawn(_split_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0 = json.load(open(tmp_path / "split0.json"))
    r1 = json.load(open(tmp_path / "split1.json"))
    timed_s = 4 * 3 * (STEP_MS / 2.0) * 1e-3     # per step: the bracket's closing record + the two gather-timing records inside it
    for workload, n_units in (("md1m", 125), ("volume464", 125), ("esp464", 125)):
        line = r0[workload]["line"]
        assert r1[workload]["line"] is None
        # rank 1 holds about half of the frame (rank 0 ends on the unsharded parity run); the work credited
        # per step is the whole frame's: shards of 125 units sum to a multiple of 125 x the per-unit work
        assert 50 <= r1[workload]["units"] <= 75 and r0[workload]["units"] == n_units
        assert line["config"]["pair_evals_per_step"] % n_units == 0
        assert line["parity_checked"] is True
        for k in CONTRACT_KEYS:
            assert k in line, k
        assert line["scaling"] == "strong" and line["n_gpus"] == 2
        assert line["config"]["units_per_step"] == n_units
        assert line["value"] == pytest.approx(line["config"]["pair_evals_per_step"] * 4 / timed_s)
        assert len(line["limiter"]["rank_kernel_ms"]) == 2
        assert line["e2e"]["d2h_bytes_per_step"] > 0 and line["e2e"]["value"] > 0


@pytest.mark.timeout(300)
def test_two_ranks_over_gloo(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    d = json.load(open(tmp_path / "line.json"))
    line = d["line"]
    for k in CONTRACT_KEYS:
        assert k in line, k
    assert "cpu_baseline" not in line             # rank 0 times the CPU reference at N = 1 only
    assert line["n_gpus"] == 2
    # every gathered histogram of the timed steps (7 steps x 2 ranks) was compared with rank 0's own
    assert line["parity_checked"] is True and "14 gathered" in line["parity"]
    assert len(line["step_ms_by_rank"]["median"]) == 2
    timed_s = (7 + 1) * (STEP_MS / 2.0) * 1e-3    # both ranks report the same fake time; the max is that time
    assert line["value"] == pytest.approx(d["pairs_all"] / timed_s)          # work summed over ranks
    assert line["units_per_s"] == pytest.approx(2 * 125 * 7 / timed_s)
    # the two ranks start the rotation on different frames
    assert json.load(open(tmp_path / "rank1.json"))["first"] != d["first"]
