"""Multi-GPU layer: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

The path shards with no data-path collective: every grid point, every streamline and every MD
frame is independent and needs only the (replicated, <= 16 MB) charge set of its frame
(SURVEY.md section 8e).  Collectives appear once, at the end:

  * fields / ESP      : all_gather of per-rank slabs of the flattened point list
  * single-frame topo : all_gather of per-rank (line id, dist, curv) rows, restored to seed order
  * histograms        : all_reduce(sum) of int64 bin counts; all_reduce(min/max) for global ranges
  * MD-frame batches  : frames dealt round-robin, per-frame results gathered by frame id

Partitioning and collectives are independent of who computes: each function takes a `compute`
callable working on the local shard (``Engine`` methods in production; the CPU tests drive the
same code over gloo with world_size 2).
"""
from __future__ import annotations

import numpy as np


def _dist():
    import torch.distributed as dist

    return dist


def world():
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def slab(n: int, rank: int, size: int):
    """Contiguous slab [lo, hi) of n items for `rank`; sizes differ by at most one."""
    base, rem = divmod(int(n), int(size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def frames_for_rank(n_frames: int, rank: int, size: int):
    """MD frames dealt round-robin (frame f -> rank f % size)."""
    return list(range(rank, int(n_frames), size))


def deal_lines_all(n_iter, size: int, block: int = 32):
    """Line ids of every rank for a single-frame topology: lines sorted by n_iter (descending,
    stable) and dealt in warp-sized blocks in serpentine order, so every rank gets the same mix of
    long and short lines (cost is proportional to the steps taken).  Deterministic in (n_iter, size):
    every rank can name every other rank's lines, so no ids travel with the results."""
    n_iter = np.asarray(n_iter).reshape(-1)
    order = np.argsort(-n_iter, kind="stable")
    blk = np.arange(len(order)) // block
    rnd, pos = np.divmod(blk, size)
    owner = np.where(rnd % 2 == 0, pos, size - 1 - pos)          # serpentine: cancels the sort bias
    return [order[owner == r] for r in range(size)]


def deal_lines(n_iter, rank: int, size: int, block: int = 32):
    """This rank's share of deal_lines_all."""
    return deal_lines_all(n_iter, size, block)[rank]


def _comm_device():
    """Tensors handed to collectives live on the GPU for NCCL and on the host for gloo."""
    import torch

    dist = _dist()
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _as_tensor(a, dev=None):
    import torch

    t = a if torch.is_tensor(a) else torch.as_tensor(np.ascontiguousarray(a))
    if dev is not None and t.device != dev:
        t = t.to(dev)
    return t.contiguous()


def all_gather_blocks(local, counts, out=None):
    """Per-rank row blocks -> one (sum(counts), ...) tensor of the same dtype on every rank, rank
    order.  `counts` (rows per rank) is known to every rank up front -- slabs, line deals and frame
    deals are all deterministic -- so nothing but the payload is exchanged:
      * equal counts: one all_gather_into_tensor straight into `out` (preallocated by the caller or
        here), no staging and no copy;
      * ragged counts: each block padded to the longest in a staging buffer (the blocks differ by
        one row for slabs and frame deals), gathered with one all_gather_into_tensor, and the
        valid rows copied into place."""
    import torch

    dist = _dist()
    rank, size = world()
    counts = [int(c) for c in counts]
    dev = _comm_device() if size > 1 else None
    t = _as_tensor(local, dev)
    if t.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {t.shape[0]} rows, the plan says {counts[rank]}")
    total = sum(counts)
    tail = tuple(t.shape[1:])
    if out is None:
        out = torch.empty((total,) + tail, dtype=t.dtype, device=t.device)
    elif tuple(out.shape) != (total,) + tail or out.dtype != t.dtype:
        raise ValueError("all_gather_blocks: `out` has the wrong shape or dtype")
    if size == 1:
        out.copy_(t)
        return out
    if min(counts) == max(counts):
        dist.all_gather_into_tensor(out, t)
        return out
    n_max = max(counts)
    mine = torch.zeros((n_max,) + tail, dtype=t.dtype, device=t.device)
    mine[: t.shape[0]] = t
    stage = torch.empty((size * n_max,) + tail, dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(stage, mine)
    lo = 0
    for r, n in enumerate(counts):
        out[lo:lo + n] = stage[r * n_max:r * n_max + n]
        lo += n
    return out


def all_gather_rows(local, counts=None):
    """Concatenate per-rank row blocks on every rank.  Without `counts` the block lengths are
    exchanged first (one int64 per rank)."""
    import torch

    dist = _dist()
    rank, size = world()
    if size == 1:
        return _as_tensor(local)
    if counts is None:
        dev = _comm_device()
        n_local = torch.tensor([len(local)], dtype=torch.int64, device=dev)
        ns = torch.empty(size, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(ns, n_local)
        counts = ns.cpu().tolist()
    return all_gather_blocks(local, counts)


def all_reduce_(t, op="sum"):
    import torch

    dist = _dist()
    _, size = world()
    if size == 1:
        return t
    ops = {"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}
    dev = _comm_device()
    buf = t.to(dev).contiguous()
    dist.all_reduce(buf, op=ops[op])
    if buf.data_ptr() != t.data_ptr():
        t.copy_(buf)
    return t


# ------------------------------------------------------------------------------------------------
def grid_sharded(compute, x0, out=None):
    """Field / ESP over a point list sharded by slabs.
    compute(points_slab) -> rows for that slab (torch tensor or ndarray, first dim = points).
    Returns the full result, identical on every rank, in point order (gathered into `out` when given)."""
    rank, size = world()
    lo, hi = slab(len(x0), rank, size)
    counts = [slab(len(x0), r, size)[1] - slab(len(x0), r, size)[0] for r in range(size)]
    return all_gather_blocks(compute(x0[lo:hi]), counts, out=out)


def lattice_sharded(compute, xs, ys, zs, out=None):
    """Field / ESP on the box mesh xs x ys x zs (z fastest, CPET/utils/calculator.py:218-233) sharded
    by slabs of x-planes -- a contiguous slab of the flattened point list, so the gathered rows are in
    the reference's point order.  compute(xs_slab, ys, zs) -> rows of that slab."""
    rank, size = world()
    lo, hi = slab(len(xs), rank, size)
    plane = len(ys) * len(zs)
    counts = [(slab(len(xs), r, size)[1] - slab(len(xs), r, size)[0]) * plane for r in range(size)]
    return all_gather_blocks(compute(xs[lo:hi], ys, zs), counts, out=out)


def topo_sharded(compute, seeds, n_iter, out=None):
    """Single-frame topology sharded by streamlines.
    compute(seeds_subset, n_iter_subset) -> (n,2) [dist|curv] rows in the order given.
    Returns (L,2) in SEED order on every rank: the rows are gathered in rank order and scattered with
    the permutation every rank derives from n_iter alone (deal_lines_all)."""
    import torch

    rank, size = world()
    seeds = np.asarray(seeds).reshape(-1, 3)
    n_iter = np.asarray(n_iter).reshape(-1)
    deal = deal_lines_all(n_iter, size)
    ids = deal[rank]
    local = compute(seeds[ids], n_iter[ids])
    gathered = all_gather_blocks(local, [len(d) for d in deal])
    perm = torch.as_tensor(np.concatenate(deal).astype(np.int64)).to(gathered.device)
    if out is None:
        out = torch.empty_like(gathered)
    out[perm] = gathered
    return out


def hist_sharded(compute_counts, values_local, d_edges, c_edges):
    """Histogram of a line set spread over ranks: local counts then all_reduce(sum)."""
    import torch

    counts = compute_counts(values_local, d_edges, c_edges)
    counts = counts if torch.is_tensor(counts) else torch.as_tensor(np.ascontiguousarray(counts))
    return all_reduce_(counts.clone(), "sum")


def global_ranges(values_local):
    """(dmin, dmax, cmin, cmax) over all ranks -- the global ranges make_histograms needs
    (CPET/utils/calculator.py:664-667)."""
    import torch

    v = values_local if torch.is_tensor(values_local) else torch.as_tensor(np.asarray(values_local))
    v = v.reshape(-1, 2).to(torch.float64)
    if v.shape[0]:
        lo = v.min(dim=0).values
        hi = v.max(dim=0).values
    else:
        lo = torch.full((2,), float("inf"), dtype=torch.float64)
        hi = torch.full((2,), float("-inf"), dtype=torch.float64)
    lo = all_reduce_(lo.clone(), "min")
    hi = all_reduce_(hi.clone(), "max")
    return float(lo[0]), float(hi[0]), float(lo[1]), float(hi[1])


def order_stats_sharded(radix_hist, ranks):
    """Exact global order statistics of values spread over the ranks (the
    global min/max + IQR of make_histograms over every line of every frame without gathering them).
    radix_hist(prefixes, prefix_bits) -> (len(prefixes), 256) histogram of THIS rank's values
    (Engine.radix_hist bound to the local tensor); the histograms are all-reduced after each of the
    four 8-bit passes, so every rank walks the same prefixes.  Returns float32 values."""
    import torch

    ranks = [int(r) for r in ranks]
    k = list(ranks)
    prefix = [0] * len(ranks)
    for p in range(4):
        pre = [0] if p == 0 else prefix
        h = np.asarray(radix_hist(np.asarray(pre, dtype=np.uint32), 8 * p)).astype(np.int64)
        h = all_reduce_(torch.from_numpy(np.ascontiguousarray(h)), "sum").cpu().numpy()
        for t in range(len(ranks)):
            row = h[0 if p == 0 else t]
            cum = np.concatenate([[0], np.cumsum(row)])
            b = int(np.searchsorted(cum, k[t], side="right") - 1)
            if b > 255:
                raise ValueError("rank beyond the number of values")
            k[t] -= int(cum[b])
            prefix[t] = (prefix[t] << 8) | b
    keys = np.asarray(prefix, dtype=np.uint32)
    bits = np.where(keys & np.uint32(0x80000000), keys & np.uint32(0x7FFFFFFF), ~keys)
    return bits.astype(np.uint32).view(np.float32)


def frames_sharded(compute_frame, n_frames):
    """MD-frame batch: this rank runs compute_frame(f) for f = rank, rank+size, ...; the per-frame
    results (equal-shaped arrays/tensors) are gathered and returned stacked in frame order."""
    return frames_batch_sharded(lambda ids: [compute_frame(f) for f in ids], n_frames)


def frames_batch_sharded(compute_batch, n_frames):
    """Same, with one call per rank: compute_batch(ids) gets this rank's frame ids (rank, rank+size,
    ...) and returns one equal-shaped result per id (a list, or an array/tensor with leading
    dimension len(ids)) -- the shape of Math_ops.topo_hist_frames, which overlaps the copies and
    kernels of neighbouring frames.  Results come back stacked in frame order on every rank, in the
    dtype they were computed in (int64 counts stay int64)."""
    import torch

    rank, size = world()
    mine = frames_for_rank(n_frames, rank, size)
    got = compute_batch(mine) if len(mine) else []
    res = [r if torch.is_tensor(r) else torch.as_tensor(np.ascontiguousarray(r)) for r in got]
    if len(res) != len(mine):
        raise ValueError(f"compute_batch returned {len(res)} results for {len(mine)} frames")
    if size == 1:
        return torch.stack(res) if res else torch.zeros(0)
    # a rank without frames (fewer frames than ranks) learns shape and dtype from the others
    meta = (tuple(res[0].shape), res[0].dtype) if res else None
    metas = [None] * size
    _dist().all_gather_object(metas, meta)
    found = [m for m in metas if m is not None]
    if not found:
        return torch.zeros(0)
    shape, dtype = found[0]
    local = torch.stack(res) if res else torch.zeros((0,) + shape, dtype=dtype)
    counts = [len(frames_for_rank(n_frames, r, size)) for r in range(size)]
    gathered = all_gather_blocks(local, counts)                   # rank-major: frames r, r+size, ...
    out = torch.empty_like(gathered)
    lo = 0
    for r, n in enumerate(counts):
        out[r::size] = gathered[lo:lo + n]
        lo += n
    return out
