"""Multi-GPU layout of the path: one process per GPU, scene pairs (or tiles)
round-robin over ranks, no collective on the data path; ONE exchange step at the
end collects the match lists and the statistics moments on every rank
(torch.distributed: NCCL over NVLink on the GPUs, gloo in the CPU tests).

The reference has no multi-process mode; the merged table is what its single
process would have appended tile by tile (karios/api/core.py:912-919): units in
index order, rows of a unit in their (x0, y0) order.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist

COLUMNS = ("x0", "y0", "dx", "dy", "score", "zncc_score")


def assign(n_units: int, rank: int, world: int) -> List[int]:
    """Units (scene pairs or tiles) of this rank: i with i % world == rank."""
    return [i for i in range(n_units) if i % world == rank]


def pack_rows(f32_rows: torch.Tensor, zncc: torch.Tensor) -> torch.Tensor:
    """[5, n] float32 (x0,y0,dx,dy,score) + [n] float64 -> [n, 6] float64 (exact)."""
    return torch.cat([f32_rows.t().to(torch.float64), zncc.reshape(-1, 1).to(torch.float64)], dim=1)


def gather_matches(unit_ids: Sequence[int], unit_rows: Sequence[torch.Tensor], n_units: int,
                   group=None) -> List[torch.Tensor]:
    """Every rank contributes the [n_i, 6] float64 tables of its units; returns, on
    every rank, the list of all n_units tables in unit order.  Two collectives:
    all_gather of the per-unit row counts, all_gather of the padded rows."""
    world = dist.get_world_size(group)
    dev = unit_rows[0].device if len(unit_rows) else _default_device(group)
    counts = torch.zeros(n_units, dtype=torch.int64, device=dev)
    for u, rows in zip(unit_ids, unit_rows):
        counts[u] = rows.shape[0]
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)       # every unit has one owner
    per_rank = (n_units + world - 1) // world
    cnt = counts.cpu().tolist()                                      # the one host synchronisation
    n_max = max(cnt) if n_units else 0
    mine = torch.zeros((per_rank, max(n_max, 1), len(COLUMNS)), dtype=torch.float64, device=dev)
    for slot, (u, rows) in enumerate(zip(unit_ids, unit_rows)):
        mine[slot, : rows.shape[0]] = rows
    allr = torch.empty((world,) + tuple(mine.shape), dtype=torch.float64, device=dev)
    dist.all_gather(list(allr.unbind(0)), mine, group=group)
    # views into the gathered block: unit u lives at [u % world, u // world]
    return [allr[u % world, u // world, : cnt[u]] for u in range(n_units)]


def gather_units(unit_ids: Sequence[int], arena_f32: torch.Tensor, arena_z: torch.Tensor,
                 counts: Sequence[int], n_units: int, group=None):
    """gather_matches for rows that already sit in one arena (SceneMatcher.match_many:
    arena_f32 [k, 5, cap] float32, arena_z [k, cap] float64, counts[i] rows of local unit i):
    the padded [per_rank, n_max, 6] float64 block is built with a handful of batched
    operations instead of several launches per unit.  Same collectives, same result.
    -> (list of all n_units tables in unit order, this rank's own rows [N, 6])."""
    world = dist.get_world_size(group)
    dev = arena_f32.device
    k = len(unit_ids)
    counts_t = torch.zeros(n_units, dtype=torch.int64, device=dev)
    if k:
        counts_t[torch.as_tensor(list(unit_ids), dtype=torch.int64, device=dev)] = \
            torch.as_tensor(list(counts), dtype=torch.int64, device=dev)
    dist.all_reduce(counts_t, op=dist.ReduceOp.SUM, group=group)
    per_rank = (n_units + world - 1) // world
    cnt = counts_t.cpu().tolist()                                    # the one host synchronisation
    n_max = max(max(cnt) if n_units else 0, 1)
    mine = torch.zeros((per_rank, n_max, len(COLUMNS)), dtype=torch.float64, device=dev)
    own = torch.zeros((0, len(COLUMNS)), dtype=torch.float64, device=dev)
    if k:
        w = min(n_max, arena_f32.shape[2])
        mine[:k, :w, :5] = arena_f32[:k, :, :w].transpose(1, 2)      # float32 -> float64 is exact
        mine[:k, :w, 5] = arena_z[:k, :w]
        valid = torch.arange(n_max, device=dev)[None, :] < \
            torch.as_tensor(list(counts), dtype=torch.int64, device=dev)[:, None]
        mine[:k].masked_fill_(~valid[:, :, None], 0.0)               # stale arena rows (may hold NaN bits)
        own = mine[:k][valid]
    allr = torch.empty((world,) + tuple(mine.shape), dtype=torch.float64, device=dev)
    dist.all_gather(list(allr.unbind(0)), mine, group=group)
    return [allr[u % world, u // world, : cnt[u]] for u in range(n_units)], own


def gather_moments(rows: torch.Tensor, group=None):
    """all_reduce of [n, sum dx, sum dy, sum dx^2, sum dy^2] and of the min / max
    of dx, dy -> dict with n, mean, std (population), min, max per component."""
    dev = rows.device
    dx, dy = rows[:, 2], rows[:, 3]
    m = torch.stack([torch.tensor(float(rows.shape[0]), dtype=torch.float64, device=dev),
                     dx.sum(), dy.sum(), (dx * dx).sum(), (dy * dy).sum()])
    inf = float("inf")
    lo = torch.stack([dx.min() if len(dx) else torch.tensor(inf, dtype=torch.float64, device=dev),
                      dy.min() if len(dy) else torch.tensor(inf, dtype=torch.float64, device=dev)])
    hi = torch.stack([dx.max() if len(dx) else torch.tensor(-inf, dtype=torch.float64, device=dev),
                      dy.max() if len(dy) else torch.tensor(-inf, dtype=torch.float64, device=dev)])
    dist.all_reduce(m, op=dist.ReduceOp.SUM, group=group)
    ext = torch.cat([-lo, hi])                                       # min and max in one collective
    dist.all_reduce(ext, op=dist.ReduceOp.MAX, group=group)
    lo, hi = -ext[:2], ext[2:]
    m = m.cpu()
    lo, hi = lo.cpu(), hi.cpu()
    n = float(m[0])
    if n == 0:
        return {"n": 0}
    mean = (m[1] / n, m[2] / n)
    var = (m[3] / n - mean[0] ** 2, m[4] / n - mean[1] ** 2)
    return {"n": int(n), "mean_dx": float(mean[0]), "mean_dy": float(mean[1]),
            "std_dx": float(var[0].clamp_min(0).sqrt()), "std_dy": float(var[1].clamp_min(0).sqrt()),
            "min_dx": float(lo[0]), "min_dy": float(lo[1]), "max_dx": float(hi[0]), "max_dy": float(hi[1])}


def _default_device(group):
    backend = dist.get_backend(group)
    if backend == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")
