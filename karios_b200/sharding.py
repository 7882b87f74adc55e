"""Multi-GPU layout of the path: one process per GPU, scene pairs (or tiles)
round-robin over ranks, no collective on the data path; ONE collective at the end
of a batch collects the match tables and the statistics moments on every rank
(torch.distributed: NCCL over NVLink on the GPUs, gloo in the CPU tests).

Exchange format.  Every unit (a tile of a scene pair) owns one fixed-size record
of `unit_words(cap)` 32-bit words:

    words [0, 32)               kr_unit_header (include/karios_b200.h): row count, flags,
                                n / sum dx / sum dy / sum dx^2 / sum dy^2 / min / max (float64)
    words [32, 32 + 5 capp)     float32 columns x0, y0, dx, dy, score (column stride capp)
    words [32 + 5 capp, + 2 capp)   float64 column zncc_score

(capp = cap rounded up to a multiple of 4).  28 bytes per row carry exactly the bits
of the DataFrame columns.  SceneMatcher.match_many lets the kernels write rows and
header straight into such records (its result arena IS the payload), so the exchange
is a single `all_gather_into_tensor` of the arena: no packing pass, no count
collective, no host synchronisation before it; the moments of the whole batch are a
sum over the gathered headers.  The collective is issued on a side stream as soon as
the last unit of the rank is enqueued.

The reference has no multi-process mode; the merged table is what its single
process would have appended tile by tile (karios/api/core.py:912-919): units in
index order, rows of a unit in their (x0, y0) order.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist

COLUMNS = ("x0", "y0", "dx", "dy", "score", "zncc_score")
HEADER_WORDS = 32                      # sizeof(kr_unit_header) / 4


def assign(n_units: int, rank: int, world: int) -> List[int]:
    """Units (scene pairs or tiles) of this rank: i with i % world == rank."""
    return [i for i in range(n_units) if i % world == rank]


def padded_cap(cap: int) -> int:
    return (int(cap) + 3) // 4 * 4


def unit_words(cap: int) -> int:
    return HEADER_WORDS + 7 * padded_cap(cap)


def new_arena(n_units: int, cap: int, device) -> torch.Tensor:
    """[n_units, unit_words(cap)] float32; headers zeroed (count 0), rows uninitialised."""
    a = torch.empty((max(int(n_units), 1), unit_words(cap)), dtype=torch.float32, device=device)
    a[:, :HEADER_WORDS].zero_()
    return a


def unit_views(arena: torch.Tensor, u: int, cap: int):
    """(header [32] float32 words, rows [5, capp] float32, zncc [capp] float64) of unit u."""
    capp = padded_cap(cap)
    rec = arena[u]
    return (rec[:HEADER_WORDS], rec[HEADER_WORDS:HEADER_WORDS + 5 * capp].view(5, capp),
            rec[HEADER_WORDS + 5 * capp:HEADER_WORDS + 7 * capp].view(torch.float64))


def write_header(arena: torch.Tensor, u: int, cap: int, n: int, flags: int = 0) -> None:
    """kr_unit_header of unit u from its first n rows, with torch operations (what
    kr_unit_header_write does on the device; used by the CPU tests and by callers whose
    rows were not produced through match_many)."""
    hdr, rows, _ = unit_views(arena, u, cap)
    dx, dy = rows[2, :n].to(torch.float64), rows[3, :n].to(torch.float64)
    inf = float("inf")
    vals = torch.tensor([float(n), float(dx.sum()), float(dy.sum()), float((dx * dx).sum()),
                         float((dy * dy).sum()),
                         float(dx.min()) if n else inf, float(dy.min()) if n else inf,
                         float(dx.max()) if n else -inf, float(dy.max()) if n else -inf]
                        + [0.0] * 6, dtype=torch.float64)
    hdr.view(torch.int32)[0] = int(n)
    hdr.view(torch.int32)[1] = int(flags)
    hdr[2:].view(torch.float64).copy_(vals.to(hdr.device))


def exchange(arena: torch.Tensor, group=None, stream=None) -> torch.Tensor:
    """The one collective: every rank contributes its [k, words] arena (same k on every
    rank) and receives [world, k, words].  With `stream` (CUDA) the collective is enqueued
    there; the caller orders it after its producers and before its consumers."""
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(arena.shape), dtype=arena.dtype, device=arena.device)
    if stream is not None:
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(out.view(-1, arena.shape[1]), arena, group=group)
    else:
        dist.all_gather_into_tensor(out.view(-1, arena.shape[1]), arena, group=group)
    return out


def match_many_exchange(sm, pairs, group=None, side_stream=None):
    """SceneMatcher.match_many over this rank's pairs + the exchange, overlapped: the
    collective is enqueued on `side_stream` as soon as the rank's last unit is enqueued (it
    waits for the unit streams on the device, the host goes on finalising units).  A unit
    that needed the exact re-run (select_incomplete, rare) was sent stale: the exchange is
    then repeated after the call.  -> (gathered [world, k, words], total rows of this rank)"""
    dev = sm.device
    side = side_stream if side_stream is not None else torch.cuda.Stream(device=dev)
    box = {}

    def hook(arena, streams):
        for st in streams:
            side.wait_stream(st)
        box["g"] = exchange(arena, group, side)

    sm.on_last_enqueued = hook
    try:
        _, total = sm.match_many(pairs)
    finally:
        sm.on_last_enqueued = None
    cur = torch.cuda.current_stream(dev)
    cur.wait_stream(side)
    if sm.n_redo or "g" not in box:
        box["g"] = exchange(sm.last_arena, group, None)
    return box["g"], total


def headers(gathered: torch.Tensor):
    """(counts [world, k] int32, flags [world, k] int32, moments [world, k, 9] float64) --
    views of the gathered records (no copy, no synchronisation)."""
    h = gathered[..., :HEADER_WORDS]
    ints = h.view(torch.int32)
    return ints[..., 0], ints[..., 1], h[..., 2:20].view(torch.float64)


def batch_moments(gathered: torch.Tensor) -> torch.Tensor:
    """[9] float64 on the device: n, sum dx, sum dy, sum dx^2, sum dy^2, min dx, min dy,
    max dx, max dy over every unit of every rank (the all_reduce of the former design,
    now a local reduction of the gathered headers)."""
    _, _, m = headers(gathered)
    m = m.reshape(-1, 9)
    empty = (m[:, :1] == 0)                 # unused records have an all-zero header
    inf = torch.full_like(m[:, 5:7], float("inf"))
    lo = torch.where(empty, inf, m[:, 5:7])
    hi = torch.where(empty, -inf, m[:, 7:9])
    return torch.cat([m[:, :5].sum(0), lo.min(0).values, hi.max(0).values])


def moments_dict(m) -> dict:
    """accuracy-statistics style summary of batch_moments (host side, after one D2H)."""
    m = [float(v) for v in (m.cpu() if isinstance(m, torch.Tensor) else m)]
    n = m[0]
    if n == 0:
        return {"n": 0}
    mean = (m[1] / n, m[2] / n)
    var = (max(m[3] / n - mean[0] ** 2, 0.0), max(m[4] / n - mean[1] ** 2, 0.0))
    return {"n": int(n), "mean_dx": mean[0], "mean_dy": mean[1], "std_dx": var[0] ** 0.5,
            "std_dy": var[1] ** 0.5, "min_dx": m[5], "min_dy": m[6], "max_dx": m[7], "max_dy": m[8]}


def unit_tables(gathered: torch.Tensor, cap: int, n_units: int | None = None):
    """Host-side view of the merged result: list, in unit order (unit u lives at
    [u % world, u // world]), of (rows [5, n] float32, zncc [n] float64) device views.
    Reads the counts (one small D2H -- after the collective, not before it)."""
    world, k = gathered.shape[0], gathered.shape[1]
    counts, _, _ = headers(gathered)
    cnt = counts.cpu().numpy()
    total = world * k if n_units is None else n_units
    flat = gathered.view(world * k, -1)
    out = []
    for u in range(total):
        r, i = u % world, u // world
        _, rows, z = unit_views(flat, r * k + i, cap)
        n = int(cnt[r, i])
        out.append((rows[:, :n], z[:n]))
    return out


def merged_table(gathered: torch.Tensor, cap: int, n_units: int | None = None) -> torch.Tensor:
    """[N, 6] float64 (COLUMNS) of every unit in unit order -- the table the reference's single
    process appends tile by tile (float32 -> float64 is exact)."""
    parts = [torch.cat([r.t().to(torch.float64), z.reshape(-1, 1)], dim=1)
             for r, z in unit_tables(gathered, cap, n_units)]
    if not parts:
        return torch.zeros((0, len(COLUMNS)), dtype=torch.float64, device=gathered.device)
    return torch.cat(parts, dim=0)


# ---------------------------------------------------------------------------
# Host side of a multi-GPU box: keep a rank's pinned staging memory on the NUMA
# node of its GPU.
def gpu_numa_cpus(index: int):
    """(numa node, sorted cpu list) of GPU `index` from sysfs, or (None, None)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:            # 00000000:1B:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None, None
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read().strip())
        return node, cpus
    except Exception:  # noqa: BLE001
        return None, None


def _parse_cpulist(s: str):
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return sorted(out)


def bind_to_gpu_numa(index: int) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node BEFORE it allocates pinned
    memory (first touch then places the pages on that node, so H2D reads do not cross
    sockets).  Returns what was done (for the bench record)."""
    node, cpus = gpu_numa_cpus(index)
    info = {"gpu": index, "numa_node": node, "bound": False}
    if cpus:
        try:
            allowed = sorted(set(cpus) & set(os.sched_getaffinity(0))) or cpus
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, n_cpus=len(allowed))
        except Exception as e:  # noqa: BLE001
            info["error"] = repr(e)
    return info


def to_numpy_table(rows: torch.Tensor, zncc: torch.Tensor) -> np.ndarray:
    """(rows [5, n], zncc [n]) -> [n, 6] float64 NumPy table."""
    return np.concatenate([rows.t().double().cpu().numpy(), zncc.cpu().numpy()[:, None]], axis=1)
