"""world_size-2 gloo test (CPU) of the one exchange step of the multi-GPU path:
unit assignment, all_gather of match tables, all_reduce of statistics moments."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from karios_b200 import sharding


def _unit_table(u):
    g = torch.Generator().manual_seed(100 + u)
    n = [7, 0, 13, 5, 1][u % 5]
    t = torch.rand((n, 6), generator=g, dtype=torch.float64)
    t[:, 0] = torch.arange(n) + 1000 * u
    return t


def _worker(rank, world, port, n_units, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = sharding.assign(n_units, rank, world)
        tables = [_unit_table(u) for u in ids]
        merged = sharding.gather_matches(ids, tables, n_units)
        allrows = torch.cat(merged) if merged else torch.zeros((0, 6), dtype=torch.float64)
        mine = torch.cat(tables) if tables else torch.zeros((0, 6), dtype=torch.float64)
        mom = sharding.gather_moments(mine)
        # the same exchange from a result arena (SceneMatcher.match_many layout): [k, 5, cap] float32
        # columns + [k, cap] float64 zncc with stale (NaN) rows beyond each unit's count
        cap = 16
        a32 = torch.full((max(len(ids), 1), 5, cap), float("nan"), dtype=torch.float32)
        az = torch.full((max(len(ids), 1), cap), float("nan"), dtype=torch.float64)
        t32 = [t.to(torch.float32).to(torch.float64) for t in tables]       # what float32 columns can hold
        for i, t in enumerate(tables):
            a32[i, :, : t.shape[0]] = t[:, :5].t().to(torch.float32)
            az[i, : t.shape[0]] = t[:, 5]
        merged2, own2 = sharding.gather_units(ids, a32, az, [t.shape[0] for t in tables], n_units)
        rows2 = torch.cat(merged2) if merged2 else torch.zeros((0, 6), dtype=torch.float64)
        own_want = torch.cat([torch.cat([a[:, :5], b[:, 5:]], 1) for a, b in zip(t32, tables)]) if tables \
            else torch.zeros((0, 6), dtype=torch.float64)
        torch.save({"rows": allrows, "mom": mom, "ids": ids, "rows2": rows2,
                    "own_ok": bool(torch.equal(own2, own_want))}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n_units", [5, 2, 1])
def test_gather_world2(tmp_path, n_units):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_units, str(tmp_path)), nprocs=world, join=True)
    want = torch.cat([_unit_table(u) for u in range(n_units)])
    seen = []
    for r in range(world):
        d = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert torch.equal(d["rows"], want)              # same table, unit order, on every rank
        want2 = torch.cat([want[:, :5].to(torch.float32).to(torch.float64), want[:, 5:]], 1)
        assert torch.equal(d["rows2"], want2) and d["own_ok"]      # arena exchange: same table
        seen += d["ids"]
        m = d["mom"]
        assert m["n"] == want.shape[0]
        assert abs(m["mean_dx"] - float(want[:, 2].mean())) < 1e-12
        assert abs(m["std_dy"] - float(want[:, 3].std(unbiased=False))) < 1e-9
        assert m["max_dx"] == float(want[:, 2].max()) and m["min_dy"] == float(want[:, 3].min())
    assert sorted(seen) == list(range(n_units))          # every unit has exactly one owner


def test_assign_round_robin():
    assert sharding.assign(7, 0, 4) == [0, 4] and sharding.assign(7, 3, 4) == [3]
    assert sum(len(sharding.assign(64, r, 8)) for r in range(8)) == 64
    f = torch.arange(10, dtype=torch.float32).reshape(5, 2)
    z = torch.tensor([0.5, float("nan")], dtype=torch.float64)
    p = sharding.pack_rows(f, z)
    assert p.shape == (2, 6) and p.dtype == torch.float64 and np.isnan(p[1, 5].item())
