"""world_size-2 gloo test (CPU) of the one exchange step of the multi-GPU path:
unit assignment, the fixed-size unit records (header + 28-byte rows), ONE
all_gather of the arena, batch moments from the gathered headers."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from karios_b200 import sharding

CAP = 18          # not a multiple of 4: the records pad their column stride to 20


def _unit_table(u):
    """[n, 6] float64 whose first five columns are float32-representable."""
    g = torch.Generator().manual_seed(100 + u)
    n = [7, 0, 13, 5, 1][u % 5]
    t = torch.rand((n, 6), generator=g, dtype=torch.float64)
    t[:, 0] = torch.arange(n) + 1000 * u
    t[:, :5] = t[:, :5].to(torch.float32).to(torch.float64)
    if n > 2:
        t[2, 5] = float("nan")                   # zncc_score may be NaN
    return t


def _fill(arena, slot, table):
    _, rows, z = sharding.unit_views(arena, slot, CAP)
    rows.fill_(float("nan"))                     # stale memory beyond the count
    z.fill_(float("nan"))
    n = table.shape[0]
    rows[:, :n] = table[:, :5].t().to(torch.float32)
    z[:n] = table[:, 5]
    sharding.write_header(arena, slot, CAP, n)


def _worker(rank, world, port, n_units, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = sharding.assign(n_units, rank, world)
        per_rank = (n_units + world - 1) // world
        arena = sharding.new_arena(per_rank, CAP, "cpu")          # unused records keep count 0
        for slot, u in enumerate(ids):
            _fill(arena, slot, _unit_table(u))
        gathered = sharding.exchange(arena)                        # the one collective
        merged = sharding.merged_table(gathered, CAP, n_units)
        mom = sharding.moments_dict(sharding.batch_moments(gathered))
        counts, flags, _ = sharding.headers(gathered)
        torch.save({"rows": merged, "mom": mom, "ids": ids, "counts": counts.clone(),
                    "flags_zero": bool((flags == 0).all()), "shape": tuple(gathered.shape)},
                   os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n_units", [5, 2, 1])
def test_exchange_world2(tmp_path, n_units):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_units, str(tmp_path)), nprocs=world, join=True)
    want = torch.cat([_unit_table(u) for u in range(n_units)])
    seen = []
    for r in range(world):
        d = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert d["shape"] == (world, (n_units + 1) // 2, sharding.unit_words(CAP))
        got = d["rows"]
        assert got.shape == want.shape                   # same table, unit order, on every rank
        assert torch.equal(torch.isnan(got), torch.isnan(want))
        assert torch.equal(torch.nan_to_num(got), torch.nan_to_num(want))      # bit-exact columns
        assert d["flags_zero"]
        for u in range(n_units):
            assert int(d["counts"][u % world, u // world]) == _unit_table(u).shape[0]
        seen += d["ids"]
        m = d["mom"]
        assert m["n"] == want.shape[0]
        assert abs(m["mean_dx"] - float(want[:, 2].mean())) < 1e-12
        assert abs(m["std_dy"] - float(want[:, 3].std(unbiased=False))) < 1e-9
        assert m["max_dx"] == float(want[:, 2].max()) and m["min_dy"] == float(want[:, 3].min())
    assert sorted(seen) == list(range(n_units))          # every unit has exactly one owner


def test_record_layout():
    assert sharding.assign(7, 0, 4) == [0, 4] and sharding.assign(7, 3, 4) == [3]
    assert sum(len(sharding.assign(64, r, 8)) for r in range(8)) == 64
    # 128-byte header + 28 bytes per (padded) row
    assert sharding.unit_words(20000) * 4 == 128 + 28 * 20000
    assert sharding.padded_cap(18) == 20 and sharding.unit_words(18) == 32 + 7 * 20
    a = sharding.new_arena(3, CAP, "cpu")
    hdr, rows, z = sharding.unit_views(a, 1, CAP)
    assert hdr.shape == (32,) and rows.shape == (5, 20) and z.shape == (20,) and z.dtype == torch.float64
    # the views alias the record (the kernels write rows straight into the payload)
    rows[3, 4] = 2.5
    z[4] = -1.25
    assert a[1, 32 + 3 * 20 + 4] == 2.5
    assert a[1, 32 + 5 * 20:].view(torch.float64)[4] == -1.25
    # an empty batch reduces to n = 0, min = +inf, max = -inf
    sharding.write_header(a, 0, CAP, 0)
    m = sharding.batch_moments(a[None, :1])
    assert m[0] == 0 and np.isinf(m[5].item()) and m[7].item() == -np.inf
    assert sharding.moments_dict(m) == {"n": 0}
