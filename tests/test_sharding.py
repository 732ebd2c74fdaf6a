"""Host-side multi-process logic on CPU: world_size-2 gloo run of the shard / all-gather path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ungar_b200 import sharding
from ungar_b200 import workloads as W


def test_shard_ranges_tile_the_batch():
    for total in (0, 1, 7, 1024, 8192, 8193):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _oracle_summaries(xp, model, N, stiffness, eps):
    """Per-trajectory summaries computed by the CPU oracle (checker): u_0, f, Zsoft, |g|_inf, max h."""
    import oracle as oracle_pkg

    orc = oracle_pkg.Oracle()
    L = orc.record_layout(model, N)
    s = orc.sizes(model, N)
    rec = orc.stage_sweep(model, N, xp, stiffness, eps)
    out = np.zeros((xp.shape[0], sharding.SUMMARY_SIZE))
    nX = s["nx"] * (N + 1)
    out[:, :s["nu"]] = xp[:, nX:nX + s["nu"]]
    out[:, 24] = rec[:, L["cost"]]
    out[:, 25] = rec[:, L["cost"] + 1]
    out[:, 26] = np.abs(rec[:, L["g"]:L["g"] + s["m_eq"]]).max(axis=1)
    out[:, 27] = rec[:, L["h"]:L["h"] + s["m_ineq"]].max(axis=1)
    return out


def _worker(rank, world, port, total, ragged, result_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model, N = W.QUADROTOR, 6
        xp_all = W.synthetic_batch(model, N, total, seed=77)  # every rank can regenerate the global batch
        a, b = sharding.shard_range(total, world, rank)
        assert (b - a != total - (b - a)) == ragged  # total % world != 0 gives ragged shards
        local = torch.from_numpy(_oracle_summaries(xp_all[a:b], model, N, 100.0, 2e-5))
        gathered = sharding.gather_summaries(local)
        expect = _oracle_summaries(xp_all, model, N, 100.0, 2e-5)
        assert gathered.shape == (total, sharding.SUMMARY_SIZE)
        assert np.array_equal(gathered.numpy(), expect)
        status = sharding.fleet_status(gathered)
        assert status["trajectories"] == total
        assert np.isclose(status["total_cost"], expect[:, 24].sum())
        assert np.isclose(status["worst_eq_inf_norm"], expect[:, 26].max())
        # overlapped exchange (double-buffered async all-gather): every posted iteration arrives intact, in rank order
        if not ragged:
            ex = sharding.SummaryExchange(b - a, torch.float64, "cpu")
            for it in range(5):
                k = ex.slot()
                ex.local[k].copy_(local + it)
                ex.post(k)
                if it % 2 == 1:
                    assert np.array_equal(ex.latest().numpy(), expect + it)
            ex.drain()
            assert np.array_equal(ex.latest().numpy(), expect + 4)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == float(world)
        open(os.path.join(result_dir, f"ok{rank}"), "w").close()
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("total,ragged", [(8, False), (7, True)])
def test_two_rank_gloo_gather(oracle, tmp_path, total, ragged):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), total, ragged, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_single_process_is_identity():
    x = torch.arange(64, dtype=torch.float64).reshape(2, 32)
    assert sharding.gather_summaries(x) is x
