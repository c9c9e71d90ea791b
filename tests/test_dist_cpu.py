"""world_size-2 (and 4) gloo test of the sample-sharded elite merge (manipulator_mujoco_b200/parallel.py):
local top-k' + all-gather + (cost, row) merge == the reference's global stable argsort top-k,
including cost ties across ranks and NaN costs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from manipulator_mujoco_b200 import parallel

NVAR = 66


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make(B, seed):
    rng = np.random.default_rng(seed)
    cost = rng.uniform(100, 200, B).astype(np.float32)
    cost[rng.integers(0, B, B // 4)] = 123.0          # many exact ties, spread over all ranks
    cost[[5, B - 3]] = np.nan                         # NaN sorts last in jnp.argsort
    xi = rng.normal(size=(B, NVAR)).astype(np.float32)
    return cost, xi


def _order(cost):
    key = np.where(np.isnan(cost), np.inf, cost)
    return np.lexsort((np.arange(len(cost)), np.isnan(cost), key))


def _worker(rank, world, port, B, k, seed, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cost, xi = _make(B, seed)
    lo, hi = parallel.shard_bounds(B, world, rank)
    kl = parallel.local_topk_size(k, hi - lo)
    loc = _order(cost[lo:hi])[:kl]                                    # what cemk_argsort_topk does per GPU
    pack = parallel.pack_elites(torch.from_numpy(xi[lo:hi][loc]), torch.from_numpy(cost[lo:hi][loc]),
                                torch.from_numpy((loc + lo).astype(np.int32)))
    g_cost, g_idx, g_xi = parallel.split_gathered(parallel.gather_elites(pack, world))
    # the planner's allocation-free variants (persistent buffers) must give the same tensors
    nv = xi.shape[1]
    pack2 = parallel.pack_elites(torch.from_numpy(xi[lo:hi][loc]), torch.from_numpy(cost[lo:hi][loc]),
                                 torch.from_numpy((loc + lo).astype(np.int32)), out=torch.empty(kl, nv + 2))
    bits = lambda t: t.contiguous().view(torch.int32)                 # NaN-safe bitwise comparison
    assert torch.equal(bits(pack), bits(pack2))
    gath2 = parallel.gather_elites(pack2, world, out=torch.empty(world * kl, nv + 2))
    c2, i2, x2 = parallel.split_gathered(gath2, out=(torch.empty(world * kl), torch.empty(world * kl, dtype=torch.int32),
                                                     torch.empty(world * kl, nv)))
    assert torch.equal(bits(c2), bits(g_cost)) and torch.equal(i2, g_idx) and torch.equal(bits(x2), bits(g_xi))
    sel = _order(g_cost.numpy())[:k]                                  # what cemk_merge_elites does (cost, row)
    ret[rank] = (g_idx.numpy()[sel].copy(), g_xi.numpy()[sel].copy(), g_cost.numpy()[sel].copy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world,B,k", [(2, 256, 12), (4, 256, 12), (2, 64, 40)])
def test_sharded_merge_equals_global_stable_topk(world, B, k):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, B, k, 7, ret), nprocs=world, join=True)
    cost, xi = _make(B, 7)
    ref = _order(cost)[:k]
    for r in range(world):
        gidx, gxi, gcost = ret[r]
        np.testing.assert_array_equal(gidx, ref)                      # bit-identical elite index lists on every rank
        np.testing.assert_array_equal(gxi, xi[ref])
        np.testing.assert_array_equal(gcost, cost[ref])


def test_index_range_check():
    parallel.check_index_range(1 << 24)
    with pytest.raises(ValueError):
        parallel.check_index_range((1 << 24) + 1)


def test_shard_bounds():
    assert parallel.shard_bounds(65536, 8, 3) == (24576, 32768)
    with pytest.raises(ValueError):
        parallel.shard_bounds(100, 8, 0)
