"""Host-side plumbing of the sample-sharded CEM.

Samples are independent through sampling, projection, rollout and cost, so a global batch of B
samples is split contiguously: rank r owns global indices [r*B/G, (r+1)*B/G).  Per iteration each
rank keeps its own k' = min(k, B/G) best (cost, global index, xi) triples, one all-gather (NCCL on
the GPUs, gloo in the CPU tests) concatenates them rank-major, and every rank selects the k best of
the G*k' candidates by (cost, row).  Because each rank's block is already (cost, index)-sorted and
blocks arrive in rank order, the row number orders equal costs exactly like the global sample index
does -- i.e. like the reference's stable ``jnp.argsort`` over the whole batch (mjx_planner.py:307).
Every rank therefore holds the identical elite list and computes the identical mean / covariance;
no broadcast is needed.
"""
from __future__ import annotations

import torch


def shard_bounds(num_batch: int, world: int, rank: int):
    if num_batch % world:
        raise ValueError("num_batch must be divisible by the number of ranks")
    bl = num_batch // world
    return rank * bl, (rank + 1) * bl


def local_topk_size(k: int, batch_local: int) -> int:
    return min(k, batch_local)


def pack_elites(xi_e: torch.Tensor, cost_e: torch.Tensor, gidx: torch.Tensor) -> torch.Tensor:
    """[k', nvar] , [k'], [k'] int -> [k', nvar + 2] float32 (global indices < 2^24 are exact in float32)."""
    if int(gidx.max()) >= 1 << 24:
        raise ValueError("global sample index does not fit a float32 mantissa")
    return torch.cat([xi_e, cost_e[:, None], gidx.to(torch.float32)[:, None]], dim=1).contiguous()


def gather_elites(pack: torch.Tensor, world: int, group=None) -> torch.Tensor:
    import torch.distributed as dist
    out = torch.empty(world * pack.shape[0], pack.shape[1], dtype=pack.dtype, device=pack.device)
    dist.all_gather_into_tensor(out, pack, group=group)
    return out


def split_gathered(gathered: torch.Tensor):
    nvar = gathered.shape[1] - 2
    return (gathered[:, nvar].contiguous(), gathered[:, nvar + 1].to(torch.int32).contiguous(),
            gathered[:, :nvar].contiguous())
