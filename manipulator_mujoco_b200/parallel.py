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


def check_index_range(num_batch: int):
    if num_batch > 1 << 24:
        raise ValueError("global sample index does not fit a float32 mantissa (num_batch > 2^24)")


def pack_elites(xi_e: torch.Tensor, cost_e: torch.Tensor, gidx: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """[k', nvar] , [k'], [k'] int -> [k', nvar + 2] float32.  Global indices must be < 2^24 to be exact in
    float32; the planner checks its batch size once at construction (`check_index_range`) -- reading the
    index tensor here would put a device synchronisation into every CEM iteration."""
    if out is None:
        return torch.cat([xi_e, cost_e[:, None], gidx.to(torch.float32)[:, None]], dim=1).contiguous()
    nvar = xi_e.shape[1]
    out[:, :nvar].copy_(xi_e)
    out[:, nvar].copy_(cost_e)
    out[:, nvar + 1].copy_(gidx)            # int32 -> float32
    return out


def gather_elites(pack: torch.Tensor, world: int, group=None, out: torch.Tensor | None = None) -> torch.Tensor:
    import torch.distributed as dist
    if out is None:
        out = torch.empty(world * pack.shape[0], pack.shape[1], dtype=pack.dtype, device=pack.device)
    dist.all_gather_into_tensor(out, pack, group=group)
    return out


def split_gathered(gathered: torch.Tensor, out=None):
    nvar = gathered.shape[1] - 2
    if out is not None:
        g_cost, g_idx, g_xi = out
        g_cost.copy_(gathered[:, nvar])
        g_idx.copy_(gathered[:, nvar + 1])      # float32 -> int32 (exact below 2^24)
        g_xi.copy_(gathered[:, :nvar])
        return g_cost, g_idx, g_xi
    return (gathered[:, nvar].contiguous(), gathered[:, nvar + 1].to(torch.int32).contiguous(),
            gathered[:, :nvar].contiguous())
