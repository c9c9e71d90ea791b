"""numpy restatement of the elite-exchange kernels (test infrastructure; csrc/cemk.cu k_make_keys / k_pack_sorted /
k_unpack_sorted behind cemk_topk_pack and cemk_merge_packed).  Used by the gloo tests on the CPU and compared bit for bit
with the CUDA kernels in tests/test_gpu_kernels.py."""
import numpy as np

NVAR = 66


def stable_order(cost):
    """jnp.argsort semantics (mjx_planner.py:307): ascending, stable, NaN last."""
    cost = np.asarray(cost)
    key = np.where(np.isnan(cost), np.inf, cost)
    return np.lexsort((np.arange(len(cost)), np.isnan(cost), key))


def topk_pack_ref(cost, idx_base, k, xi):
    """cemk_topk_pack: records [k][NVAR + 2] = xi, cost, global index (as float32) of the k best local samples."""
    loc = stable_order(cost)[:k]
    out = np.empty((k, NVAR + 2), np.float32)
    out[:, :NVAR] = xi[loc]
    out[:, NVAR] = np.asarray(cost, np.float32)[loc]
    out[:, NVAR + 1] = (loc + idx_base).astype(np.float32)
    return out


def merge_packed_ref(packed, k):
    """cemk_merge_packed: the k best candidate rows by (cost, row) -> xi_elite, cost_elite, gidx_elite."""
    sel = stable_order(packed[:, NVAR])[:k]
    return packed[sel, :NVAR].copy(), packed[sel, NVAR].copy(), packed[sel, NVAR + 1].astype(np.int32)
