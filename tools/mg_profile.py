"""Where a multi-GPU CEM iteration spends its time (run under torchrun on N GPUs)."""
import contextlib, io, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manipulator_mujoco_b200 import cem_planner, jax_prng  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
pg = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD
B, T = 4096 * world, 100
with contextlib.redirect_stdout(io.StringIO()):
    pl = cem_planner(num_dof=6, num_batch=B, num_steps=T, timestep=0.05, maxiter_cem=1, num_elite=0.05, w_pos=20.0, w_rot=3.0,
                     w_col=80.0, maxiter_projection=10, device=dev, process_group=pg)
Bl = B // world
z6 = torch.zeros(6, device=dev)
q0 = torch.tensor([1.5, -1.8, 1.75, -1.25, -1.6, 0.0], device=dev)
tp = torch.tensor([-0.3, -0.3, 0.5], device=dev)
tr = torch.tensor([0.0, 1.0, 0.0, 0.0], device=dev)
st = torch.cat([q0, z6, z6, z6, z6]).unsqueeze(0).expand(Bl, 30).contiguous()
mean0, cov0 = torch.zeros(pl.nvar, device=dev), 10 * torch.eye(pl.nvar, device=dev)
key1 = jax_prng.split(pl.key)[0]
carry = (q0, z6, tp, tr, mean0, cov0, key1, st)
for _ in range(5):
    pl.cem_iter(carry, None)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
n = 20
t0 = time.perf_counter()
for _ in range(n):
    pl.cem_iter(carry, None)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"rank {rank}: host enqueue {1e3 * (t1 - t0) / n:.2f} ms/iter, total {1e3 * (t2 - t0) / n:.2f} ms/iter", flush=True)
# sections, synchronised
def sect(fn):
    torch.cuda.synchronize(); a = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return r, 1e3 * (time.perf_counter() - a)
tot = {}
for _ in range(5):
    (xi, key), t = sect(lambda: pl.compute_xi_samples(key1, mean0, cov0)); tot["sample"] = tot.get("sample", 0) + t
    (xf, td), t = sect(lambda: pl._project(xi, st, True)); tot["project"] = tot.get("project", 0) + t
    (out), t = sect(lambda: pl._rollout(td, q0, z6, tp, tr, False)); tot["rollout"] = tot.get("rollout", 0) + t
    (el), t = sect(lambda: pl._select_elites(out[1], xi)); tot["select+gather+merge"] = tot.get("select+gather+merge", 0) + t
    (mc), t = sect(lambda: pl.compute_mean_cov(el[1], mean0, cov0, el[0])); tot["mean_cov"] = tot.get("mean_cov", 0) + t
if rank == 0:
    print({k: round(v / 5, 3) for k, v in tot.items()})
if world > 1:
    import ctypes as C
    from manipulator_mujoco_b200 import parallel, _lib
    from manipulator_mujoco_b200.planner import _ptr
    k = pl.ellite_num; kl = min(k, Bl); nv = pl.nvar; n = world * kl
    sub = {}
    def add(name, t): sub[name] = sub.get(name, 0) + t
    for _ in range(5):
        (r1), t = sect(lambda: pl._argsort_topk(out[1], 4, Bl, kl, xi, idx_base=rank * Bl)); add("topk", t)
        xi_e, idx, cost_e = r1
        pack = pl._buf("elite_pack", (kl, nv + 2))
        (_), t = sect(lambda: parallel.pack_elites(xi_e, cost_e, idx[:kl], out=pack)); add("pack", t)
        gathered = pl._buf("elite_gathered", (world * kl, nv + 2))
        (_), t = sect(lambda: parallel.gather_elites(pack, world, pg, out=gathered)); add("gather", t)
        (sp), t = sect(lambda: parallel.split_gathered(gathered, out=(pl._buf("elite_gcost", (n,)), pl._buf("elite_gidx", (n,), torch.int32), pl._buf("elite_gxi", (n, nv))))); add("split", t)
        g_cost, g_idx, g_xi = sp
        def alloc():
            np2 = 1 << max(0, (n - 1).bit_length())
            return pl._buf("keys_merge", (np2,), torch.int64), torch.empty(k, nv, device=dev), torch.empty(k, device=dev), torch.empty(k, dtype=torch.int32, device=dev)
        (bufs), t = sect(alloc); add("alloc", t)
        keys, xi_m, cost_m, gidx_m = bufs
        (_), t = sect(lambda: _lib.check(pl._lib.cemk_merge_elites(pl._h, n, _ptr(g_cost), _ptr(g_idx), _ptr(g_xi), _ptr(keys), k, _ptr(xi_m), _ptr(cost_m), _ptr(gidx_m), pl._stream()), pl._lib)); add("merge", t)
        (_), t = sect(lambda: pl._select_elites(out[1], xi)); add("whole", t)
    if rank == 0:
        print({k2: round(v / 5, 3) for k2, v in sub.items()}, "n", n, "k", k)
if world > 1:
    dist.destroy_process_group()
