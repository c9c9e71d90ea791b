"""N-GPU == 1-GPU check (run under torchrun on N GPUs of one box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_multi_gpu.py
Every rank plans with the sharded planner (global batch B); rank 0 also plans with an unsharded
planner on its own GPU.  The CEM results (per-iteration best cost, new mean, best trajectory, elite
global indices) must be bit-identical: samples are a function of the global index, rollouts are
deterministic per sample and the merge reproduces the global stable argsort."""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manipulator_mujoco_b200 import cem_planner  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    B, T, iters = 2048, 32, 3
    kw = dict(num_dof=6, num_batch=B, num_steps=T, timestep=0.05, maxiter_cem=iters, num_elite=0.05, w_pos=20.0, w_rot=3.0,
              w_col=80.0, maxiter_projection=10)
    q0, tp, tr = np.array([1.5, -1.8, 1.75, -1.25, -1.6, 0]), np.array([-0.3, -0.3, 0.5]), np.array([0.0, 1, 0, 0])
    with contextlib.redirect_stdout(io.StringIO()):
        pl = cem_planner(**kw, process_group=dist.group.WORLD)
        ref = cem_planner(**kw) if rank == 0 else None
    ok = True
    names = ["cost", "best_cost_g", "best_cost_r", "best_cost_c", "best_vels", "best_traj", "xi_mean"]
    xi_mean = np.zeros(66)
    # four ticks: two eager, the third is captured into a CUDA graph (NCCL collectives included), the fourth replays it
    for tick in range(4):
        out = pl.compute_cem(xi_mean, q0, np.zeros(6), np.zeros(6), tp, tr)
        elite = pl._last_elite[1].cpu().numpy()
        if rank == 0:
            o1 = ref.compute_cem(xi_mean, q0, np.zeros(6), np.zeros(6), tp, tr)
            e1 = ref._last_elite[1].cpu().numpy()
            for n, a, b in zip(names, out[:7], o1[:7]):
                same = np.array_equal(np.asarray(a), np.asarray(b))
                ok &= same
                print(f"tick {tick} {n:12s} identical: {same}")
            same = np.array_equal(elite, e1)
            ok &= same
            print(f"tick {tick} elite index list identical ({len(e1)} of {B}): {same}")
            same = bool(torch.equal(out[8][:, :B // world], o1[8][:, :B // world]))
            ok &= same
            print(f"tick {tick} rank-0 theta shard identical: {same}")
        xi_mean = out[6]
    if rank == 0:
        print("graph captured:", pl._graph is not None, "| overflow samples:", pl.overflow_samples)
        ok &= pl._graph is not None
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(flag.item()) else "FAIL", f"(world={world})")
    sys.stdout.flush()
    pl.close()                      # the captured graph holds NCCL kernels: release it before the communicator
    if ref is not None:
        ref.close()
    torch.cuda.synchronize()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
