"""bench.py -- rollout env-steps/s and CEM-iteration latency of the B200-native CEM planner.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one full CEM iteration of the planner hot path (sample -> projection filter ->
Bernstein evaluation -> T-step rollout of every sample -> cost -> elite top-k [+ NCCL all-gather
merge] -> mean/covariance update) on the BASELINE.json configuration quoted for the metric:
4096 samples x 100-step horizon per GPU (weak scaling: global batch = 4096 * N).
  value   env-steps/s with all inputs resident in HBM, CUDA-event timed, max over ranks
  e2e     the same through the public API (`cem_planner.compute_cem`, maxiter_cem=1) with HOST numpy
          inputs: pinned H2D of the tick inputs and D2H of the results inside the timed region
  roofline  FP32-FMA roofline of the rollout kernel (dominant kernel), timed live with CUDA events
  cpu_baseline  the in-repo CPU restatement (oracle, float32 build, OpenMP over samples) on the
          host cores of the same box -- NOT the reference's MJX (jax / mujoco cannot be installed here)
`--impl reference` times that CPU restatement as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, T, DT = 4096, 100, 0.05
W_POS, W_ROT, W_COL, ELITE, PROJ_IT = 20.0, 3.0, 80.0, 0.05, 10
Q0 = np.array([1.5, -1.8, 1.75, -1.25, -1.6, 0.0])
TP, TR = np.array([-0.3, -0.3, 0.5]), np.array([0.0, 1.0, 0.0, 0.0])
FLOP_PER_ENV_STEP = 8.0e4        # algorithmic work F_A per env-step, scene A (BASELINE.md section 4 / SURVEY.md 8d)
HBM_BYTES_PER_ENV_STEP = 48.0    # theta + thetadot written per env-step
METRIC = "rollout env-steps/sec (CEM iteration, UR5e+Hand-E scene, 4096 samples x 100 steps per GPU)"


class ClockSampler(threading.Thread):
    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop = gpu, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_rollout_baseline(nthreads, B, seed=0):
    """Time the oracle (float32 build) on B samples x T steps; returns (env-steps/s, seconds)."""
    from manipulator_mujoco_b200.mjcf import load_model
    from oracle.oracle import Oracle
    mc = load_model()
    ora = Oracle(mc, DT, dtype="f32")
    rng = np.random.default_rng(seed)
    # smooth, bounded joint-velocity profiles of the planner's magnitude (|thetadot| <= 0.8)
    ph = rng.uniform(0, 2 * np.pi, size=(B, 6, 1))
    am = rng.uniform(0.1, 0.8, size=(B, 6, 1))
    tt = np.linspace(0, 1, T)[None, None, :]
    td = (am * np.sin(2 * np.pi * tt + ph)).reshape(B, 6 * T)
    t0 = time.perf_counter()
    ora.rollout(td, Q0, np.zeros(6), nthreads=nthreads, want_collision=True)
    dt = time.perf_counter() - t0
    return B * T / dt, dt


def ncu_traffic(batch):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_rollout launch from the committed
    `ncu --set full` capture (profiles/), valid for the default 4096-sample workload only."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_final_k_rollout_metrics.json")) as f:
            return float(json.load(f)["dram_traffic_bytes_per_launch"]) if batch == B_PER_GPU else None
    except Exception:
        return None


def run_reference(args):
    """Reference arm: the CPU restatement of the path (oracle) on all host cores, same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    Bs = 1024                                       # bounded sample of the 4096-sample workload per step
    for _ in range(args.warmup):
        cpu_rollout_baseline(cores, 256)
    times = []
    for _ in range(args.steps):
        v, dt = cpu_rollout_baseline(cores, Bs)
        times.append(dt)
    tot = sum(times)
    val = Bs * T * args.steps / tot
    line = {"metric": METRIC, "value": val, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot / args.steps * (B_PER_GPU / Bs), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"UR5e+Hand-E scene A, {B_PER_GPU} samples x {T} steps, dt={DT}; CPU arm times a {Bs}-sample slice per step",
                       "note": "reference MJX/JAX cannot be installed in this image; this is the in-repo CPU restatement (oracle/mjstep.c, float32, OpenMP)"},
            "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": cores, "kind": "port",
                             "sample": f"{Bs} samples x {T} steps per step (rollout + collision distances), all host threads"},
            "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=B_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    from manipulator_mujoco_b200 import cem_planner
    Bl = args.batch_per_gpu
    Bg = Bl * world
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        pl = cem_planner(num_dof=6, num_batch=Bg, num_steps=T, timestep=DT, maxiter_cem=1, num_elite=ELITE, w_pos=W_POS, w_rot=W_ROT,
                         w_col=W_COL, maxiter_projection=PROJ_IT, device=dev, process_group=pg)
    lib, h = pl._lib, pl._h
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)            # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- device-resident CEM iteration ----------------
    z6 = torch.zeros(6, device=dev)
    q0 = torch.as_tensor(Q0, dtype=torch.float32, device=dev)
    tp = torch.as_tensor(TP, dtype=torch.float32, device=dev)
    tr = torch.as_tensor(TR, dtype=torch.float32, device=dev)
    state_term = torch.cat([q0, z6, z6, z6, z6]).unsqueeze(0).expand(Bl, 30).contiguous()
    mean0 = torch.zeros(pl.nvar, device=dev)
    cov0 = 10 * torch.eye(pl.nvar, device=dev)

    from manipulator_mujoco_b200 import jax_prng
    key1 = jax_prng.split(pl.key)[0]                                   # compute_cem's key (mjx_planner.py:388)

    def device_step():
        carry = (q0, z6, tp, tr, mean0, cov0, key1, state_term)
        return pl.cem_iter(carry, None)

    def timed(fn, k, flush_l2=True):
        evs = []
        for _ in range(k):
            if flush_l2:
                flush.fill_(1.0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize(dev)
        return [a.elapsed_time(b) for a, b in evs]

    for _ in range(args.warmup):
        device_step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    l0 = lib.cemk_launch_count(h)
    ms = timed(device_step, args.steps)
    barrier()
    launches = lib.cemk_launch_count(h) - l0
    tot_ms = torch.tensor([sum(ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot_ms, op=dist.ReduceOp.MAX)
    tot_ms = float(tot_ms.item())
    value = Bg * T * args.steps / (tot_ms * 1e-3)

    # ---------------- end to end through the public API (host inputs, host results) ----------------
    xi_mean_host = np.zeros(pl.nvar, dtype=np.float64)

    def e2e_step():
        return pl.compute_cem(xi_mean_host, Q0, np.zeros(6), np.zeros(6), TP, TR)

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    e2e_ms = []
    for _ in range(args.steps):
        flush.fill_(1.0)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        e2e_step()                                   # returns after its own stream synchronisation
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
    barrier()
    e2e_tot = torch.tensor([sum(e2e_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_tot, op=dist.ReduceOp.MAX)
    e2e_value = Bg * T * args.steps / (float(e2e_tot.item()) * 1e-3)

    # ---------------- dominant kernel alone: the fused rollout + cost ----------------
    xi, _ = pl.compute_xi_samples(key1, mean0, cov0)
    xi_f, thetadot = pl._project(xi, state_term, True)

    def rollout_only():
        pl._rollout(thetadot, q0, z6, tp, tr, False)

    for _ in range(2):
        rollout_only()
    roll_ms = timed(rollout_only, max(3, args.steps))
    if rank == 0:
        sampler.stop = True
        sampler.join(timeout=2)
    roll_avg = float(np.mean(roll_ms))
    steps_per_s_kernel = Bl * T / (roll_avg * 1e-3)
    prop = torch.cuda.get_device_properties(dev)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    sm_max = peaks.get("sm_max_mhz", 1965.0)
    fp32_nominal = prop.multi_processor_count * 128 * 2 * sm_max * 1e6 / 1e12
    import ctypes as C
    meas = C.c_double(0.0)
    lib.cemk_fp32_fma_peak(h, C.byref(meas))
    fp32_peak = meas.value if meas.value > 0 else fp32_nominal
    achieved = steps_per_s_kernel * FLOP_PER_ENV_STEP / 1e12

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"UR5e+Hand-E scene A (ur5e_hande_mjx/scene.xml constants), {Bl} samples/GPU x {T} steps, dt={DT}, "
                               f"order-10 Bernstein, {PROJ_IT} projection iterations, elite {ELITE}, 1 CEM iteration per step",
                   "global_batch": Bg, "horizon": T, "parallelism": f"sample-sharded x{world}", "l2": "flushed (256 MB write) before every timed step"},
        "cem_iter_latency_ms": tot_ms / args.steps,
        "e2e": {"value": e2e_value, "unit": "env-steps/s", "ms_per_step": float(e2e_tot.item()) / args.steps,
                "h2d_bytes_per_step": int(pl.h2d_bytes), "d2h_bytes_per_step": int(pl.d2h_bytes)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "fp32", "kernel": "k_rollout", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp32_peak, "traffic": ncu_traffic(Bl),
                     "peak_source": "FP32 FMA throughput measured live by cemk_fp32_fma_peak (8 independent register chains/thread); "
                                    f"nominal {prop.multi_processor_count} SMs x 128 lanes x 2 x {sm_max} MHz = {fp32_nominal:.1f} TFLOP/s (MEASURED_PEAKS.json holds no FP32 figure)",
                     "peak_nominal": fp32_nominal,
                     "flop_per_env_step": FLOP_PER_ENV_STEP, "kernel_ms": roll_avg, "kernel_env_steps_per_s": steps_per_s_kernel,
                     "hbm_gbs": steps_per_s_kernel * HBM_BYTES_PER_ENV_STEP / 1e9, "hbm_peak_gbs": peaks.get("hbm_gbs")},
        "clocks": sampler.summary(),
    }
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, dt = cpu_rollout_baseline(cores, 1024)
        line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                "sample": f"1024 samples x {T} steps of the same workload (oracle/mjstep.c float32 build, OpenMP, {dt:.1f} s)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
