"""Latency of the elite all-gather alone (torchrun, N GPUs)."""
import os, time
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=dev)
for rows in (204, 3276):
    x = torch.randn(rows, 68, device=dev)
    out = torch.empty(world * rows, 68, device=dev)
    for _ in range(10):
        dist.all_gather_into_tensor(out, x)
    torch.cuda.synchronize()
    evs = []
    for _ in range(50):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); dist.all_gather_into_tensor(out, x); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    # synchronous single calls
    t = []
    for _ in range(20):
        torch.cuda.synchronize(); t0 = time.perf_counter(); dist.all_gather_into_tensor(out, x); torch.cuda.synchronize(); t.append(1e3 * (time.perf_counter() - t0))
    if rank == 0:
        print(f"all_gather {rows}x68 f32, world {world}: device median {ms[len(ms) // 2] * 1e3:.0f} us (min {ms[0] * 1e3:.0f}), host-synchronous median {sorted(t)[10] * 1e3:.0f} us", flush=True)
dist.destroy_process_group()
