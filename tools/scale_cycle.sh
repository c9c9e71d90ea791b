#!/bin/bash
# The multi-GPU leg of a measurement cycle on an N-GPU box:  gpurun --gpus N -- 'bash tools/scale_cycle.sh N'
# bench.py under torchrun (weak-scaling line + the 65 536-sample leg) -> gpurun_out/f_scaleN.json; at N = 8 also the per-rank stage
# times (tools/mg_profile.py), at N = 2 the sharded-vs-unsharded bit-identity test.  Every command runs under its own timeout.
N=$1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f_scale$N.json 2> gpurun_out/f_scale$N.err; echo "bench rc $?"
if [ "$N" = "8" ]; then timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/mg_profile.py > gpurun_out/f_mgprof$N.txt 2>&1; echo "mgprof rc $?"; fi
if [ "$N" = "2" ]; then timeout 280 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3; fi
tail -1 gpurun_out/f_scale$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N',d['n_gpus'],'value %.4e'%d['value'],'ms',d['ms_per_step'],'e2e %.4e'%d['e2e']['value'],'c5 %.4e'%d['config5']['value'], d['per_rank']['k_rollout_ms'])"
grep -v Warn gpurun_out/f_mgprof$N.txt 2>/dev/null | grep "^rank" | head -8
