#!/bin/bash
# Compile k_rollout variants (samples per CTA x min CTAs/SM => register cap) on the GPU box and time them.
# usage: tools/sweep_rollout.sh "14:1 12:1 8:1"      (warps per CTA : min CTAs per SM; at most 14 warps = 28 samples fit the shared memory)
set -e
cd "$(dirname "$0")/.."
for v in $1; do
  W=${v%%:*}; M=${v##*:}
  (cd manipulator_mujoco_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 ${SYNCFLAGS--DCEMK_STEP_SYNC -DCEMK_PHASE_SYNC=18} \
     -use_fast_math -DROLLOUT_WARPS=$W -DROLLOUT_MINB=$M $EXTRA -Xptxas -v -shared -Xcompiler -fPIC -o ../libcemk.so cemk.cu 2>&1 \
     | grep -A2 "k_rolloutILi20" | grep -E "registers|spill" | tr '\n' ' ')
  python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('  WARPS=$W MINB=$M  kernel_ms %.2f  iter_ms %.2f  env-steps/s %.3e' % (d['roofline']['kernel_ms'], d['ms_per_step'], d['value']))"
done
