#!/bin/bash
# A/B of extra in-step CTA re-alignments (bit 1: before the narrow phase, 2: before the constraint rows, 4: before the free-box pairs, 8: before the warm-start choice, 16: before the line search)
cd "$(dirname "$0")/.."
for v in 0 1 2 4 6; do
  SYNCFLAGS="-DCEMK_STEP_SYNC -DCEMK_PHASE_SYNC=$v" tools/sweep_rollout.sh "14:1" | sed "s/^.*WARPS/PHASE_SYNC=$v WARPS/"
done
