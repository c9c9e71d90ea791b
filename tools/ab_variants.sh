#!/bin/bash
# A/B timing of prebuilt library variants on the GPU box: tools/ab_variants.sh "v0:0,14 v12:0" [batch list]
# Each variant build_variants/<name>.so is copied over the in-tree libcemk.so and timed with tools/time_rollout.py.
cd "$(dirname "$0")/.."
cp manipulator_mujoco_b200/libcemk.so /tmp/libcemk_keep.so
for v in $1; do
  name=${v%%:*}; cta=${v##*:}
  cp build_variants/$name.so manipulator_mujoco_b200/libcemk.so
  echo "== $name"
  python tools/time_rollout.py --cta "$cta" --batch "${2:-4096}" 2>&1 | grep -v Warning
done
cp /tmp/libcemk_keep.so manipulator_mujoco_b200/libcemk.so
