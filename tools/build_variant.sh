#!/bin/bash
# Build a named library variant into build_variants/<name>.so:  tools/build_variant.sh <name> [extra nvcc flags...]
# (set NOFAST=1 to drop -use_fast_math).  Use with CEMK_LIB_PATH=build_variants/<name>.so.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build_variants
FM="-use_fast_math"; [ -n "$NOFAST" ] && FM=""
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 ${SYNCFLAGS--DCEMK_STEP_SYNC -DCEMK_PHASE_SYNC=18} $FM "$@" \
  -Xptxas -v -shared -Xcompiler -fPIC -o build_variants/$name.so manipulator_mujoco_b200/csrc/cemk.cu 2>&1 \
  | grep -A2 "k_rolloutILi20ELi14" | grep -E "registers|spill" | tr '\n' ' '
echo " -> build_variants/$name.so"
