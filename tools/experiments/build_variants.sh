#!/bin/bash
# tools/experiments/build_variants.sh NAME="-DFLAG=.. -DFLAG2=.." ...
# Builds one librtc_b200.so per variant into variants/NAME/ (git-ignored; travels to the GPU box) for A/B
# runs with tools/experiments/run_variants.sh.
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
for spec in "$@"; do
  name=${spec%%=*}; defs=${spec#*=}
  mkdir -p "$ROOT/variants/$name"
  make -s -j8 -C "$ROOT/raytracing-course_b200/csrc" OUT="$ROOT/variants/$name" BUILD="$ROOT/variants/$name/build" DEFS="$defs" "$ROOT/variants/$name/librtc_b200.so" 2>&1 | grep -i "error" || true
  ls -la "$ROOT/variants/$name/librtc_b200.so"
done
