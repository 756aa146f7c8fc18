#!/bin/bash
# Build kernel variants of the library into build/variants/ (shipped to the GPU box, git-ignored), one per NAME=FLAGS pair:
#   tools/sweep_variants.sh base="" fence="-DKT_WAIT_MODE=0" nopre="-DKT_PRETRANSLATED=0" t8="-DKT_PASS_THREADS=1024"
# then on the GPU:  tools/sweep_run.sh [bench args]     (every variant on the same box, back to back, twice)
# Switches (kt_kernels.cuh): KT_WAIT_MODE, KT_PRETRANSLATED, KT_PASS_THREADS, KT_SLOT_CAP, KT_TILE_RECONCILE, KT_HEAVY_PODS.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for v in "$@"; do
  name=${v%%=*}
  flags=${v#*=}
  out=$PWD/build/variants/libkt_$name.so
  make -s -C kube_throttler_b200/csrc -B OUT=$out NVCCFLAGS_EXTRA="$flags" > /dev/null 2>&1
  echo "built $out  [$flags]"
done
