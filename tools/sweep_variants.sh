#!/bin/bash
# Build kernel-parameter variants of the library into build/variants/ (shipped to the GPU box, git-ignored).
#   tools/sweep_variants.sh "128 6" "64 6" ...     then on the GPU: tools/sweep_run.sh
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for v in "$@"; do
  set -- $v
  out=$PWD/build/variants/libkt_t$1_h$2.so
  make -s -C kube_throttler_b200/csrc -B OUT=$out NVCCFLAGS_EXTRA="-DKT_TILE_RECONCILE=$1 -DKT_HEAVY_PODS=$2" > /dev/null 2>&1
  echo built $out
done
