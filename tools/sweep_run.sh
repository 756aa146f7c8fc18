#!/bin/bash
# Run bench.py for every variant library under build/variants/ (two rounds, interleaved, so that drift of the box shows up as
# a difference between the rounds rather than between the variants) and print pass time and the per-kernel device times.
cd "$(dirname "$0")/.."
for round in 1 2; do
  for so in build/variants/*.so; do
    KT_B200_LIB=$PWD/$so python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras --e2e-steps 1 "$@" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline'].get('chained_kernel_ms') or {}; print('round $round', '$so', 'pass_us %.2f' % (d['ms_per_step']*1e3), 'flush_us %.2f' % (d['roofline']['other_timing']['ms_per_step']*1e3), 'e2e %.3g' % d['e2e']['value'], {a: round(b*1e3,1) for a,b in k.items()})"
  done
done
