#!/bin/bash
# Run bench.py once per variant library under build/variants/ and print the per-kernel device times.
cd "$(dirname "$0")/.."
for so in build/variants/*.so; do
  KT_B200_LIB=$PWD/$so python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms']; print('$so', 'pass_us %.1f' % (d['ms_per_step']*1e3), {a: round(b*1e3,1) for a,b in k.items()})"
done
