#!/bin/bash
# A/B of library builds on one box: tools/ab_bench.sh <lib.so>... (kernel-only numbers, short runs)
for lib in "$@"; do
  echo "== $lib"
  GTP_B200_LIB=$lib python bench.py --no-e2e --no-cpu-baseline --steps 10 --warmup 3 --sustained-seconds 1.0 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if not line.startswith('{'): print(line.rstrip()); continue
    d = json.loads(line)
    print('value %.4g px/s  kernel_ms %.3f  frac %.3f  sustained frac %.3f  parity dlat %.1e dlon %.1e over %s' % (
        d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d.get('sustained', {}).get('roofline_frac', 0),
        d['parity'].get('max_dlat', -1), d['parity'].get('max_dlon_outside_band', -1), d['parity'].get('n_in_band_over_conditioning_bound')))
"
done
