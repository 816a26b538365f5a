#!/bin/bash
# like ab_bench.sh but for one --config: tools/ab_bench_cfg.sh <config> <lib.so>...
cfg=$1; shift
for lib in "$@"; do
  echo "== $cfg $lib"
  GTP_B200_LIB=$lib python bench.py --config $cfg --no-e2e --no-cpu-baseline --steps 10 --warmup 3 --sustained-seconds 0 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if not line.startswith('{'): continue
    d = json.loads(line)
    print('value %.4g px/s  kernel_ms %.3f  frac %.3f  parity dlat %.1e dlon %.1e over %s' % (
        d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'],
        d['parity'].get('max_dlat', -1), d['parity'].get('max_dlon_outside_band', -1), d['parity'].get('n_in_band_over_conditioning_bound')))
"
done
