#!/bin/bash
# Kernel-only numbers for every SURVEY 8(d) configuration and dtype (one line each):
#   tools/secondary_configs.sh > gpurun_out/secondary_configs.txt
for cfg in cviirs_1km_250m cviirs_1km_500m simple_1km_250m simple_1km_500m cviirs_5km_1km cviirs_5km_1km_270; do
  for dt in f64 f32; do
    python bench.py --config $cfg --dtype $dt --no-e2e --no-cpu-baseline --steps 5 --warmup 3 --sustained-seconds 0 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if not line.startswith('{'): continue
    d = json.loads(line); p = d['parity']; r = d['roofline']
    print('%-20s %s  %.3e px/s  frac %.3f  kernel %.3f ms  B/px %.2f  dlat %.1e dlon %.1e band %.1e over %s nan_ok %s bitexact %.3f clk %s' % (
        '$cfg', '$dt', d['value'], r['frac'], r['kernel_ms'], r['algorithmic_bytes_per_px'], p.get('max_dlat', -1),
        p.get('max_dlon_outside_band', p.get('max_dlon', -1)), p.get('max_dlon_in_band', -1), p.get('n_in_band_over_conditioning_bound'),
        p.get('nan_pattern_equal'), p.get('bit_exact_fraction', -1), d['clocks']))
"
  done
done
