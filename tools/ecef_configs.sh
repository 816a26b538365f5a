#!/bin/bash
# Kernel-only numbers of the ECEF-output configurations, one line each:
#   tools/ecef_configs.sh [lib.so] > gpurun_out/ecef_configs.txt
lib=${1:-}
run() {
  GTP_B200_LIB=$lib python bench.py "$@" --no-e2e --no-cpu-baseline --steps 5 --warmup 3 --sustained-seconds 0 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if not line.startswith('{'): continue
    d = json.loads(line); p = d['parity']; r = d['roofline']
    print('%-44s %.3e px/s  frac %.3f  kernel %.3f ms  B/px %.2f  max |d| %.1e m  %.1f ulp  nan_mismatch %s bitexact %.3f' % (
        '$*', d['value'], r['frac'], r['kernel_ms'], r['algorithmic_bytes_per_px'], p.get('max_abs_m', -1), p.get('max_ulps', -1),
        p.get('nan_mismatches'), p.get('bit_exact_fraction', -1)))
"
}
run --config ecef_1km_250m --dtype f64
run --config ecef_1km_250m --dtype f64 --kernel generic
run --config ecef_1km_250m --dtype f32
run --config mod03_ecef_1km_250m
