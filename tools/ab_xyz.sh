for l in build/variants/*.so; do
  for cfg in ecef_1km_250m mod03_ecef_1km_250m; do
    GTP_B200_LIB=$l python bench.py --config $cfg --dtype f64 --no-e2e --no-cpu-baseline --steps 10 --warmup 3 --sustained-seconds 0 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if not line.startswith('{'): continue
    d = json.loads(line); r = d['roofline']
    print('%-40s %-22s %.3e px/s frac %.3f kernel %.3f ms' % ('$l', '$cfg', d['value'], r['frac'], r['kernel_ms']))
"
  done
done
