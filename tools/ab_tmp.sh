for cfg in simple_1km_250m simple_1km_500m; do
  python bench.py --config $cfg --dtype f32 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg f32', d['value'], d['roofline']['frac'], d['parity']['bit_exact_fraction'], d['parity']['n_over_tol_and_ulp'], d.get('sustained')['value'])"
done
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "simple" 2>&1 | tail -2
