#!/usr/bin/env python
"""Print the numbers of a bench.py JSON line that the docs quote:  python tools/show_bench.py FILE"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.4g} px/s  frac {d['roofline']['frac']:.3f}  kernel_ms {d['roofline']['kernel_ms']:.3f}  n_gpus {d['n_gpus']}")
sus = d["roofline"].get("sustained") or d.get("sustained")
if sus:
    print(f"sustained {sus['value']:.4g} px/s  frac {sus['roofline_frac']:.3f}  clocks {sus.get('clocks')}")
print("clocks", d.get("clocks"))
print("library", d["config"].get("library"))
e = d.get("e2e") or {}
print("e2e", {k: (f"{v['value']:.4g}" if isinstance(v, dict) and "value" in v else v) for k, v in e.items()
              if k in ("value", "dask", "dask_one_array", "direct", "mod03_f32", "ecef", "h2d_bytes_per_step", "d2h_bytes_per_step")})
for s in d.get("secondary") or []:
    p = s.get("parity") or {}
    print(f"  {s.get('config'):18s} {str(s.get('dtype')):24s} {s.get('value', 0):.4g} px/s  frac {s.get('roofline_frac', 0):.3f}  "
          f"ms {s.get('kernel_ms', 0):.3f}  bit_exact {p.get('bit_exact_fraction')}  lat {p.get('max_dlat')}  lon {p.get('max_dlon_outside_band')}  "
          f"band_over {p.get('n_in_band_over_conditioning_bound')}  f32_over {p.get('n_over_tol_and_ulp')} {s.get('error', '')}")
print("parity", d.get("parity"))
print("cpu_baseline", d.get("cpu_baseline"))
