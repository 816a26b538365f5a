#!/bin/bash
# What the round-end checks run, plus the tables under profiles/:  bash tools/round_end.sh  (one GPU)
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt | cut -c1-200
tools/secondary_configs.sh > gpurun_out/secondary_configs.txt 2>&1
tools/ecef_configs.sh > gpurun_out/ecef_configs.txt 2>&1
python bench.py > gpurun_out/bench_default.jsonl 2> gpurun_out/bench_default.err; cut -c1-400 gpurun_out/bench_default.jsonl
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.jsonl 2> gpurun_out/bench_reference.err; cut -c1-300 gpurun_out/bench_reference.jsonl
