set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
B="--no-e2e --no-cpu-baseline --steps 1 --warmup 3 --sustained-seconds 0 --granules-per-gpu 8"
ncu --set full --clock-control none --import-source on -k regex:modis_fast_1km_xyz -c 1 -o gpurun_out/r01_xyz python bench.py --config ecef_1km_250m $B > gpurun_out/ncu_xyz.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:modis_fast_5km -c 1 -o gpurun_out/r01_fast5 python bench.py --config cviirs_5km_1km --no-e2e --no-cpu-baseline --steps 1 --warmup 3 --sustained-seconds 0 > gpurun_out/ncu_fast5.log 2>&1
ls -la gpurun_out/*.ncu-rep
