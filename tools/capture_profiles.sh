#!/bin/bash
# Run on the GPU box (gpurun): captures the ncu evidence committed under profiles/.
#   1. launch list of the bench command (every kernel with its device time; cold-cache, serialised)
#   2. one --set full capture of the fused kernels of one step
set -e
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1 || true
ncu --set full --clock-control none --import-source on -k "regex:k_scan_(aggregate|coop)" -s 8 -c 4 -o gpurun_out/${TAG}_scan \
    python tools/profile_scan.py --series 1000000 --steps 4 > gpurun_out/${TAG}_profile_scan.log 2>&1 || true
ncu -i gpurun_out/${TAG}_scan.ncu-rep --page raw --csv > gpurun_out/${TAG}_scan_raw.csv 2>/dev/null || true
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || true
tail -c 600 gpurun_out/${TAG}_bench.json
