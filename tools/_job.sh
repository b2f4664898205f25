timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
TSKV_DEBUG_BINS=1 python tools/profile_scan.py --series 1000000 --steps 4 2>&1 | tail -5
TSKV_DEBUG_BINS=1 python tools/profile_scan.py --series 125000 --steps 4 2>&1 | tail -5
