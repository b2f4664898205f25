mkdir -p gpurun_out
TSKV_DEBUG_BINS=1 TSKV_NO_GRAPH=1 timeout 200 python tools/profile_scan.py --series 1000000 --steps 4 2>&1 | grep -E "prologue|scan " | tail -2
TSKV_DEBUG_BINS=1 TSKV_NO_GRAPH=1 timeout 200 python tools/profile_scan.py --series 125000 --steps 4 2>&1 | grep -E "prologue|scan " | tail -2
timeout 200 python tools/profile_scan.py --series 1000000 --steps 6 2>&1 | tail -1
timeout 700 bash tools/sanitize.sh memcheck "parts or malformed or merge or overlapping or bool or c4_shape or long_pages"
timeout 500 bash tools/sanitize.sh racecheck "random_pages or merge_tables or scan_bool or c4_shape_cut"
