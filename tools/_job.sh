timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
TSKV_DEBUG_BINS=1 python tools/profile_scan.py --series 1000000 --steps 4 2>&1 | tail -5
TSKV_DEBUG_BINS=1 python tools/profile_scan.py --series 125000 --steps 4 2>&1 | tail -5
M="smsp__inst_executed.sum,gpu__time_duration.sum"
ncu --metrics $M --clock-control none -k "regex:k_scan_(aggregate|coop)" -s 8 -c 4 --csv --log-file gpurun_out/q13_inst.csv python tools/profile_scan.py --series 1000000 --steps 3 > /dev/null 2>&1
grep -E "inst_executed|time_duration" gpurun_out/q13_inst.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-120
