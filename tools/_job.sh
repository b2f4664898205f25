mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
grep -n "^E " gpurun_out/t_all.log | head -10
