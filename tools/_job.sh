timeout 400 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; tail -12 gpurun_out/t_all.log
