mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bool.py -m gpu -q > gpurun_out/t_bool.log 2>&1; tail -25 gpurun_out/t_bool.log
timeout 900 bash tools/capture_profiles.sh r02b 2>&1 | tail -5
