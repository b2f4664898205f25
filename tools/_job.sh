timeout 300 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; tail -8 gpurun_out/t_all.log
bash tools/quick_prof.sh q6 2>&1 | tail -16
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/q6_bench.json 2> gpurun_out/q6_bench.err; tail -c 2500 gpurun_out/q6_bench.json; tail -5 gpurun_out/q6_bench.err
