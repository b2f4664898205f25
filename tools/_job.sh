timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; tail -3 gpurun_out/t_all.log
timeout 400 python bench.py --workload C3 --steps 10 --warmup 3 > gpurun_out/r02_bench_C3.json 2> gpurun_out/r02_bench_C3.err; tail -c 300 gpurun_out/r02_bench_C3.err
python - <<'PY'
import json
for w in ("C3",):
    d=json.loads([l for l in open("gpurun_out/r02_bench_%s.json"%w).read().splitlines() if l.startswith("{")][-1])
    print(w, "value %.3g ms %.3f e2e %.3g frac %.3f cpu %.3g parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["parity_sample"]))
PY
