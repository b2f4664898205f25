mkdir -p gpurun_out
timeout 115 python bench.py --workload C5 --series 500000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_C5.json 2> gpurun_out/r02c_bench_C5.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02c_bench_C5.json").read().splitlines() if l.startswith("{")][-1])
    print("C5", "value %.4g %s, %.3f ms/step, e2e %.4g, frac %.4f, parity %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d.get("parity_sample")))
    print(d["config"])
except Exception as e:
    print("C5 parse failed", e)
PY
tail -3 gpurun_out/r02c_bench_C5.err
