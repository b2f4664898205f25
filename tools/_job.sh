mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 bash tools/capture_profiles.sh r02c 2>&1 | tail -3
for w in C1 C2 C3; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_$w.json 2> gpurun_out/r02c_bench_$w.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02c_bench_$w.json").read().splitlines() if l.startswith("{")][-1])
    print("$w", "value %.4g %s, %.3f ms/step, e2e %.4g, frac %.4f, parity %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d.get("parity_sample")))
except Exception as e:
    print("$w parse failed", e)
PY
done
