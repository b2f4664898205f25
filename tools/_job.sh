mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
run() {
  echo "== $*"
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/b.err | tail -1 > gpurun_out/b.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/b.json").read())
print('value %.4g pts/s, %.3f ms/step, fused %.3f ms parity %s e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['parity_sample'], d['e2e']['value']))
PY
}
run A=1
run TSKV_WORKLIST=items
TSKV_DEBUG_BINS=1 TSKV_NO_GRAPH=1 timeout 200 python tools/profile_scan.py --series 1000000 --steps 4 2>&1 | grep -E "prologue|scan " | tail -2
TSKV_DEBUG_BINS=1 TSKV_NO_GRAPH=1 timeout 200 python tools/profile_scan.py --series 125000 --steps 4 2>&1 | grep -E "prologue|scan " | tail -2
timeout 200 python tools/profile_scan.py --series 125000 --steps 6 2>&1 | tail -1
