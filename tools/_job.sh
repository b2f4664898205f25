mkdir -p gpurun_out
run() {
  echo "== $*"
  env "$@" TSKV_DEBUG_BINS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/b.err | tail -1 > gpurun_out/b.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/b.json").read())
print('value %.3g pts/s, %.3f ms/step, fused %.3f ms' % (d['value'], d['ms_per_step'], d['roofline']['ms_per_launch']))
PY
  grep "\[tskv\] bin" gpurun_out/b.err | tail -4 | awk '{printf "%s %s g%s run %s | ", $2,$3,$5,$10} END {print ""}'
}
run TSKV_PARTS=1
run TSKV_PARTS=4
run TSKV_PARTS=4 TSKV_GRID_OVERSUB=1.5
run TSKV_PARTS=4 TSKV_GRID_OVERSUB=2
run TSKV_PARTS=4 TSKV_GRID_OVERSUB=4
run TSKV_PARTS=4 TSKV_GRID_MODE=1
run TSKV_PARTS=8 TSKV_GRID_OVERSUB=2
run TSKV_PARTS=8 TSKV_GRID_MODE=1
run TSKV_PARTS=2 TSKV_GRID_OVERSUB=2
run TSKV_PARTS=2 TSKV_GRID_MODE=1
