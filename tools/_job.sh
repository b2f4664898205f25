mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
run() {
  echo "== $*"
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/b.err | tail -1 > gpurun_out/b.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/b.json").read())
print('value %.4g pts/s, %.3f ms/step, fused %.3f ms parity %s' % (d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['parity_sample']))
PY
}
run A=1
run TSKV_GPU_LIB=$PWD/cnosdb_b200/libtskv_gpu_skip64.so
run TSKV_GPU_LIB=$PWD/cnosdb_b200/libtskv_gpu_skip64.so TSKV_PARTS=8
for lib in "" skip64; do
  echo "== 1/8 shard lib=$lib"
  if [ -n "$lib" ]; then export TSKV_GPU_LIB=$PWD/cnosdb_b200/libtskv_gpu_$lib.so; fi
  timeout 200 python tools/profile_scan.py --series 125000 --steps 6 2>&1 | tail -1
  TSKV_PARTS_TARGET=8 timeout 200 python tools/profile_scan.py --series 125000 --steps 6 2>&1 | tail -1
done
