for v in "$@"; do
  echo "== variant $v"
  TSKV_GPU_LIB=$PWD/cnosdb_b200/libtskv_gpu_$v.so python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.3g pts/s, %.3f ms/step, fused %.3f ms' % (d['value'], d['ms_per_step'], d['roofline']['ms_per_launch']))"
done
