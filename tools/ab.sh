for mb in 4 5 6; do
  echo "== SCAN_MIN_BLOCKS=$mb"
  TSKV_GPU_LIB=$PWD/cnosdb_b200/libtskv_gpu_mb$mb.so python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.3g pts/s, %.3f ms/step, fused %.3f ms, dom %s %.3f ms' % (d['value'], d['ms_per_step'], d['roofline']['fused_phase']['ms'], d['roofline']['kernel'], d['roofline']['ms_per_launch']))"
done
