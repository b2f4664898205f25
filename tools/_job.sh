mkdir -p gpurun_out
for v in 2 4 8; do
TSKV_CRC_BLOCKS_PER_SM=$v timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/b.err | tail -1 > gpurun_out/b.json
python - <<PY
import json
d=json.loads(open("gpurun_out/b.json").read())
print('crc blocks/SM $v: value %.4g pts/s, %.3f ms/step; with CRC per step %.4g pts/s, %.3f ms' % (d['value'], d['ms_per_step'], d['value_crc_per_step'], d['ms_per_step_crc_per_step']))
PY
done
