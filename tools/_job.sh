mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; tail -4 gpurun_out/t_all.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/b.err | tail -1 > gpurun_out/b.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/b.json").read())
print('value %.4g pts/s, %.3f ms/step, fused %.3f ms; crc/step %.4g; e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['value_crc_per_step'], d['e2e']['value']))
PY
timeout 600 bash tools/sanitize.sh memcheck "parts or malformed or merge or overlapping or bool or verify_on_read or item_driven or many_column or small or tombstone or predicates or pruning"
timeout 400 bash tools/sanitize.sh racecheck "item_driven or many_column or c4_shape_cut or merge_tables or verify_on_read or scan_bool"
