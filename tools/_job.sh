mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_overlap_merge.py -m gpu -q -x > gpurun_out/t_merge.log 2>&1; tail -25 gpurun_out/t_merge.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_overlap_merge.py > gpurun_out/t_all.log 2>&1; tail -8 gpurun_out/t_all.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 400 gpurun_out/r02b_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02b_bench.json").read().splitlines() if l.startswith("{")][-1])
print("value %.3g ms %.3f crc %.3g e2e %.3g frac %.3f cpu %.3g parity %s launches %s" % (d["value"], d["ms_per_step"], d.get("value_crc_per_step",0), d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["parity_sample"], d["gpu_launches"]))
print(d["roofline"])
PY
echo "== 1/8 shard"
TSKV_DEBUG_BINS=1 timeout 200 python tools/profile_scan.py --series 125000 --steps 6 2>&1 | tail -5
