timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_n8.json 2> gpurun_out/r02_n8.err; python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02_n8.json").read().splitlines() if l.startswith("{")][-1])
    print({k:d[k] for k in ("n_gpus","value","ms_per_step","value_crc_per_step","parity_sample")}, d["run"]["exchange"])
    print(d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["step_ms"], d["roofline"]["slowest_bin"])
except Exception as e:
    print("parse failed", e)
PY
tail -3 gpurun_out/r02_n8.err
