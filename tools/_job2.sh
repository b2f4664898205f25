mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_concurrency.py -m gpu -q > gpurun_out/t_n2.log 2>&1; tail -5 gpurun_out/t_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02c_n2.json 2> gpurun_out/r02c_n2.err; python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02c_n2.json").read().splitlines() if l.startswith("{")][-1])
    print({k:d[k] for k in ("n_gpus","value","ms_per_step","value_crc_per_step","parity_sample")}, d["run"]["exchange"])
    print(d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["step_ms"])
except Exception as e:
    print("parse failed", e)
PY
tail -5 gpurun_out/r02c_n2.err
