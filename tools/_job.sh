mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_page_parts.py -m gpu -q -x -k "item_driven or c4_shape" > gpurun_out/t_p.log 2>&1; tail -3 gpurun_out/t_p.log
run() {
  echo "== $*"
  for n in 1000000 500000; do
    env "$@" timeout 200 python tools/profile_scan.py --series $n --steps 8 2>&1 | tail -1
  done
}
run A=1
run TSKV_PARTS_TS=4 TSKV_PARTS_TARGET=4
run TSKV_PARTS_TS=8 TSKV_PARTS_TARGET=4
run TSKV_PARTS=8 TSKV_PARTS_TS=8
run TSKV_PARTS_TARGET=6
run TSKV_PARTS_TARGET=8
