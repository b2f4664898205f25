"""Thread-safety of the C ABI (several vnode scans run concurrently on tokio worker threads):
4 host threads issue scans at the same time on ONE context - device-resident and host-resident page sets, through the
one-call entry point and through prepare / run / finalize - and every result must equal the oracle's."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from cnosdb_b200 import cabi, datagen
from cnosdb_b200.engine import PushedAggregate, QueryOption
from oracle import pyoracle as orc
from tests.helpers import ALL_AGGS, assert_results_equal, bucket_spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_concurrent_scans_on_one_context(engine):
    g = datagen.generate(1500, n_fields=2, n_points=400, value_kind=datagen.MIXED, seed=9, jitter_permille=250, jitter_max=999,
                         null_page_permille=50, null_row_permille=100)
    arena = g.arena.copy()
    w = 60_000_000_000
    fbs, nb = bucket_spec(datagen.TSBS_T0 - 1000, datagen.TSBS_T0 + 399 * datagen.TSBS_STEP + 1000, w)
    dev = engine.upload_pages(arena, g.descs)
    host = engine.upload_pages(arena, g.descs, verify_crc=True, host_resident=True)
    queries = []
    for k in range(4):
        sel = np.arange(k, 1500, 3 + k, dtype=np.uint32)
        aggs = ALL_AGGS if k % 2 else ("count", "sum", "min", "max", "mean")
        queries.append(QueryOption([PushedAggregate(1, cabi.TSKV_PT_I64, aggs), PushedAggregate(3, cabi.TSKV_PT_F64, aggs)],
                                   series_ids=sel, width=w, first_bucket_start=fbs, n_buckets=nb,
                                   group_by_series=(k == 3)))
    expected = [orc.scan_aggregate(arena, g.descs, q, n_threads=2) for q in queries]
    errors = []

    def worker(k):
        try:
            for it in range(6):
                pages = host if (it + k) % 2 else dev
                if it % 3 == 2:  # the split entry points, interleaved with the other threads' calls
                    s = engine.prepare(pages, queries[k])
                    s.run()
                    got = s.finalize()
                    s.close()
                else:
                    got = engine.scan_aggregate(pages, queries[k])
                assert_results_equal(got, expected[k], what="thread %d iteration %d" % (k, it))
        except Exception as e:  # noqa: BLE001 - reported by the main thread
            errors.append("thread %d: %r" % (k, e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    dev.close()
    host.close()


@pytest.mark.gpu
def test_two_gpu_scan_parity_on_hardware():
    """N = 2 on real devices (skipped on a single-GPU box): bench.py's multi-rank parity check - a sample with series
    from every shard goes through scan + NCCL exchange + merge on all ranks and is compared with the oracle."""
    import json
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "3",
           "--series", "60000"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 2 and str(line["parity_sample"]).startswith("ok"), line["parity_sample"]
