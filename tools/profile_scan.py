#!/usr/bin/env python
"""Minimal driver for ncu: builds the bench workload and runs a few device-resident scan steps.
   ncu --set full --clock-control none --import-source on -k regex:k_scan_ -s 2 -c 1 \
       -o gpurun_out/prof python tools/profile_scan.py --series 1000000 --steps 4"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cnosdb_b200.engine import Engine  # noqa: E402
from cnosdb_b200.parallel import select_tag_subset  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--series", type=int, default=1_000_000)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--kind", default="c4", help="c4 | c3 | i64 | f64a | f64b")
ap.add_argument("--jitter", type=int, default=-1, help="permille of series with jittered timestamps")
ap.add_argument("--aggs", default="count,sum,min,max,mean")
ap.add_argument("--host-resident", type=int, default=0, help="1: pages stay in host memory; 2: + CRC on every read")
a = ap.parse_args()
from cnosdb_b200 import cabi, datagen  # noqa: E402
from cnosdb_b200.engine import PushedAggregate, QueryOption  # noqa: E402
if a.kind == "c4":
    g = bench.generate_shard(a.series, 0, 1)
elif a.kind == "c3":  # TSBS devops cpu-only: hosts x 10 fields, 1-min mean/max over all hosts
    g = datagen.generate(a.series, n_fields=10, n_points=1000, value_kind=datagen.I64_WALK, seed=3)
else:
    kind = {"i64": datagen.I64_WALK, "f64a": datagen.F64_INT, "f64b": datagen.F64_NOISE}[a.kind]
    g = datagen.generate(a.series, n_fields=1, n_points=1000, value_kind=kind, seed=4,
                         jitter_permille=max(a.jitter, 0), jitter_max=999_999)
eng = Engine(0)
pages = eng.upload_pages(g.arena, g.descs, verify_crc=a.host_resident == 2, host_resident=a.host_resident > 0)
if a.kind == "c4":
    q = bench.make_query(select_tag_subset(a.series, 10))
elif a.kind == "c3":
    fbs, nb = bench.bucket_spec()
    q = QueryOption([PushedAggregate(c, cabi.TSKV_PT_I64, ["mean", "max"]) for c in range(1, 11)],
                    width=bench.W_NS, first_bucket_start=fbs, n_buckets=nb)
else:
    fbs, nb = bench.bucket_spec()
    pt = cabi.TSKV_PT_I64 if a.kind == "i64" else cabi.TSKV_PT_F64
    q = QueryOption([PushedAggregate(1, pt, a.aggs.split(","))], series_ids=select_tag_subset(a.series, 10),
                    width=bench.W_NS, first_bucket_start=fbs, n_buckets=nb)
scan = eng.prepare(pages, q)
for _ in range(a.steps):
    scan.run()
c = eng.counters()
print("scan %.3f ms fused %.3f ms, %d points, %d bytes" % (c["elapsed_scan_ms"], c["elapsed_fused_ms"],
                                                           c["points_decoded"], c["page_read_bytes"]))
