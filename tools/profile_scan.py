#!/usr/bin/env python
"""Minimal driver for ncu: builds the bench workload and runs a few device-resident scan steps.
   ncu --set full --clock-control none --import-source on -k regex:k_scan_aggregate -s 2 -c 1 \
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
a = ap.parse_args()
g = bench.generate_shard(a.series, 0, 1)
eng = Engine(0)
pages = eng.upload_pages(g.arena, g.descs, verify_crc=False)
scan = eng.prepare(pages, bench.make_query(select_tag_subset(a.series, 10)))
for _ in range(a.steps):
    scan.run()
c = eng.counters()
print("scan %.3f ms fused %.3f ms, %d points, %d bytes" % (c["elapsed_scan_ms"], c["elapsed_fused_ms"],
                                                           c["points_decoded"], c["page_read_bytes"]))
