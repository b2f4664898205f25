#!/usr/bin/env python
"""bench.py — decoded+aggregated points/s of the tskv scan hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    the reference algorithm on the host CPU cores
                                                           (oracle port: the Rust reference cannot be built here)

Workload (config.workload): BASELINE config C4 — 1 000 000 series x 1 000 points, mixed i64 (Delta/simple8b)
and f64 (Gorilla, full-mantissa) columns, 20 % of the series with jittered timestamps (simple8b time pages),
1 % of the pages with 5 % nulls, tag predicate selecting 10 % of the series, GROUP BY 1-minute bucket with
count/sum/min/max/mean. Series are sharded in contiguous id ranges over N GPUs (strong scaling: total work fixed); the
only collective is the all-reduce of the per-bucket partials.

One JSON line on stdout (rank 0). A "step" is one full pass: series selection -> work list -> fused
decode/filter/bucket-reduce kernels -> (all-reduce) -> dense result.
  value     whole-job points/s with the pages already resident in HBM (device time, CUDA events, max over ranks)
  e2e       the same through the public call with the pages in HOST memory: query args H2D, PCIe gather of
            the selected pages, scan, result D2H (wall clock around synchronised calls, max over ranks)
  roofline  dominant fused kernel: encoded bytes it reads / its CUDA-event time vs the measured HBM copy peak
  cpu_baseline  the oracle (port of the reference algorithm) on the host cores, bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from cnosdb_b200 import cabi, datagen  # noqa: E402
from cnosdb_b200.engine import PushedAggregate, QueryOption  # noqa: E402
from cnosdb_b200.parallel import select_tag_subset, shard_range  # noqa: E402

METRIC = "decoded+aggregated points/s"
W_NS = 60_000_000_000
AGGS = ["count", "sum", "min", "max", "mean"]
BIN_NAMES = {0: "ts=RLE,val=simple8b", 1: "ts=RLE,val=gorilla", 2: "ts=RLE,val=generic", 3: "ts=simple8b,val=simple8b",
             4: "ts=simple8b,val=gorilla", 5: "ts=simple8b,val=generic", 6: "ts=generic,val=simple8b",
             7: "ts=generic,val=gorilla", 8: "ts=generic,val=generic", 9: "ts=RLE,val=simple8b (<=1024 rows)",
             10: "ts=simple8b,val=simple8b (<=1024 rows)", 11: "ts=RLE,val=gorilla (<=1024 rows)",
             12: "ts=simple8b,val=gorilla (<=1024 rows)"}


def workload_name(n_series):
    return ("C4: %d series x 1000 pts, mixed i64 Delta / f64 Gorilla, 20%% jittered ts, 1%% pages with 5%% nulls, "
            "10%% tag selection, group by 1-min bucket (count,sum,min,max,mean)" % n_series)


def bucket_spec():
    lo = datagen.TSBS_T0 - 1_000_000
    hi = datagen.TSBS_T0 + 999 * datagen.TSBS_STEP + 1_000_000
    start = lo - (lo % W_NS)
    return start, int((hi - start) // W_NS + 1)


def make_query(series_ids):
    fbs, nb = bucket_spec()
    return QueryOption([PushedAggregate(1, cabi.TSKV_PT_I64, AGGS), PushedAggregate(2, cabi.TSKV_PT_F64, AGGS)],
                       series_ids=series_ids, width=W_NS, first_bucket_start=fbs, n_buckets=nb)


def generate_shard(n_total, rank, world):
    lo, hi = shard_range(n_total, rank, world)
    return datagen.generate(hi - lo, n_fields=1, n_points=1000, value_kind=datagen.MIXED, seed=4,
                            first_series_id=lo, series_stride=1, jitter_permille=200, jitter_max=999_999,
                            null_page_permille=10, null_row_permille=50)


class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe). The region is
    ~20 ms long, so NVML is polled from a thread every millisecond (what nvidia-smi reads, without its 100 ms period);
    `nvidia-smi -lms` is the fallback when pynvml is missing."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    REASON_BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                   0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.handle, self.stop_flag = index, [], None, None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml, self.handle = pynvml, handle
        except Exception:
            self.nvml = None

    def start(self):
        if self.nvml:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll(self):
        n, h = self.nvml, self.handle
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
                mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(h) / 1000.0
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((time.perf_counter(), float(sm), float(mx), pw, int(rs)))
            except Exception:
                pass
            time.sleep(0.001)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self, t_begin=None, t_end=None):
        """t_begin / t_end (time.perf_counter): host-side bounds of the timed region; NVML samples outside are dropped."""
        if self.nvml:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            rows = [r for r in self.rows if (t_begin is None or r[0] >= t_begin) and (t_end is None or r[0] <= t_end)]
            if not rows:
                rows = self.rows[-3:]
            reasons = sorted({name for r in rows for bit, name in self.REASON_BITS.items() if r[4] & bit})
            return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None,
                    "sm_max_mhz": max(r[2] for r in rows) if rows else None,
                    "power_w_max": max(r[3] for r in rows) if rows else None, "reasons": reasons, "samples": len(rows),
                    "source": "nvml, 1 ms period, inside the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm), "source": "nvidia-smi -lms 100"}


def host_threads():
    """Threads the CPU arm may use: the affinity / cgroup view, not the machine's core count."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_arm(arena, descs, sel_all, steps, warmup, max_s=150.0):
    """Times the oracle (port of the reference algorithm; CRC32 of every page verified on every read like
    Page::crc_validation) on the WHOLE workload: all selected series, every step. The page set is opened once
    (series index + persistent worker pool, like the reference's cached TsmReader metadata and live runtime threads);
    a step is one query. Steps are cut short only if the run would exceed max_s seconds (stated in `sample`).
    Returns (info, selection used, result, seconds per step)."""
    from oracle import pyoracle as orc
    cores = host_threads()
    op = orc.OpenPages(arena, descs, cores)
    q = make_query(sel_all)
    times, pts, res = [], 0, None
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        res, pts = op.scan(q, verify_crc=True, return_points=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_begin + dt > max_s and len(times) >= 2:
            break
    op.close()
    med = float(np.median(times))
    info = {"value": pts / med, "unit": "points/s", "cores": cores, "kind": "port",
            "sample": "%d of %d selected series (%d points) per step, %d of %d steps timed, %d threads (persistent pool, "
                      "series index built once), CRC32 verified per page per step" % (
                          len(sel_all), len(sel_all), pts, len(times), steps, cores),
            "ms_per_step_median": med * 1e3, "ms_per_step_min": min(times) * 1e3, "ms_per_step_max": max(times) * 1e3}
    return info, sel_all, res, med


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--series", type=int, default=1_000_000, help="total series (BASELINE C4: 1e6)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warmup = args.steps, max(args.warmup, 3 if args.impl == "ours" else 0)
    sel_all = select_tag_subset(args.series, 10)
    config = {"workload": workload_name(args.series), "series_total": args.series, "points_per_series": 1000,
              "selectivity": 0.1, "selected_series": int(len(sel_all)), "buckets": bucket_spec()[1],
              "aggregates": AGGS, "sharding": "contiguous series-id ranges, one per GPU"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        g = generate_shard(args.series, 0, 1)
        info, sample, _, step_s = cpu_arm(g.arena, g.descs, sel_all, steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": info["value"], "unit": "points/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "i64/f64",
                "data": "synthetic", "config": config, "cpu_baseline": info,
                "e2e": {"value": info["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "oracle port of the reference algorithm (the Rust reference cannot be compiled in this image)"}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from cnosdb_b200.engine import Engine
    from cnosdb_b200.parallel import GatherExchange

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    t_gen = time.perf_counter()
    g = generate_shard(args.series, rank, world)
    t_gen = time.perf_counter() - t_gen
    eng = Engine(local_rank)
    stream = torch.cuda.ExternalStream(eng.stream(), device=device)
    pages = eng.upload_pages(g.arena, g.descs, verify_crc=True)
    q = make_query(sel_all)
    scan = eng.prepare(pages, q)
    exchange = GatherExchange(scan, eng, world) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def one_step():
        scan.enqueue()
        if world > 1:
            exchange.run()
        scan.finalize_device()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- device-resident throughput (value) ------------------------------------------------------
    for _ in range(warmup):
        one_step()
    scan.sync()
    c = eng.counters()
    points_local = c["points_decoded"]
    launches_per_step = c["kernel_launches"] + 1
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    t_region = time.perf_counter()
    with torch.cuda.stream(stream):
        for a, b in ev:
            flush.zero_()          # evict the previous step's pages from L2 (outside the timed interval)
            a.record(stream)
            one_step()
            b.record(stream)
    scan.sync()
    barrier()
    clocks = sampler.stop(t_region, time.perf_counter()) if rank == 0 else None
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms, float(points_local)], dtype=torch.float64, device=device)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dev_ms, points_total = float(tmax[0]), float(t[1])
    else:
        points_total = float(points_local)
    value = points_total * steps / (dev_ms / 1e3)

    # ---- roofline of the dominant fused kernel (live CUDA events inside the library) ---------------
    dom_ms, dom_bytes, fused_ms, scan_ms, dom_bin = [], 0, [], [], 0
    for _ in range(max(3, min(steps, 10))):
        with torch.cuda.stream(stream):
            flush.zero_()
        scan.run()
        c = eng.counters()
        dom_ms.append(c["dominant_kernel_ms"])
        fused_ms.append(c["elapsed_fused_ms"])
        scan_ms.append(c["elapsed_scan_ms"])
        dom_bytes, dom_bin = c["dominant_kernel_bytes"], c["dominant_kernel_bin"]
    page_bytes = c["page_read_bytes"]
    L = scan.layout
    algo_bytes = page_bytes + 24 * c["page_read_count"] + 4 * len(sel_all) + 8 * int(L.n_out * L.n_cells)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # The dominant kernel is the fused decode/filter/reduce kernel k_scan_aggregate: one template, launched
    # once per decode-kind bin present (C4: 4 instantiations) on concurrent streams. `achieved` = the
    # algorithmic bytes of one step / the CUDA-event time from the fork to the join of those launches.
    fused_s = float(np.mean(fused_ms)) * 1e-3
    traffic = None
    try:  # DRAM bytes of the same launches from the committed `ncu --set full` capture (profiles/)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_scan_ncu_summary.json")))["step_dram_traffic_bytes"]
        traffic = int(traffic * page_bytes / 565936507) if world > 1 else traffic
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "hbm", "kernel": "k_scan_aggregate<TK,VK,SEL> (fused decode+filter+bucket-reduce; %d concurrent bin launches)" % 4,
                "achieved": algo_bytes / fused_s / 1e9, "peak": peak, "unit": "GB/s",
                "frac": algo_bytes / fused_s / 1e9 / peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "traffic": traffic, "bytes_per_launch": int(algo_bytes), "ms_per_launch": fused_s * 1e3,
                "decoded_equivalent_frac": 16 * points_local / fused_s / 1e9 / peak,
                "slowest_bin": {"kernel": "fused scan kernel of bin <%s> (k_scan_aggregate, or k_scan_coop on small scans)" % BIN_NAMES[int(dom_bin)], "ms": float(np.mean(dom_ms)),
                                "page_bytes": int(dom_bytes)},
                "step_ms": float(np.mean(scan_ms)),
                "bound_note": "issue/latency-bound lane-serial decode (ncu: issue-active 19-37%, DRAM 1.1x algorithmic bytes); see DESIGN.md section 5"}

    # ---- end to end: pages in host memory, PCIe gather inside the timed region ---------------------
    # Page CRC32s are re-checked on the device after every transfer (what the reference does on every page read, and
    # what the CPU arm does); TSKV_BENCH_E2E_CRC=0 measures the transfer + scan alone.
    e2e_crc = os.environ.get("TSKV_BENCH_E2E_CRC", "1") != "0"
    hp = eng.upload_pages(g.arena, g.descs, verify_crc=e2e_crc, host_resident=True)

    def e2e_step():
        s = eng.prepare(hp, q)
        s.enqueue()
        if world > 1:
            GatherExchange(s, eng, world).run()
        res = s.finalize()
        s.sync()
        cc = eng.counters()
        s.close()
        return res, cc

    for _ in range(max(1, warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        res, cc = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    e2e = {"value": points_total * steps / e2e_s, "unit": "points/s",
           "crc_verified": "device, every step" if e2e_crc else "off",
           "h2d_bytes_per_step": int(cc["page_read_bytes"] + cc["h2d_bytes"]),
           "d2h_bytes_per_step": int(L.values_bytes + L.validity_bytes + 12 + 13 * 8),
           "ms_per_step": e2e_s / steps * 1e3,
           "path": "prepare(H2D args) + select + PCIe gather of selected pages + fused scan + finalize(D2H)"}

    # ---- CPU baseline + parity on the sample (rank 0, N = 1) ---------------------------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, sample, cpu_res, _ = cpu_arm(g.arena, g.descs, sel_all, 2, 1)
        got = eng.scan_aggregate(pages, make_query(sample))
        ok = True
        for j, (col, agg) in enumerate(got.names):
            ok &= bool((got.validity[j] == cpu_res.validity[j]).all())
            m = cpu_res.validity[j]
            if agg == "mean" or (agg == "sum" and col == 2):
                a, b = got.values[j][m].view(np.float64), cpu_res.values[j][m].view(np.float64)
                ok &= bool((np.abs(a - b) <= 1e-6 * np.maximum(np.abs(b), 1e-300)).all())
            else:
                ok &= bool((got.values[j][m] == cpu_res.values[j][m]).all())
        parity = "ok" if ok else "MISMATCH"

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "i64/f64", "data": "synthetic",
                "config": dict(config, l2_flush_between_steps=True, pages_resident="HBM",
                               encoded_bytes_selected_per_rank=int(page_bytes), generate_s=round(t_gen, 2)),
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * steps),
                "roofline": roofline, "cpu_baseline": cpu, "parity_sample": parity,
                "points_per_step": points_total}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
