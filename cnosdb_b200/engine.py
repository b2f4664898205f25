"""Host-side mirror of the reference's reader interface for the scan path, over the C ABI.

Reference seam (paths relative to the reference tree):
  QueryOption / time ranges / aggregates   tskv/src/reader/iterator.rs:713-741
  BatchReader::process                      tskv/src/reader/mod.rs:159-164
  TsmReader::read_adjacent_pages + CRC      tskv/src/tsm/reader.rs:236-264
  Page::to_arrow_array                      tskv/src/tsm/page.rs:96-98
This module only marshals arguments: all compute happens in libtskv_gpu.so (CUDA, sm_100a).
"""
import ctypes as C

import numpy as np

from . import cabi
from .cabi import (TSKV_AGG_COUNT, TSKV_AGG_FIRST, TSKV_AGG_LAST, TSKV_AGG_MAX, TSKV_AGG_MEAN,
                   TSKV_AGG_MIN, TSKV_AGG_SUM, TSKV_PT_F64, TSKV_PT_I64, TSKV_PT_TIME, TSKV_PT_U64)

AGG_BITS = {"count": TSKV_AGG_COUNT, "sum": TSKV_AGG_SUM, "min": TSKV_AGG_MIN, "max": TSKV_AGG_MAX,
            "mean": TSKV_AGG_MEAN, "avg": TSKV_AGG_MEAN, "first": TSKV_AGG_FIRST, "last": TSKV_AGG_LAST}


class TskvError(RuntimeError):
    """Mirrors TskvError::Decode / TsmPageFileHashCheckFailed: carries the status code."""

    def __init__(self, status, message, page=-1):
        super().__init__("%s (status %d = %s, page %d)" % (
            message, status, cabi.STATUS_NAMES.get(status, "?"), page))
        self.status = status
        self.page = page


class PushedAggregate:
    """One projected value column with its aggregate set (extends PushedAggregateFunction::Count)."""

    def __init__(self, column_id, phys_type, aggs):
        self.column_id = int(column_id)
        self.phys_type = int(phys_type)
        if isinstance(aggs, int):
            self.agg_mask = aggs
        else:
            self.agg_mask = 0
            for a in aggs:
                self.agg_mask |= AGG_BITS[a]

    def agg_list(self):
        return [b for b in (1, 2, 4, 8, 16, 32, 64) if self.agg_mask & b]


class QueryOption:
    """The pushed-down scan: series selection, closed time ranges, bucket expression, aggregates."""

    def __init__(self, columns, series_ids=None, time_ranges=(), origin=0, width=0,
                 first_bucket_start=0, n_buckets=1, group_by_series=False, multi_rank=False, predicates=()):
        self.columns = list(columns)
        self.series_ids = None if series_ids is None else np.ascontiguousarray(series_ids, dtype=np.uint32)
        self.time_ranges = [(int(a), int(b)) for a, b in time_ranges]
        self.origin = int(origin)
        self.width = int(width)
        self.first_bucket_start = int(first_bucket_start)
        self.n_buckets = int(n_buckets)
        self.group_by_series = bool(group_by_series)
        # field comparisons AND-ed into the row filter: (column_id, phys_type, op, constant), op in == != < <= > >=
        self.predicates = [(int(c), int(pt), op if isinstance(op, int) else cabi.CMP_OPS[op], v) for c, pt, op, v in predicates]
        self.multi_rank = bool(multi_rank)  # TSKV_QUERY_MULTI_RANK: partials get merged with other ranks'
        self._keep = None

    def to_c(self):
        q = cabi.Query()
        if self.series_ids is not None:
            q.series_ids = self.series_ids.ctypes.data_as(C.POINTER(C.c_uint32))
            q.n_series = len(self.series_ids)
        tr = (cabi.TimeRange * max(1, len(self.time_ranges)))()
        for i, (a, b) in enumerate(self.time_ranges):
            tr[i].min_ts, tr[i].max_ts = a, b
        q.time_ranges = tr
        q.n_time_ranges = len(self.time_ranges)
        q.origin, q.width = self.origin, self.width
        q.first_bucket_start, q.n_buckets = self.first_bucket_start, self.n_buckets
        q.group_by_series = 1 if self.group_by_series else 0
        q.reserved = cabi.TSKV_QUERY_MULTI_RANK if self.multi_rank else 0
        cols = (cabi.AggColumn * len(self.columns))()
        for i, c in enumerate(self.columns):
            cols[i].column_id, cols[i].phys_type, cols[i].agg_mask = c.column_id, c.phys_type, c.agg_mask
        q.columns = cols
        q.n_columns = len(self.columns)
        preds = (cabi.FieldPredicate * max(1, len(self.predicates)))()
        for i, (c, pt, op, v) in enumerate(self.predicates):
            preds[i].column_id, preds[i].phys_type, preds[i].op = c, pt, op
            if pt == TSKV_PT_F64:
                preds[i].value = int(np.float64(v).view(np.uint64))
            else:
                preds[i].value = int(v) & 0xFFFFFFFFFFFFFFFF
        if self.predicates:
            q.predicates = preds
            q.n_predicates = len(self.predicates)
        self._keep = (tr, cols, preds)  # keep the ctypes arrays alive as long as the query
        return q

    def output_names(self):
        return [(c.column_id, cabi.AGG_NAMES[a]) for c in self.columns for a in c.agg_list()]


class ScanResult:
    """Dense result: values[j, cell] (u64 bit patterns) + validity[j, cell] (bool)."""

    def __init__(self, query, layout, values, bitmaps):
        self.names = query.output_names()
        self.n_groups = int(layout.n_groups)
        self.n_buckets = query.n_buckets
        self.values = values.reshape(int(layout.n_out), int(layout.n_cells))
        bits = np.unpackbits(bitmaps.reshape(int(layout.n_out), int(layout.bitmap_stride)), axis=1,
                             bitorder="little")
        self.validity = bits[:, : int(layout.n_cells)].astype(bool)
        self.phys = {(c.column_id): c.phys_type for c in query.columns}

    def column(self, column_id, agg):
        """(typed values, validity) of one output column, shaped [n_groups, n_buckets]."""
        j = self.names.index((column_id, agg))
        raw = self.values[j]
        pt = self.phys[column_id]
        if agg == "count":
            v = raw.view(np.uint64)
        elif agg == "mean" or pt == TSKV_PT_F64:
            v = raw.view(np.float64)
        elif pt == TSKV_PT_I64:
            v = raw.view(np.int64)
        else:
            v = raw.view(np.uint64)
        return (v.reshape(self.n_groups, self.n_buckets),
                self.validity[j].reshape(self.n_groups, self.n_buckets))


class PageSet:
    """Device-resident page arena + descriptor tables (the engine's view of cached TsmReaders)."""

    def __init__(self, engine, handle, n_pages):
        self.engine = engine
        self.handle = handle
        self.n_pages = n_pages

    def set_tombstones(self, tombs):
        """Attach the TsmTombstone ranges (cabi.TOMBSTONE_DTYPE array or cabi.tombstones([...]) input);
        replaces the previous set, an empty one clears it. Applies to every later scan of this page set."""
        if not isinstance(tombs, np.ndarray):
            tombs = cabi.tombstones(list(tombs))
        tombs = np.ascontiguousarray(tombs, dtype=cabi.TOMBSTONE_DTYPE)
        self.engine._check(self.engine.lib.tskvgpu_pages_set_tombstones(
            self.engine.ctx, self.handle, tombs.ctypes.data if len(tombs) else None, len(tombs)))

    def set_chunk_files(self, cg_file_ids):
        """File id of every column group (descriptor order): later scans merge + de-duplicate the chunks of a series
        whose time ranges overlap (DataMerger semantics: rows with equal time collapse, the newest file's non-null value
        wins per column). An empty list clears."""
        ids = np.ascontiguousarray(cg_file_ids, dtype=np.uint64)
        self.engine._check(self.engine.lib.tskvgpu_pages_set_chunk_files(
            self.engine.ctx, self.handle, ids.ctypes.data if len(ids) else None, len(ids)))

    def set_value_stats(self, stats):
        """PageMeta.statistics per descriptor (cabi.VALUE_STATS_DTYPE: min / max bit patterns + TSKV_STATS_MINMAX): scans
        with field predicates skip the column groups the bounds rule out, also on host-resident page sets."""
        st = np.ascontiguousarray(stats, dtype=cabi.VALUE_STATS_DTYPE)
        self.engine._check(self.engine.lib.tskvgpu_pages_set_value_stats(self.engine.ctx, self.handle, st.ctypes.data, len(st)))

    def set_time_bounds(self, bounds):
        """Per-column-group (min_ts, max_ts), ColumnGroup::time_range() order = descriptor order: lets scans with time
        ranges skip whole column groups (statistics pruning). bounds: [(lo, hi), ...] or an int64 array of shape [n, 2]."""
        b = np.ascontiguousarray(bounds, dtype=np.int64).reshape(-1, 2)
        self.engine._check(self.engine.lib.tskvgpu_pages_set_time_bounds(self.engine.ctx, self.handle, b.ctypes.data, len(b)))

    def close(self):
        if self.handle and self.engine.ctx:
            self.engine.lib.tskvgpu_pages_destroy(self.engine.ctx, self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PreparedScan:
    def __init__(self, engine, pages, query, handle, layout):
        self.engine, self.pages, self.query, self.handle, self.layout = engine, pages, query, handle, layout

    def run(self):
        self.engine._check(self.engine.lib.tskvgpu_scan_run(self.engine.ctx, self.handle))

    def enqueue(self):
        """Launches one full pass on the engine stream without synchronising the host."""
        self.engine._check(self.engine.lib.tskvgpu_scan_enqueue(self.engine.ctx, self.handle))

    def sync(self):
        self.engine._check(self.engine.lib.tskvgpu_scan_sync(self.engine.ctx, self.handle))

    def partials(self):
        v = cabi.PartialsView()
        self.engine._check(self.engine.lib.tskvgpu_scan_partials(self.engine.ctx, self.handle, C.byref(v)))
        return v

    def exchange_view(self):
        """(device pointer, length in 8-byte words) of the region a multi-GPU run all-gathers."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        self.engine._check(self.engine.lib.tskvgpu_scan_exchange_view(self.engine.ctx, self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    def exchange(self):
        """The multi-GPU exchange inside the library: one ncclAllGather of the exchange region + the local merge
        (Engine.comm_init first)."""
        self.engine._check(self.engine.lib.tskvgpu_scan_exchange(self.engine.ctx, self.handle))

    def merge_gathered(self, gathered_ptr, n_ranks):
        self.engine._check(self.engine.lib.tskvgpu_scan_merge_gathered(self.engine.ctx, self.handle, gathered_ptr, n_ranks))

    def snapshot_keys(self):
        self.engine._check(self.engine.lib.tskvgpu_scan_snapshot_keys(self.engine.ctx, self.handle))

    def mask_values(self):
        self.engine._check(self.engine.lib.tskvgpu_scan_mask_values(self.engine.ctx, self.handle))

    def finalize(self):
        L = self.layout
        values = np.empty(int(L.n_out * L.n_cells), dtype=np.uint64)
        bitmaps = np.empty(int(L.validity_bytes), dtype=np.uint8)
        self.engine._check(self.engine.lib.tskvgpu_scan_finalize(
            self.engine.ctx, self.handle, values.ctypes.data, bitmaps.ctypes.data))
        return ScanResult(self.query, L, values, bitmaps)

    def finalize_device(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self.engine._check(self.engine.lib.tskvgpu_scan_finalize_device(
            self.engine.ctx, self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    def close(self):
        if self.handle and self.engine.ctx:  # (an engine closed first has already released the device)
            self.engine.lib.tskvgpu_scan_destroy(self.engine.ctx, self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One CUDA device + stream (tskv_ctx)."""

    def __init__(self, device=0):
        self.lib = cabi.load_gpu_library()
        ctx = C.c_void_p()
        st = self.lib.tskvgpu_ctx_create(int(device), C.byref(ctx))
        if st != cabi.TSKV_OK:
            raise TskvError(st, "tskvgpu_ctx_create(device=%d) failed: no usable CUDA device" % device)
        self.ctx = ctx
        self.device = int(device)

    def close(self):
        if self.ctx:
            self.lib.tskvgpu_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != cabi.TSKV_OK:
            msg = self.lib.tskvgpu_last_error(self.ctx).decode()
            raise TskvError(st, msg, self.lib.tskvgpu_last_error_page(self.ctx))

    def version(self):
        return self.lib.tskvgpu_version().decode()

    def comm_unique_id(self):
        """rank 0: the 128-byte NCCL unique id the other ranks need for comm_init."""
        buf = (C.c_uint8 * 128)()
        st = self.lib.tskvgpu_comm_unique_id(buf)
        if st != cabi.TSKV_OK:
            raise TskvError(st, "tskvgpu_comm_unique_id: NCCL unavailable")
        return bytes(buf)

    def comm_init(self, unique_id, rank, n_ranks):
        """ncclCommInitRank on this engine's device (collective: every rank calls it)."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._check(self.lib.tskvgpu_comm_init(self.ctx, buf, int(rank), int(n_ranks)))

    def stream(self):
        return int(self.lib.tskvgpu_ctx_stream(self.ctx))

    def counters(self):
        c = cabi.Counters()
        self._check(self.lib.tskvgpu_get_counters(self.ctx, C.byref(c)))
        return {k: getattr(c, k) for k, _ in cabi.Counters._fields_ if k != "reserved"}

    def upload_pages(self, arena, descs, verify_crc=True, host_resident=False, verify_on_read=False):
        """arena: uint8 array (or (ptr, len)); descs: array of PAGE_DESC_DTYPE.
        host_resident=True keeps the page bytes in (page-locked) host memory: every scan then pulls the
        selected pages over PCIe itself; the caller must keep `arena` alive until PageSet.close().
        verify_on_read=True re-checks the CRC32 of every page a scan reads on the device, every scan."""
        if isinstance(arena, tuple):
            aptr, alen = arena
        else:
            arena = np.ascontiguousarray(arena, dtype=np.uint8)
            aptr, alen = arena.ctypes.data, arena.size
        descs = np.ascontiguousarray(descs, dtype=cabi.PAGE_DESC_DTYPE)
        h = C.c_void_p()
        flags = ((cabi.TSKV_UPLOAD_VERIFY_CRC if verify_crc else 0) | (cabi.TSKV_UPLOAD_HOST_RESIDENT if host_resident else 0) |
                 (cabi.TSKV_UPLOAD_VERIFY_ON_READ if verify_on_read else 0))
        st = self.lib.tskvgpu_upload_pages(self.ctx, aptr, alen, descs.ctypes.data, len(descs), flags, C.byref(h))
        self._check(st)
        ps = PageSet(self, h, len(descs))
        ps._keep = arena if host_resident else None
        return ps

    def decode_pages(self, pages, descs, first_page=0, n_pages=None):
        """Page::to_arrow_array for a page range: returns a list of (u64 values, bool validity)."""
        descs = np.ascontiguousarray(descs, dtype=cabi.PAGE_DESC_DTYPE)
        n_pages = len(descs) - first_page if n_pages is None else n_pages
        rows = descs["num_values"][first_page:first_page + n_pages].astype(np.uint64)
        bm = (rows + 63) // 64 * 8
        values = np.zeros(int(rows.sum()), dtype=np.uint64)
        bitmaps = np.zeros(int(bm.sum()), dtype=np.uint8)
        self._check(self.lib.tskvgpu_decode_pages(self.ctx, pages.handle, first_page, n_pages,
                                                  values.ctypes.data, bitmaps.ctypes.data))
        out, ro, bo = [], 0, 0
        for r, b in zip(rows, bm):
            r, b = int(r), int(b)
            valid = np.unpackbits(bitmaps[bo:bo + b], bitorder="little")[:r].astype(bool)
            out.append((values[ro:ro + r], valid))
            ro += r
            bo += b
        return out

    def output_layout(self, pages, query):
        L = cabi.OutputLayout()
        q = query.to_c()
        st = self.lib.tskvgpu_query_output_layout(pages.handle, C.byref(q), C.byref(L))
        if st != cabi.TSKV_OK:
            raise TskvError(st, "invalid query")
        return L

    def scan_aggregate(self, pages, query):
        """End-to-end call: query args H2D, fused scan, result D2H (BatchReader::process analogue)."""
        L = self.output_layout(pages, query)
        values = np.empty(int(L.n_out * L.n_cells), dtype=np.uint64)
        bitmaps = np.empty(int(L.validity_bytes), dtype=np.uint8)
        q = query.to_c()
        self._check(self.lib.tskvgpu_scan_aggregate(self.ctx, pages.handle, C.byref(q),
                                                    values.ctypes.data, bitmaps.ctypes.data))
        return ScanResult(query, L, values, bitmaps)

    def prepare(self, pages, query):
        L = self.output_layout(pages, query)
        q = query.to_c()
        h = C.c_void_p()
        self._check(self.lib.tskvgpu_scan_prepare(self.ctx, pages.handle, C.byref(q), C.byref(h)))
        return PreparedScan(self, pages, query, h, L)
