"""ctypes front end of the CPU oracle (oracle/libtskv_oracle.so). TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's CPU arms — never by cnosdb_b200/."""
import ctypes as C
import os

import numpy as np

from cnosdb_b200 import cabi  # struct definitions only (shared header include/tskv_gpu.h)
from cnosdb_b200.engine import ScanResult

_DIR = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_DIR, "libtskv_oracle.so")
        if not os.path.exists(path):
            raise ImportError("oracle/libtskv_oracle.so missing: run `make -C oracle`")
        L = C.CDLL(path)
        vp = C.c_void_p
        for name in ("orc_simple8b_encode", "orc_simple8b_decode", "orc_ts_encode", "orc_i64_encode",
                     "orc_f64_encode", "orc_raw_encode", "orc_bool_encode", "orc_bool_raw_encode"):
            getattr(L, name).argtypes = [vp, C.c_uint64, vp, C.c_uint64]
            getattr(L, name).restype = C.c_int64
        L.orc_zigzag_encode.argtypes = [C.c_int64]
        L.orc_zigzag_encode.restype = C.c_uint64
        L.orc_zigzag_decode.argtypes = [C.c_uint64]
        L.orc_zigzag_decode.restype = C.c_int64
        L.orc_decode_column.argtypes = [C.c_uint32, vp, C.c_uint64, vp, C.c_uint64, vp, vp]
        L.orc_decode_column.restype = C.c_int32
        L.orc_crc32.argtypes = [vp, C.c_uint64]
        L.orc_crc32.restype = C.c_uint32
        L.orc_page_build.argtypes = [vp, C.c_uint32, C.c_uint64, vp, C.c_uint64, vp]
        L.orc_page_build.restype = C.c_uint64
        L.orc_page_decode.argtypes = [C.c_uint32, vp, C.c_uint64, C.c_int, vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
        L.orc_page_decode.restype = C.c_int32
        for name in ("orc_ceil_sliding_window", "orc_floor_sliding_window"):
            getattr(L, name).argtypes = [C.c_int64] * 4 + [C.POINTER(C.c_int64)] * 2
            getattr(L, name).restype = None
        L.orc_sliding_window.argtypes = [C.c_int64] * 5 + [C.POINTER(C.c_int64)] * 2
        L.orc_sliding_window.restype = None
        L.orc_decode_pages.argtypes = [vp, C.c_uint64, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, vp, vp]
        L.orc_decode_pages.restype = C.c_int32
        L.orc_query_output_layout.argtypes = [vp, C.c_uint64, C.POINTER(cabi.Query), C.POINTER(cabi.OutputLayout)]
        L.orc_query_output_layout.restype = C.c_int32
        L.orc_scan_aggregate.argtypes = [vp, C.c_uint64, vp, C.c_uint64, C.POINTER(cabi.Query), C.c_int, C.c_int,
                                         vp, vp, C.POINTER(C.c_uint64)]
        L.orc_scan_aggregate.restype = C.c_int32
        L.orc_scan_aggregate_tomb.argtypes = [vp, C.c_uint64, vp, C.c_uint64, C.POINTER(cabi.Query), vp, C.c_uint64,
                                              C.c_int, C.c_int, vp, vp, C.POINTER(C.c_uint64)]
        L.orc_scan_aggregate_tomb.restype = C.c_int32
        L.orc_open.argtypes = [vp, C.c_uint64, vp, C.c_uint64, C.c_int, C.POINTER(vp)]
        L.orc_open.restype = C.c_int32
        L.orc_scan.argtypes = [vp, C.POINTER(cabi.Query), vp, C.c_uint64, C.c_int, vp, vp, C.POINTER(C.c_uint64)]
        L.orc_scan.restype = C.c_int32
        L.orc_set_chunk_files.argtypes = [vp, vp, C.c_uint64]
        L.orc_set_chunk_files.restype = C.c_int32
        L.orc_close.argtypes = [vp]
        L.orc_close.restype = None
        L.orc_last_error.restype = C.c_char_p
        L.orc_set_value_stats_pruning.argtypes = [C.c_int]
        L.orc_set_value_stats_pruning.restype = None
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, status, msg=""):
        super().__init__("oracle status %d (%s) %s" % (status, cabi.STATUS_NAMES.get(status, "?"), msg))
        self.status = status


def _enc(name, arr, dtype):
    a = np.ascontiguousarray(arr, dtype=dtype)
    cap = 64 + 20 * max(1, a.size)
    out = np.empty(cap, dtype=np.uint8)
    n = getattr(lib(), name)(a.ctypes.data, a.size, out.ctypes.data, cap)
    if n < 0:
        raise OracleError(int(-n), name)
    return out[:n].copy()


def simple8b_encode(v):
    return _enc("orc_simple8b_encode", v, np.uint64)


def simple8b_decode(b):
    b = np.ascontiguousarray(b, dtype=np.uint8)
    out = np.empty(240 * (b.size // 8) + 1, dtype=np.uint64)
    n = lib().orc_simple8b_decode(b.ctypes.data, b.size, out.ctypes.data, out.size)
    return out[:n].copy()


def ts_encode(v):
    return _enc("orc_ts_encode", v, np.int64)


def i64_encode(v):
    return _enc("orc_i64_encode", v, np.int64)


def f64_encode(v):
    return _enc("orc_f64_encode", v, np.float64)


def raw_encode(v):
    return _enc("orc_raw_encode", v, np.uint64)


def bool_encode(v):
    return _enc("orc_bool_encode", np.asarray(v, dtype=bool).astype(np.uint8), np.uint8)


def bool_raw_encode(v):
    return _enc("orc_bool_raw_encode", np.asarray(v, dtype=bool).astype(np.uint8), np.uint8)


def decode_column(phys_type, data, n_rows, validity=None):
    """-> (u64 values, bool validity). validity: bool array (None = all valid)."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    valid = np.ones(n_rows, dtype=bool) if validity is None else np.asarray(validity, dtype=bool)
    bm = np.packbits(valid, bitorder="little") if n_rows else np.zeros(1, dtype=np.uint8)
    vals = np.zeros(max(n_rows, 1), dtype=np.uint64)
    ov = np.zeros(max(n_rows, 1), dtype=np.uint8)
    st = lib().orc_decode_column(phys_type, data.ctypes.data, data.size, bm.ctypes.data, n_rows,
                                 vals.ctypes.data, ov.ctypes.data)
    if st != 0:
        raise OracleError(st)
    return vals[:n_rows], ov[:n_rows].astype(bool)


def crc32(data):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    return int(lib().orc_crc32(data.ctypes.data, data.size))


def page_build(data, n_rows, validity=None):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    valid = np.ones(n_rows, dtype=bool) if validity is None else np.asarray(validity, dtype=bool)
    bm = np.packbits(valid, bitorder="little") if n_rows else np.zeros(0, dtype=np.uint8)
    out = np.empty(16 + bm.size + data.size, dtype=np.uint8)
    n = lib().orc_page_build(bm.ctypes.data, bm.size, n_rows, data.ctypes.data, data.size, out.ctypes.data)
    assert n == out.size
    return out


def decode_pages(arena, descs, first_page=0, n_pages=None, verify_crc=True):
    arena = np.ascontiguousarray(arena, dtype=np.uint8)
    descs = np.ascontiguousarray(descs, dtype=cabi.PAGE_DESC_DTYPE)
    n_pages = len(descs) - first_page if n_pages is None else n_pages
    rows = descs["num_values"][first_page:first_page + n_pages].astype(np.uint64)
    bm = (rows + 63) // 64 * 8
    values = np.zeros(max(int(rows.sum()), 1), dtype=np.uint64)
    bitmaps = np.zeros(max(int(bm.sum()), 1), dtype=np.uint8)
    st = lib().orc_decode_pages(arena.ctypes.data, arena.size, descs.ctypes.data, len(descs), first_page,
                                n_pages, 1 if verify_crc else 0, values.ctypes.data, bitmaps.ctypes.data)
    if st != 0:
        raise OracleError(st, lib().orc_last_error().decode())
    out, ro, bo = [], 0, 0
    for r, b in zip(rows, bm):
        r, b = int(r), int(b)
        out.append((values[ro:ro + r], np.unpackbits(bitmaps[bo:bo + b], bitorder="little")[:r].astype(bool)))
        ro += r
        bo += b
    return out


def scan_aggregate(arena, descs, query, verify_crc=True, n_threads=1, return_points=False, tombstones=None,
                   chunk_files=None):
    """Same inputs / ScanResult as cnosdb_b200.engine.Engine.scan_aggregate, computed on the CPU.
    tombstones: cabi.TOMBSTONE_DTYPE array, what PageSet.set_tombstones takes.
    chunk_files: file id of every column group (what PageSet.set_chunk_files takes): overlapping chunks are merged."""
    if chunk_files is not None:
        op = OpenPages(arena, descs, n_threads=n_threads)
        try:
            op.set_chunk_files(chunk_files)
            return op.scan(query, verify_crc=verify_crc, return_points=return_points, tombstones=tombstones)
        finally:
            op.close()
    if isinstance(arena, tuple):
        aptr, alen = arena
    else:
        arena = np.ascontiguousarray(arena, dtype=np.uint8)
        aptr, alen = arena.ctypes.data, arena.size
    descs = np.ascontiguousarray(descs, dtype=cabi.PAGE_DESC_DTYPE)
    q = query.to_c()
    L = cabi.OutputLayout()
    st = lib().orc_query_output_layout(descs.ctypes.data, len(descs), C.byref(q), C.byref(L))
    if st != 0:
        raise OracleError(st)
    values = np.zeros(max(int(L.n_out * L.n_cells), 1), dtype=np.uint64)
    bitmaps = np.zeros(max(int(L.validity_bytes), 1), dtype=np.uint8)
    pts = C.c_uint64(0)
    tombs = np.ascontiguousarray(tombstones if tombstones is not None else [], dtype=cabi.TOMBSTONE_DTYPE)
    st = lib().orc_scan_aggregate_tomb(aptr, alen, descs.ctypes.data, len(descs), C.byref(q),
                                       tombs.ctypes.data if len(tombs) else None, len(tombs),
                                       1 if verify_crc else 0, n_threads, values.ctypes.data, bitmaps.ctypes.data,
                                       C.byref(pts))
    if st != 0:
        raise OracleError(st, lib().orc_last_error().decode())
    res = ScanResult(query, L, values[: int(L.n_out * L.n_cells)], bitmaps[: int(L.validity_bytes)])
    return (res, int(pts.value)) if return_points else res


def set_value_stats_pruning(on):
    lib().orc_set_value_stats_pruning(1 if on else 0)


class OpenPages:
    """An opened page set (orc_open): series index built once, persistent worker pool — what the reference's cached
    TsmReader metadata + live runtime threads amount to. scan() = one query, same results as scan_aggregate()."""

    def __init__(self, arena, descs, n_threads=1):
        self.arena = np.ascontiguousarray(arena, dtype=np.uint8)
        self.descs = np.ascontiguousarray(descs, dtype=cabi.PAGE_DESC_DTYPE)
        self.n_threads = n_threads
        h = C.c_void_p()
        st = lib().orc_open(self.arena.ctypes.data, self.arena.size, self.descs.ctypes.data, len(self.descs),
                            n_threads, C.byref(h))
        if st != 0:
            raise OracleError(st, lib().orc_last_error().decode())
        self.h = h

    def set_chunk_files(self, cg_file_ids):
        ids = np.ascontiguousarray(cg_file_ids, dtype=np.uint64)
        st = lib().orc_set_chunk_files(self.h, ids.ctypes.data if len(ids) else None, len(ids))
        if st != 0:
            raise OracleError(st, lib().orc_last_error().decode())

    def scan(self, query, verify_crc=True, return_points=False, tombstones=None):
        q = query.to_c()
        L = cabi.OutputLayout()
        st = lib().orc_query_output_layout(self.descs.ctypes.data, len(self.descs), C.byref(q), C.byref(L))
        if st != 0:
            raise OracleError(st)
        values = np.zeros(max(int(L.n_out * L.n_cells), 1), dtype=np.uint64)
        bitmaps = np.zeros(max(int(L.validity_bytes), 1), dtype=np.uint8)
        pts = C.c_uint64(0)
        tombs = np.ascontiguousarray(tombstones if tombstones is not None else [], dtype=cabi.TOMBSTONE_DTYPE)
        st = lib().orc_scan(self.h, C.byref(q), tombs.ctypes.data if len(tombs) else None, len(tombs),
                            1 if verify_crc else 0, values.ctypes.data, bitmaps.ctypes.data, C.byref(pts))
        if st != 0:
            raise OracleError(st, lib().orc_last_error().decode())
        res = ScanResult(query, L, values[: int(L.n_out * L.n_cells)], bitmaps[: int(L.validity_bytes)])
        return (res, int(pts.value)) if return_points else res

    def close(self):
        if self.h and _lib is not None:
            _lib.orc_close(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass


def sliding_window(t, window, slide, start_time, i=0):
    a, b = C.c_int64(), C.c_int64()
    lib().orc_sliding_window(t, window, slide, start_time, i, C.byref(a), C.byref(b))
    return a.value, b.value


def ceil_sliding_window(t, window, slide, start_time):
    a, b = C.c_int64(), C.c_int64()
    lib().orc_ceil_sliding_window(t, window, slide, start_time, C.byref(a), C.byref(b))
    return a.value, b.value


def floor_sliding_window(t, window, slide, start_time):
    a, b = C.c_int64(), C.c_int64()
    lib().orc_floor_sliding_window(t, window, slide, start_time, C.byref(a), C.byref(b))
    return a.value, b.value
