"""TSM file -> page arena + descriptor table, over libtskv_hostgen.so
(cnosdb_b200/csrc/host/tsm_file.cc: the layout and the reference lines it follows are documented there).

  load(data)                     parse a .tsm file image (TsmReader::open order: footer, metadata, chunk groups, chunks,
                                 column groups, page specs) and repack its pages 16-byte aligned for the engine
  write(arena, descs, bounds)    the inverse, for tests: the same metadata layout (TsmVersion V1, or V2 with snappy-coded
                                 metadata)
"""
import ctypes as C

import numpy as np

from . import cabi


class _Result(C.Structure):
    _fields_ = [("arena", C.c_void_p), ("arena_len", C.c_uint64), ("descs", C.c_void_p), ("n_descs", C.c_uint64),
                ("cg_bounds", C.c_void_p), ("n_column_groups", C.c_uint64), ("value_stats", C.c_void_p), ("n_skipped_pages", C.c_uint64),
                ("min_ts", C.c_int64), ("max_ts", C.c_int64), ("version", C.c_uint32), ("reserved", C.c_uint32)]


class TsmFile:
    """arena (uint8), descs (PAGE_DESC_DTYPE, column group by column group), cg_bounds (int64 [n_cg, 2] =
    ColumnGroup::time_range()), value_stats (cabi.VALUE_STATS_DTYPE per descriptor = PageMeta.statistics), time_range (Footer),
    version (1 | 2), n_skipped_pages (tag / string / geometry pages)."""

    def __init__(self, arena, descs, cg_bounds, time_range, version, n_skipped_pages, value_stats=None):
        self.arena, self.descs, self.cg_bounds = arena, descs, cg_bounds
        self.time_range, self.version, self.n_skipped_pages = time_range, version, n_skipped_pages
        self.value_stats = value_stats


class TsmFormatError(ValueError):
    def __init__(self, status, message):
        super().__init__("%s (status %d = %s)" % (message, status, cabi.STATUS_NAMES.get(status, "?")))
        self.status = status


def _lib():
    L = cabi.load_hostgen_library()
    L.tskvtsm_load.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p, C.POINTER(_Result)]
    L.tskvtsm_load.restype = C.c_int32
    L.tskvtsm_free.argtypes = [C.POINTER(_Result)]
    L.tskvtsm_free.restype = None
    L.tskvtsm_last_error.restype = C.c_char_p
    L.tskvtsm_write.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint32,
                                C.c_void_p, C.c_uint64]
    L.tskvtsm_write.restype = C.c_uint64
    L.tskvtsm_write_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint32,
                                      C.c_void_p, C.c_void_p, C.c_uint64]
    L.tskvtsm_write_stats.restype = C.c_uint64
    return L


def load(data, table=None):
    L = _lib()
    buf = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data, dtype=np.uint8)
    r = _Result()
    st = L.tskvtsm_load(buf.ctypes.data, buf.size, (table or "").encode(), C.byref(r))
    if st != 0:
        raise TsmFormatError(st, L.tskvtsm_last_error().decode())
    try:
        arena = np.ctypeslib.as_array(C.cast(r.arena, C.POINTER(C.c_uint8)), shape=(max(int(r.arena_len), 1),))[: int(r.arena_len)].copy()
        raw = np.ctypeslib.as_array(C.cast(r.descs, C.POINTER(C.c_uint8)), shape=(max(int(r.n_descs), 1) * 24,))
        descs = raw[: int(r.n_descs) * 24].copy().view(cabi.PAGE_DESC_DTYPE)
        b = np.ctypeslib.as_array(C.cast(r.cg_bounds, C.POINTER(C.c_int64)), shape=(max(int(r.n_column_groups), 1) * 2,))
        bounds = b[: int(r.n_column_groups) * 2].copy().reshape(-1, 2)
        vs = np.ctypeslib.as_array(C.cast(r.value_stats, C.POINTER(C.c_uint8)), shape=(max(int(r.n_descs), 1) * 24,))
        stats = vs[: int(r.n_descs) * 24].copy().view(cabi.VALUE_STATS_DTYPE)
        return TsmFile(arena, descs, bounds, (int(r.min_ts), int(r.max_ts)), int(r.version), int(r.n_skipped_pages), stats)
    finally:
        L.tskvtsm_free(C.byref(r))


def write(arena, descs, cg_bounds, table="test0", meta_encoding="null", value_stats=None):
    """-> bytes of a TSM file holding the given pages. meta_encoding: "null" (TsmVersion V1) or "snappy" (V2).
    value_stats: cabi.VALUE_STATS_DTYPE per descriptor (PageMeta.statistics), or None (min / max None everywhere)."""
    L = _lib()
    arena = np.ascontiguousarray(arena, dtype=np.uint8)
    descs = np.ascontiguousarray(descs, dtype=cabi.PAGE_DESC_DTYPE)
    bounds = np.ascontiguousarray(cg_bounds, dtype=np.int64).reshape(-1, 2)
    enc = {"null": 1, "snappy": 7}[meta_encoding]
    vs = None if value_stats is None else np.ascontiguousarray(value_stats, dtype=cabi.VALUE_STATS_DTYPE)
    vp = None if vs is None else vs.ctypes.data
    n = L.tskvtsm_write_stats(arena.ctypes.data, descs.ctypes.data, len(descs), bounds.ctypes.data, len(bounds), table.encode(), enc,
                              vp, None, 0)
    if n == 0:
        raise ValueError(L.tskvtsm_last_error().decode())
    out = np.empty(n, dtype=np.uint8)
    L.tskvtsm_write_stats(arena.ctypes.data, descs.ctypes.data, len(descs), bounds.ctypes.data, len(bounds), table.encode(), enc,
                          vp, out.ctypes.data, n)
    return out.tobytes()
