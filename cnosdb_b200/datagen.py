"""Python front end of the seeded synthetic TSM page generator (libtskv_hostgen.so) and of the
host-side page writer. Workload presets follow BASELINE.md section 3 (C1..C5)."""
import ctypes as C
import os

import numpy as np

from . import cabi

I64_WALK, F64_INT, F64_NOISE, MIXED, U64_WALK = 0, 1, 2, 3, 4
TSBS_T0 = 1_640_995_200_000_000_000  # 2022-01-01T00:00:00Z in ns
TSBS_STEP = 10_000_000_000           # 10 s


class Generated:
    """Arena + descriptors owned by the native generator (freed on close / GC)."""

    def __init__(self, lib, res):
        self._lib, self._res = lib, res
        self.arena = np.ctypeslib.as_array(C.cast(res.arena, C.POINTER(C.c_uint8)), shape=(max(int(res.arena_len), 1),))[: int(res.arena_len)]
        raw = np.ctypeslib.as_array(C.cast(res.descs, C.POINTER(C.c_uint8)), shape=(max(int(res.n_descs), 1) * 24,))
        self.descs = raw[: int(res.n_descs) * 24].view(cabi.PAGE_DESC_DTYPE)
        self.n_points = int(res.n_points)

    def close(self):
        if self._res is not None:
            self._lib.tskvgen_free(C.byref(self._res))
            self._res = None
            self.arena = self.descs = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def generate(n_series, n_fields=1, n_points=1000, value_kind=I64_WALK, seed=1, first_series_id=0,
             series_stride=1, t0=TSBS_T0, step=TSBS_STEP, jitter_permille=0, jitter_max=0,
             null_page_permille=0, null_row_permille=0, raw_encoding_permille=0, n_threads=None):
    lib = cabi.load_hostgen_library()
    spec = cabi.GenSpec(seed=seed, n_series=n_series, first_series_id=first_series_id,
                        series_stride=series_stride, n_fields=n_fields, n_points=n_points,
                        value_kind=value_kind, t0=t0, step=step, jitter_permille=jitter_permille,
                        jitter_max=jitter_max, null_page_permille=null_page_permille,
                        null_row_permille=null_row_permille, raw_encoding_permille=raw_encoding_permille)
    res = cabi.GenResult()
    nt = n_threads or os.cpu_count() or 1
    rc = lib.tskvgen_generate(C.byref(spec), nt, C.byref(res))
    if rc != 0:
        raise RuntimeError("tskvgen_generate failed (%d)" % rc)
    return Generated(lib, res)


def _encode(fn_name, arr, dtype):
    lib = cabi.load_hostgen_library()
    a = np.ascontiguousarray(arr, dtype=dtype)
    cap = 32 + 10 * max(1, a.size) + a.size * 9
    out = np.empty(cap, dtype=np.uint8)
    n = getattr(lib, fn_name)(a.ctypes.data, a.size, out.ctypes.data, cap)
    if n < 0:
        raise ValueError("%s: value not encodable" % fn_name)
    return out[:n].copy()


def encode_timestamps(v):
    return _encode("tskvw_encode_timestamps", v, np.int64)


def encode_integers(v):
    return _encode("tskvw_encode_integers", v, np.int64)


def encode_floats(v):
    return _encode("tskvw_encode_floats", v, np.float64)


def encode_raw(v):
    return _encode("tskvw_encode_raw", np.asarray(v).view(np.uint64) if np.asarray(v).dtype.itemsize == 8 else v, np.uint64)


def encode_bools(v):
    """Boolean bit-pack (tskv/src/tsm/codec/boolean.rs:24-64): BitPack id | 0x10 | LEB128 count | 1 bit per value, MSB first.
    (numpy only: independent of the CPU checker's restatement.)"""
    v = np.asarray(v, dtype=bool)
    if v.size == 0:
        return np.zeros(0, dtype=np.uint8)
    n, var = int(v.size), []
    while n >= 0x80:
        var.append((n & 0x7F) | 0x80)
        n >>= 7
    var.append(n)
    return np.concatenate([np.array([cabi.TSKV_ENC_BITPACK, 0x10] + var, dtype=np.uint8), np.packbits(v, bitorder="big")])


def encode_bools_raw(v):
    """Boolean column under Encoding::Null (boolean.rs:66-76): one byte per value."""
    return np.concatenate([np.array([cabi.TSKV_ENC_NULL], dtype=np.uint8), np.asarray(v, dtype=bool).astype(np.uint8)])


def simple8b_pack(v):
    return _encode("tskvw_simple8b_pack", v, np.uint64)


def build_page(data, rows, validity=None):
    """Frames one page. validity: bool array of `rows` (None = all valid)."""
    lib = cabi.load_hostgen_library()
    data = np.ascontiguousarray(data, dtype=np.uint8)
    bm = None
    if validity is not None:
        bm = np.packbits(np.asarray(validity, dtype=bool), bitorder="little")
    cap = 16 + (rows + 7) // 8 + data.size
    out = np.empty(cap, dtype=np.uint8)
    n = lib.tskvw_build_page(bm.ctypes.data if bm is not None else None, rows, data.ctypes.data,
                             data.size, out.ctypes.data, cap)
    assert n == cap, (n, cap)
    return out


class ArenaBuilder:
    """Assembles hand-made pages into an arena + descriptor table (tests, small tools)."""

    def __init__(self):
        self.chunks, self.descs, self.size = [], [], 0

    def add_page(self, page, series_id, column_id, phys_type, num_values):
        pad = (-self.size) % 16
        if pad:
            self.chunks.append(np.zeros(pad, dtype=np.uint8))
            self.size += pad
        self.descs.append((self.size, len(page), num_values, series_id, column_id, phys_type, 0))
        self.chunks.append(np.asarray(page, dtype=np.uint8))
        self.size += len(page)

    def add_column_group(self, series_id, timestamps, fields, time_validity=None):
        """fields: list of (column_id, phys_type, values, validity-or-None[, encoder])."""
        n = len(timestamps)
        tv = None if time_validity is None else np.asarray(time_validity, dtype=bool)
        tvals = np.asarray(timestamps, dtype=np.int64)
        self.add_page(build_page(encode_timestamps(tvals if tv is None else tvals[tv]), n, tv),
                      series_id, 0, cabi.TSKV_PT_TIME, n)
        for f in fields:
            column_id, pt, vals, valid = f[:4]
            enc = f[4] if len(f) > 4 else None
            vals = np.asarray(vals)
            vv = None if valid is None else np.asarray(valid, dtype=bool)
            kept = vals if vv is None else vals[vv]
            if enc is None:
                enc = encode_floats if pt == cabi.TSKV_PT_F64 else encode_bools if pt == cabi.TSKV_PT_BOOL else encode_integers
            if pt == cabi.TSKV_PT_U64 and enc is encode_integers:
                kept = np.asarray(kept, dtype=np.uint64).view(np.int64)
            self.add_page(build_page(enc(kept), n, vv), series_id, column_id, pt, n)

    def finish(self):
        arena = np.concatenate(self.chunks) if self.chunks else np.zeros(0, dtype=np.uint8)
        descs = np.array(self.descs, dtype=cabi.PAGE_DESC_DTYPE)
        return arena, descs
