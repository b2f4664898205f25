"""TSM file loader (cnosdb_b200/csrc/host/tsm_file.cc). The reference tree holds no .tsm fixture
and cannot run here, so the FILE layout is pinned structurally: the bincode rule set against the reference's FOOTER_SIZE
constant and field order, and a writer / loader round trip over the reference's own reader-test table
(tskv/src/reader/column_group/mod.rs:266-368: time 1,3,5,7; c1 u64 1,3,5,7; c2 f64 1.0,3.0,5.0,7.0)."""
import struct

import numpy as np
import pytest

from cnosdb_b200 import cabi, datagen, tsmfile
from oracle import pyoracle as orc
from tests.helpers import assert_results_equal, make_query, random_arena

FOOTER_SIZE = 131140  # tskv/src/tsm/mod.rs:18


def reference_table():
    b = datagen.ArenaBuilder()
    ts = np.array([1, 3, 5, 7], dtype=np.int64)
    b.add_column_group(1, ts, [(1, cabi.TSKV_PT_U64, np.array([1, 3, 5, 7], dtype=np.uint64), None),
                               (2, cabi.TSKV_PT_F64, np.array([1.0, 3.0, 5.0, 7.0]), None)])
    arena, descs = b.finish()
    return arena, descs, np.array([[1, 7]], dtype=np.int64)


def group_bounds(truth):
    return np.array([[int(ts.min()), int(ts.max())] for _, cgs in truth.items() for ts, _ in cgs], dtype=np.int64)


@pytest.mark.parametrize("enc", ["null", "snappy"])
def test_reference_reader_table_round_trips_through_a_tsm_file(enc):
    arena, descs, bounds = reference_table()
    data = tsmfile.write(arena, descs, bounds, table="test0", meta_encoding=enc)
    assert data[:4] == bytes.fromhex("012cda16")                       # TSM_MAGIC, writer.rs:38
    footer = data[-FOOTER_SIZE:]
    version, tmin, tmax = struct.unpack_from("<Iqq", footer, 0)
    assert version == (0 if enc == "null" else 1) and (tmin, tmax) == (1, 7)   # TsmVersion variant index, Footer.time_range
    assert struct.unpack_from("<Q", footer, 4 + 16 + 16)[0] == 1024 * 1024 // 8  # BloomFilter.b length (BLOOM_FILTER_BITS)
    f = tsmfile.load(data)
    assert f.version == (1 if enc == "null" else 2) and f.time_range == (1, 7) and f.n_skipped_pages == 0
    assert (f.cg_bounds == bounds).all() and len(f.descs) == 3
    assert [int(d["phys_type"]) for d in f.descs] == [cabi.TSKV_PT_TIME, cabi.TSKV_PT_U64, cabi.TSKV_PT_F64]
    pages = orc.decode_pages(f.arena, f.descs)
    assert list(pages[0][0].view(np.int64)) == [1, 3, 5, 7]
    assert list(pages[1][0]) == [1, 3, 5, 7]
    assert list(pages[2][0].view(np.float64)) == [1.0, 3.0, 5.0, 7.0]


@pytest.mark.parametrize("enc", ["null", "snappy"])
def test_loaded_file_scans_like_the_original_arena(enc):
    rng = np.random.default_rng(8)
    fields = ((1, cabi.TSKV_PT_I64), (2, cabi.TSKV_PT_F64), (3, cabi.TSKV_PT_U64))
    arena, descs, truth = random_arena(rng, n_series=60, n_points=150, fields=fields, multi_cg=True, null_frac=0.1, jitter=50)
    data = tsmfile.write(arena, descs, group_bounds(truth), meta_encoding=enc)
    f = tsmfile.load(data, table="test0")
    assert len(f.descs) == len(descs) and (f.descs["offset"] % 16 == 0).all()
    for a, b in zip(descs, f.descs):   # same pages, byte for byte, in the same order
        assert (a["size"], a["num_values"], a["series_id"], a["column_id"], a["phys_type"]) == \
               (b["size"], b["num_values"], b["series_id"], b["column_id"], b["phys_type"])
        assert (arena[a["offset"]:a["offset"] + a["size"]] == f.arena[b["offset"]:b["offset"] + b["size"]]).all()
    q = make_query(fields, aggs=("count", "sum", "min", "max"), width=17_000, origin=3, first_bucket_start=1_000_000 - 17_000 * 2 + 5,
                   n_buckets=40) if False else make_query(fields, aggs=("count", "sum", "min", "max"))
    assert_results_equal(orc.scan_aggregate(f.arena, f.descs, q), orc.scan_aggregate(arena, descs, q), what="tsm round trip")
    assert tsmfile.load(data, table="other").descs.size == 0


def test_malformed_files_are_rejected():
    arena, descs, bounds = reference_table()
    data = bytearray(tsmfile.write(arena, descs, bounds))
    with pytest.raises(tsmfile.TsmFormatError):
        tsmfile.load(bytes(data[:1000]))                       # "file is too small" (reader.rs:400-404)
    bad = bytearray(data)
    bad[0] ^= 0xff
    with pytest.raises(tsmfile.TsmFormatError):
        tsmfile.load(bytes(bad))
    bad = bytearray(data)
    struct.pack_into("<Q", bad, len(bad) - FOOTER_SIZE + 4 + 16, 1 << 40)   # TableMeta.chunk_group_offset out of bounds
    with pytest.raises(tsmfile.TsmFormatError):
        tsmfile.load(bytes(bad))
    bad = bytearray(tsmfile.write(arena, descs, bounds, meta_encoding="snappy"))
    meta_at = struct.unpack_from("<Q", bad, len(bad) - 16)[0]
    bad[meta_at] = 8                                              # Encoding::Zstd: not decodable in this build
    with pytest.raises(tsmfile.TsmFormatError) as e:
        tsmfile.load(bytes(bad))
    assert e.value.status == cabi.TSKV_ERR_UNSUPPORTED


@pytest.mark.gpu
def test_tsm_file_feeds_the_engine(engine):
    rng = np.random.default_rng(9)
    fields = ((1, cabi.TSKV_PT_I64), (2, cabi.TSKV_PT_F64))
    arena, descs, truth = random_arena(rng, n_series=80, n_points=300, fields=fields, multi_cg=True, null_frac=0.05)
    f = tsmfile.load(tsmfile.write(arena, descs, group_bounds(truth), meta_encoding="snappy"))
    pages = engine.upload_pages(f.arena, f.descs)
    pages.set_time_bounds(f.cg_bounds)            # ColumnGroup::time_range() from the file's metadata
    q = make_query(fields, time_ranges=[(1_000_000 + 100_000, 1_000_000 + 160_000)], group_by_series=True)
    assert_results_equal(engine.scan_aggregate(pages, q), orc.scan_aggregate(arena, descs, q), what="scan of a loaded TSM file")
    assert engine.counters()["pruned_page_count"] > 0
    pages.close()


def page_value_stats(arena, descs):
    """min / max of every field page's non-null values (what the reference's writer keeps in PageMeta.statistics)."""
    st = np.zeros(len(descs), dtype=cabi.VALUE_STATS_DTYPE)
    pages = orc.decode_pages(arena, descs)
    for i, (d, (vals, valid)) in enumerate(zip(descs, pages)):
        pt = int(d["phys_type"])
        if pt == cabi.TSKV_PT_TIME:
            continue
        if not valid.any():
            st[i] = (1, 0, cabi.TSKV_STATS_MINMAX, 0)   # min > max: the page holds no value
            continue
        v = vals[valid]
        if pt == cabi.TSKV_PT_I64:
            lo, hi = v.view(np.int64).min(), v.view(np.int64).max()
            st[i] = (np.array(lo, dtype=np.int64).view(np.uint64), np.array(hi, dtype=np.int64).view(np.uint64), cabi.TSKV_STATS_MINMAX, 0)
        elif pt == cabi.TSKV_PT_F64:
            f = v.view(np.float64)
            f = f[~np.isnan(f)]
            if f.size:
                st[i] = (np.array(f.min()).view(np.uint64), np.array(f.max()).view(np.uint64), cabi.TSKV_STATS_MINMAX, 0)
            else:
                st[i] = (1, 0, cabi.TSKV_STATS_MINMAX, 0)
        else:  # u64, bool
            st[i] = (v.min(), v.max(), cabi.TSKV_STATS_MINMAX, 0)
    return st


@pytest.mark.parametrize("enc", ["null", "snappy"])
def test_page_statistics_and_boolean_columns_round_trip(enc):
    """PageMeta.statistics (page.rs:599-613): Some(min) / Some(max) of i64 / u64 / f64 / bool pages come back per descriptor,
    None stays unknown; ValueType::Boolean columns are loaded as TSKV_PT_BOOL pages."""
    rng = np.random.default_rng(3)
    b = datagen.ArenaBuilder()
    for sid in range(12):
        n = int(rng.integers(1, 200))
        ts = 1000 + np.arange(n, dtype=np.int64) * 10
        b.add_column_group(sid, ts, [(1, cabi.TSKV_PT_I64, rng.integers(-500, 500, n), rng.random(n) > 0.2),
                                     (2, cabi.TSKV_PT_F64, rng.normal(size=n) * 100, None),
                                     (3, cabi.TSKV_PT_U64, np.uint64(2**63) + rng.integers(0, 1000, n).astype(np.uint64), None),
                                     (4, cabi.TSKV_PT_BOOL, rng.random(n) < 0.5, rng.random(n) > 0.1)])
    arena, descs = b.finish()
    bounds = np.array([[1000, 1000 + (int(d["num_values"]) - 1) * 10] for d in descs if d["phys_type"] == cabi.TSKV_PT_TIME], dtype=np.int64)
    stats = page_value_stats(arena, descs)
    stats["flags"][5::5] = 0   # some pages without statistics (None / None)
    f = tsmfile.load(tsmfile.write(arena, descs, bounds, meta_encoding=enc, value_stats=stats))
    assert f.n_skipped_pages == 0 and [int(x) for x in f.descs["phys_type"][:5]] == [0, 1, 3, 2, 4]
    assert len(f.value_stats) == len(descs)
    for a, bst, d in zip(stats, f.value_stats, descs):
        if d["phys_type"] == cabi.TSKV_PT_TIME or not a["flags"]:
            assert bst["flags"] == 0
        else:
            assert (int(bst["min"]), int(bst["max"]), int(bst["flags"])) == (int(a["min"]), int(a["max"]), cabi.TSKV_STATS_MINMAX)
    # without statistics everything is unknown
    assert (tsmfile.load(tsmfile.write(arena, descs, bounds, meta_encoding=enc)).value_stats["flags"] == 0).all()
    got, exp = orc.decode_pages(f.arena, f.descs), orc.decode_pages(arena, descs)
    assert all((g[0] == e[0]).all() and (g[1] == e[1]).all() for g, e in zip(got, exp))
