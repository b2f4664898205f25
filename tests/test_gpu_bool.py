"""Boolean pages (tskv/src/tsm/codec/boolean.rs) through the C ABI: decode and count / min / max / first / last scans
against the oracle, which tests/test_oracle_bool.py pins to the reference's vectors."""
import numpy as np
import pytest

from cnosdb_b200 import cabi, datagen
from cnosdb_b200.engine import PushedAggregate, QueryOption, TskvError
from oracle import pyoracle as orc
from tests.helpers import assert_results_equal, bucket_spec

pytestmark = pytest.mark.gpu
AGGS = ("count", "min", "max", "first", "last")


def bool_arena(rng, n_series=50, jitter=False):
    b = datagen.ArenaBuilder()
    for sid in range(n_series):
        n = int(rng.integers(1, 1500))
        ts = 1_000_000 + np.arange(n, dtype=np.int64) * 1000 + (rng.integers(-300, 301, n) if jitter and sid % 2 else 0)
        valid = rng.random(n) < 0.85 if sid % 3 else None
        vals = rng.random(n) < (0.1 if sid % 4 == 0 else 0.6)
        raw = sid % 5 == 0
        fields = [(1, cabi.TSKV_PT_BOOL, vals, valid, datagen.encode_bools_raw if raw else datagen.encode_bools),
                  (2, cabi.TSKV_PT_I64, np.cumsum(rng.integers(-5, 6, n)), valid)]
        if sid % 7 == 0:
            fields.append((3, cabi.TSKV_PT_BOOL, np.zeros(n, dtype=bool), np.zeros(n, dtype=bool)))  # all null: empty data buffer
        b.add_column_group(sid, ts, fields)
    return b.finish()


def test_decode_bool_pages(engine):
    rng = np.random.default_rng(1)
    arena, descs = bool_arena(rng)
    pages = engine.upload_pages(arena, descs)
    got, exp = engine.decode_pages(pages, descs), orc.decode_pages(arena, descs)
    for i, ((gv, gb), (ev, eb)) in enumerate(zip(got, exp)):
        assert (gb == eb).all() and (gv == ev).all(), "page %d" % i
    pages.close()


@pytest.mark.parametrize("jitter", [False, True])
def test_scan_bool_columns(engine, jitter):
    rng = np.random.default_rng(2 + jitter)
    arena, descs = bool_arena(rng, jitter=jitter)
    pages = engine.upload_pages(arena, descs)
    t_lo, t_hi = 1_000_000 - 500, 1_000_000 + 1_600_000
    fbs, nb = bucket_spec(t_lo, t_hi, 17_000, origin=3)
    cols = [PushedAggregate(1, cabi.TSKV_PT_BOOL, AGGS), PushedAggregate(3, cabi.TSKV_PT_BOOL, ("count", "max")),
            PushedAggregate(2, cabi.TSKV_PT_I64, ("count", "sum", "first"))]
    for gbs in (False, True):
        for ranges in ([], [(t_lo + 30_000, t_lo + 250_000), (t_lo + 700_000, t_lo + 700_900)]):
            q = QueryOption(cols, time_ranges=ranges, origin=3, width=17_000, first_bucket_start=fbs, n_buckets=nb, group_by_series=gbs)
            got = engine.scan_aggregate(pages, q)
            exp, pts = orc.scan_aggregate(arena, descs, q, return_points=True)
            assert_results_equal(got, exp, what="bool gbs=%s %s" % (gbs, ranges))
            assert engine.counters()["points_decoded"] == pts
    q = QueryOption([PushedAggregate(1, cabi.TSKV_PT_BOOL, ("count", "min", "max"))])  # unbucketed, no first / last
    assert_results_equal(engine.scan_aggregate(pages, q), orc.scan_aggregate(arena, descs, q), what="bool unbucketed")
    with pytest.raises(TskvError) as e:
        engine.scan_aggregate(pages, QueryOption([PushedAggregate(1, cabi.TSKV_PT_BOOL, ("sum",))]))
    assert e.value.status == cabi.TSKV_ERR_INVALID_ARG
    with pytest.raises(TskvError) as e:   # an i64 page under a boolean query column
        engine.scan_aggregate(pages, QueryOption([PushedAggregate(2, cabi.TSKV_PT_BOOL, ("count",))]))
    assert e.value.status == cabi.TSKV_ERR_INVALID_ARG
    pages.close()


@pytest.mark.parametrize("case,status", [("short_count", cabi.TSKV_ERR_BITSET_MISMATCH), ("bad_header", cabi.TSKV_ERR_BAD_ENCODING),
                                         ("open_varint", cabi.TSKV_ERR_SHORT_BLOCK), ("count_beyond_block", cabi.TSKV_ERR_BITSET_MISMATCH),
                                         ("raw_short", cabi.TSKV_ERR_BITSET_MISMATCH)])
def test_malformed_bool_blocks_match_the_oracle(engine, case, status):
    n = 21
    ts = 1_000_000 + np.arange(n, dtype=np.int64) * 1000
    data = {"short_count": datagen.encode_bools([True] * 20),
            "bad_header": np.array([10, 0x20, 21, 0xFF, 0xFF, 0xFF], dtype=np.uint8),
            "open_varint": np.array([10, 16, 0x80], dtype=np.uint8),
            "count_beyond_block": np.array([10, 16, 100, 0xFF], dtype=np.uint8),
            "raw_short": datagen.encode_bools_raw([True] * 20)}[case]
    b = datagen.ArenaBuilder()
    b.add_page(datagen.build_page(datagen.encode_timestamps(ts), n), 5, 0, cabi.TSKV_PT_TIME, n)
    b.add_page(datagen.build_page(data, n), 5, 1, cabi.TSKV_PT_BOOL, n)
    arena, descs = b.finish()
    q = QueryOption([PushedAggregate(1, cabi.TSKV_PT_BOOL, ("count", "max"))])
    with pytest.raises(orc.OracleError) as oe:
        orc.scan_aggregate(arena, descs, q)
    assert oe.value.status == status
    pages = engine.upload_pages(arena, descs)
    with pytest.raises(TskvError) as ge:
        engine.scan_aggregate(pages, q)
    assert ge.value.status == status and ge.value.page == 1
    with pytest.raises(TskvError) as ge:
        engine.decode_pages(pages, descs)
    assert ge.value.status == status
    pages.close()
