"""Pages cut at restart points (cnosdb_b200/csrc/skip_kernels.cuh): a scan that enters every simple8b / gorilla page at
several rows at once must give exactly what the oracle (which, like the reference, decodes every page from its first
byte) gives - for every number of parts, with nulls, jitter, time ranges, tombstones, row filters, long pages and
malformed streams."""
import numpy as np
import pytest

from cnosdb_b200 import cabi, datagen
from cnosdb_b200.engine import PushedAggregate, QueryOption, TskvError
from oracle import pyoracle as orc
from tests.helpers import assert_results_equal, bucket_spec, make_query, random_arena
from tests.test_gpu_parity import random_tombstones

pytestmark = pytest.mark.gpu

AGGS = ("count", "sum", "min", "max", "mean")  # FIRST / LAST scans never cut pages
FIELDS = ((1, cabi.TSKV_PT_I64), (2, cabi.TSKV_PT_F64), (3, cabi.TSKV_PT_U64))
PARTS = ["1", "2", "3", "8", "auto"]


def set_parts(monkeypatch, parts):
    if parts == "auto":
        monkeypatch.delenv("TSKV_PARTS", raising=False)
    else:
        monkeypatch.setenv("TSKV_PARTS", parts)


@pytest.mark.parametrize("parts", PARTS)
def test_c4_shape_cut_into_parts(engine, parts, monkeypatch):
    set_parts(monkeypatch, parts)
    g = datagen.generate(3000, n_fields=2, n_points=1000, value_kind=datagen.MIXED, seed=40, jitter_permille=300,
                         jitter_max=999_999, null_page_permille=200, null_row_permille=80)
    pages = engine.upload_pages(g.arena, g.descs)
    w = 60_000_000_000
    fbs, nb = bucket_spec(datagen.TSBS_T0 - 1_000_000, datagen.TSBS_T0 + 999 * datagen.TSBS_STEP + 1_000_000, w)
    sel = np.arange(0, 3000, 3, dtype=np.uint32)
    cols = [PushedAggregate(c, cabi.TSKV_PT_I64, AGGS) for c in (1, 2)] + [PushedAggregate(c, cabi.TSKV_PT_F64, AGGS) for c in (3, 4)]
    t0, st = datagen.TSBS_T0, datagen.TSBS_STEP
    for ranges in ([], [(t0 + 130 * st + 1, t0 + 777 * st)], [(t0 + 5 * st, t0 + 100 * st), (t0 + 250 * st, t0 + 260 * st), (t0 + 900 * st, t0 + 2000 * st)]):
        for gbs in (False, True):
            q = QueryOption(cols, series_ids=sel, time_ranges=ranges, width=w, first_bucket_start=fbs, n_buckets=nb, group_by_series=gbs)
            got = engine.scan_aggregate(pages, q)
            exp, pts = orc.scan_aggregate(g.arena, g.descs, q, n_threads=8, return_points=True)
            assert_results_equal(got, exp, what="parts=%s ranges=%s gbs=%s" % (parts, ranges, gbs))
            assert engine.counters()["points_decoded"] == pts
    q = QueryOption(cols, series_ids=sel, time_ranges=[(t0 + 300 * st, t0 + 301 * st)])  # unbucketed
    assert_results_equal(engine.scan_aggregate(pages, q), orc.scan_aggregate(g.arena, g.descs, q, n_threads=8), what="unbucketed")
    pages.close()


@pytest.mark.parametrize("parts", ["2", "8", "auto"])
@pytest.mark.parametrize("variant", ["nulls", "jitter", "multi_cg", "raw"])
def test_random_pages_tombstones_and_row_filters_cut_into_parts(engine, variant, parts, monkeypatch):
    set_parts(monkeypatch, parts)
    rng = np.random.default_rng(77 + len(variant))
    kw = dict(n_series=60, n_points=700, fields=FIELDS, null_frac=0.25)
    if variant == "jitter":
        kw.update(jitter=300, null_frac=0.02)
    if variant == "multi_cg":
        kw.update(multi_cg=True, null_frac=0.05)
    if variant == "raw":
        kw["raw_frac"] = 0.4
    arena, descs, _ = random_arena(rng, **kw)
    pages = engine.upload_pages(arena, descs)
    t_lo, t_hi = 1_000_000 - 400, 1_000_000 + 1_500_000
    fbs, nb = bucket_spec(t_lo, t_hi, 17_000, origin=3)
    sel = np.array(sorted(rng.choice(np.arange(60), 40, replace=False)), dtype=np.uint32)
    preds = [(1, cabi.TSKV_PT_I64, ">", -20), (2, cabi.TSKV_PT_F64, "<=", 9.5)]
    for gbs in (False, True):
        for ranges in ([], [(t_lo + 130_000, t_lo + 131_000), (t_lo + 300_500, t_lo + 655_000)]):
            for p in ([], preds):
                q = make_query(FIELDS, aggs=AGGS, series_ids=sel, time_ranges=ranges, origin=3, width=17_000, first_bucket_start=fbs,
                               n_buckets=nb, group_by_series=gbs, predicates=p)
                assert_results_equal(engine.scan_aggregate(pages, q), orc.scan_aggregate(arena, descs, q),
                                     what="%s parts=%s gbs=%s %s %s" % (variant, parts, gbs, ranges, p))
    tombs = random_tombstones(rng, descs, t_lo, 1_000_000 + 700_000)
    pages.set_tombstones(tombs)
    for gbs in (False, True):
        q = make_query(FIELDS, aggs=AGGS, series_ids=sel, time_ranges=[(t_lo + 30_000, t_lo + 650_000)], origin=3, width=17_000,
                       first_bucket_start=fbs, n_buckets=nb, group_by_series=gbs)
        assert_results_equal(engine.scan_aggregate(pages, q), orc.scan_aggregate(arena, descs, q, tombstones=tombs), what="tombstones")
    pages.close()


@pytest.mark.parametrize("parts", ["4", "64", "auto"])
def test_long_pages_and_mixed_lengths(engine, parts, monkeypatch):
    """Pages of 1 .. 50 000 rows in one bin: short pages have fewer parts than the bin's launch (idle lanes), the long
    ones are cut into up to 64 parts."""
    set_parts(monkeypatch, parts)
    rng = np.random.default_rng(5)
    b = datagen.ArenaBuilder()
    lens = [1, 127, 128, 129, 255, 256, 257, 1000, 4097, 50_000, 128 * 40, 128 * 40 + 1]
    for sid, n in enumerate(lens * 3):
        ts = 10_000 + np.arange(n, dtype=np.int64) * 1000 + (rng.integers(-300, 301, n) if sid % 2 else 0)
        valid = rng.random(n) > 0.1 if sid % 3 == 0 else None
        b.add_column_group(sid, ts, [(1, cabi.TSKV_PT_I64, np.cumsum(rng.integers(-9, 10, n)), valid),
                                     (2, cabi.TSKV_PT_F64, np.cumsum(rng.integers(-3, 4, n)) + rng.random(n), valid)])
    arena, descs = b.finish()
    pages = engine.upload_pages(arena, descs)
    fbs, nb = bucket_spec(0, 10_000 + 50_001 * 1000, 250_000)
    for gbs in (False, True):
        for ranges in ([], [(10_000 + 128_000, 10_000 + 3_000_500)]):
            q = make_query(FIELDS[:2], aggs=AGGS, time_ranges=ranges, width=250_000, first_bucket_start=fbs, n_buckets=nb, group_by_series=gbs)
            got = engine.scan_aggregate(pages, q)
            exp, pts = orc.scan_aggregate(arena, descs, q, return_points=True)
            assert_results_equal(got, exp, what="lengths parts=%s gbs=%s %s" % (parts, gbs, ranges))
            assert engine.counters()["points_decoded"] == pts
    pages.close()


@pytest.mark.parametrize("case,status", [("extra_values", 0), ("early_sentinel", cabi.TSKV_ERR_BITSET_MISMATCH),
                                         ("truncated_tail", cabi.TSKV_ERR_SHORT_BLOCK), ("truncated_middle", cabi.TSKV_ERR_SHORT_BLOCK),
                                         ("short_simple8b", cabi.TSKV_ERR_BITSET_MISMATCH)])
def test_malformed_streams_report_the_reference_errors_when_cut(engine, case, status, monkeypatch):
    """A stream that breaks before its last restart point gets no restart points (decoded whole, error and all); one
    that breaks after it is caught by the lane of the last part."""
    monkeypatch.setenv("TSKV_PARTS", "8")
    n = 1000
    vals = np.cumsum(np.arange(n) % 5).astype(np.float64) * 0.37 + 1.5
    ivals = np.cumsum(np.arange(n) % 7 - 3).astype(np.int64)
    ts = datagen.TSBS_T0 + np.arange(n, dtype=np.int64) * datagen.TSBS_STEP
    b = datagen.ArenaBuilder()
    b.add_page(datagen.build_page(datagen.encode_timestamps(ts), n), 5, 0, cabi.TSKV_PT_TIME, n)
    pt = cabi.TSKV_PT_F64
    if case == "extra_values":
        data = datagen.encode_floats(np.concatenate([vals, np.arange(30) * 3.25]))
    elif case == "early_sentinel":
        data = datagen.encode_floats(vals[:700])
    elif case == "truncated_tail":
        data = datagen.encode_floats(vals)[:-24]
    elif case == "truncated_middle":
        data = datagen.encode_floats(vals)[:600]
    else:
        pt = cabi.TSKV_PT_I64
        data = datagen.encode_integers(ivals[:520])
    b.add_page(datagen.build_page(data, n), 5, 1, pt, n)
    b.add_column_group(6, ts, [(1, pt, np.arange(n) * (0.5 if pt == cabi.TSKV_PT_F64 else 2), None)])
    arena, descs = b.finish()
    w = 60_000_000_000
    fbs, nb = bucket_spec(datagen.TSBS_T0, datagen.TSBS_T0 + 999 * datagen.TSBS_STEP, w)
    q = QueryOption([PushedAggregate(1, pt, AGGS)], width=w, first_bucket_start=fbs, n_buckets=nb)
    pages = engine.upload_pages(arena, descs)
    if status == 0:
        assert_results_equal(engine.scan_aggregate(pages, q), orc.scan_aggregate(arena, descs, q), what=case)
    else:
        with pytest.raises(orc.OracleError) as oe:
            orc.scan_aggregate(arena, descs, q)
        assert oe.value.status == status
        with pytest.raises(TskvError) as ge:
            engine.scan_aggregate(pages, q)
        assert ge.value.status == status and ge.value.page == 1
    pages.close()


def test_cut_and_whole_scans_agree_bit_for_bit_on_integers(engine, monkeypatch):
    """Size-independent property at a size the oracle is not run on: every number of parts gives the same integer
    aggregates and the same decoded-point count as the uncut scan."""
    g = datagen.generate(40_000, n_fields=1, n_points=1000, value_kind=datagen.MIXED, seed=9, jitter_permille=200, jitter_max=999_999,
                         null_page_permille=10, null_row_permille=50)
    pages = engine.upload_pages(g.arena, g.descs)
    w = 60_000_000_000
    fbs, nb = bucket_spec(datagen.TSBS_T0 - 1_000_000, datagen.TSBS_T0 + 999 * datagen.TSBS_STEP + 1_000_000, w)
    cols = [PushedAggregate(1, cabi.TSKV_PT_I64, AGGS), PushedAggregate(2, cabi.TSKV_PT_F64, ("count", "min", "max"))]
    q = QueryOption(cols, width=w, first_bucket_start=fbs, n_buckets=nb)
    ref, ref_pts = None, None
    for parts in ("1", "2", "4", "8"):
        monkeypatch.setenv("TSKV_PARTS", parts)
        r = engine.scan_aggregate(pages, q)
        pts = engine.counters()["points_decoded"]
        if ref is None:
            ref, ref_pts = r, pts
            continue
        assert pts == ref_pts
        for j in range(len(r.names)):
            assert (r.validity[j] == ref.validity[j]).all()
            if r.names[j][1] != "mean":
                assert (r.values[j] == ref.values[j]).all(), (parts, r.names[j])
    pages.close()


def test_item_driven_work_list_still_agrees(engine, monkeypatch):
    """TSKV_WORKLIST=items: the round-1 work list (every field page of the page set flagged, ordered compaction) must
    give what the selection-driven one gives - with a selection list, without one, with pruning and a row filter."""
    g = datagen.generate(4000, n_fields=2, n_points=700, value_kind=datagen.MIXED, seed=77, jitter_permille=300, jitter_max=999_999,
                         null_page_permille=100, null_row_permille=80)
    pages = engine.upload_pages(g.arena, g.descs)
    w = 60_000_000_000
    fbs, nb = bucket_spec(datagen.TSBS_T0 - 1_000_000, datagen.TSBS_T0 + 999 * datagen.TSBS_STEP + 1_000_000, w)
    cols = [PushedAggregate(1, cabi.TSKV_PT_I64, AGGS + ("first", "last")), PushedAggregate(3, cabi.TSKV_PT_F64, AGGS)]
    t0, st = datagen.TSBS_T0, datagen.TSBS_STEP
    for sel in (np.arange(0, 4000, 7, dtype=np.uint32), None):
        for kw in (dict(), dict(time_ranges=[(t0 + 100 * st, t0 + 300 * st)]), dict(predicates=[(1, cabi.TSKV_PT_I64, ">", 10)])):
            q = QueryOption(cols, series_ids=sel, width=w, first_bucket_start=fbs, n_buckets=nb, **kw)
            exp = orc.scan_aggregate(g.arena, g.descs, q, n_threads=8)
            for mode in ("items", "series"):
                monkeypatch.setenv("TSKV_WORKLIST", mode)
                got = engine.scan_aggregate(pages, q)
                assert_results_equal(got, exp, what="work list %s sel=%s %s" % (mode, sel is not None, kw))
                c = engine.counters()
                if mode == "items":
                    ref_counts = (c["page_read_count"], c["page_read_bytes"], c["pruned_page_count"])
                else:
                    assert (c["page_read_count"], c["page_read_bytes"], c["pruned_page_count"]) == ref_counts
    pages.close()


def test_few_series_with_many_column_groups(engine):
    """3 series x 90 column groups: the work list is built by the pass over the field pages (a thread per selected series
    would walk 90 groups serially); same results as the oracle either way."""
    rng = np.random.default_rng(8)
    b = datagen.ArenaBuilder()
    for sid in (4, 9, 11):
        t = 1_000_000
        for _ in range(90):
            n = int(rng.integers(1, 400))
            ts = t + np.arange(n, dtype=np.int64) * 1000
            t = int(ts[-1]) + 1000
            b.add_column_group(sid, ts, [(1, cabi.TSKV_PT_I64, np.cumsum(rng.integers(-9, 10, n)), rng.random(n) > 0.1),
                                         (2, cabi.TSKV_PT_F64, np.cumsum(rng.integers(-3, 4, n)) + rng.random(n), None)])
    arena, descs = b.finish()
    pages = engine.upload_pages(arena, descs)
    fbs, nb = bucket_spec(1_000_000, 1_000_000 + 90 * 400 * 1000, 500_000)
    for sel in (None, np.array([4, 11], dtype=np.uint32)):
        for ranges in ([], [(1_000_000 + 3_000_000, 1_000_000 + 9_000_000)]):
            q = make_query(FIELDS[:2], aggs=AGGS + ("first", "last"), series_ids=sel, time_ranges=ranges, width=500_000, first_bucket_start=fbs,
                           n_buckets=nb, group_by_series=True)
            got = engine.scan_aggregate(pages, q)
            exp, pts = orc.scan_aggregate(arena, descs, q, return_points=True)
            assert_results_equal(got, exp, what="many groups sel=%s %s" % (sel is not None, ranges))
            assert engine.counters()["points_decoded"] == pts
    pages.close()
