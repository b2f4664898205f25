"""Oracle scan/aggregate semantics: checked against the reference's SQL goldens (func_tb2) and against an
independent numpy brute force. No GPU needed."""
import numpy as np
import pytest

from cnosdb_b200 import cabi, datagen
from cnosdb_b200.engine import PushedAggregate, QueryOption
from oracle import pyoracle as orc
from tests.helpers import ALL_AGGS, bucket_spec, make_query, random_arena


def func_tb2(golden):
    t = golden["sql_goldens"]["func_tb2"]
    b = datagen.ArenaBuilder()
    b.add_column_group(1, np.array(t["time"], dtype=np.int64), [
        (1, cabi.TSKV_PT_U64, np.array(t["f0_u64"], dtype=np.uint64), None),
        (2, cabi.TSKV_PT_F64, np.array(t["f1_f64"], dtype=np.float64), None),
        (5, cabi.TSKV_PT_I64, np.array(t["f4_i64"], dtype=np.int64), None)])
    return b.finish()


def test_sql_goldens_func_tb2(golden):
    """sqllogicaltests/cases/function/common/{sum,avg,min,max,count,first,last}.slt over setup.slt:46-56."""
    arena, descs = func_tb2(golden)
    q = make_query([(1, cabi.TSKV_PT_U64), (2, cabi.TSKV_PT_F64), (5, cabi.TSKV_PT_I64)])
    res = orc.scan_aggregate(arena, descs, q)
    exp = golden["sql_goldens"]["expected"]
    name = {1: "f0", 2: "f1", 5: "f4"}
    sqlname = {"mean": "avg"}
    checked = 0
    for col in (1, 2, 5):
        for agg in ALL_AGGS:
            key = "%s(%s)" % (sqlname.get(agg, agg), name[col])
            if key not in exp:
                continue
            v, valid = res.column(col, agg)
            assert valid.all()
            assert float(v[0, 0]) == float(exp[key]), key
            checked += 1
    assert checked >= 19


def test_time_window_sql_golden():
    """function/time_window.slt:101-110: rows at 1970-01-01 and 1980-01-01, tumbling window of 10 ms:
    each row is alone in its window; sum(f0)=111, count(f1)=1 per window."""
    b = datagen.ArenaBuilder()
    t1980 = 315_532_800_000_000_000
    b.add_column_group(7, np.array([0, t1980], dtype=np.int64),
                       [(1, cabi.TSKV_PT_I64, np.array([111, 111]), None), (2, cabi.TSKV_PT_I64, np.array([5, 6]), None)])
    arena, descs = b.finish()
    w = 10_000_000
    for t in (0, t1980):
        start, end = orc.sliding_window(t, w, w, 0)
        assert start == t and end == t + w
        q = QueryOption([PushedAggregate(1, cabi.TSKV_PT_I64, ["sum"]), PushedAggregate(2, cabi.TSKV_PT_I64, ["count"])],
                        time_ranges=[(t, t + w - 1)], width=w, first_bucket_start=start, n_buckets=1)
        r = orc.scan_aggregate(arena, descs, q)
        assert r.column(1, "sum")[0][0, 0] == 111 and r.column(2, "count")[0][0, 0] == 1


def brute_force(truth, query, slots):
    """Independent numpy/python re-computation of the dense result: dict (col, agg) -> (values, valid)."""
    nb = query.n_buckets
    n_groups = len(slots) if query.group_by_series else 1
    out = {}
    for c in query.columns:
        acc = [[dict(count=0, vals=[], first=None, last=None) for _ in range(nb)] for _ in range(n_groups)]
        for slot, sid in enumerate(slots):
            g = slot if query.group_by_series else 0
            for ts, cols in truth.get(sid, []):
                if c.column_id not in cols:
                    continue
                vals, valid = cols[c.column_id]
                keep = np.zeros(len(ts), dtype=bool) if query.time_ranges else np.ones(len(ts), dtype=bool)
                for lo, hi in query.time_ranges:
                    keep |= (ts >= lo) & (ts <= hi)
                if query.width > 0:
                    o = query.origin % query.width
                    start = ts - ((ts - o + query.width) % query.width)  # floor regime only (t >= 0)
                    b = (start - query.first_bucket_start) // query.width
                else:
                    b = np.zeros(len(ts), dtype=np.int64)
                for bi in np.unique(b[keep]):
                    rows = np.nonzero(keep & (b == bi))[0]
                    cell = acc[g][int(bi)]
                    vrows = rows[valid[rows]]
                    cell["count"] += len(vrows)
                    cell["vals"].extend(vals[vrows].tolist())
                    r0, r1 = rows[np.argmin(ts[rows])], rows[np.argmax(ts[rows])]
                    if valid[r0] and (cell["first"] is None or ts[r0] < cell["first"][0]):
                        cell["first"] = (ts[r0], vals[r0])
                    if valid[r1] and (cell["last"] is None or ts[r1] > cell["last"][0]):
                        cell["last"] = (ts[r1], vals[r1])
        out[c.column_id] = acc
    return out


@pytest.mark.parametrize("group_by_series", [False, True])
@pytest.mark.parametrize("null_frac", [0.0, 0.15])
def test_oracle_matches_brute_force(group_by_series, null_frac):
    rng = np.random.default_rng(100 + int(group_by_series) + int(null_frac * 100))
    fields = ((1, cabi.TSKV_PT_I64), (2, cabi.TSKV_PT_F64), (3, cabi.TSKV_PT_U64))
    arena, descs, truth = random_arena(rng, n_series=12, n_points=120, fields=fields, null_frac=null_frac,
                                       jitter=200, multi_cg=True)
    sel = np.array(sorted(rng.choice(np.arange(14), 9, replace=False)), dtype=np.uint32)
    t_lo, t_hi = 1_000_000 + 7_000, 1_000_000 + 150_000
    width = 13_000
    fbs, nb = bucket_spec(t_lo, t_hi, width, origin=5)
    q = make_query(fields, series_ids=sel, time_ranges=[(t_lo, t_hi)], origin=5, width=width,
                   first_bucket_start=fbs, n_buckets=nb, group_by_series=group_by_series)
    res = orc.scan_aggregate(arena, descs, q)
    bf = brute_force(truth, q, sel.tolist())
    for c in q.columns:
        for agg in ALL_AGGS:
            v, valid = res.column(c.column_id, agg)
            for g in range(v.shape[0]):
                for b in range(nb):
                    cell = bf[c.column_id][g][b]
                    vals = cell["vals"]
                    if agg == "count":
                        assert v[g, b] == cell["count"] and valid[g, b]
                        continue
                    if agg in ("first", "last"):
                        assert valid[g, b] == (cell[agg] is not None)
                        if valid[g, b]:
                            assert v[g, b] == cell[agg][1]
                        continue
                    assert valid[g, b] == (len(vals) > 0)
                    if not vals:
                        continue
                    if agg == "sum":
                        if c.phys_type == cabi.TSKV_PT_F64:
                            assert abs(v[g, b] - sum(vals)) <= 1e-9 * max(1, abs(sum(vals)))
                        else:
                            assert int(v[g, b]) == sum(int(x) for x in vals) % 2**64 or int(v[g, b]) == sum(int(x) for x in vals)
                    elif agg == "min":
                        assert v[g, b] == min(vals)
                    elif agg == "max":
                        assert v[g, b] == max(vals)
                    elif agg == "mean":
                        assert abs(v[g, b] - sum(float(x) for x in vals) / len(vals)) <= 1e-9 * max(1, abs(v[g, b]))
    # multi-threaded driver (series chunks like iterator.rs:232-235) agrees
    res_mt = orc.scan_aggregate(arena, descs, q, n_threads=4)
    from tests.helpers import assert_results_equal
    assert_results_equal(res_mt, res, what="mt")


def test_first_skips_null_valued_head_row():
    """first.rs:139-148 + :91-94: the min-time row of a (page, bucket) group decides; when its VALUE is null
    that group contributes nothing to first() - it does not fall through to the next row."""
    b = datagen.ArenaBuilder()
    ts = np.array([10, 20, 30, 40], dtype=np.int64)
    b.add_column_group(1, ts, [(1, cabi.TSKV_PT_I64, np.array([1, 2, 3, 4]), np.array([False, True, True, True]))])
    b.add_column_group(2, ts + 1, [(1, cabi.TSKV_PT_I64, np.array([5, 6, 7, 8]), None)])
    arena, descs = b.finish()
    r = orc.scan_aggregate(arena, descs, make_query([(1, cabi.TSKV_PT_I64)], aggs=("first", "last", "count")))
    assert r.column(1, "first")[0][0, 0] == 5      # series 1's head row is null => series 2 wins although later
    assert r.column(1, "last")[0][0, 0] == 8
    assert r.column(1, "count")[0][0, 0] == 7


def test_first_last_ties_keep_earlier_series():
    """first.rs:103-107: replace only when strictly less/greater => the earlier-seen (lower slot) point wins."""
    b = datagen.ArenaBuilder()
    ts = np.array([100, 200], dtype=np.int64)
    for sid, (a, z) in enumerate([(11, 12), (21, 22), (31, 32)]):
        b.add_column_group(sid + 5, ts, [(1, cabi.TSKV_PT_I64, np.array([a, z]), None)])
    arena, descs = b.finish()
    r = orc.scan_aggregate(arena, descs, make_query([(1, cabi.TSKV_PT_I64)], aggs=("first", "last")))
    assert r.column(1, "first")[0][0, 0] == 11 and r.column(1, "last")[0][0, 0] == 12
    r = orc.scan_aggregate(arena, descs, make_query([(1, cabi.TSKV_PT_I64)], aggs=("first", "last"),
                                                    series_ids=np.array([6, 7], dtype=np.uint32)))
    assert r.column(1, "first")[0][0, 0] == 21 and r.column(1, "last")[0][0, 0] == 22


def test_negative_time_window_quirk_is_kept():
    """Rust % keeps the dividend's sign: for t < origin%w - w the window start is not a floor
    (time_window.rs:184-198). (0,5,2,0) -> (-4,1) in the reference KAT is the same arithmetic."""
    assert orc.sliding_window(-7, 5, 5, 0) == (-5, 0)      # dividend -2 -> rem -2 -> start = -7+2
    assert orc.sliding_window(-5, 5, 5, 0) == (-5, 0)      # dividend 0
    assert orc.sliding_window(-3, 5, 5, 0) == (-5, 0)      # floor regime
    assert orc.sliding_window(-12, 5, 5, 0) == (-10, -5)   # ceil-like: -12 lands in (-15,-10]


def test_bucket_range_error_and_bad_args():
    rng = np.random.default_rng(3)
    arena, descs, _ = random_arena(rng, n_series=3, n_points=50, fields=((1, cabi.TSKV_PT_I64),))
    q = make_query([(1, cabi.TSKV_PT_I64)], width=1000, first_bucket_start=1_000_000, n_buckets=3)
    with pytest.raises(orc.OracleError) as e:
        orc.scan_aggregate(arena, descs, q)
    assert e.value.status == cabi.TSKV_ERR_BUCKET_RANGE
    q = make_query([(1, cabi.TSKV_PT_F64)])
    with pytest.raises(orc.OracleError) as e:
        orc.scan_aggregate(arena, descs, q)
    assert e.value.status == cabi.TSKV_ERR_INVALID_ARG


def test_tombstone_golden_compaction_3():
    """compaction/compact/compact_test.rs:421-521 (test_compaction_3): a column tombstone (series 1, column 1) over
    [2, 6] on the files holding t = 2..6 turns those values into NULL and keeps the rows:
    expected [111, None x5, 418, 419] at t = [1, 2, 3, 4, 5, 6, 8, 9]."""
    b = datagen.ArenaBuilder()
    for ts, vals in (([1], [111]), ([2, 3, 4], [212, 213, 214]), ([4, 5, 6], [314, 315, 316]), ([8, 9], [418, 419])):
        b.add_column_group(1, np.array(ts, dtype=np.int64), [(1, cabi.TSKV_PT_I64, np.array(vals, dtype=np.int64), None)])
    arena, descs = b.finish()
    q = make_query([(1, cabi.TSKV_PT_I64)])
    res = orc.scan_aggregate(arena, descs, q, tombstones=cabi.tombstones([(1, 1, 2, 6)]))
    assert int(res.column(1, "count")[0][0, 0]) == 3
    assert int(res.column(1, "sum")[0][0, 0].view(np.int64)) == 111 + 418 + 419
    assert int(res.column(1, "min")[0][0, 0].view(np.int64)) == 111
    assert int(res.column(1, "max")[0][0, 0].view(np.int64)) == 419
    assert int(res.column(1, "first")[0][0, 0].view(np.int64)) == 111
    assert int(res.column(1, "last")[0][0, 0].view(np.int64)) == 419
    # per-bucket: the rows survive as NULLs, so buckets [2,3], [4,5] have no values at all
    q = make_query([(1, cabi.TSKV_PT_I64)], aggs=("count", "sum"), width=2, first_bucket_start=0, n_buckets=5)
    res = orc.scan_aggregate(arena, descs, q, tombstones=cabi.tombstones([(1, 1, 2, 6)]))
    assert res.column(1, "count")[0][0].tolist() == [1, 0, 0, 0, 2]


def test_tombstone_semantics_restated_from_decode_pages():
    """tsm/reader.rs:507-551: all-fields ranges drop ROWS (filter_record_batch), column ranges null the VALUES
    (updated_nullbuffer); first()/last() then see a NULL at the run's end row and skip the run (first.rs:91-94)."""
    b = datagen.ArenaBuilder()
    ts = np.arange(10, dtype=np.int64) * 10
    b.add_column_group(7, ts, [(1, cabi.TSKV_PT_I64, np.arange(1, 11, dtype=np.int64), None),
                               (2, cabi.TSKV_PT_F64, np.arange(1, 11, dtype=np.float64) / 2, None)])
    b.add_column_group(8, ts, [(1, cabi.TSKV_PT_I64, np.arange(101, 111, dtype=np.int64), None)])
    arena, descs = b.finish()
    q = make_query([(1, cabi.TSKV_PT_I64), (2, cabi.TSKV_PT_F64)], series_ids=np.array([7], dtype=np.uint32))

    def scan(tombs, query=q):
        return orc.scan_aggregate(arena, descs, query, tombstones=cabi.tombstones(tombs))

    r = scan([(7, 1, 20, 40)])  # values at t = 20, 30, 40 of column 1 are NULL; column 2 untouched
    assert int(r.column(1, "count")[0][0, 0]) == 7 and int(r.column(1, "sum")[0][0, 0].view(np.int64)) == 55 - 12
    assert int(r.column(2, "count")[0][0, 0]) == 10
    r = scan([(7, 1, 20, 40), (None, None, 65, 75)])  # + the row t = 70 is dropped from every column
    assert int(r.column(1, "count")[0][0, 0]) == 6 and int(r.column(1, "sum")[0][0, 0].view(np.int64)) == 43 - 8
    assert int(r.column(2, "count")[0][0, 0]) == 9
    r = scan([(7, None, 0, 5)])  # series-scoped row drop: first() moves to t = 10
    assert int(r.column(1, "first")[0][0, 0].view(np.int64)) == 2
    r = scan([(7, 1, 0, 0)])  # NULL at the min-time row: first() has no value for this run, last() unaffected
    assert not r.column(1, "first")[1][0, 0]
    assert r.column(1, "last")[1][0, 0] and int(r.column(1, "last")[0][0, 0].view(np.int64)) == 10
    r = scan([(8, 1, 0, 1000), (9, None, 0, 1000)])  # other series' tombstones do not apply
    assert int(r.column(1, "count")[0][0, 0]) == 10
    both = make_query([(1, cabi.TSKV_PT_I64)], series_ids=np.array([7, 8], dtype=np.uint32), group_by_series=True)
    r = scan([(None, None, 0, 49)], both)
    assert r.column(1, "count")[0][:, 0].tolist() == [5, 5]


def test_reference_reader_tables():
    """The tables of the reference's own reader tests, through the writer and the oracle.
    reader/column_group/mod.rs:266-368: time [1,3,5,7] with u64 / f64 fields written to a TSM file and read back;
    reader/filter.rs:160-253: `time > 2` (normalised to the closed range [3, i64::MAX], predicate/domain.rs:41-57)
    over the rows time = [-1, 2, 4, 18, 8] keeps 4, 18 and 8 (the field comparisons of that test are outside the path)."""
    b = datagen.ArenaBuilder()
    b.add_column_group(1, np.array([1, 3, 5, 7], dtype=np.int64), [
        (1, cabi.TSKV_PT_U64, np.array([1, 3, 5, 7], dtype=np.uint64), None),
        (2, cabi.TSKV_PT_F64, np.array([1.0, 3.0, 5.0, 7.0]), None)])
    arena, descs = b.finish()
    (t, tv), (c1, v1), (c2, v2) = orc.decode_pages(arena, descs)
    assert tv.all() and v1.all() and v2.all()
    assert t.view(np.int64).tolist() == [1, 3, 5, 7] and c1.tolist() == [1, 3, 5, 7]
    assert c2.view(np.float64).tolist() == [1.0, 3.0, 5.0, 7.0]
    r = orc.scan_aggregate(arena, descs, make_query([(1, cabi.TSKV_PT_U64), (2, cabi.TSKV_PT_F64)]))
    assert int(r.column(1, "sum")[0][0, 0]) == 16 and float(r.column(2, "mean")[0][0, 0]) == 4.0
    assert int(r.column(1, "first")[0][0, 0]) == 1 and float(r.column(2, "last")[0][0, 0]) == 7.0

    b = datagen.ArenaBuilder()
    for i, (ts, c1v, c2v) in enumerate(zip([-1, 2, 4, 18, 8], [1, 2, 4, 18, 8], [1.0, 2.0, 4.0, 18.0, 8.0])):  # unsorted rows: one page each
        b.add_column_group(1, np.array([ts], dtype=np.int64), [(1, cabi.TSKV_PT_U64, np.array([c1v], dtype=np.uint64), None),
                                                                  (2, cabi.TSKV_PT_F64, np.array([c2v]), None)])
    arena, descs = b.finish()
    q = make_query([(1, cabi.TSKV_PT_U64), (2, cabi.TSKV_PT_F64)], aggs=("count", "sum", "min", "max"),
                   time_ranges=[(3, np.iinfo(np.int64).max)])
    r = orc.scan_aggregate(arena, descs, q)
    assert int(r.column(1, "count")[0][0, 0]) == 3 and int(r.column(1, "sum")[0][0, 0]) == 30
    assert float(r.column(2, "min")[0][0, 0]) == 4.0 and float(r.column(2, "max")[0][0, 0]) == 18.0


def test_value_statistics_pruning_never_changes_the_result():
    """filter_column_groups with the pages' min / max (reader/chunk.rs:12-50): skipping the column groups a predicate
    rules out gives what evaluating the predicate on every row gives, and decodes fewer points."""
    rng = np.random.default_rng(17)
    b = datagen.ArenaBuilder()
    for sid in range(60):
        t = 1_000_000
        for _ in range(int(rng.integers(1, 4))):
            n = int(rng.integers(1, 300))
            ts = t + np.arange(n, dtype=np.int64) * 1000
            t = int(ts[-1]) + 1000
            base = int(rng.integers(-4, 5)) * 1000
            fv = (base + rng.integers(0, 900, n)).astype(np.float64) * 0.5
            if sid % 9 == 0:
                fv[rng.integers(0, n)] = np.nan
            if sid % 10 == 0:
                fv[:] = -0.0
            b.add_column_group(sid, ts, [(1, cabi.TSKV_PT_I64, base + rng.integers(0, 900, n), rng.random(n) > 0.1),
                                         (2, cabi.TSKV_PT_F64, fv, None)])
    arena, descs = b.finish()
    fbs, nb = bucket_spec(1_000_000, 1_000_000 + 1_000_000, 50_000)
    fields = [(1, cabi.TSKV_PT_I64), (2, cabi.TSKV_PT_F64)]
    try:
        for preds in ([(1, cabi.TSKV_PT_I64, ">", 1500)], [(1, cabi.TSKV_PT_I64, "<=", -2000), (2, cabi.TSKV_PT_F64, "<", 0.0)],
                      [(2, cabi.TSKV_PT_F64, ">=", 0.0)], [(2, cabi.TSKV_PT_F64, "!=", -0.0)], [(2, cabi.TSKV_PT_F64, "==", 250.5)],
                      [(1, cabi.TSKV_PT_I64, "!=", 7)], [(2, cabi.TSKV_PT_F64, "<", float("nan"))], [(1, cabi.TSKV_PT_I64, ">=", 4899)]):
            q = make_query(fields, width=50_000, first_bucket_start=fbs, n_buckets=nb, group_by_series=True, predicates=preds)
            orc.set_value_stats_pruning(True)
            a, pa = orc.scan_aggregate(arena, descs, q, return_points=True)
            orc.set_value_stats_pruning(False)
            b2, pb = orc.scan_aggregate(arena, descs, q, return_points=True)
            assert a.names == b2.names and pa <= pb
            for j in range(len(a.names)):
                assert (a.validity[j] == b2.validity[j]).all() and (a.values[j] == b2.values[j]).all(), (preds, a.names[j])
    finally:
        orc.set_value_stats_pruning(True)


def test_pruning_against_the_reference_chunk_tests():
    """The column-group statistics table of tskv/src/reader/chunk.rs:69-176 (time [1,3] [4,5] [7,8] [9,14]; field1 [0,5]
    [4,6] ...): `time > 5` keeps groups 3 and 4 (test_filter_time_column_groups_indices, :200-217), `field1 == 10` is
    ruled out by the statistics [0,5] and [4,6] (the field branch of test_filter_multi_column_groups_indices, :240-276).
    Pruned groups are not read: their values are not counted as decoded points."""
    b = datagen.ArenaBuilder()
    groups = [([1, 3], [0, 5]), ([4, 5], [4, 6]), ([7, 8], [2, 4]), ([9, 14], [3, 30])]
    for ts, vals in groups:
        b.add_column_group(1, np.array(ts, dtype=np.int64), [(2, cabi.TSKV_PT_I64, np.array(vals, dtype=np.int64), None)])
    arena, descs = b.finish()
    col = [PushedAggregate(2, cabi.TSKV_PT_I64, ["count", "sum"])]
    # time > 5  ==  the closed range [6, +inf): groups 1 and 2 are never read
    r, pts = orc.scan_aggregate(arena, descs, QueryOption(col, time_ranges=[(6, 2**62)]), return_points=True)
    assert int(r.column(2, "count")[0][0, 0]) == 4 and pts == 4
    # field1 == 10: [0,5], [4,6] and [2,4] are ruled out, [3,30] is not (it holds no 10: read, no row passes)
    r, pts = orc.scan_aggregate(arena, descs, QueryOption(col, predicates=[(2, cabi.TSKV_PT_I64, "==", 10)]), return_points=True)
    assert int(r.column(2, "count")[0][0, 0]) == 0 and pts == 2
    # field1 >= 5: the groups whose max is below 5 go
    r, pts = orc.scan_aggregate(arena, descs, QueryOption(col, predicates=[(2, cabi.TSKV_PT_I64, ">=", 5)]), return_points=True)
    assert int(r.column(2, "count")[0][0, 0]) == 3 and int(r.column(2, "sum")[0][0, 0].view(np.int64)) == 5 + 6 + 30 and pts == 6
