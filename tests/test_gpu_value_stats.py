"""Value-statistics pruning (filter_column_groups with the pages' min / max, tskv/src/reader/chunk.rs:12-50 +
reader/column_group/statistics.rs:11-80): a column group whose page statistics rule a pushed `column <op> constant`
out for every row is never read - same results, fewer pages, against the oracle's restatement."""
import numpy as np
import pytest

from cnosdb_b200 import cabi, datagen
from cnosdb_b200.engine import PushedAggregate, QueryOption
from oracle import pyoracle as orc
from tests.helpers import ALL_AGGS, assert_results_equal, bucket_spec, make_query

pytestmark = pytest.mark.gpu
FIELDS = ((1, cabi.TSKV_PT_I64), (2, cabi.TSKV_PT_F64), (3, cabi.TSKV_PT_U64))


def banded_arena(rng, n_series=120, multi_cg=True):
    """Every column group's values live in a band of their own, so range predicates rule whole groups out."""
    b = datagen.ArenaBuilder()
    for sid in range(n_series):
        t = 1_000_000
        for _ in range(int(rng.integers(1, 4)) if multi_cg else 1):
            n = int(rng.integers(1, 500))
            ts = t + np.arange(n, dtype=np.int64) * 1000
            t = int(ts[-1]) + 1000
            base = int(rng.integers(-5, 6)) * 1000
            iv = base + rng.integers(0, 900, n)
            fv = (base + rng.integers(0, 900, n)).astype(np.float64) * 0.5
            qv = fv.copy()                               # column 4: only ever a predicate column (NaN is never aggregated)
            if sid % 11 == 0:
                qv[rng.integers(0, n)] = np.nan          # NaN never satisfies a comparison and is not part of min / max
            if sid % 13 == 0:
                qv[:] = -0.0                             # -0.0 == +0.0 for `>= 0.0`
            uv = (np.uint64(2**63) + np.uint64(base + 6000) + rng.integers(0, 900, n).astype(np.uint64))
            valid = rng.random(n) > 0.1 if sid % 5 == 0 else None
            fields = [(1, cabi.TSKV_PT_I64, iv, valid), (2, cabi.TSKV_PT_F64, fv, valid)]
            if sid % 7:
                fields.append((3, cabi.TSKV_PT_U64, uv, None))   # some groups do not hold column 3
            fields.append((4, cabi.TSKV_PT_F64, qv, None))
            if sid % 17 == 0:
                fields[0] = (1, cabi.TSKV_PT_I64, iv, np.zeros(n, dtype=bool))   # an all-null predicate column
            b.add_column_group(sid, ts, fields)
    return b.finish()


def test_value_statistics_prune_column_groups(engine):
    rng = np.random.default_rng(31)
    arena, descs = banded_arena(rng)
    pages = engine.upload_pages(arena, descs)
    fbs, nb = bucket_spec(1_000_000, 1_000_000 + 1_600_000, 50_000)
    cases = [
        [(1, cabi.TSKV_PT_I64, ">", 2500)],
        [(1, cabi.TSKV_PT_I64, "<=", -3000), (2, cabi.TSKV_PT_F64, "<", 0.0)],
        [(4, cabi.TSKV_PT_F64, ">=", 0.0)],                      # keeps the -0.0 groups
        [(4, cabi.TSKV_PT_F64, "==", 1250.5)],
        [(4, cabi.TSKV_PT_F64, "!=", -0.0)],
        [(4, cabi.TSKV_PT_F64, "<", -1000.25), (1, cabi.TSKV_PT_I64, "!=", 0)],
        [(3, cabi.TSKV_PT_U64, ">", 2**63 + 9000)],
        [(1, cabi.TSKV_PT_I64, "==", 10**12)],                    # rules everything out
        [(4, cabi.TSKV_PT_F64, ">", float("nan"))],               # a NaN constant is never TRUE
    ]
    total_pruned = 0
    for preds in cases:
        for gbs in (False, True):
            for sel in (None, np.arange(0, 120, 2, dtype=np.uint32)):
                q = make_query(FIELDS[:2], aggs=ALL_AGGS if gbs else ("count", "sum", "min", "max", "mean"), series_ids=sel, width=50_000,
                               first_bucket_start=fbs, n_buckets=nb, group_by_series=gbs, predicates=preds)
                got = engine.scan_aggregate(pages, q)
                exp, pts = orc.scan_aggregate(arena, descs, q, return_points=True)
                assert_results_equal(got, exp, what="value stats %s gbs=%s sel=%s" % (preds, gbs, sel is not None))
                c = engine.counters()
                assert c["points_decoded"] == pts, (preds, c["points_decoded"], pts)
                total_pruned += c["pruned_page_count"]
    assert total_pruned > 0
    # the predicate that rules everything out reads nothing at all
    q = make_query(FIELDS[:2], aggs=("count",), predicates=[(1, cabi.TSKV_PT_I64, "==", 10**12)])
    engine.scan_aggregate(pages, q)
    assert engine.counters()["page_read_count"] == 0 and engine.counters()["points_decoded"] == 0
    pages.close()


def test_statistics_handed_in_by_the_caller_prune_host_resident_page_sets(engine):
    """tskvgpu_pages_set_value_stats: PageMeta.statistics from the TSM file. A host-resident page set (whose pages the
    library never reads ahead of a scan) then prunes by values too: fewer bytes cross PCIe, same results; loose bounds
    and pages without statistics are safe."""
    from cnosdb_b200 import tsmfile
    from tests.test_tsm_file import page_value_stats
    rng = np.random.default_rng(5)
    arena, descs = banded_arena(rng, n_series=80)
    bounds = []
    for i, d in enumerate(descs):
        if d["phys_type"] == cabi.TSKV_PT_TIME:
            t = orc.decode_pages(arena, descs, i, 1)[0][0].view(np.int64)
            bounds.append((int(t.min()), int(t.max())))
    stats = page_value_stats(arena, descs)
    f = tsmfile.load(tsmfile.write(arena, descs, np.array(bounds, dtype=np.int64), value_stats=stats))   # through a real file image
    assert (f.value_stats["flags"] == stats["flags"]).all()
    fbs, nb = bucket_spec(1_000_000, 1_000_000 + 1_600_000, 50_000)
    preds = [(1, cabi.TSKV_PT_I64, ">", 2500), (4, cabi.TSKV_PT_F64, ">=", 0.0)]
    q = make_query(FIELDS[:2], aggs=("count", "sum", "min", "max", "mean"), width=50_000, first_bucket_start=fbs, n_buckets=nb, predicates=preds)
    exp, pts = orc.scan_aggregate(f.arena, f.descs, q, return_points=True)
    hp = engine.upload_pages(f.arena, f.descs, verify_crc=True, host_resident=True)
    assert_results_equal(engine.scan_aggregate(hp, q), exp, what="host-resident, no statistics")
    bytes_without = engine.counters()["page_read_bytes"]
    assert engine.counters()["pruned_page_count"] == 0
    hp.set_value_stats(f.value_stats)
    assert_results_equal(engine.scan_aggregate(hp, q), exp, what="host-resident, caller statistics")
    c = engine.counters()
    assert c["pruned_page_count"] > 0 and c["page_read_bytes"] < bytes_without and c["points_decoded"] == pts
    loose = f.value_stats.copy()   # bounds may be loose, and some pages may come without statistics
    i64 = f.descs["phys_type"] == cabi.TSKV_PT_I64
    loose["min"][i64] = (loose["min"][i64].view(np.int64) - 700).view(np.uint64)
    loose["max"][i64] = (loose["max"][i64].view(np.int64) + 700).view(np.uint64)
    loose["flags"][::3] = 0
    hp.set_value_stats(loose)
    assert_results_equal(engine.scan_aggregate(hp, q), exp, what="host-resident, loose statistics")
    assert 0 < engine.counters()["pruned_page_count"] <= c["pruned_page_count"]
    hp.close()
    # an HBM-resident page set takes the caller's statistics instead of computing its own
    pages = engine.upload_pages(f.arena, f.descs)
    pages.set_value_stats(f.value_stats)
    assert_results_equal(engine.scan_aggregate(pages, q), exp, what="HBM-resident, caller statistics")
    assert engine.counters()["points_decoded"] == pts
    pages.close()
