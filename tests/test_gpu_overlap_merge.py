"""Overlapping chunks through the merge pass (cnosdb_b200/csrc/merge_kernels.cuh) against the oracle's restatement of
DataMerger / sort_merge / BatchMergeBuilder (which tests/test_oracle_merge.py pins to the reference's own tables)."""
import numpy as np
import pytest

from cnosdb_b200 import cabi, datagen
from cnosdb_b200.engine import PushedAggregate, QueryOption, TskvError
from oracle import pyoracle as orc
from tests.helpers import ALL_AGGS, assert_results_equal, bucket_spec, make_query
from tests.test_gpu_parity import random_tombstones

pytestmark = pytest.mark.gpu

FIELDS = ((1, cabi.TSKV_PT_I64), (2, cabi.TSKV_PT_F64), (3, cabi.TSKV_PT_U64))


def overlapping_arena(rng, n_series=60, raw_memcache=True):
    """Per series 1..4 files, each with 1..2 column groups; timestamps on a common 1000-step grid (plus a little jitter
    for some), so chunks of different files share many timestamps; 25 % nulls; the newest file of some series is a
    'memcache row group': raw-encoded pages. Returns (arena, descs, file id per column group)."""
    b = datagen.ArenaBuilder()
    files = []
    for sid in range(n_series):
        n_files = int(rng.integers(1, 5))
        for f in range(n_files):
            file_id = 10 * (f + 1) + int(rng.integers(0, 3))
            start = int(rng.integers(0, 400))
            for part in range(int(rng.integers(1, 3))):
                n = int(rng.integers(1, 300))
                idx = start + np.sort(rng.choice(np.arange(0, 2 * n + 5), n, replace=False))
                start = int(idx[-1]) + 1
                ts = 1_000_000 + idx.astype(np.int64) * 1000
                fl = []
                for col, pt in FIELDS:
                    if rng.random() < 0.15:
                        continue  # a column this column group does not hold
                    valid = rng.random(n) >= 0.25
                    if pt == cabi.TSKV_PT_F64:
                        vals = np.cumsum(rng.integers(-3, 4, n)).astype(np.float64) + rng.random(n)
                    elif pt == cabi.TSKV_PT_U64:
                        vals = np.cumsum(rng.integers(0, 5, n)).astype(np.uint64) + np.uint64(2**63 - 100)
                    else:
                        vals = np.cumsum(rng.integers(-50, 51, n)).astype(np.int64)
                    enc = datagen.encode_raw if (raw_memcache and f == n_files - 1 and sid % 3 == 0) else None
                    fl.append((col, pt, vals, valid, enc))
                if not fl:
                    fl.append((1, cabi.TSKV_PT_I64, np.arange(n, dtype=np.int64), None, None))
                b.add_column_group(sid, ts, fl)
                files.append(file_id)
    arena, descs = b.finish()
    return arena, descs, np.array(files, dtype=np.uint64)


@pytest.mark.parametrize("seed", [1, 2])
def test_overlapping_chunks_match_the_oracle_merge(engine, seed):
    rng = np.random.default_rng(seed)
    arena, descs, files = overlapping_arena(rng)
    pages = engine.upload_pages(arena, descs)
    pages.set_chunk_files(files)
    t_lo, t_hi = 1_000_000 - 500, 1_000_000 + 1_700_000
    fbs, nb = bucket_spec(t_lo, t_hi, 17_000, origin=3)
    sel = np.array(sorted(rng.choice(np.arange(70), 40, replace=False)), dtype=np.uint32)
    for gbs in (False, True):
        for series_ids in (sel, None):
            for ranges in ([], [(t_lo + 130_000, t_lo + 131_000), (t_lo + 300_500, t_lo + 655_000)]):
                q = make_query(FIELDS, series_ids=series_ids, time_ranges=ranges, origin=3, width=17_000, first_bucket_start=fbs,
                               n_buckets=nb, group_by_series=gbs)
                got = engine.scan_aggregate(pages, q)
                exp, pts = orc.scan_aggregate(arena, descs, q, chunk_files=files, return_points=True)
                assert_results_equal(got, exp, what="merge gbs=%s sel=%s %s" % (gbs, series_ids is not None, ranges))
                assert engine.counters()["points_decoded"] == pts
        q = make_query(FIELDS, series_ids=sel, group_by_series=gbs, time_ranges=[(t_lo + 100_000, t_lo + 900_000)])  # unbucketed
        assert_results_equal(engine.scan_aggregate(pages, q), orc.scan_aggregate(arena, descs, q, chunk_files=files), what="unbucketed")
    # merged vs unmerged really differ here, and clearing the file ids goes back to the one-file semantics
    q = make_query(FIELDS, aggs=("count", "sum"))
    merged = engine.scan_aggregate(pages, q)
    prepared = engine.prepare(pages, q)
    pages.set_chunk_files([])
    with pytest.raises(TskvError):  # prepared before the change
        prepared.run()
    prepared.close()
    plain = engine.scan_aggregate(pages, q)
    assert_results_equal(plain, orc.scan_aggregate(arena, descs, q), what="file ids cleared")
    assert int(merged.column(1, "count")[0][0, 0]) < int(plain.column(1, "count")[0][0, 0])
    pages.close()


def test_row_filter_runs_before_the_merge_and_tombstones_apply(engine):
    rng = np.random.default_rng(11)
    arena, descs, files = overlapping_arena(rng, n_series=40)
    pages = engine.upload_pages(arena, descs)
    pages.set_chunk_files(files)
    t_lo, t_hi = 1_000_000 - 500, 1_000_000 + 1_700_000
    fbs, nb = bucket_spec(t_lo, t_hi, 17_000, origin=3)
    proj = FIELDS[:2]
    for preds in ([(1, cabi.TSKV_PT_I64, ">", 0)], [(2, cabi.TSKV_PT_F64, "<=", 1.5), (3, cabi.TSKV_PT_U64, ">=", 2**63 - 50)]):
        for gbs in (False, True):
            q = make_query(proj, origin=3, width=17_000, first_bucket_start=fbs, n_buckets=nb, group_by_series=gbs, predicates=preds)
            assert_results_equal(engine.scan_aggregate(pages, q), orc.scan_aggregate(arena, descs, q, chunk_files=files),
                                 what="merge + predicates %s gbs=%s" % (preds, gbs))
    tombs = random_tombstones(rng, descs, t_lo, 1_000_000 + 700_000)
    pages.set_tombstones(tombs)
    for gbs in (False, True):
        q = make_query(FIELDS, origin=3, width=17_000, first_bucket_start=fbs, n_buckets=nb, group_by_series=gbs)
        assert_results_equal(engine.scan_aggregate(pages, q), orc.scan_aggregate(arena, descs, q, chunk_files=files, tombstones=tombs),
                             what="merge + tombstones gbs=%s" % gbs)
    pages.close()


def test_reference_merge_tables_on_the_gpu(engine):
    """sort_merge.rs:449-539 through the C ABI (the oracle is pinned to the same tables in tests/test_oracle_merge.py)."""
    def run(streams):
        b = datagen.ArenaBuilder()
        files = []
        for fid, ts, vals in streams:
            valid = np.array([v is not None for v in vals])
            b.add_column_group(7, np.array(ts, dtype=np.int64), [(1, cabi.TSKV_PT_I64, np.array([0 if x is None else x for x in vals], dtype=np.int64), valid)])
            files.append(fid)
        arena, descs = b.finish()
        pages = engine.upload_pages(arena, descs)
        pages.set_chunk_files(np.array(files, dtype=np.uint64))
        q = QueryOption([PushedAggregate(1, cabi.TSKV_PT_I64, ["count", "sum", "first", "last"])], width=1, first_bucket_start=1, n_buckets=2)
        r = engine.scan_aggregate(pages, q)
        pages.close()
        return r
    r = run([(1, [1, 1, 1], [1, 2, 3]), (2, [1, 1, 2], [4, 5, 6]), (3, [1, 2, 2], [7, 8, 9])])
    assert r.column(1, "sum")[0].ravel().view(np.int64).tolist() == [7, 9] and r.column(1, "count")[0].ravel().tolist() == [1, 1]
    r = run([(1, [1, 1, 1], [1, None, 3]), (2, [1, 1, 2], [None, 5, None]), (3, [1, 2, 2], [None, 8, None])])
    assert r.column(1, "sum")[0].ravel().view(np.int64).tolist() == [5, 8]
    assert r.column(1, "first")[0].ravel().view(np.int64).tolist() == [5, 8] and r.column(1, "last")[0].ravel().view(np.int64).tolist() == [5, 8]
    r = run([(1, [1, 1, 1], [None] * 3), (2, [1, 1, 2], [None] * 3), (3, [1, 2, 2], [10, 20, 30])])
    assert r.column(1, "count")[0].ravel().tolist() == [1, 1] and r.column(1, "sum")[0].ravel().view(np.int64).tolist() == [10, 30]
