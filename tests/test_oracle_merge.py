"""The oracle's restatement of the overlap merge (reader/iterator.rs:463-560 -> DataMerger -> sort_merge.rs +
batch_builder.rs) against the reference's own tests: the three merge tables of sort_merge.rs:449-539 and the grouping
table of reader/utils.rs:330-353. CPU only."""
import numpy as np

from cnosdb_b200 import cabi, datagen
from cnosdb_b200.engine import PushedAggregate, QueryOption
from oracle import pyoracle as orc


def build(streams, col_valid=None):
    """streams: list of (file_id, [times], [values or None]) - one column group per stream, series 7, column 1."""
    b = datagen.ArenaBuilder()
    files = []
    for fid, ts, vals in streams:
        ts = np.array(ts, dtype=np.int64)
        valid = np.array([v is not None for v in vals])
        v = np.array([0 if x is None else x for x in vals], dtype=np.int64)
        b.add_column_group(7, ts, [(1, cabi.TSKV_PT_I64, v, valid)])
        files.append(fid)
    arena, descs = b.finish()
    return arena, descs, np.array(files, dtype=np.uint64)


def per_time(arena, descs, files, times):
    """value of column 1 at every merged timestamp, through a 1-wide bucket scan: (count, sum) per time"""
    lo, hi = min(times), max(times)
    q = QueryOption([PushedAggregate(1, cabi.TSKV_PT_I64, ["count", "sum", "first", "last"])], width=1, first_bucket_start=lo, n_buckets=hi - lo + 1)
    return orc.scan_aggregate(arena, descs, q, chunk_files=files)


def test_merge_tree_table():
    # sort_merge.rs:449-478: column1 = time [1,1,1] [1,1,2] [1,2,2], column2 = [1,2,3] [4,5,6] [7,8,9] -> time [1,2], value [7,9]
    arena, descs, files = build([(1, [1, 1, 1], [1, 2, 3]), (2, [1, 1, 2], [4, 5, 6]), (3, [1, 2, 2], [7, 8, 9])])
    r = per_time(arena, descs, files, [1, 2])
    assert r.column(1, "count")[0].ravel().tolist() == [1, 1]
    assert r.column(1, "sum")[0].ravel().view(np.int64).tolist() == [7, 9]


def test_merge_column_table():
    # sort_merge.rs:481-510: values [1,None,3] [None,5,None] [None,8,None] -> [5, 8]
    arena, descs, files = build([(1, [1, 1, 1], [1, None, 3]), (2, [1, 1, 2], [None, 5, None]), (3, [1, 2, 2], [None, 8, None])])
    r = per_time(arena, descs, files, [1, 2])
    assert r.column(1, "sum")[0].ravel().view(np.int64).tolist() == [5, 8]
    assert r.column(1, "first")[0].ravel().view(np.int64).tolist() == [5, 8]


def test_merge_time_only_dedups_rows():
    # sort_merge.rs:513-539: times [1,1,1] [1,1,2] [1,2,2] -> [1, 2]: two merged rows
    arena, descs, files = build([(1, [1, 1, 1], [None] * 3), (2, [1, 1, 2], [None] * 3), (3, [1, 2, 2], [10, 20, 30])])
    q = QueryOption([PushedAggregate(1, cabi.TSKV_PT_I64, ["count", "sum"])])
    r = orc.scan_aggregate(arena, descs, q, chunk_files=files)
    assert int(r.column(1, "count")[0][0, 0]) == 2 and int(r.column(1, "sum")[0][0, 0]) == 10 + 30


def test_file_order_decides_not_arena_order():
    # the newer file (higher id) wins whatever the order of the column groups in the descriptor table
    for order in ([(9, [5, 6], [1, 2]), (3, [5, 6], [100, None])], [(3, [5, 6], [100, None]), (9, [5, 6], [1, 2])]):
        arena, descs, files = build(order)
        r = per_time(arena, descs, files, [5, 6])
        assert r.column(1, "sum")[0].ravel().view(np.int64).tolist() == [1, 2]
    # without file ids: one file, nothing merged (rows with equal times all count)
    q = QueryOption([PushedAggregate(1, cabi.TSKV_PT_I64, ["count"])])
    assert int(orc.scan_aggregate(arena, descs, q).column(1, "count")[0][0, 0]) == 3


def test_group_overlapping_segments_table():
    # reader/utils.rs:330-353: ranges (0,10) (1,3) (4,7) (6,10) | (11,14) (12,15) | (16,18) -> groups of 4, 2, 1 chunks.
    # Every chunk holds the same timestamp-free marker rows at its two ends; rows of different groups never merge.
    trs = [(0, 10), (1, 3), (4, 7), (6, 10), (11, 14), (12, 15), (16, 18)]
    streams = [(i + 1, [a, b], [1, 1]) for i, (a, b) in enumerate(trs)]
    arena, descs, files = build(streams)
    q = QueryOption([PushedAggregate(1, cabi.TSKV_PT_I64, ["count"])])
    merged = int(orc.scan_aggregate(arena, descs, q, chunk_files=files).column(1, "count")[0][0, 0])
    distinct_per_group = [len({0, 10, 1, 3, 4, 7, 6}), len({11, 14, 12, 15}), len({16, 18})]
    assert merged == sum(distinct_per_group)
    # a chunk that only touches the previous group's maximum joins it (min_ts <= running max): (10, 12) bridges groups 1 and 2
    streams.append((8, [10, 12], [1, 1]))
    arena, descs, files = build(streams)
    merged = int(orc.scan_aggregate(arena, descs, q, chunk_files=files).column(1, "count")[0][0, 0])
    assert merged == len({0, 10, 1, 3, 4, 7, 6, 11, 14, 12, 15}) + 2
