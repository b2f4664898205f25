"""world_size-2 CPU (gloo) test of the multi-GPU exchange step: series sharded `id % N`, per-rank partial
sections combined element-wise (cnosdb_b200/parallel.py). The partials come from the oracle run on each
shard; the combined result must equal the oracle on the whole arena."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cnosdb_b200 import cabi, datagen
from cnosdb_b200.engine import PushedAggregate, QueryOption
from cnosdb_b200.parallel import KEY_FIRST_IDENTITY, KEY_LAST_IDENTITY, allreduce_sections, select_tag_subset, shard_range

W = 60_000_000_000


def _query(sel):
    lo = datagen.TSBS_T0
    start = lo - lo % W
    nb = (datagen.TSBS_T0 + 299 * datagen.TSBS_STEP - start) // W + 1
    return QueryOption([PushedAggregate(1, cabi.TSKV_PT_I64, ["count", "sum", "min", "max"]),
                        PushedAggregate(2, cabi.TSKV_PT_F64, ["count", "sum", "min", "max"])],
                       series_ids=sel, width=W, first_bucket_start=start, n_buckets=int(nb))


def _okey_f64(bits):
    b = bits.astype(np.int64)
    return b ^ ((b >> 63) & np.int64(0x7FFFFFFFFFFFFFFF))


def _sections_from_oracle(res, rng, rank, n_first):
    """Partial sections in the layout of tskvgpu_scan_partials (counts+int sums | f64 sums | min keys | max keys),
    plus synthetic FIRST/LAST (key, value) cells to exercise the masked value exchange."""
    c1, v1 = res.column(1, "count")
    c2, _ = res.column(2, "count")
    s1, _ = res.column(1, "sum")
    s2, sv2 = res.column(2, "sum")
    mn1, mv1 = res.column(1, "min")
    mx1, _ = res.column(1, "max")
    mn2, mv2 = res.column(2, "min")
    mx2, _ = res.column(2, "max")
    i64max, i64min = np.iinfo(np.int64).max, np.iinfo(np.int64).min
    mins = np.concatenate([np.where(mv1, mn1, i64max).ravel(), np.where(mv2, _okey_f64(mn2.view(np.uint64)), i64max).ravel()])
    maxs = np.concatenate([np.where(mv1, mx1, i64min).ravel(), np.where(mv2, _okey_f64(mx2.view(np.uint64)), i64min).ravel()])
    # synthetic first/last cells: key = ts << 4 | slot with slot parity == rank (unique across ranks)
    ts = rng.integers(1, 1000, n_first)
    slot = rng.integers(0, 8, n_first) * 2 + rank
    present = rng.random(n_first) < 0.7
    fkeys = np.where(present, (ts << 4) | slot, KEY_FIRST_IDENTITY)
    lkeys = np.where(present, (ts << 4) | (15 - slot), KEY_LAST_IDENTITY)
    fvals = np.where(present, rng.integers(-10**12, 10**12, n_first), 0)
    lvals = np.where(present, rng.integers(-10**12, 10**12, n_first), 0)
    return {
        "sum_i64": torch.from_numpy(np.concatenate([c1.ravel().astype(np.int64), c2.ravel().astype(np.int64), s1.ravel().astype(np.int64)])),
        "sum_f64": torch.from_numpy(np.where(sv2, s2, 0.0).ravel().copy()),
        "min_i64": torch.from_numpy(np.concatenate([mins, fkeys]).astype(np.int64)),
        "max_i64": torch.from_numpy(np.concatenate([maxs, lkeys]).astype(np.int64)),
        "sel_val": torch.from_numpy(np.concatenate([fvals, lvals]).astype(np.int64)),
        "first_len": n_first, "last_len": n_first,
    }, (fkeys, fvals, lkeys, lvals)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as orc
    n_series, n_first = 64, 40
    sel = select_tag_subset(n_series, 2)
    # each rank generates only its shard (a contiguous id range), like bench.py does per GPU
    lo, hi = shard_range(n_series, rank, world)
    g = datagen.generate(hi - lo, n_fields=1, n_points=300, value_kind=datagen.MIXED,
                         seed=21, first_series_id=lo, null_page_permille=200, null_row_permille=100)
    assert all(lo <= int(s) < hi for s in g.descs["series_id"])
    res = orc.scan_aggregate(g.arena, g.descs, _query(sel))   # global selection list, local pages
    rng = np.random.default_rng(100 + rank)
    sections, local_sel = _sections_from_oracle(res, rng, rank, n_first)
    allreduce_sections(sections)
    np.save(os.path.join(out_dir, "sel_%d.npy" % rank), np.stack(local_sel))
    if rank == 0:
        np.savez(os.path.join(out_dir, "reduced.npz"), **{k: v.numpy() for k, v in sections.items() if hasattr(v, "numpy")})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partial_exchange(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    red = np.load(tmp_path / "reduced.npz")
    from oracle import pyoracle as orc
    full = datagen.generate(64, n_fields=1, n_points=300, value_kind=datagen.MIXED, seed=21,
                            null_page_permille=200, null_row_permille=100)
    exp = orc.scan_aggregate(full.arena, full.descs, _query(select_tag_subset(64, 2)))
    nc = exp.column(1, "count")[0].size
    c1, c2, s1 = red["sum_i64"][:nc], red["sum_i64"][nc:2 * nc], red["sum_i64"][2 * nc:3 * nc]
    assert (c1 == exp.column(1, "count")[0].ravel().astype(np.int64)).all()
    assert (c2 == exp.column(2, "count")[0].ravel().astype(np.int64)).all()
    v = exp.column(1, "sum")[1].ravel()
    assert (s1[v] == exp.column(1, "sum")[0].ravel()[v]).all()
    v2 = exp.column(2, "sum")[1].ravel()
    assert np.allclose(red["sum_f64"][v2], exp.column(2, "sum")[0].ravel()[v2], rtol=1e-9)
    assert (red["min_i64"][:nc][v] == exp.column(1, "min")[0].ravel()[v]).all()
    assert (red["max_i64"][:nc][v] == exp.column(1, "max")[0].ravel()[v]).all()
    assert (red["min_i64"][nc:2 * nc][v2] == _okey_f64(exp.column(2, "min")[0].ravel().view(np.uint64))[v2]).all()
    assert (red["max_i64"][nc:2 * nc][v2] == _okey_f64(exp.column(2, "max")[0].ravel().view(np.uint64))[v2]).all()
    # first/last: the (key, value) of the rank holding the winning key survives, everything else is masked
    a, b = np.load(tmp_path / "sel_0.npy"), np.load(tmp_path / "sel_1.npy")
    nf = a.shape[1]
    fk = np.minimum(a[0], b[0])
    fv = np.where(a[0] < b[0], a[1], b[1])
    fv = np.where(fk == KEY_FIRST_IDENTITY, 0, fv)
    lk = np.maximum(a[2], b[2])
    lv = np.where(a[2] > b[2], a[3], b[3])
    lv = np.where(lk == KEY_LAST_IDENTITY, 0, lv)
    assert (red["min_i64"][2 * nc:] == fk).all() and (red["max_i64"][2 * nc:] == lk).all()
    assert (red["sel_val"][:nf] == fv).all() and (red["sel_val"][nf:] == lv).all()
