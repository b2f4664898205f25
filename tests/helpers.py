"""Shared test helpers: random arenas, numpy brute-force aggregation, result comparison."""
import numpy as np

from cnosdb_b200 import cabi, datagen
from cnosdb_b200.engine import PushedAggregate, QueryOption

ALL_AGGS = ("count", "sum", "min", "max", "mean", "first", "last")


def bucket_spec(t_lo, t_hi, width, origin=0):
    """first_bucket_start / n_buckets covering [t_lo, t_hi] for t >= origin % width - width (floor regime)."""
    o = origin % width if origin >= 0 else -((-origin) % width)
    start = t_lo - ((t_lo - o + width) % width)
    n = (t_hi - start) // width + 1
    return start, int(n)


def assert_results_equal(got, exp, rtol=1e-6, what=""):
    """got/exp: ScanResult. Integers, counts, min/max/first/last bit-exact; f64 sum/mean within rtol."""
    assert got.names == exp.names
    for j, (col, agg) in enumerate(got.names):
        gv, ev = got.validity[j], exp.validity[j]
        assert (gv == ev).all(), "%s validity differs for col %s %s at %s" % (what, col, agg, np.nonzero(gv != ev)[0][:5])
        g, e = got.values[j][ev], exp.values[j][ev]
        pt = got.phys[col]
        if agg == "mean" or (agg == "sum" and pt == cabi.TSKV_PT_F64):
            gf, ef = g.view(np.float64), e.view(np.float64)
            ok = np.abs(gf - ef) <= rtol * np.maximum(np.abs(ef), 1e-300)
            assert ok.all(), "%s col %s %s: max rel err %g" % (what, col, agg, np.max(np.abs(gf - ef) / np.maximum(np.abs(ef), 1e-300)))
        else:
            bad = np.nonzero(g != e)[0]
            assert bad.size == 0, "%s col %s %s differs at %s: got %s exp %s" % (what, col, agg, bad[:5], g[bad[:5]], e[bad[:5]])
        # invalid cells hold 0
        assert (got.values[j][~ev] == 0).all()


def random_arena(rng, n_series=40, n_points=300, fields=((1, cabi.TSKV_PT_I64), (2, cabi.TSKV_PT_F64)),
                 null_frac=0.0, t0=1_000_000, step=1000, jitter=0, ids=None, raw_frac=0.0, multi_cg=False):
    """Hand-built arena via the product writer; returns (arena, descs, truth) where truth[series] is a
    list of column groups (ts, {col: (values, valid)})."""
    b = datagen.ArenaBuilder()
    truth = {}
    ids = list(range(n_series)) if ids is None else list(ids)
    for sid in ids:
        n_cg = 2 if (multi_cg and rng.random() < 0.3) else 1
        t_start = t0
        for _ in range(n_cg):
            n = int(n_points if not multi_cg else rng.integers(1, n_points + 1))
            ts = t_start + np.arange(n, dtype=np.int64) * step
            if jitter:
                ts = ts + rng.integers(-jitter, jitter + 1, n)
            t_start = int(ts[-1]) + step
            fl, cols = [], {}
            for col, pt in fields:
                valid = rng.random(n) >= null_frac if null_frac else None
                if pt == cabi.TSKV_PT_F64:
                    vals = np.cumsum(rng.integers(-3, 4, n)).astype(np.float64) + (rng.random(n) if rng.random() < 0.5 else 0)
                elif pt == cabi.TSKV_PT_U64:
                    vals = np.cumsum(rng.integers(0, 5, n)).astype(np.uint64) + np.uint64(2**63 - 100)
                else:
                    vals = np.cumsum(rng.integers(-50, 51, n)).astype(np.int64)
                enc = None
                if raw_frac and rng.random() < raw_frac:
                    enc = datagen.encode_raw
                fl.append((col, pt, vals, valid, enc))
                cols[col] = (vals, np.ones(n, dtype=bool) if valid is None else valid)
            b.add_column_group(sid, ts, fl)
            truth.setdefault(sid, []).append((ts, cols))
    arena, descs = b.finish()
    return arena, descs, truth


def make_query(fields, aggs=ALL_AGGS, **kw):
    cols = [PushedAggregate(c, pt, aggs) for c, pt in fields]
    return QueryOption(cols, **kw)
