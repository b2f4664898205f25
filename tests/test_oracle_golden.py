"""Pins the CPU oracle against the reference's own golden vectors (tests/golden/*.json, extracted from
/root/reference by tests/golden/make_golden.py). No GPU needed."""
import numpy as np
import pytest

from cnosdb_b200 import cabi
from oracle import pyoracle as orc


def bits_to_f64(bits):
    return np.array([int(b, 16) for b in bits], dtype=np.uint64).view(np.float64)


def test_zigzag_table(golden):  # integer.rs:270-280
    g = golden["codec_vectors"]["i64_zigzag_table"]
    for v, e in zip(g["input"], g["expected"]):
        assert orc.lib().orc_zigzag_encode(v) == e
        assert orc.lib().orc_zigzag_decode(e) == v


def test_influx_rle_bytes(golden):  # integer.rs:438-459, byte-exact
    g = golden["codec_vectors"]["i64_rle_regression"]
    vals = np.full(g["count"], g["value"], dtype=np.int64)
    enc = orc.i64_encode(vals)
    assert enc[0] == cabi.TSKV_ENC_DELTA
    assert enc[1:].tolist() == g["enc_after_id"]
    out, valid = orc.decode_column(cabi.TSKV_PT_I64, enc, len(vals))
    assert valid.all() and (out.view(np.int64) == vals).all()


def test_influx_simple8b_short_bytes(golden):  # integer.rs:463-483, byte-exact
    g = golden["codec_vectors"]["i64_simple8b_short_regression"]
    enc = orc.i64_encode(g["values"])
    assert enc[1:].tolist() == g["enc_after_id"]
    out, _ = orc.decode_column(cabi.TSKV_PT_I64, enc, 1)
    assert out.view(np.int64).tolist() == g["values"]


def test_u64_rle_bytes(golden):  # unsigned.rs:198-213, byte-exact
    g = golden["codec_vectors"]["u64_rle_bytes"]
    vals = np.full(g["count"], g["value"], dtype=np.uint64)
    enc = orc.i64_encode(vals.view(np.int64))
    assert enc[1:].tolist() == g["enc_after_id"]
    out, _ = orc.decode_column(cabi.TSKV_PT_U64, enc, len(vals))
    assert (out == vals).all()


def test_simple8b_lengths_and_bounds(golden):  # simple8b.rs:231-261
    for case in golden["codec_vectors"]["simple8b_lengths"]:
        enc = orc.simple8b_encode(case["input"])
        assert len(enc) == case["encoded_len"]
        assert orc.simple8b_decode(enc).tolist() == case["input"]
    with pytest.raises(orc.OracleError):
        orc.simple8b_encode(golden["codec_vectors"]["simple8b_too_big"]["input"])


def test_simple8b_every_width():  # simple8b.rs:264-371 (the reference seeds rand::StdRng, not reproducible
    rng = np.random.default_rng(231)  # here: same structure, own seed)
    for bits in (1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 15, 20, 30, 60):
        v = rng.integers(0, 1 << bits, 100, dtype=np.uint64)
        v |= (np.arange(100, dtype=np.uint64) & 1) << np.uint64(bits - 1)
        assert (orc.simple8b_decode(orc.simple8b_encode(v)) == v).all()
    ones = np.ones(240, dtype=np.uint64)
    assert len(orc.simple8b_encode(ones)) == 8
    for n, idx in ((240, 120), (240, 119), (241, 239)):
        v = np.ones(n, dtype=np.uint64)
        v[idx] = 5
        assert (orc.simple8b_decode(orc.simple8b_encode(v)) == v).all()


@pytest.mark.parametrize("group,kind", [("i64_rle", 2), ("i64_simple8b", 1)])
def test_i64_corpora(golden, group, kind):  # integer.rs:348-434
    for t in golden["codec_vectors"][group]["tests"]:
        vals = np.array(t["input"], dtype=np.int64)
        enc = orc.i64_encode(vals)
        assert enc[1] >> 4 == kind, t["name"]
        out, valid = orc.decode_column(cabi.TSKV_PT_I64, enc, len(vals))
        assert valid.all() and (out.view(np.int64) == vals).all(), t["name"]


def test_i64_uncompressed(golden):  # integer.rs:297-315
    vals = np.array(golden["codec_vectors"]["i64_uncompressed"]["input"], dtype=np.int64)
    enc = orc.i64_encode(vals)
    assert enc[1] >> 4 == 0
    out, _ = orc.decode_column(cabi.TSKV_PT_I64, enc, len(vals))
    assert (out.view(np.int64) == vals).all()


@pytest.mark.parametrize("group,kind", [("ts_rle", 2), ("ts_simple8b", 1)])
def test_ts_corpora(golden, group, kind):  # timestamp.rs:415-512
    for t in golden["codec_vectors"][group]["tests"]:
        vals = np.array(t["input"], dtype=np.int64)
        enc = orc.ts_encode(vals)
        assert enc[0] == cabi.TSKV_ENC_DELTA_TS and enc[1] >> 4 == kind, t["name"]
        out, valid = orc.decode_column(cabi.TSKV_PT_TIME, enc, len(vals))
        assert valid.all() and (out.view(np.int64) == vals).all(), t["name"]


def test_ts_uncompressed(golden):  # timestamp.rs:370-386
    vals = np.array(golden["codec_vectors"]["ts_uncompressed"]["input"], dtype=np.int64)
    enc = orc.ts_encode(vals)
    assert enc[1] >> 4 == 0
    out, _ = orc.decode_column(cabi.TSKV_PT_TIME, enc, len(vals))
    assert (out.view(np.int64) == vals).all()


def test_u64_corpora(golden):  # unsigned.rs:105-240
    g = golden["codec_vectors"]
    cases = [g["u64_uncompressed"]["input"]] + [t["input"] for t in g["u64_rle"]["tests"]] + \
            [t["input"] for t in g["u64_simple8b"]["tests"]]
    for c in cases:
        vals = np.array(c, dtype=np.uint64)
        out, _ = orc.decode_column(cabi.TSKV_PT_U64, orc.i64_encode(vals.view(np.int64)), len(vals))
        assert (out == vals).all()


def test_f64_special_values(golden):  # float.rs:635-665: NaN payloads, +-inf, stale NaN: bit-exact
    vals = bits_to_f64(golden["codec_vectors"]["f64_special_values"]["input_bits"])
    enc = orc.f64_encode(vals)
    assert enc[0] == cabi.TSKV_ENC_GORILLA and enc[1] == 0x10
    out, valid = orc.decode_column(cabi.TSKV_PT_F64, enc, len(vals))
    assert valid.all() and (out == vals.view(np.uint64)).all()


def test_f64_corpora(golden):  # float.rs:698-1811 incl. "1000 real CPU values"
    tests = golden["codec_vectors"]["f64_roundtrip"]["tests"]
    assert [t["name"] for t in tests][-1] == "1000 real CPU values"
    for t in tests:
        vals = bits_to_f64(t["input_bits"])
        enc = orc.f64_encode(vals)
        out, valid = orc.decode_column(cabi.TSKV_PT_F64, enc, len(vals))
        assert valid.all() and (out == vals.view(np.uint64)).all(), t["name"]


def test_f64_sentinel_rejected():  # float.rs:58-60
    with pytest.raises(orc.OracleError):
        orc.f64_encode(np.array([1.0, np.array([0x7ff80000000000ff], dtype=np.uint64).view(np.float64)[0]]))


def test_window_kats(golden):  # time_window.rs:318-368
    for c in golden["window_kat"]["cases"]:
        fn = orc.ceil_sliding_window if c["ceil"] else orc.floor_sliding_window
        assert fn(c["t"], c["window"], c["slide"], c["start_time"]) == (c["start"], c["end"]), c


def test_crc32_known_answer():  # CRC-32/IEEE check value
    assert orc.crc32(np.frombuffer(b"123456789", dtype=np.uint8)) == 0xCBF43926


def test_page_roundtrip_with_nulls():  # page.rs:334-345 + :58-94
    rng = np.random.default_rng(5)
    vals = rng.integers(-1000, 1000, 77)
    valid = rng.random(77) > 0.3
    page = orc.page_build(orc.i64_encode(vals[valid]), 77, valid)
    assert int.from_bytes(page[0:4].tobytes(), "big") == 10 and int.from_bytes(page[4:12].tobytes(), "big") == 77
    descs = np.array([(0, len(page), 77, 1, 1, cabi.TSKV_PT_I64, 0)], dtype=cabi.PAGE_DESC_DTYPE)
    (out, ov), = orc.decode_pages(page, descs)
    assert (ov == valid).all() and (out.view(np.int64)[valid] == vals[valid]).all() and (out[~valid] == 0).all()
    page[-1] ^= 1  # corrupt the data => TsmPageFileHashCheckFailed
    with pytest.raises(orc.OracleError) as e:
        orc.decode_pages(page, descs)
    assert e.value.status == cabi.TSKV_ERR_CRC_MISMATCH


def test_crc32_matches_an_independent_implementation():
    """crc32fast (page.rs:62-65, test page.rs:651-690 over b"hello world") is CRC-32/IEEE: zlib's is the same polynomial.
    Every tail length of the slicing-by-8 loop, plus the reference test's page bytes."""
    import zlib
    rng = np.random.default_rng(32)
    for n in list(range(0, 40)) + [255, 256, 257, 4095, 7001]:
        buf = rng.integers(0, 256, n, dtype=np.uint8)
        assert orc.crc32(buf) == zlib.crc32(buf.tobytes()), n
    hello = np.frombuffer(b"hello world", dtype=np.uint8)
    assert orc.crc32(hello) == zlib.crc32(b"hello world") == 0x0D4A1185
    # create_test_page(): bitset_len 0 | rows 1 | crc | (no bitset) | data
    page = np.frombuffer((0).to_bytes(4, "big") + (1).to_bytes(8, "big") + (0x0D4A1185).to_bytes(4, "big") + b"hello world", dtype=np.uint8)
    assert int.from_bytes(page[12:16].tobytes(), "big") == orc.crc32(page[16:])
