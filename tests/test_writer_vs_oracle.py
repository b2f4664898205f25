"""The product's page writer (cnosdb_b200/csrc/host/tsm_writer.cc) is an independent implementation;
it must produce byte-identical output to the oracle's line-faithful restatement of the reference
encoders, and hit the reference's byte-exact vectors itself."""
import numpy as np
import pytest

from cnosdb_b200 import cabi, datagen
from oracle import pyoracle as orc


def test_writer_hits_reference_byte_vectors(golden):
    g = golden["codec_vectors"]
    r = g["i64_rle_regression"]
    assert datagen.encode_integers(np.full(r["count"], r["value"], dtype=np.int64))[1:].tolist() == r["enc_after_id"]
    s = g["i64_simple8b_short_regression"]
    assert datagen.encode_integers(s["values"])[1:].tolist() == s["enc_after_id"]
    u = g["u64_rle_bytes"]
    assert datagen.encode_integers(np.full(u["count"], u["value"], dtype=np.int64))[1:].tolist() == u["enc_after_id"]
    for case in g["simple8b_lengths"]:
        assert len(datagen.simple8b_pack(case["input"])) == case["encoded_len"]


def _rand_cases(rng):
    yield np.array([5], dtype=np.int64)
    yield np.array([5, 9], dtype=np.int64)
    yield np.array([7, 7, 7], dtype=np.int64)
    yield np.arange(0, 1000, dtype=np.int64) * 10_000_000_000 + 1_640_995_200_000_000_000
    yield np.cumsum(rng.integers(0, 3, 1000)).astype(np.int64)
    yield np.cumsum(rng.integers(-3, 4, 1000)).astype(np.int64)
    yield rng.integers(-2**62, 2**62, 50).astype(np.int64)          # raw (delta > 2^60)
    yield np.cumsum(np.ones(700, dtype=np.int64))                   # runs of ones: selectors 0/1
    yield np.cumsum(rng.integers(0, 2**40, 333)).astype(np.int64) * 1000  # scaler 10^3
    yield np.cumsum(rng.choice([1, 1, 1, 1, 5], 900)).astype(np.int64)
    for bits in (1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 15, 20, 30, 59):
        yield np.cumsum(rng.integers(0, 2**bits, 257)).astype(np.int64)


def test_timestamp_and_integer_writer_bytes_match_oracle():
    rng = np.random.default_rng(11)
    for v in _rand_cases(rng):
        assert datagen.encode_timestamps(v).tobytes() == orc.ts_encode(v).tobytes()
        assert datagen.encode_integers(v).tobytes() == orc.i64_encode(v).tobytes()


def test_float_writer_bytes_match_oracle(golden):
    rng = np.random.default_rng(12)
    cases = [np.array([int(b, 16) for b in t["input_bits"]], dtype=np.uint64).view(np.float64)
             for t in golden["codec_vectors"]["f64_roundtrip"]["tests"]]
    cases.append(np.array([int(b, 16) for b in golden["codec_vectors"]["f64_special_values"]["input_bits"]],
                          dtype=np.uint64).view(np.float64))
    cases.append(np.cumsum(rng.integers(-2, 3, 1000)).astype(np.float64))
    cases.append(np.cumsum(rng.integers(-2, 3, 1000)) + rng.random(1000))
    cases.append(rng.integers(0, 2**64, 500, dtype=np.uint64).view(np.float64))  # random bit patterns
    cases.append(np.array([1.5]))
    for v in cases:
        v = v[v.view(np.uint64) != 0x7ff80000000000ff]
        assert datagen.encode_floats(v).tobytes() == orc.f64_encode(v).tobytes()


def test_page_and_crc_match_oracle():
    rng = np.random.default_rng(13)
    data = rng.integers(0, 256, 999, dtype=np.uint8)
    assert cabi.load_hostgen_library().tskvw_crc32(data.ctypes.data, data.size) == orc.crc32(data)
    for rows in (1, 7, 8, 9, 1000):
        valid = rng.random(rows) > 0.2
        assert datagen.build_page(data, rows, valid).tobytes() == orc.page_build(data, rows, valid).tobytes()
        assert datagen.build_page(data, rows).tobytes() == orc.page_build(data, rows).tobytes()


def test_generator_is_deterministic_and_thread_independent():
    a = datagen.generate(64, n_fields=2, n_points=200, value_kind=datagen.MIXED, seed=9, n_threads=1,
                         jitter_permille=300, jitter_max=999, null_page_permille=200, null_row_permille=100)
    b = datagen.generate(64, n_fields=2, n_points=200, value_kind=datagen.MIXED, seed=9, n_threads=5,
                         jitter_permille=300, jitter_max=999, null_page_permille=200, null_row_permille=100)
    assert a.arena.tobytes() == b.arena.tobytes() and a.descs.tobytes() == b.descs.tobytes()
    # sharded generation (id % 2) produces the same pages as the matching ids of the full run
    s0 = datagen.generate(32, n_fields=2, n_points=200, value_kind=datagen.MIXED, seed=9, first_series_id=0,
                          series_stride=2, jitter_permille=300, jitter_max=999, null_page_permille=200,
                          null_row_permille=100)
    full = {}
    for d in a.descs:
        full[(int(d["series_id"]), int(d["column_id"]))] = a.arena[int(d["offset"]):int(d["offset"]) + int(d["size"])].tobytes()
    for d in s0.descs:
        assert int(d["series_id"]) % 2 == 0
        assert s0.arena[int(d["offset"]):int(d["offset"]) + int(d["size"])].tobytes() == full[(int(d["series_id"]), int(d["column_id"]))]
    # every generated page decodes in the oracle with a valid CRC
    pages = orc.decode_pages(a.arena, a.descs)
    assert len(pages) == len(a.descs)
    kinds = {(int(a.arena[int(d["offset"]) + 16 + (int(d["num_values"]) + 7) // 8]),
              int(a.arena[int(d["offset"]) + 17 + (int(d["num_values"]) + 7) // 8]) >> 4) for d in a.descs if d["phys_type"] == 0}
    assert (11, 2) in kinds and (11, 1) in kinds  # RLE and simple8b time pages both present
