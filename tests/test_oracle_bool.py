"""Boolean codec (tskv/src/tsm/codec/boolean.rs): the oracle's restatement against the reference's own vectors
(boolean.rs:150-260) and against the product's independent numpy encoder. CPU only."""
import numpy as np
import pytest

from cnosdb_b200 import cabi, datagen
from oracle import pyoracle as orc


def test_bitpack_encoder_vectors():
    # boolean.rs encode_no_values / encode_single_true / encode_single_false / encode_multi_compressed
    assert orc.bool_encode([]).size == 0
    assert orc.bool_encode([True])[1:].tolist() == [16, 1, 128]
    assert orc.bool_encode([False])[1:].tolist() == [16, 1, 0]
    assert orc.bool_encode([i % 2 == 0 for i in range(10)])[1:].tolist() == [16, 10, 170, 128]
    assert orc.bool_encode([True])[0] == cabi.TSKV_ENC_BITPACK


def test_bitpack_decoder_vectors():
    # decode_single_true / decode_single_false / decode_multi_compressed (src = id byte + the vectors above)
    for src, exp in (([0, 16, 1, 128], [True]), ([0, 16, 1, 0], [False]), ([0, 16, 10, 170, 128], [i % 2 == 0 for i in range(10)])):
        v, ok = orc.decode_column(cabi.TSKV_PT_BOOL, np.array(src, dtype=np.uint8), len(exp))
        assert ok.all() and v.astype(bool).tolist() == exp
    v, ok = orc.decode_column(cabi.TSKV_PT_BOOL, np.zeros(0, dtype=np.uint8), 5)  # empty buffer: all null
    assert not ok.any()


def test_round_trips_nulls_and_the_independent_writer():
    rng = np.random.default_rng(3)
    for n in (1, 7, 8, 9, 127, 128, 129, 1000, 20_000):
        vals = rng.random(n) < 0.4
        valid = rng.random(n) < 0.8
        kept = vals[valid]
        for enc_o, enc_w in ((orc.bool_encode, datagen.encode_bools), (orc.bool_raw_encode, datagen.encode_bools_raw)):
            if kept.size == 0:
                continue
            a, b = enc_o(kept), enc_w(kept)
            assert a.tolist() == b.tolist()       # the product's writer and the oracle's encoder agree byte for byte
            v, ok = orc.decode_column(cabi.TSKV_PT_BOOL, a, n, valid)
            assert (ok == valid).all() and (v[valid].astype(bool) == kept).all() and (v[~valid] == 0).all()


def test_malformed_blocks():
    good = orc.bool_encode([True] * 20)
    with pytest.raises(orc.OracleError) as e:   # fewer values than valid rows: "Insufficient data for decoding"
        orc.decode_column(cabi.TSKV_PT_BOOL, good, 21)
    assert e.value.status == cabi.TSKV_ERR_BITSET_MISMATCH
    bad = good.copy()
    bad[1] = 0x20
    with pytest.raises(orc.OracleError) as e:   # assert_eq!(src[0], 1 << 4)
        orc.decode_column(cabi.TSKV_PT_BOOL, bad, 20)
    assert e.value.status == cabi.TSKV_ERR_BAD_ENCODING
    with pytest.raises(orc.OracleError) as e:   # the count's varint never ends
        orc.decode_column(cabi.TSKV_PT_BOOL, np.array([10, 16, 0x80], dtype=np.uint8), 1)
    assert e.value.status == cabi.TSKV_ERR_SHORT_BLOCK
    with pytest.raises(orc.OracleError) as e:   # count says 100, the block holds 8 bits
        orc.decode_column(cabi.TSKV_PT_BOOL, np.array([10, 16, 100, 0xFF], dtype=np.uint8), 9)
    assert e.value.status == cabi.TSKV_ERR_BITSET_MISMATCH
