"""The two-phase gorilla decode of the cooperative kernels (cnosdb_b200/csrc/coop_kernels.cuh), as a small Python model
checked against the oracle's line-faithful decoder: phase 1 walks only the control bits and emits one record per element
{window bit offset, width, trailing zeros}; phase 2 extracts the windows independently and XOR-scans them. Pins the
record format and the end-of-stream rules (float.rs:480-591) the kernel implements. No GPU needed."""
import numpy as np
import pytest

from cnosdb_b200 import cabi, datagen
from oracle import pyoracle as orc

SENTINEL = 0x7FF80000000000FF


def bits_at(stream, pos, n):
    out = 0
    for k in range(n):
        byte, bit = (pos + k) >> 3, 7 - ((pos + k) & 7)
        out = (out << 1) | (int(stream[byte] >> bit) & 1 if byte < len(stream) else 0)
    return out


def two_phase(data, n_valid):
    """Returns (status, values): what scan_page_coop<.., VK_GOR, ..> decides for a page with n_valid valid rows."""
    stream, total = data[10:], (len(data) - 10) * 8
    first = int.from_bytes(bytes(data[2:10]), "big")
    recs, bitpos, mean, trail, n_parsed = [None], 0, 64, 0, n_valid
    for e in range(1, n_valid + 1):                       # phase 1 (gor_parse_ctrl)
        x = bits_at(stream, min(bitpos, total), 13)
        c0, c1, lead, m = x & 0x1000, x & 0x0800, (x >> 6) & 31, x & 63
        if c0 and c1:
            mean, trail = (m, (64 - lead - m) & 0xFF) if m else (64, 0)
        ln, sig = (13 if c1 else 2, mean) if c0 else (1, 0)
        recs.append((bitpos + ln, sig, trail & 63))
        bitpos += ln + sig
        if bitpos > total and n_parsed == n_valid:
            n_parsed = e - 1
    vals, v, first_sentinel, end_ok = [first], first, None, False
    for i in range(1, n_parsed + 1):                      # phase 2 (window extraction + XOR scan)
        pos, sig, tr = recs[i]
        if sig:
            v ^= ((bits_at(stream, pos, 64) >> (64 - sig)) << tr) & (2**64 - 1)
        is_end = sig != 0 and v == SENTINEL               # the first value and "repeat" elements are never tested
        if i < n_valid:
            vals.append(v)
            if is_end and first_sentinel is None:
                first_sentinel = i
        else:
            end_ok = is_end
    if first_sentinel is not None:
        return cabi.TSKV_ERR_BITSET_MISMATCH, None
    if n_parsed < n_valid:
        return cabi.TSKV_ERR_SHORT_BLOCK, None
    while not end_ok:                                     # more elements than valid rows: walk on to the sentinel
        x = bits_at(stream, bitpos, 13)
        c0, c1, lead, m = x & 0x1000, x & 0x0800, (x >> 6) & 31, x & 63
        if c0 and c1:
            mean, trail = (m, (64 - lead - m) & 0xFF) if m else (64, 0)
        ln, sig = (13 if c1 else 2, mean) if c0 else (1, 0)
        if sig:
            v ^= ((bits_at(stream, bitpos + ln, 64) >> (64 - sig)) << (trail & 63)) & (2**64 - 1)
        bitpos += ln + sig
        if bitpos > total:
            return cabi.TSKV_ERR_SHORT_BLOCK, None
        end_ok = sig != 0 and v == SENTINEL
    return cabi.TSKV_OK, np.array(vals[:n_valid], dtype=np.uint64)


def oracle_decode(data, n_rows):
    b = datagen.ArenaBuilder()
    b.add_page(datagen.build_page(datagen.encode_timestamps(np.arange(n_rows)), n_rows), 0, 0, cabi.TSKV_PT_TIME, n_rows)
    b.add_page(datagen.build_page(data, n_rows), 0, 1, cabi.TSKV_PT_F64, n_rows)
    arena, descs = b.finish()
    return orc.decode_pages(arena, descs)[1][0].view(np.uint64)


@pytest.mark.parametrize("kind", ["integer_walk", "noise", "wide_exponents", "constant"])
def test_two_phase_decode_equals_the_serial_reference_decode(kind):
    rng = np.random.default_rng(len(kind))
    for _ in range(25):
        n = int(rng.integers(1, 300))
        if kind == "integer_walk":
            v = np.cumsum(rng.integers(-3, 4, n)).astype(np.float64)
        elif kind == "noise":
            v = np.cumsum(rng.integers(-3, 4, n)).astype(np.float64) + rng.random(n)
        elif kind == "wide_exponents":
            v = rng.standard_normal(n) * 10.0 ** rng.integers(-300, 300, n)
        else:
            v = np.full(n, 42.5)
        data = datagen.encode_floats(v)
        st, vals = two_phase(data, n)
        assert st == cabi.TSKV_OK
        assert (vals == v.view(np.uint64)).all() and (vals == oracle_decode(data, n)).all()


def test_two_phase_end_of_stream_rules_match_the_oracle():
    n = 120
    v = np.cumsum(np.arange(n) % 5).astype(np.float64) * 0.37 + 1.5
    cases = {
        "extra_values": (datagen.encode_floats(np.concatenate([v, np.arange(30) * 3.25])), cabi.TSKV_OK),
        "early_sentinel": (datagen.encode_floats(v[:90]), cabi.TSKV_ERR_BITSET_MISMATCH),
        "truncated": (datagen.encode_floats(v)[:-24], cabi.TSKV_ERR_SHORT_BLOCK),
        "first_value_only": (np.frombuffer(bytes([6, 0x10]) + SENTINEL.to_bytes(8, "big"), dtype=np.uint8), cabi.TSKV_ERR_SHORT_BLOCK),
    }
    for name, (data, want) in cases.items():
        st, vals = two_phase(data, n)
        assert st == want, name
        if want == cabi.TSKV_OK:
            assert (vals == oracle_decode(data, n)).all(), name
        else:
            with pytest.raises(orc.OracleError) as e:
                oracle_decode(data, n)
            assert e.value.status == want, name
    # sentinel-valued data: the first value and repeats are pushed untested (float.rs:437,493-497)
    v25 = int(np.float64(2.5).view(np.uint64))
    bits = "0" * 6 + "11" + "00000" + "000000" + format(SENTINEL ^ v25, "064b") + "11" + "00000" + "000000" + format(v25 ^ SENTINEL, "064b")
    data = np.frombuffer(bytes([6, 0x10]) + SENTINEL.to_bytes(8, "big") + int(bits, 2).to_bytes(len(bits) // 8, "big"), dtype=np.uint8)
    st, vals = two_phase(data, 8)
    assert st == cabi.TSKV_OK and vals.tolist() == [SENTINEL] * 7 + [v25]
    assert (vals == oracle_decode(data, 8)).all()
