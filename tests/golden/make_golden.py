#!/usr/bin/env python
"""Extract the reference's own golden vectors / known-answer tests for the hot path into JSON.

Run in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Only DATA is extracted (test inputs, expected bytes, expected values) - never code. Sources:
  tskv/src/tsm/codec/{integer,timestamp,unsigned,float,simple8b}.rs  `mod tests`
  query_server/query/src/extension/expr/window/time_window.rs:318-368
  query_server/sqllogicaltests/cases/function/{setup.slt,common/*.slt}
f64 values are stored as u64 bit patterns (hex) so that NaN payloads survive JSON.
"""
import json
import math
import os
import re
import struct
import sys

REF = os.environ.get("TSKV_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
S8B_MAX = (1 << 60) - 1


def f64_bits(x: float) -> str:
    return "0x%016x" % struct.unpack("<Q", struct.pack("<d", x))[0]


def parse_scalar(tok: str, as_float: bool):
    tok = tok.strip()
    tok = re.sub(r"\s+as\s+(i64|u64|f64)$", "", tok)
    tok = re.sub(r"(i64|u64|f64)$", "", tok) if re.match(r"^-?[0-9_]+(i64|u64)$", tok) else tok
    if tok == "simple8b::MAX_VALUE":
        return S8B_MAX
    m = re.match(r"^f64::from_bits\((0x[0-9a-fA-F_]+)\)$", tok)
    if m:
        return ("bits", int(m.group(1).replace("_", ""), 16))
    if tok == "f64::NAN":
        return ("bits", 0x7FF8000000000000)
    if tok == "f64::INFINITY":
        return ("bits", 0x7FF0000000000000)
    if tok == "f64::NEG_INFINITY":
        return ("bits", 0xFFF0000000000000)
    tok = tok.replace("_", "")
    if as_float:
        return float(tok)
    return int(tok, 0)


def split_top(s: str):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x for x in (y.strip() for y in out) if x and not x.startswith("//")]


def strip_comments(s: str) -> str:
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    return re.sub(r"//[^\n]*", "", s)


def parse_vec(body: str, as_float: bool):
    body = strip_comments(body).strip()
    m = re.match(r"^(.*);\s*([0-9_]+)$", body, flags=re.S)
    if m:  # vec![x; n]
        return [parse_scalar(m.group(1), as_float)] * int(m.group(2).replace("_", ""))
    return [parse_scalar(t, as_float) for t in split_top(body)]


def find_vec(text: str, start: int):
    """Return (body, end) of the vec![...] starting at/after `start`."""
    i = text.index("vec![", start) + 5
    depth, j = 1, i
    while depth:
        if text[j] == "[":
            depth += 1
        elif text[j] == "]":
            depth -= 1
        j += 1
    return text[i : j - 1], j


def named_tests(text: str, fn_name: str, as_float: bool):
    """All `Test { name: String::from("..."), input: vec![...] }` inside `fn fn_name()`."""
    m = re.search(r"fn %s\(\)" % re.escape(fn_name), text)
    assert m, fn_name
    nxt = re.search(r"\n    fn |\n    #\[test\]", text[m.end() :])
    seg = text[m.end() : m.end() + (nxt.start() if nxt else len(text))]
    out = []
    for tm in re.finditer(r'name:\s*String::from\("([^"]*)"\),\s*input:', seg):
        body, _ = find_vec(seg, tm.end())
        out.append({"name": tm.group(1), "input": parse_vec(body, as_float)})
    return out


def to_f64_bits(vals):
    return [("0x%016x" % v[1]) if isinstance(v, tuple) else f64_bits(v) for v in vals]


def read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def main():
    codec = {}
    # ---- integer.rs -------------------------------------------------------------------------
    t = read("tskv/src/tsm/codec/integer.rs")
    codec["i64_zigzag_table"] = {  # integer.rs:270-280
        "input": parse_vec(find_vec(t, t.index("fn zig_zag_encoding"))[0], False),
        "expected": parse_vec(find_vec(t, t.index("let exp = "))[0].replace("[", "").replace("]", ""), False)
        if False
        else [int(x) for x in re.search(r"let exp = \[([^\]]*)\]", t).group(1).split(",")],
        "src": "tskv/src/tsm/codec/integer.rs:270-280",
    }
    codec["i64_zigzag_table"]["input"] = [
        int(x) for x in re.search(r"let input = \[([^\]]*)\]", t).group(1).split(",")
    ]
    codec["i64_uncompressed"] = {
        "input": parse_vec(find_vec(t, t.index("fn encode_uncompressed"))[0], False),
        "kind": 0,
        "src": "tskv/src/tsm/codec/integer.rs:297-315",
    }
    codec["i64_rle"] = {"tests": named_tests(t, "encode_rle", False), "kind": 2,
                        "src": "tskv/src/tsm/codec/integer.rs:348-396"}
    codec["i64_simple8b"] = {"tests": named_tests(t, "encode_simple8b", False), "kind": 1,
                             "src": "tskv/src/tsm/codec/integer.rs:399-434"}
    m = re.search(r"vec!\[(\d+)i64; (\d+)\];.*?let enc_influx = \[([^\]]*)\]", t, flags=re.S)
    codec["i64_rle_regression"] = {  # byte-exact InfluxDB vector
        "value": int(m.group(1)), "count": int(m.group(2)),
        "enc_after_id": [int(x) for x in m.group(3).split(",")],
        "src": "tskv/src/tsm/codec/integer.rs:438-459",
    }
    m = re.search(r"fn simple8b_short_regression.*?vec!\[(\d+)\];.*?let enc_influx = \[([^\]]*)\]", t, flags=re.S)
    codec["i64_simple8b_short_regression"] = {
        "values": [int(m.group(1))],
        "enc_after_id": [int(x) for x in m.group(2).split(",")],
        "src": "tskv/src/tsm/codec/integer.rs:463-483",
    }
    # ---- unsigned.rs ------------------------------------------------------------------------
    t = read("tskv/src/tsm/codec/unsigned.rs")
    m = re.search(r"vec!\[(\d+)u64; (\d+)\];.*?let expected_encoded = vec!\[([^\]]*)\]", t, flags=re.S)
    codec["u64_rle_bytes"] = {
        "value": int(m.group(1)), "count": int(m.group(2)),
        "enc_after_id": [int(x) for x in m.group(3).split(",")],
        "src": "tskv/src/tsm/codec/unsigned.rs:198-213",
    }
    codec["u64_uncompressed"] = {
        "input": parse_vec(find_vec(t, t.index("fn encode_uncompressed"))[0], False),
        "src": "tskv/src/tsm/codec/unsigned.rs:105-123",
    }
    codec["u64_rle"] = {"tests": named_tests(t, "encode_rle", False),
                        "src": "tskv/src/tsm/codec/unsigned.rs:152-196"}
    codec["u64_simple8b"] = {"tests": named_tests(t, "encode_simple8b", False),
                             "src": "tskv/src/tsm/codec/unsigned.rs:215-240"}
    # ---- timestamp.rs -----------------------------------------------------------------------
    t = read("tskv/src/tsm/codec/timestamp.rs")
    codec["ts_uncompressed"] = {
        "input": parse_vec(find_vec(t, t.index("fn encode_uncompressed"))[0], False), "kind": 0,
        "src": "tskv/src/tsm/codec/timestamp.rs:370-386"}
    codec["ts_rle"] = {"tests": named_tests(t, "encode_rle", False), "kind": 2,
                       "src": "tskv/src/tsm/codec/timestamp.rs:415-472"}
    codec["ts_simple8b"] = {"tests": named_tests(t, "encode_simple8b", False), "kind": 1,
                            "src": "tskv/src/tsm/codec/timestamp.rs:475-512"}
    # ---- simple8b.rs ------------------------------------------------------------------------
    t = read("tskv/src/tsm/codec/simple8b.rs")
    codec["simple8b_lengths"] = [
        {"input": parse_vec(find_vec(t, t.index("fn test_encode_mixed_sizes()"))[0], False),
         "encoded_len": 16, "src": "tskv/src/tsm/codec/simple8b.rs:231-240"},
        {"input": parse_vec(find_vec(t, t.index("fn test_encode_mixed_sizes_alt()"))[0], False),
         "encoded_len": 24, "src": "tskv/src/tsm/codec/simple8b.rs:243-252"},
    ]
    assert "assert_eq!(encoded.len(), 16)" in t and "assert_eq!(encoded.len(), 24)" in t
    codec["simple8b_too_big"] = {"input": [7, 6, 2 << 60, 4, 3, 2, 1],
                                 "src": "tskv/src/tsm/codec/simple8b.rs:255-261"}
    # ---- float.rs ---------------------------------------------------------------------------
    t = read("tskv/src/tsm/codec/float.rs")
    sv = parse_vec(find_vec(t, t.index("fn encode_special_values"))[0], True)
    codec["f64_special_values"] = {"input_bits": to_f64_bits(sv),
                                   "src": "tskv/src/tsm/codec/float.rs:635-665"}
    ft = named_tests(t, "encode", True)
    codec["f64_roundtrip"] = {
        "tests": [{"name": x["name"], "input_bits": to_f64_bits(x["input"])} for x in ft],
        "src": "tskv/src/tsm/codec/float.rs:698-1811"}
    assert len(ft) == 7 and len(ft[-1]["input"]) == 1000, [len(x["input"]) for x in ft]
    with open(os.path.join(OUT, "codec_vectors.json"), "w") as f:
        json.dump(codec, f, indent=0, separators=(",", ":"))

    # ---- time_window.rs KATs ------------------------------------------------------------------
    t = read("query_server/query/src/extension/expr/window/time_window.rs")
    kats = []
    for fn, first in (("test_first_sliding_window_start_bound", True),
                      ("test_last_sliding_window_start_bound", False)):
        i = t.index("fn " + fn)
        a, e1 = find_vec(t, i)
        b, _ = find_vec(t, e1)
        tup = lambda s: [tuple(int(x) for x in m.group(1).split(",")) for m in
                         re.finditer(r"\(([-0-9, ]+)\)", strip_comments(s))]
        for arg, exp in zip(tup(a), tup(b)):
            kats.append({"ceil": first, "t": arg[0], "window": arg[1], "slide": arg[2],
                         "start_time": arg[3], "start": exp[0], "end": exp[1]})
    assert len(kats) == 14, len(kats)
    with open(os.path.join(OUT, "window_kat.json"), "w") as f:
        json.dump({"src": "query_server/query/src/extension/expr/window/time_window.rs:318-368",
                   "cases": kats}, f, indent=0)

    # ---- SQL goldens (func_tb2: ns timestamps 100..107) ---------------------------------------
    setup = read("query_server/sqllogicaltests/cases/function/setup.slt")
    rows = re.findall(r"\((\d+), (\d+), (\d+), (true|false), '[^']*', (-?\d+), '[^']*', '[^']*', '[^']*'\)",
                      setup)
    assert len(rows) == 8, len(rows)
    tb2 = {"time": [int(r[0]) for r in rows], "f0_u64": [int(r[1]) for r in rows],
           "f1_f64": [float(r[2]) for r in rows], "f4_i64": [int(r[4]) for r in rows]}
    goldens = {}
    for agg in ("sum", "avg", "min", "max", "count", "first", "last"):
        txt = read("query_server/sqllogicaltests/cases/function/common/%s.slt" % agg)
        for col in ("f0", "f1", "f4"):
            arg = r"time,\s*%s" % col if agg in ("first", "last") else col
            m = re.search(r"select\s+%s\(%s\)\s+from\s+func_tb2;\n----\n([^\n]+)\n" % (agg, arg), txt,
                          flags=re.I)
            if m:
                goldens["%s(%s)" % (agg, col)] = m.group(1).strip()
    with open(os.path.join(OUT, "sql_goldens.json"), "w") as f:
        json.dump({"src": "query_server/sqllogicaltests/cases/function/setup.slt:46-56 + common/*.slt",
                   "func_tb2": tb2, "expected": goldens}, f, indent=0)
    print("codec groups:", len(codec), "window KATs:", len(kats), "sql goldens:", len(goldens))


if __name__ == "__main__":
    sys.exit(main())
