// oracle/codecs.cc — CPU restatement of the reference column codecs. TEST INFRASTRUCTURE ONLY
// (see tskv_oracle.h). Every function cites the reference lines it follows; arithmetic is
// wrapping because the reference ships with overflow-checks = false (Cargo.toml:190-197).
#include <cstring>
#include <vector>

#include "tskv_oracle.h"

namespace {

inline uint64_t be64(const uint8_t *p) {
  uint64_t v = 0;
  for (int i = 0; i < 8; i++) v = (v << 8) | p[i];
  return v;
}
inline void put_be64(uint8_t *p, uint64_t v) {
  for (int i = 7; i >= 0; i--) {
    p[i] = (uint8_t)v;
    v >>= 8;
  }
}

// Growable byte sink that mimics Vec<u8> so the encoders can be restated line by line.
struct Bytes {
  std::vector<uint8_t> v;
  void push(uint8_t b) { v.push_back(b); }
  void extend_be64(uint64_t x) {
    uint8_t b[8];
    put_be64(b, x);
    v.insert(v.end(), b, b + 8);
  }
  size_t len() const { return v.size(); }
  void resize(size_t n) { v.resize(n, 0); }
  void truncate(size_t n) {
    if (n < v.size()) v.resize(n);
  }
};

// integer-encoding 4.0.2 `VarInt::encode_var` for u64: unsigned LEB128.
size_t encode_var(uint64_t x, uint8_t *dst) {
  size_t i = 0;
  while (x >= 0x80) {
    dst[i++] = (uint8_t)(x | 0x80);
    x >>= 7;
  }
  dst[i++] = (uint8_t)x;
  return i;
}
// integer-encoding 4.0.2 `VarInt::decode_var` for u64: returns false when the slice ends before a
// terminating byte (or after 10 continuation bytes).
bool decode_var(const uint8_t *src, size_t len, uint64_t *out) {
  uint64_t result = 0;
  unsigned shift = 0;
  bool success = false;
  for (size_t i = 0; i < len; i++) {
    uint8_t b = src[i];
    uint64_t msb_dropped = b & 0x7f;
    if (shift < 64) result |= msb_dropped << shift;
    shift += 7;
    if ((b & 0x80) == 0 || shift > 9 * 7) {
      success = (b & 0x80) == 0;
      break;
    }
  }
  *out = result;
  return success;
}

const int MAX_VAR_INT_64 = 10;  // codec/mod.rs

// simple8b.rs:7-22
const uint8_t NUM_BITS[14][2] = {{60, 1}, {30, 2}, {20, 3}, {15, 4}, {12, 5}, {10, 6}, {8, 7},
                                 {7, 8},  {6, 10}, {5, 12}, {4, 15}, {3, 20}, {2, 30}, {1, 60}};
const uint64_t S8B_MAX_VALUE = (1ull << 60) - 1;  // simple8b.rs:5

// simple8b.rs:26-76
bool s8b_encode(const uint64_t *src, size_t n, Bytes &dst) {
  size_t i = 0;
  while (i < n) {
    size_t remain = n - i;
    if (remain >= 120) {
      size_t a_len = remain >= 240 ? 240 : 120;
      size_t k = 0;
      while (k < a_len && src[i + k] == 1) k++;
      if (k == 240) {
        i += 240;
        dst.resize(dst.len() + 8);
        continue;
      } else if (k >= 120) {
        i += 120;
        dst.extend_be64(1ull << 60);
        continue;
      }
    }
    bool packed = false;
    for (int idx = 0; idx < 14 && !packed; idx++) {
      size_t int_n = NUM_BITS[idx][0], bit_n = NUM_BITS[idx][1];
      if (int_n > remain) continue;
      uint64_t max_val = 1ull << (bit_n & 0x3f);
      uint64_t val = ((uint64_t)idx + 2) << 60;
      bool fits = true;
      for (size_t k = 0; k < int_n; k++) {
        uint64_t in_v = src[i + k];
        if (in_v >= max_val) {
          fits = false;
          break;
        }
        val |= in_v << ((k * bit_n) & 0x3f);
      }
      if (!fits) continue;
      dst.extend_be64(val);
      i += int_n;
      packed = true;
    }
    if (!packed) return false;  // "value out of bounds"
  }
  return true;
}

// simple8b.rs:95-208 (decode_value) folded into :80-93 (decode)
void s8b_decode(const uint8_t *src, size_t len, std::vector<uint64_t> &dst) {
  size_t i = 0;
  while (i + 8 <= len) {
    uint64_t v = be64(src + i);
    unsigned sel = (unsigned)(v >> 60);
    if (sel == 0) {
      dst.insert(dst.end(), 240, 1);
    } else if (sel == 1) {
      dst.insert(dst.end(), 120, 1);
    } else {
      unsigned n = NUM_BITS[sel - 2][0], b = NUM_BITS[sel - 2][1];
      uint64_t mask = b == 60 ? 0x0fffffffffffffffull : ((1ull << b) - 1);
      for (unsigned k = 0; k < n; k++) dst.push_back((v >> (k * b)) & mask);
    }
    i += 8;
  }
}

inline uint64_t zz_enc(int64_t v) { return ((uint64_t)v << 1) ^ (uint64_t)(v >> 63); }       // integer.rs:102-104
inline int64_t zz_dec(uint64_t v) { return (int64_t)((v >> 1) ^ (0 - (v & 1))); }             // integer.rs:108-110

inline uint8_t log10_of_pow10(uint64_t div) {  // `((div as f64).log10()) as u8` for div = 10^k
  uint8_t k = 0;
  while (div >= 10) {
    div /= 10;
    k++;
  }
  return k;
}

// timestamp.rs:136-175
void ts_encode_rle(uint64_t v, uint64_t delta, uint64_t count, Bytes &dst) {
  dst.push(0);
  size_t n = 2;
  dst.extend_be64(v);
  n += 8;
  uint64_t div = 1000000000000ull;
  while (div > 1 && delta % div != 0) div /= 10;
  if (dst.len() <= n + MAX_VAR_INT_64) dst.resize(n + MAX_VAR_INT_64);
  if (div > 1) {
    uint8_t scaler = log10_of_pow10(div);
    dst.v[1] |= scaler;
    n += encode_var(delta / div, &dst.v[n]);
  } else {
    n += encode_var(delta, &dst.v[n]);
  }
  if (dst.len() - n <= (size_t)MAX_VAR_INT_64) dst.resize(n + MAX_VAR_INT_64);
  n += encode_var(count, &dst.v[n]);
  dst.truncate(n);
}

// timestamp.rs:51-122
bool ts_encode(const int64_t *src, size_t n, Bytes &dst) {
  if (n == 0) return true;
  dst.push(TSKV_ENC_DELTA_TS);
  uint64_t max = 0;
  std::vector<uint64_t> deltas(n);
  for (size_t i = 0; i < n; i++) deltas[i] = (uint64_t)src[i];
  if (n > 1) {
    for (size_t i = n - 1; i >= 1; i--) {
      deltas[i] = deltas[i] - deltas[i - 1];
      if (deltas[i] > max) max = deltas[i];
    }
    bool use_rle = true;
    for (size_t i = 2; i < n; i++)
      if (deltas[1] != deltas[i]) {
        use_rle = false;
        break;
      }
    if (use_rle) {
      ts_encode_rle(deltas[0], deltas[1], n, dst);
      dst.v[1] |= 2 << 4;
      return true;
    }
  }
  if (max > S8B_MAX_VALUE) {
    dst.push(0 << 4);
    for (size_t i = 0; i < n; i++) dst.extend_be64(deltas[i]);
    return true;
  }
  uint64_t div = 1000000000000ull;
  for (size_t i = 1; i < n; i++) {
    if (div <= 1) break;
    while (div > 1 && deltas[i] % div != 0) div /= 10;
  }
  if (div > 1)
    for (size_t i = 1; i < n; i++) deltas[i] /= div;
  dst.push(1 << 4);
  dst.v[1] |= log10_of_pow10(div);
  dst.extend_be64(deltas[0]);
  return s8b_encode(deltas.data() + 1, n - 1, dst);
}

// integer.rs:124-140
void i64_encode_rle(uint64_t v, uint64_t delta, uint64_t count, Bytes &dst) {
  dst.push(0);
  dst.extend_be64(v);
  size_t n = 10;
  if (dst.len() - n <= (size_t)MAX_VAR_INT_64) dst.resize(n + MAX_VAR_INT_64);
  n += encode_var(delta, &dst.v[n]);
  if (dst.len() - n <= (size_t)MAX_VAR_INT_64) dst.resize(n + MAX_VAR_INT_64);
  n += encode_var(count, &dst.v[n]);
  dst.truncate(n);
}

// integer.rs:40-96
bool i64_encode(const int64_t *src, size_t n, Bytes &dst) {
  if (n == 0) return true;
  dst.push(TSKV_ENC_DELTA);
  uint64_t max = 0;
  std::vector<uint64_t> deltas(n);
  for (size_t i = 0; i < n; i++) deltas[i] = (uint64_t)src[i];
  for (size_t i = n - 1; i >= 1; i--) {
    deltas[i] = zz_enc((int64_t)(deltas[i] - deltas[i - 1]));
    if (deltas[i] > max) max = deltas[i];
  }
  deltas[0] = zz_enc(src[0]);
  if (n > 2) {
    bool use_rle = true;
    for (size_t i = 2; i < n; i++)
      if (deltas[1] != deltas[i]) {
        use_rle = false;
        break;
      }
    if (use_rle) {
      i64_encode_rle(deltas[0], deltas[1], n - 1, dst);
      dst.v[1] |= 2 << 4;
      return true;
    }
  }
  if (max > S8B_MAX_VALUE) {
    dst.push(0 << 4);
    for (size_t i = 0; i < n; i++) dst.extend_be64(deltas[i]);
    return true;
  }
  dst.push(1 << 4);
  dst.extend_be64(deltas[0]);
  return s8b_encode(deltas.data() + 1, n - 1, dst);
}

const uint64_t SENTINEL = 0x7ff80000000000ffull;  // float.rs:16

// float.rs:32-243. `dst.v[0]` is the Encoding id; bit index n counts from dst.v[1].
bool f64_encode(const double *src, size_t len, Bytes &dst) {
  if (len == 0) return true;
  dst.push(TSKV_ENC_GORILLA);
  size_t n = 8;
  dst.push(1 << 4);
  uint64_t prev;
  memcpy(&prev, &src[0], 8);
  dst.extend_be64(prev);
  n += 64;
  uint64_t prev_leading = ~0ull, prev_trailing = 0;
  auto &d = dst.v;
  for (size_t i = 1; i <= len; i++) {
    uint64_t cur;
    if (i < len) {
      memcpy(&cur, &src[i], 8);
      if (cur == SENTINEL) return false;  // "unsupported value"
    } else {
      cur = SENTINEL;
    }
    uint64_t v_delta = cur ^ prev;
    if (v_delta == 0) {
      n += 1;
      prev = cur;
      continue;
    }
    while ((n >> 3) + 1 >= d.size()) d.push_back(0);
    d[(n >> 3) + 1] |= (uint8_t)(128 >> (n & 7));
    n += 1;
    uint64_t leading = (uint64_t)__builtin_clzll(v_delta);
    uint64_t trailing = (uint64_t)__builtin_ctzll(v_delta);
    leading &= 0x1f;
    if (((n + 2) >> 3) + 1 >= d.size()) d.push_back(0);
    if (prev_leading != ~0ull && leading >= prev_leading && trailing >= prev_trailing) {
      n += 1;
      uint64_t l = 64 - prev_leading - prev_trailing;
      while (((n + 1) >> 3) + 1 >= d.size()) d.push_back(0);
      uint64_t v = (v_delta >> prev_trailing) << (64 - l);
      uint64_t m = n & 7;
      uint64_t written = 0;
      if (m > 0) {
        written = l < 8 - m ? l : 8 - m;
        uint64_t mask = v >> 56;
        d[(n >> 3) + 1] |= (uint8_t)(mask >> m);
        n += written;
        if (l - written == 0) {
          prev = cur;
          continue;
        }
      }
      uint64_t vv = v << written;
      while (((n >> 3) + 8) + 1 >= d.size()) d.push_back(0);
      size_t k = (n >> 3) + 1;
      put_be64(&d[k], vv);
      n += l - written;
    } else {
      prev_leading = leading;
      prev_trailing = trailing;
      d[(n >> 3) + 1] |= (uint8_t)(128 >> (n & 7));
      n += 1;
      if (((n + 5) >> 3) + 1 >= d.size()) d.push_back(0);
      size_t m = n & 7;
      size_t l = 5;
      uint64_t v = leading << 59;
      uint64_t mask = v >> 56;
      if (m <= 3) {
        d[(n >> 3) + 1] |= (uint8_t)(mask >> m);
        n += l;
      } else {
        size_t written = 8 - m;
        d[(n >> 3) + 1] |= (uint8_t)(mask >> m);
        n += written;
        mask = v << written;
        mask >>= 56;
        m = n & 7;
        d[(n >> 3) + 1] |= (uint8_t)(mask >> m);
        n += l - written;
      }
      uint64_t sig_bits = 64 - leading - trailing;
      if (((n + 6) >> 3) + 1 >= d.size()) d.push_back(0);
      m = n & 7;
      l = 6;
      v = sig_bits << 58;
      mask = v >> 56;
      if (m <= 2) {
        d[(n >> 3) + 1] |= (uint8_t)(mask >> m);
        n += l;
      } else {
        size_t written = 8 - m;
        d[(n >> 3) + 1] |= (uint8_t)(mask >> m);
        n += written;
        mask = v << written;
        mask >>= 56;
        m = n & 7;
        d[(n >> 3) + 1] |= (uint8_t)(mask >> m);
        n += l - written;
      }
      m = n & 7;
      l = (size_t)sig_bits;
      v = (v_delta >> trailing) << (64 - l);
      while (((n + l) >> 3) + 1 >= d.size()) d.push_back(0);
      size_t written = 0;
      if (m > 0) {
        written = l < 8 - m ? l : 8 - m;
        mask = v >> 56;
        d[(n >> 3) + 1] |= (uint8_t)(mask >> m);
        n += written;
        if (l - written == 0) {
          prev = cur;
          continue;
        }
      }
      uint64_t vv = v << written;
      while (((n >> 3) + 8) + 1 >= d.size()) d.push_back(0);
      size_t k = (n >> 3) + 1;
      put_be64(&d[k], vv);
      n += l - written;
    }
    prev = cur;
  }
  size_t length = (n >> 3) + 1;
  if ((n & 7) > 0) length += 1;
  dst.truncate(length);
  return true;
}

inline uint64_t rotl(uint64_t v, unsigned r) {
  r &= 63;
  return r ? (v << r) | (v >> (64 - r)) : v;
}
inline uint64_t rotr(uint64_t v, unsigned r) {
  r &= 63;
  return r ? (v >> r) | (v << (64 - r)) : v;
}
inline uint64_t bit_mask(unsigned idx) {  // float.rs:283-348 BIT_MASK[idx & 0x3f]
  idx &= 0x3f;
  return idx == 0 ? ~0ull : ((1ull << idx) - 1);
}

// float.rs:418-606 decode_with_sentinel; `src` starts at the 0x10 byte (float.rs:357).
// Deviation (documented in DESIGN.md): where the reference's u8 counters would wrap on a truncated
// stream and keep decoding garbage, this returns TSKV_ERR_SHORT_BLOCK.
tskv_status gorilla_decode(const uint8_t *src, size_t len, std::vector<uint64_t> &dst) {
  if (len == 0) return TSKV_OK;  // float.rs:423-425: empty array (length mismatch later)
  if (len < 9) return TSKV_ERR_SHORT_BLOCK;  // src[i..i+8] would panic
  size_t i = 1;
  uint64_t val = be64(src + i);
  i += 8;
  dst.push_back(val);
  uint64_t br_cached_val = 0;
  unsigned br_valid_bits = 0;
  auto refill = [&](void) -> bool {
    size_t remaining = len - i;
    if (remaining >= 8) {
      br_cached_val = be64(src + i);
      br_valid_bits = 64;
      i += 8;
      return true;
    } else if (remaining > 0) {
      uint64_t c = 0;
      unsigned vb = (unsigned)(remaining * 8);
      for (size_t k = i; k < len; k++) c = (c << 8) | src[k];
      br_cached_val = rotr(c, vb);
      br_valid_bits = vb;
      i = len;
      return true;
    }
    return false;  // "unexpected end of block"
  };
  if (!refill()) return TSKV_ERR_SHORT_BLOCK;
  unsigned trailing_n = 0, meaningful_n = 64;
  for (;;) {
    if (br_valid_bits == 0 && !refill()) return TSKV_ERR_SHORT_BLOCK;
    br_valid_bits -= 1;
    br_cached_val = rotl(br_cached_val, 1);
    if ((br_cached_val & 1) == 0) {
      dst.push_back(val);
      continue;
    }
    if (br_valid_bits == 0 && !refill()) return TSKV_ERR_SHORT_BLOCK;
    br_valid_bits -= 1;
    br_cached_val = rotl(br_cached_val, 1);
    if ((br_cached_val & 1) > 0) {
      const unsigned ltbc = 11;
      uint64_t lm_bits = 0;
      if (br_valid_bits >= ltbc) {
        br_valid_bits -= ltbc;
        br_cached_val = rotl(br_cached_val, ltbc);
        lm_bits = br_cached_val;
      } else {
        unsigned bits_01 = 11;
        if (br_valid_bits > 0) {
          bits_01 -= br_valid_bits;
          lm_bits = rotl(br_cached_val, 11);
        }
        if (!refill()) return TSKV_ERR_SHORT_BLOCK;
        if (br_valid_bits < bits_01) return TSKV_ERR_SHORT_BLOCK;  // reference: u8 wrap
        br_cached_val = rotl(br_cached_val, bits_01);
        br_valid_bits -= bits_01;
        lm_bits &= ~bit_mask(bits_01);
        lm_bits |= br_cached_val & bit_mask(bits_01);
      }
      lm_bits &= 0x7ff;
      unsigned leading_n = (unsigned)(lm_bits >> 6) & 0x1f;
      meaningful_n = (unsigned)(lm_bits & 0x3f);
      if (meaningful_n > 0) {
        trailing_n = (uint8_t)(64 - leading_n - meaningful_n);  // u8 arithmetic
      } else {
        trailing_n = 0;
        meaningful_n = 64;
      }
    }
    uint64_t s_bits = 0;
    if (br_valid_bits >= meaningful_n) {
      br_valid_bits -= meaningful_n;
      br_cached_val = rotl(br_cached_val, meaningful_n);
      s_bits = br_cached_val;
    } else {
      unsigned m_bits = meaningful_n;
      if (br_valid_bits > 0) {
        m_bits -= br_valid_bits;
        s_bits = rotl(br_cached_val, meaningful_n);
      }
      if (!refill()) return TSKV_ERR_SHORT_BLOCK;
      if (br_valid_bits < m_bits) return TSKV_ERR_SHORT_BLOCK;  // reference: wrapping_sub
      br_cached_val = rotl(br_cached_val, m_bits);
      br_valid_bits -= m_bits;
      s_bits &= ~bit_mask(m_bits);
      s_bits |= br_cached_val & bit_mask(m_bits);
    }
    s_bits &= bit_mask(meaningful_n);
    val ^= s_bits << (trailing_n & 0x3f);
    if (val == SENTINEL) break;
    dst.push_back(val);
  }
  return TSKV_OK;
}

inline bool bit_is_set(const uint8_t *bitset, uint64_t i) { return (bitset[i >> 3] >> (i & 7)) & 1; }

// Feed decoded values to the valid rows (the `for is_valid in bit_set.iter()` tails of every codec).
tskv_status scatter_valid(const std::vector<uint64_t> &vals, const uint8_t *bitset, uint64_t n_rows,
                          uint64_t *out_vals, uint8_t *out_valid) {
  size_t k = 0;
  for (uint64_t r = 0; r < n_rows; r++) {
    if (bit_is_set(bitset, r)) {
      if (k >= vals.size()) return TSKV_ERR_BITSET_MISMATCH;
      out_vals[r] = vals[k++];
      out_valid[r] = 1;
    } else {
      out_vals[r] = 0;
      out_valid[r] = 0;
    }
  }
  return TSKV_OK;
}

// timestamp.rs:177-299 (`src` = data after the Encoding id byte)
tskv_status ts_delta_decode(const uint8_t *src, size_t len, const uint8_t *bitset, uint64_t n_rows,
                            uint64_t *out_vals, uint8_t *out_valid) {
  if (len < 1) return TSKV_ERR_SHORT_BLOCK;  // src[0] would panic
  unsigned encoding = src[0] >> 4;
  std::vector<uint64_t> vals;
  if (encoding == 0) {  // :201-224, called with &src[1..]
    const uint8_t *p = src + 1;
    size_t l = len - 1;
    if (l == 0 || (l & 7) != 0) return TSKV_ERR_BAD_LENGTH;
    uint64_t prev = 0;
    for (size_t i = 0; i < l; i += 8) {
      prev += be64(p + i);
      vals.push_back(prev);
    }
    return scatter_valid(vals, bitset, n_rows, out_vals, out_valid);
  } else if (encoding == 2) {  // :226-259
    if (len < 9) return TSKV_ERR_SHORT_BLOCK;
    uint64_t scaler = 1;
    for (unsigned k = 0; k < (unsigned)(src[0] & 0xf); k++) scaler *= 10;
    uint64_t first = be64(src + 1);
    uint64_t delta;
    if (!decode_var(src + 9, len - 9, &delta)) return TSKV_ERR_SHORT_BLOCK;  // "unable to decode delta"
    delta *= scaler;
    bool is_first = true;
    uint64_t cur = 0;
    for (uint64_t r = 0; r < n_rows; r++) {
      if (!bit_is_set(bitset, r)) {
        out_vals[r] = 0;
        out_valid[r] = 0;
        continue;
      }
      if (is_first) {
        cur = first;
        is_first = false;
      } else {
        cur += delta;
      }
      out_vals[r] = cur;
      out_valid[r] = 1;
    }
    return TSKV_OK;
  } else if (encoding == 1) {  // :261-299
    if (len < 9) return TSKV_ERR_SHORT_BLOCK;
    uint64_t scaler = 1;
    for (unsigned k = 0; k < (unsigned)(src[0] & 0xf); k++) scaler *= 10;
    uint64_t next = be64(src + 1);
    if (((len - 9) & 7) != 0) return TSKV_ERR_SHORT_BLOCK;  // src[i..i+8] would panic
    s8b_decode(src + 9, len - 9, vals);
    size_t k = 0;
    if (n_rows == 0) return TSKV_OK;
    // Row 0: value if valid, else null — the first value is NOT re-offered to a later row (:273-279).
    if (bit_is_set(bitset, 0)) {
      out_vals[0] = next;
      out_valid[0] = 1;
    } else {
      out_vals[0] = 0;
      out_valid[0] = 0;
    }
    for (uint64_t r = 1; r < n_rows; r++) {
      if (bit_is_set(bitset, r)) {
        if (k >= vals.size()) return TSKV_ERR_BITSET_MISMATCH;  // reference: silently shorter array
        next += scaler > 1 ? vals[k] * scaler : vals[k];
        k++;
        out_vals[r] = next;
        out_valid[r] = 1;
      } else {
        out_vals[r] = 0;
        out_valid[r] = 0;
      }
    }
    return TSKV_OK;
  }
  return TSKV_ERR_BAD_ENCODING;
}

// integer.rs:142-248 (`src` = data after the Encoding id byte)
tskv_status i64_delta_decode(const uint8_t *src, size_t len, const uint8_t *bitset, uint64_t n_rows,
                             uint64_t *out_vals, uint8_t *out_valid) {
  if (len < 1) return TSKV_ERR_SHORT_BLOCK;
  unsigned encoding = src[0] >> 4;
  const uint8_t *p = src + 1;
  size_t l = len - 1;
  std::vector<uint64_t> vals;
  if (encoding == 0) {  // :165-184
    if (l == 0 || (l & 7) != 0) return TSKV_ERR_BAD_LENGTH;
    uint64_t prev = 0;
    for (size_t i = 0; i < l; i += 8) {
      prev += (uint64_t)zz_dec(be64(p + i));
      vals.push_back(prev);
    }
    return scatter_valid(vals, bitset, n_rows, out_vals, out_valid);  // OOB slice panic => mismatch
  } else if (encoding == 2) {  // :186-214
    if (l < 8) return TSKV_ERR_SHORT_BLOCK;
    uint64_t delta;
    if (!decode_var(p + 8, l - 8, &delta)) return TSKV_ERR_SHORT_BLOCK;
    uint64_t first = (uint64_t)zz_dec(be64(p));
    uint64_t delta_z = (uint64_t)zz_dec(delta);
    bool is_first = true;
    for (uint64_t r = 0; r < n_rows; r++) {
      if (!bit_is_set(bitset, r)) {
        out_vals[r] = 0;
        out_valid[r] = 0;
        continue;
      }
      if (is_first)
        is_first = false;
      else
        first += delta_z;
      out_vals[r] = first;
      out_valid[r] = 1;
    }
    return TSKV_OK;
  } else if (encoding == 1) {  // :216-248
    if (l < 8) return TSKV_ERR_SHORT_BLOCK;
    if (((l - 8) & 7) != 0) return TSKV_ERR_SHORT_BLOCK;
    uint64_t next = (uint64_t)zz_dec(be64(p));
    vals.push_back(next);
    std::vector<uint64_t> packed;
    s8b_decode(p + 8, l - 8, packed);
    for (uint64_t u : packed) {
      next += (uint64_t)zz_dec(u);
      vals.push_back(next);
    }
    return scatter_valid(vals, bitset, n_rows, out_vals, out_valid);
  }
  return TSKV_ERR_BAD_ENCODING;
}

// timestamp.rs:301-323 / float.rs:387-413 (`src` = data after the Encoding id byte)
tskv_status raw_decode(const uint8_t *src, size_t len, const uint8_t *bitset, uint64_t n_rows,
                       uint64_t *out_vals, uint8_t *out_valid) {
  if ((len & 7) != 0) return TSKV_ERR_BAD_LENGTH;  // decode_be_i64 on a short chunk is UB
  std::vector<uint64_t> vals;
  for (size_t i = 0; i < len; i += 8) vals.push_back(be64(src + i));
  return scatter_valid(vals, bitset, n_rows, out_vals, out_valid);
}

// bool_bitpack_decode (boolean.rs:79-111), `src` = data after the Encoding id byte
tskv_status bool_bitpack_decode(const uint8_t *src, size_t len, const uint8_t *bitset, uint64_t n_rows,
                                uint64_t *out_vals, uint8_t *out_valid) {
  if (len < 1) return TSKV_ERR_SHORT_BLOCK;          // src[0] on an empty slice panics
  if (src[0] != (1u << 4)) return TSKV_ERR_BAD_ENCODING;  // assert_eq!(src[0], BOOLEAN_COMPRESSED_BIT_PACKED << 4)
  src += 1;
  len -= 1;
  uint64_t count = 0;
  if (!decode_var(src, len, &count)) return TSKV_ERR_SHORT_BLOCK;  // "boolean decoder: invalid count"
  size_t nread = 0;
  while (src[nread] & 0x80) nread++;
  nread++;
  src += nread;
  len -= nread;
  uint64_t bit_index = 0;
  for (uint64_t r = 0; r < n_rows; r++) {
    if (bit_is_set(bitset, r)) {
      if (bit_index >= count) return TSKV_ERR_BITSET_MISMATCH;   // "Insufficient data for decoding"
      if (bit_index / 8 >= len) return TSKV_ERR_BITSET_MISMATCH;  // src[bit_index / 8] out of bounds: panic in the reference
      out_vals[r] = (src[bit_index / 8] >> (7 - (bit_index % 8))) & 1;
      out_valid[r] = 1;
      bit_index++;
    } else {
      out_vals[r] = 0;
      out_valid[r] = 0;
    }
  }
  return TSKV_OK;
}
// bool_without_compress_decode (boolean.rs:112-140), `src` = data after the Encoding id byte
tskv_status bool_raw_decode(const uint8_t *src, size_t len, const uint8_t *bitset, uint64_t n_rows, uint64_t *out_vals,
                            uint8_t *out_valid) {
  size_t k = 0;
  for (uint64_t r = 0; r < n_rows; r++) {
    if (bit_is_set(bitset, r)) {
      // `if let Some(v) = iter.next()`: a valid row past the data appends NOTHING - the array comes out shorter than
      // the bitset and the RecordBatch build fails; reported as the bitset / data mismatch it is
      if (k >= len) return TSKV_ERR_BITSET_MISMATCH;
      out_vals[r] = src[k++] == 1 ? 1 : 0;
      out_valid[r] = 1;
    } else {
      out_vals[r] = 0;
      out_valid[r] = 0;
    }
  }
  return TSKV_OK;
}

}  // namespace

extern "C" {

static int64_t emit_bytes(const Bytes &b, uint8_t *dst, uint64_t cap) {
  if (b.v.size() > cap) return -TSKV_ERR_OOM;
  if (!b.v.empty()) memcpy(dst, b.v.data(), b.v.size());
  return (int64_t)b.v.size();
}
// bool_bitpack_encode (boolean.rs:24-64): id | 0x10 | varint n | 1 bit per value, MSB first
int64_t orc_bool_encode(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap) {
  Bytes b;
  if (n) {
    b.push(TSKV_ENC_BITPACK);
    const size_t size = 1 + 8 + (n + 7) / 8;  // header + count + data
    b.v.resize(size + 1, 0);
    b.v[1] = 1u << 4;
    uint64_t nbit = 8;
    nbit += 8 * encode_var(n, &b.v[2]);
    for (uint64_t i = 0; i < n; i++) {
      const size_t index = (size_t)(nbit >> 3);
      if (src[i]) b.v[index + 1] |= (uint8_t)(128 >> (nbit & 7));
      else b.v[index + 1] &= (uint8_t)~(128 >> (nbit & 7));
      nbit++;
    }
    uint64_t length = nbit >> 3;
    if (nbit & 7) length++;
    b.v.resize(length + 1);
  }
  return emit_bytes(b, dst, cap);
}
// bool_without_compress_encode (boolean.rs:66-76)
int64_t orc_bool_raw_encode(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap) {
  Bytes b;
  b.push(TSKV_ENC_NULL);
  for (uint64_t i = 0; i < n; i++) b.push(src[i] ? 1 : 0);
  return emit_bytes(b, dst, cap);
}

uint64_t orc_zigzag_encode(int64_t v) { return zz_enc(v); }
int64_t orc_zigzag_decode(uint64_t v) { return zz_dec(v); }

static int64_t emit(const Bytes &b, uint8_t *dst, uint64_t cap) {
  if (b.v.size() > cap) return -TSKV_ERR_OOM;
  if (!b.v.empty()) memcpy(dst, b.v.data(), b.v.size());
  return (int64_t)b.v.size();
}

int64_t orc_simple8b_encode(const uint64_t *src, uint64_t n, uint8_t *dst, uint64_t cap) {
  Bytes b;
  if (!s8b_encode(src, n, b)) return -TSKV_ERR_INVALID_ARG;
  return emit(b, dst, cap);
}
int64_t orc_simple8b_decode(const uint8_t *src, uint64_t len, uint64_t *dst, uint64_t cap) {
  std::vector<uint64_t> v;
  s8b_decode(src, len, v);
  if (v.size() > cap) return -TSKV_ERR_OOM;
  if (!v.empty()) memcpy(dst, v.data(), v.size() * 8);
  return (int64_t)v.size();
}
int64_t orc_ts_encode(const int64_t *src, uint64_t n, uint8_t *dst, uint64_t cap) {
  Bytes b;
  if (!ts_encode(src, n, b)) return -TSKV_ERR_INVALID_ARG;
  return emit(b, dst, cap);
}
int64_t orc_i64_encode(const int64_t *src, uint64_t n, uint8_t *dst, uint64_t cap) {
  Bytes b;
  if (!i64_encode(src, n, b)) return -TSKV_ERR_INVALID_ARG;
  return emit(b, dst, cap);
}
int64_t orc_f64_encode(const double *src, uint64_t n, uint8_t *dst, uint64_t cap) {
  Bytes b;
  if (!f64_encode(src, n, b)) return -TSKV_ERR_UNSUPPORTED;
  return emit(b, dst, cap);
}
// timestamp.rs:21-31 / float.rs:256-267
int64_t orc_raw_encode(const uint64_t *src, uint64_t n, uint8_t *dst, uint64_t cap) {
  Bytes b;
  if (n) {
    b.push(TSKV_ENC_NULL);
    for (uint64_t i = 0; i < n; i++) b.extend_be64(src[i]);
  }
  return emit(b, dst, cap);
}

// tsm/reader.rs:658-731 (type dispatch) + codec/instance.rs:358-401 (encoding dispatch)
tskv_status orc_decode_column(uint32_t phys_type, const uint8_t *data, uint64_t data_len,
                              const uint8_t *bitset, uint64_t n_rows, uint64_t *out_vals,
                              uint8_t *out_valid) {
  if (data_len == 0) {  // every codec: `if src.is_empty()` => all-null array of bit_set.len()
    for (uint64_t r = 0; r < n_rows; r++) {
      out_vals[r] = 0;
      out_valid[r] = 0;
    }
    return TSKV_OK;
  }
  unsigned enc = data[0];
  if (enc > 11) enc = 15;  // Encoding::Unknown (codec.rs:119-137)
  const uint8_t *src = data + 1;
  size_t len = (size_t)data_len - 1;
  if (phys_type == TSKV_PT_BOOL) {  // get_bool_codec (instance.rs:415-421): Null => bytes, everything else => bit-pack
    if (enc == TSKV_ENC_NULL) return bool_raw_decode(src, len, bitset, n_rows, out_vals, out_valid);
    return bool_bitpack_decode(src, len, bitset, n_rows, out_vals, out_valid);
  }
  if (enc == TSKV_ENC_QUANTILE) return TSKV_ERR_UNSUPPORTED;  // pco: out of scope
  if (enc == TSKV_ENC_NULL) return raw_decode(src, len, bitset, n_rows, out_vals, out_valid);
  switch (phys_type) {
    case TSKV_PT_TIME:  // get_ts_codec: Delta => integer codec, everything else => DeltaTs
      if (enc == TSKV_ENC_DELTA) return i64_delta_decode(src, len, bitset, n_rows, out_vals, out_valid);
      return ts_delta_decode(src, len, bitset, n_rows, out_vals, out_valid);
    case TSKV_PT_I64:  // get_i64_codec: DeltaTs => ts codec, everything else => Delta
      if (enc == TSKV_ENC_DELTA_TS) return ts_delta_decode(src, len, bitset, n_rows, out_vals, out_valid);
      return i64_delta_decode(src, len, bitset, n_rows, out_vals, out_valid);
    case TSKV_PT_U64:  // get_u64_codec: everything else => Delta (unsigned.rs:30-45)
      return i64_delta_decode(src, len, bitset, n_rows, out_vals, out_valid);
    case TSKV_PT_F64: {  // get_f64_codec: everything else => Gorilla
      std::vector<uint64_t> vals;
      tskv_status st = gorilla_decode(src, len, vals);
      if (st != TSKV_OK) return st;
      return scatter_valid(vals, bitset, n_rows, out_vals, out_valid);
    }
    default:
      return TSKV_ERR_UNSUPPORTED;
  }
}

}  // extern "C"
