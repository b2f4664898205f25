// tsm_writer.cc — see tsm_writer.h.
#include "tsm_writer.h"

#include <cstring>

#include "../host_util.h"

namespace tskv {
namespace {

inline void put_u64be(Bytes &out, uint64_t v) {
  for (int s = 56; s >= 0; s -= 8) out.push_back((uint8_t)(v >> s));
}
inline void put_varint(Bytes &out, uint64_t v) {
  while (v >= 0x80) {
    out.push_back((uint8_t)(v | 0x80));
    v >>= 7;
  }
  out.push_back((uint8_t)v);
}
inline unsigned bit_length(uint64_t v) { return v ? 64u - (unsigned)__builtin_clzll(v) : 0u; }
inline uint64_t zigzag(int64_t v) { return ((uint64_t)v << 1) ^ (uint64_t)(v >> 63); }

// (values per word, bits per value) for selectors 2..15
const unsigned kPackN[14] = {60, 30, 20, 15, 12, 10, 8, 7, 6, 5, 4, 3, 2, 1};
const unsigned kPackB[14] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 15, 20, 30, 60};
const uint64_t kMaxPacked = (1ull << 60) - 1;

// Largest 10^k (k <= 12) dividing every value of `d[0..n)`; 10^12 for an empty set.
uint64_t common_pow10(const uint64_t *d, size_t n, unsigned *exp) {
  uint64_t div = 1000000000000ull;
  unsigned k = 12;
  for (size_t i = 0; i < n && div > 1; i++)
    while (div > 1 && d[i] % div != 0) {
      div /= 10;
      k--;
    }
  *exp = k;
  return div;
}

// MSB-first bit sink.
struct BitSink {
  Bytes &out;
  uint64_t acc = 0;
  unsigned fill = 0;  // bits in acc (from the top)
  explicit BitSink(Bytes &o) : out(o) {}
  void put(uint64_t value, unsigned nbits) {  // low `nbits` of value, nbits in [1,64]
    while (nbits) {
      unsigned room = 64 - fill;
      unsigned take = nbits < room ? nbits : room;
      uint64_t chunk = take == 64 ? value : ((value >> (nbits - take)) & ((1ull << take) - 1));
      acc |= take == 64 ? chunk : chunk << (room - take);
      fill += take;
      nbits -= take;
      if (fill == 64) {
        put_u64be(out, acc);
        acc = 0;
        fill = 0;
      }
    }
  }
  void finish() {  // zero-padded to a byte boundary
    for (unsigned done = 0; done < fill; done += 8) out.push_back((uint8_t)(acc >> (56 - done)));
    acc = 0;
    fill = 0;
  }
};

}  // namespace

bool simple8b_pack(const uint64_t *src, size_t n, Bytes &out) {
  size_t i = 0;
  while (i < n) {
    const size_t remain = n - i;
    if (remain >= 120) {
      const size_t span = remain >= 240 ? 240 : 120;
      size_t ones = 0;
      while (ones < span && src[i + ones] == 1) ones++;
      if (ones == 240) {
        put_u64be(out, 0);
        i += 240;
        continue;
      }
      if (ones >= 120) {
        put_u64be(out, 1ull << 60);
        i += 120;
        continue;
      }
    }
    bool done = false;
    for (unsigned s = 0; s < 14 && !done; s++) {
      const unsigned cnt = kPackN[s], width = kPackB[s];
      if (cnt > remain) continue;
      uint64_t word = (uint64_t)(s + 2) << 60;
      bool fits = true;
      for (unsigned k = 0; k < cnt; k++) {
        const uint64_t v = src[i + k];
        if (bit_length(v) > width) {
          fits = false;
          break;
        }
        word |= v << (k * width);
      }
      if (!fits) continue;
      put_u64be(out, word);
      i += cnt;
      done = true;
    }
    if (!done) return false;
  }
  return true;
}

bool encode_timestamps(const int64_t *src, size_t n, Bytes &out) {
  if (n == 0) return true;
  out.push_back(11);
  std::vector<uint64_t> d(n);
  d[0] = (uint64_t)src[0];
  uint64_t dmax = 0;
  bool all_equal = n >= 2;
  for (size_t i = 1; i < n; i++) {
    d[i] = (uint64_t)src[i] - (uint64_t)src[i - 1];
    if (d[i] > dmax) dmax = d[i];
    if (d[i] != d[1]) all_equal = false;
  }
  if (all_equal) {  // run-length: first | varint(delta / 10^k) | varint(count)
    unsigned k;
    uint64_t div = common_pow10(&d[1], 1, &k);
    out.push_back((uint8_t)((2u << 4) | k));
    put_u64be(out, d[0]);
    put_varint(out, div > 1 ? d[1] / div : d[1]);
    put_varint(out, (uint64_t)n);
    return true;
  }
  if (dmax > kMaxPacked) {  // raw deltas
    out.push_back(0);
    for (size_t i = 0; i < n; i++) put_u64be(out, d[i]);
    return true;
  }
  unsigned k;
  uint64_t div = common_pow10(d.data() + 1, n - 1, &k);
  if (div > 1)
    for (size_t i = 1; i < n; i++) d[i] /= div;
  out.push_back((uint8_t)((1u << 4) | k));
  put_u64be(out, d[0]);
  return simple8b_pack(d.data() + 1, n - 1, out);
}

bool encode_integers(const int64_t *src, size_t n, Bytes &out) {
  if (n == 0) return true;
  out.push_back(2);
  std::vector<uint64_t> z(n);
  z[0] = zigzag(src[0]);
  uint64_t zmax = 0;
  bool all_equal = n > 2;
  for (size_t i = 1; i < n; i++) {
    z[i] = zigzag((int64_t)((uint64_t)src[i] - (uint64_t)src[i - 1]));
    if (z[i] > zmax) zmax = z[i];
    if (z[i] != z[1]) all_equal = false;
  }
  if (all_equal) {
    out.push_back(2u << 4);
    put_u64be(out, z[0]);
    put_varint(out, z[1]);
    put_varint(out, (uint64_t)n - 1);
    return true;
  }
  if (zmax > kMaxPacked) {
    out.push_back(0);
    for (size_t i = 0; i < n; i++) put_u64be(out, z[i]);
    return true;
  }
  out.push_back(1u << 4);
  put_u64be(out, z[0]);
  return simple8b_pack(z.data() + 1, n - 1, out);
}

bool encode_floats(const double *src, size_t n, Bytes &out) {
  if (n == 0) return true;
  const uint64_t kSentinel = 0x7ff80000000000ffull;
  out.push_back(6);
  BitSink bits(out);
  bits.put(0x10, 8);
  uint64_t prev;
  memcpy(&prev, &src[0], 8);
  bits.put(prev, 64);
  bool have_window = false;
  unsigned win_lead = 0, win_trail = 0;
  for (size_t i = 1; i <= n; i++) {
    uint64_t cur;
    if (i < n) {
      memcpy(&cur, &src[i], 8);
      if (cur == kSentinel) return false;
    } else {
      cur = kSentinel;
    }
    const uint64_t x = cur ^ prev;
    prev = cur;
    if (x == 0) {
      bits.put(0, 1);
      continue;
    }
    const unsigned lead = (unsigned)__builtin_clzll(x) & 31u;  // masked, not clamped
    const unsigned trail = (unsigned)__builtin_ctzll(x);
    if (have_window && lead >= win_lead && trail >= win_trail) {
      bits.put(0b10, 2);
      bits.put(x >> win_trail, 64 - win_lead - win_trail);
    } else {
      have_window = true;
      win_lead = lead;
      win_trail = trail;
      const unsigned sig = 64 - lead - trail;
      bits.put(0b11, 2);
      bits.put(lead, 5);
      bits.put(sig & 63u, 6);  // 64 is stored as 0
      bits.put(x >> trail, sig);
    }
  }
  bits.finish();
  return true;
}

void encode_raw(const uint64_t *src, size_t n, Bytes &out) {
  if (n == 0) return;
  out.push_back(1);
  for (size_t i = 0; i < n; i++) put_u64be(out, src[i]);
}

void append_page(const uint8_t *validity, uint64_t rows, const Bytes &data, Bytes &arena) {
  const uint32_t bitset_len = (uint32_t)((rows + 7) / 8);
  const uint32_t crc = crc32_ieee(data.data(), data.size());
  for (int s = 24; s >= 0; s -= 8) arena.push_back((uint8_t)(bitset_len >> s));
  put_u64be(arena, rows);
  for (int s = 24; s >= 0; s -= 8) arena.push_back((uint8_t)(crc >> s));
  if (validity) {
    arena.insert(arena.end(), validity, validity + bitset_len);
  } else {
    arena.insert(arena.end(), bitset_len, 0xff);
    // Arrow's BooleanBuffer::new_set leaves the bits past `rows` cleared
    if (rows & 7) arena.back() = (uint8_t)((1u << (rows & 7)) - 1);
  }
  arena.insert(arena.end(), data.begin(), data.end());
}

}  // namespace tskv
