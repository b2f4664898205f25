// cursors.cuh — per-lane streaming decoders for TSM column pages (sm_100a).
//
// Design: one lane owns one page and walks it value by value, so decode -> filter -> bucket
// reduce happens in registers and no decoded value is ever written to HBM. The formats are those of
// the reference codecs (all paths relative to the reference tree):
//   page framing     tskv/src/tsm/page.rs:31-94
//   simple8b         tskv/src/tsm/codec/simple8b.rs:80-208
//   timestamp delta  tskv/src/tsm/codec/timestamp.rs:177-299   (deltas NOT zig-zagged, 10^k scaler)
//   integer delta    tskv/src/tsm/codec/integer.rs:142-248     (zig-zag deltas)
//   gorilla          tskv/src/tsm/codec/float.rs:418-606
//   raw ("Null")     tskv/src/tsm/codec/timestamp.rs:301-323, float.rs:387-413
// Arithmetic wraps like the reference's release build (Cargo.toml:190-197).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/tskv_gpu.h"
#include "kinds.h"

namespace tskv {

__device__ __forceinline__ uint64_t bswap64(uint64_t v) {
  uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}
__device__ __forceinline__ int64_t zigzag_dec(uint64_t v) {
  return (int64_t)((v >> 1) ^ (0 - (v & 1)));
}

// Streams an arbitrarily aligned byte range as big-endian u64 words. Two implementations with the
// same interface:
//   BeStream      aligned 8-byte global loads (one per word; previous word kept and funnel-shifted).
//   RingStream    the same words, prefetched RING_WORDS ahead into a per-lane shared-memory ring with
//                 cp.async (LDGSTS): a lane-serial decoder otherwise exposes a full HBM/L2 round trip on
//                 every word it consumes.
// Both may read up to RING_WORDS*8+8 bytes past the end of the page: the arena carries 128 bytes of slack.
struct BeStream {
  const uint64_t *ap;  // next aligned word
  uint64_t cur;        // last aligned word, little-endian
  uint32_t sh;         // misalignment in bits
  __device__ __forceinline__ void init(const uint8_t *p, uint32_t /*smem_slot*/ = 0) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    sh = (uint32_t)(a & 7) * 8;
    ap = reinterpret_cast<const uint64_t *>(a & ~(uintptr_t)7);
    cur = __ldg(ap++);
  }
  __device__ __forceinline__ uint64_t next() {
    uint64_t nxt = __ldg(ap++);
    uint64_t raw = (cur >> sh) | ((nxt << 1) << (63 - sh));
    cur = nxt;
    return bswap64(raw);
  }
};

constexpr int RING_WORDS = 8;  // 64 bytes of lookahead per stream per lane

// Ring layout: word k of lane l lives at slot_base + k * 256 + l * 8 (shared space): whatever ring
// position each lane is at, lane l always hits banks 2l, 2l+1 => conflict-free 64-bit LDS.
struct RingStream {
  const uint64_t *gp;  // next aligned global word to prefetch
  uint32_t sbase;      // shared-space address of this lane's word 0
  uint32_t rd;         // ring position of the next word to consume
  uint64_t cur;
  uint32_t sh;
  __device__ __forceinline__ void prefetch(uint32_t slot) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n\tcp.async.commit_group;\n"
                 :: "r"(sbase + slot * 256u), "l"(gp) : "memory");
    gp++;
  }
  __device__ __forceinline__ uint64_t pop() {
    // every consumed word was followed by exactly one newer commit group, so at most RING_WORDS - 1
    // groups may still be in flight when the oldest one is needed
    asm volatile("cp.async.wait_group %0;\n" :: "n"(RING_WORDS - 2) : "memory");
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];\n" : "=l"(v) : "r"(sbase + rd * 256u) : "memory");
    // refill the slot consumed by the PREVIOUS pop (its read retired long ago)
    prefetch((rd + RING_WORDS - 1) & (RING_WORDS - 1));
    rd = (rd + 1) & (RING_WORDS - 1);
    return v;
  }
  // `slot_base`: shared-space address reserved for this stream of this lane (see scan_chunk)
  __device__ __forceinline__ void init(const uint8_t *p, uint32_t slot_base) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    sh = (uint32_t)(a & 7) * 8;
    gp = reinterpret_cast<const uint64_t *>(a & ~(uintptr_t)7);
    sbase = slot_base;
    rd = 0;
    // fill slots 0 .. RING_WORDS-2; slot RING_WORDS-1 is refilled by the first pop
#pragma unroll
    for (int k = 0; k < RING_WORDS - 1; k++) prefetch(k);
    cur = pop();
  }
  __device__ __forceinline__ uint64_t next() {
    uint64_t nxt = pop();
    uint64_t raw = (cur >> sh) | ((nxt << 1) << (63 - sh));
    cur = nxt;
    return bswap64(raw);
  }
};

__device__ __forceinline__ uint64_t load_be64(const uint8_t *p) {
  BeStream s;
  s.init(p);
  return s.next();
}
__device__ __forceinline__ uint32_t load_be32_aligned(const uint8_t *p) {
  return __byte_perm(__ldg(reinterpret_cast<const uint32_t *>(p)), 0, 0x0123);
}

// Parsed page header (page.rs:78-94).
struct PageView {
  const uint8_t *bitset;  // 16-byte aligned (page offset is)
  const uint8_t *data;
  uint32_t data_len;
  uint32_t n_rows;
  __device__ __forceinline__ void open(const uint8_t *arena, const tskv_page_desc &d) {
    const uint8_t *pg = arena + d.offset;
    uint4 h = __ldg(reinterpret_cast<const uint4 *>(pg));
    uint32_t bitset_len = __byte_perm(h.x, 0, 0x0123);
    // rows: u64 BE at [4..12); the host validated it equals desc.num_values (< 2^32)
    n_rows = d.num_values;
    bitset = pg + 16;
    data = pg + 16 + bitset_len;
    data_len = d.size - 16 - bitset_len;
  }
};

// Validity bitmap, Arrow LSB-first (page.rs:78-84).
struct BitCursor {
  const uint32_t *wp;
  uint32_t word, ahead;  // `ahead` = the next 32 rows, loaded one word early to hide the latency
  __device__ __forceinline__ void init(const uint8_t *bitset) {
    wp = reinterpret_cast<const uint32_t *>(bitset);
    word = 0;
    ahead = __ldg(wp++);
  }
  // `row` must advance by one per call starting at 0. Reads at most 8 bytes past the bitmap.
  __device__ __forceinline__ bool next(uint32_t row) {
    if ((row & 31) == 0) {
      word = ahead;
      ahead = __ldg(wp++);
    }
    bool b = word & 1;
    word >>= 1;
    return b;
  }
};

__constant__ uint8_t c_s8b_count[16] = {240, 120, 60, 30, 20, 15, 12, 10, 8, 7, 6, 5, 4, 3, 2, 1};
__constant__ uint8_t c_s8b_bits[16] = {0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 15, 20, 30, 60};

// LEB128 (integer-encoding 4.0.2 decode_var); returns false when the slice ends first.
__device__ inline bool decode_varint(const uint8_t *p, uint32_t len, uint64_t *out) {
  uint64_t r = 0;
  uint32_t shift = 0;
  for (uint32_t i = 0; i < len; i++) {
    uint8_t b = __ldg(p + i);
    if (shift < 64) r |= (uint64_t)(b & 0x7f) << shift;
    shift += 7;
    if ((b & 0x80) == 0) {
      *out = r;
      return true;
    }
    if (shift > 63) return false;
  }
  return false;
}

__device__ __forceinline__ uint64_t pow10_u64(uint32_t k) {
  uint64_t s = 1;
  for (uint32_t i = 0; i < k; i++) s *= 10;
  return s;
}

// ------------------------------------------------------------------------------------------------
// Delta-family cursor: RLE / simple8b / raw prefix sum / raw BE, zig-zag or scaled.
// KIND is one of the DK_* delta kinds, or -1 for a runtime switch on `kind` (generic path).
// next() has no "first value" special case: open() arranges the state so that the first call
// yields the page's first value (RLE: v = first - delta; simple8b: one fake zero delta queued).
// Running out of encoded values sets the sticky `exhausted` flag (and yields garbage): callers check
// it once per segment / page instead of once per value.
// ------------------------------------------------------------------------------------------------
template <int KIND, typename STREAM = BeStream>
struct DeltaCursor {
  STREAM bs;
  uint64_t v;           // running value (raw bits)
  uint64_t delta;       // RLE delta (already scaled / zig-zag decoded)
  uint64_t scaler;      // S8B_SC
  uint64_t w;           // current simple8b word, consumed from the low bits
  uint64_t mask;        // low `bits` bits
  uint32_t words_left;  // 8-byte words not yet loaded
  uint32_t in_word;     // values left in `w`
  uint32_t bits;        // width of one value in `w`
  uint32_t ones;        // 1 for the run-of-ones selectors (payload ignored), else 0
  uint8_t kind;         // runtime kind (== KIND when KIND >= 0)
  bool exhausted;

  __device__ __forceinline__ int k() const { return KIND >= 0 ? KIND : kind; }

  // Returns TSKV_OK or a decode error. `pv.data` starts at the Encoding id byte; the host already
  // classified the page, so lengths needed by the fixed header are guaranteed.
  __device__ inline tskv_status open(const PageView &pv, uint8_t kind_, uint32_t smem_slot = 0) {
    kind = KIND >= 0 ? (uint8_t)KIND : kind_;
    exhausted = false;
    v = 0;
    delta = 0;
    scaler = 1;
    w = 0;
    mask = 0;
    in_word = 0;
    bits = 0;
    ones = 0;
    words_left = 0;
    const uint8_t *d = pv.data;
    switch (k()) {
      case DK_RLE_SC: {  // timestamp.rs:226-259: data = id | kind/scaler | first(8) | varint delta | varint n
        uint64_t dl;
        if (!decode_varint(d + 10, pv.data_len - 10, &dl)) return TSKV_ERR_SHORT_BLOCK;
        delta = dl * pow10_u64(__ldg(d + 1) & 0xf);
        v = load_be64(d + 2) - delta;
        break;
      }
      case DK_RLE_ZZ: {  // integer.rs:186-214
        uint64_t dl;
        if (!decode_varint(d + 10, pv.data_len - 10, &dl)) return TSKV_ERR_SHORT_BLOCK;
        delta = (uint64_t)zigzag_dec(dl);
        v = (uint64_t)zigzag_dec(load_be64(d + 2)) - delta;
        break;
      }
      case DK_S8B_SC:  // timestamp.rs:261-299
        scaler = pow10_u64(__ldg(d + 1) & 0xf);
        bs.init(d + 2, smem_slot);
        v = bs.next();
        words_left = (pv.data_len - 10) >> 3;
        in_word = 1;  // fake zero delta in front of the packed ones
        break;
      case DK_S8B_ZZ:  // integer.rs:216-248
        bs.init(d + 2, smem_slot);
        v = (uint64_t)zigzag_dec(bs.next());
        words_left = (pv.data_len - 10) >> 3;
        in_word = 1;
        break;
      case DK_RAW_SC:  // timestamp.rs:201-224
      case DK_RAW_ZZ:  // integer.rs:165-184
        bs.init(d + 2, smem_slot);
        words_left = (pv.data_len - 2) >> 3;
        break;
      case DK_RAWBE:  // timestamp.rs:301-323
        bs.init(d + 1, smem_slot);
        words_left = (pv.data_len - 1) >> 3;
        break;
      default:
        break;
    }
    return TSKV_OK;
  }

  __device__ __forceinline__ void refill() {
    if (words_left == 0) {
      exhausted = true;
      in_word = 0x7fffffff;  // keep yielding zeros without refilling again
      w = 0; mask = 0; bits = 0; ones = 0;
      return;
    }
    words_left--;
    w = bs.next();
    const uint32_t sel = (uint32_t)(w >> 60);
    in_word = c_s8b_count[sel];
    bits = c_s8b_bits[sel];
    ones = sel < 2 ? 1u : 0u;
    mask = bits ? (~0ull >> (64 - bits)) : 0ull;
  }
  // Next simple8b payload value (simple8b.rs:95-208).
  __device__ __forceinline__ uint64_t next_packed() {
    if (in_word == 0) refill();
    in_word--;
    const uint64_t u = (w & mask) | ones;
    w >>= bits;  // bits <= 60
    return u;
  }
  __device__ __forceinline__ uint64_t next_word() {
    if (words_left == 0) {
      exhausted = true;
      return 0;
    }
    words_left--;
    return bs.next();
  }

  // Value for the next VALID row.
  __device__ __forceinline__ uint64_t next() {
    switch (k()) {
      case DK_RLE_SC:
      case DK_RLE_ZZ: v += delta; return v;
      case DK_S8B_SC: v += next_packed() * scaler; return v;
      case DK_S8B_ZZ: v += (uint64_t)zigzag_dec(next_packed()); return v;
      case DK_RAW_SC: v += next_word(); return v;
      case DK_RAW_ZZ: v += (uint64_t)zigzag_dec(next_word()); return v;
      case DK_RAWBE: return next_word();
      default: exhausted = true; return 0;  // DK_ALLNULL never reaches here with a valid bit
    }
  }

  // timestamp.rs:273-279 quirk: with simple8b timestamps a NULL row 0 swallows the first value.
  __device__ __forceinline__ void skip_first_if_s8b_sc() {
    if (k() == DK_S8B_SC) in_word = 0;
  }
};

// ------------------------------------------------------------------------------------------------
// Gorilla cursor (float.rs:418-606): MSB-first bit stream after id | 0x10 | first(8).
// Terminates on the sentinel 0x7ff8_0000_0000_00ff (float.rs:16). Like DeltaCursor, the first next()
// needs no special case: a fake "repeat" control bit is queued in front of the stream.
// ------------------------------------------------------------------------------------------------
template <typename STREAM = BeStream>
struct GorillaCursor {
  STREAM bs;
  uint64_t val;
  uint64_t hi, lo;     // 128-bit window of the MSB-first bit stream; `pos` bits of hi are consumed
  uint32_t pos;        // 0..63
  uint32_t bits_used;  // bits consumed so far (incl. the fake one)
  uint32_t bits_total; // (data_len - 10) * 8 + 1
  uint32_t trailing, meaningful;
  bool done, err;      // done: sentinel reached or error; err: stream ended before the sentinel

  __device__ inline tskv_status open(const PageView &pv, uint32_t smem_slot = 0) {
    done = false;
    err = false;
    trailing = 0;
    meaningful = 64;
    const uint8_t *d = pv.data;
    bs.init(d + 2, smem_slot);
    val = bs.next();
    hi = 0;  // its last bit is the fake control bit 0 = "repeat the previous value"
    lo = bs.next();
    pos = 63;
    bits_used = 0;
    bits_total = (pv.data_len - 10) * 8 + 1;
    return TSKV_OK;
  }
  __device__ __forceinline__ bool consumed_any() const { return bits_used != 0; }

  // Next 64 bits of the stream, MSB-aligned, without consuming them.
  __device__ __forceinline__ uint64_t peek() const { return (hi << pos) | ((lo >> 1) >> (63 - pos)); }
  // Consumes n in [1,64] bits.
  __device__ __forceinline__ void skip(uint32_t n) {
    pos += n;
    bits_used += n;
    if (pos >= 64) {
      pos -= 64;
      hi = lo;
      lo = bs.next();
    }
  }

  // Decodes the next stream element; returns false at the sentinel / on error
  // ("unexpected end of block": the stream ended before the sentinel).
  __device__ __forceinline__ bool advance() {
    const uint32_t x = (uint32_t)(peek() >> 51);  // 13 bits: c0 c1 lead[5] sig[6]
    if (!(x & 0x1000)) {
      skip(1);  // repeat previous value: pushed without a sentinel test (float.rs:493-497), like the first value
      if (bits_used > bits_total) {
        err = true;
        return false;
      }
      return true;
    } else {
      if (!(x & 0x0800)) {
        skip(2);  // reuse the previous (leading, trailing) window
      } else {
        skip(13);
        const uint32_t leading = (x >> 6) & 0x1f;
        meaningful = x & 0x3f;
        if (meaningful > 0) {
          trailing = (uint8_t)(64 - leading - meaningful);  // u8 arithmetic like the reference
        } else {
          trailing = 0;
          meaningful = 64;
        }
      }
      const uint64_t s = peek() >> (64 - meaningful);
      skip(meaningful);
      val ^= s << (trailing & 0x3f);
    }
    if (bits_used > bits_total) {
      err = true;
      return false;
    }
    return val != 0x7ff80000000000ffull;
  }

  // Value for the next VALID row; sets `done` when the stream ended before the bitset did.
  __device__ __forceinline__ uint64_t next() {
    if (done || !advance()) done = true;
    return val;
  }
  // The reference decodes to the sentinel (float.rs:480-591): a stream without one is an error even
  // when enough values were produced. Returns false on "unexpected end of block".
  __device__ __forceinline__ bool drain() {
    while (!done) {
      if (!advance()) done = true;
    }
    return !err;
  }
};

// Runtime-dispatched cursor over every supported kind (decode-only kernel).
template <typename STREAM = BeStream>
struct AnyCursor {
  DeltaCursor<-1, STREAM> d;
  GorillaCursor<STREAM> g;
  bool is_gorilla;
  __device__ inline tskv_status open(const PageView &pv, uint8_t kind, uint32_t smem_slot = 0) {
    is_gorilla = kind == DK_GORILLA;
    if (is_gorilla) return g.open(pv, smem_slot);
    return d.open(pv, kind, smem_slot);
  }
  __device__ __forceinline__ uint64_t next() { return is_gorilla ? g.next() : d.next(); }
  // more valid rows than encoded values (BITSET_MISMATCH) / truncated stream (SHORT_BLOCK)
  __device__ __forceinline__ bool failed() const { return is_gorilla ? g.done : d.exhausted; }
  __device__ __forceinline__ bool stream_error() const { return is_gorilla && g.err; }
};

// Error status a page kind maps to before any decoding (host classification).
__device__ __forceinline__ tskv_status kind_status(uint8_t kind) {
  switch (kind) {
    case DK_BAD_ENCODING: return TSKV_ERR_BAD_ENCODING;
    case DK_UNSUPPORTED: return TSKV_ERR_UNSUPPORTED;
    case DK_SHORT: return TSKV_ERR_SHORT_BLOCK;
    case DK_BAD_LENGTH: return TSKV_ERR_BAD_LENGTH;
    case DK_BAD_PAGE: return TSKV_ERR_PAGE_FORMAT;
    default: return TSKV_OK;
  }
}

}  // namespace tskv
