// tsm_writer.h — host-side TSM page writer (encode side of the column codecs + page framing).
// Needed to synthesise benchmark/test inputs; the reference's write path itself is out of scope.
// Formats (reference tree):
//   simple8b   tskv/src/tsm/codec/simple8b.rs:26-76      timestamp  codec/timestamp.rs:51-175
//   integer    tskv/src/tsm/codec/integer.rs:40-140      gorilla    codec/float.rs:32-243
//   raw/Null   codec/timestamp.rs:21-31, float.rs:256-267           page  tsm/page.rs:334-345
// This is an independent implementation (bit-writer based), cross-checked byte-for-byte against the
// oracle's line-faithful restatement in tests/test_writer_vs_oracle.py.
#pragma once
#include <cstdint>
#include <vector>

namespace tskv {

using Bytes = std::vector<uint8_t>;

// Appends simple8b words for `src` to `out`; false when a value needs more than 60 bits.
bool simple8b_pack(const uint64_t *src, size_t n, Bytes &out);

// Column encoders: append [Encoding id | payload] to `out` (nothing for n == 0).
bool encode_timestamps(const int64_t *src, size_t n, Bytes &out);  // Encoding::DeltaTs = 11
bool encode_integers(const int64_t *src, size_t n, Bytes &out);    // Encoding::Delta = 2
bool encode_floats(const double *src, size_t n, Bytes &out);       // Encoding::Gorilla = 6
void encode_raw(const uint64_t *src, size_t n, Bytes &out);        // Encoding::Null = 1

// Appends one framed page: u32be bitset_len | u64be rows | u32be crc32(data) | bitset | data.
// `validity` is an Arrow LSB-first bitmap of ceil(rows/8) bytes, or nullptr for all-valid.
void append_page(const uint8_t *validity, uint64_t rows, const Bytes &data, Bytes &arena);

}  // namespace tskv
