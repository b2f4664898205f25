// kinds.h — decode kinds shared by host classification and device cursors.
#pragma once
#include <cstdint>

namespace tskv {

// Decode kinds stored in tskv_page_desc.reserved on the device (filled by the host at upload from
// data[0] = Encoding id and data[1] >> 4 = sub-encoding; dispatch of codec/instance.rs:365-401).
enum : uint8_t {
  DK_ALLNULL = 0,    // empty data buffer: every row is null
  DK_RLE_SC = 1,     // DeltaTs kind 2: first + i * (varint * 10^k)
  DK_S8B_SC = 2,     // DeltaTs kind 1: simple8b deltas * 10^k
  DK_RAW_SC = 3,     // DeltaTs kind 0: prefix sum of 8-byte BE deltas
  DK_RLE_ZZ = 4,     // Delta kind 2: zigzag first + i * zigzag^-1(varint)
  DK_S8B_ZZ = 5,     // Delta kind 1: zigzag simple8b deltas
  DK_RAW_ZZ = 6,     // Delta kind 0: prefix sum of zigzag^-1(8-byte BE)
  DK_RAWBE = 7,      // Encoding::Null: 8-byte BE values
  DK_GORILLA = 8,    // f64 XOR stream
  DK_BAD_ENCODING = 9,   // "invalid block encoding" (sub-encoding nibble > 2)
  DK_UNSUPPORTED = 10,   // Quantile (pco)
  DK_SHORT = 11,         // data too short for its header ("not enough data to decode ...")
  DK_BAD_LENGTH = 12,    // "invalid uncompressed block length"
  DK_BAD_PAGE = 13,      // page shorter than header + bitset
  DK_BOOL_PACK = 14,     // boolean bit-pack (boolean.rs:79-111): 0x10 | varint count | bits, MSB first
  DK_BOOL_RAW = 15       // boolean under Encoding::Null (boolean.rs:112-140): one byte per value, 1 = true
};
inline bool dk_is_error(uint8_t k) { return k >= DK_BAD_ENCODING && k <= DK_BAD_PAGE; }

}  // namespace tskv
