// host_util.h — host-side helpers of the C-ABI library: page framing validation, decode-kind
// classification (the codec dispatch of tskv/src/tsm/codec/instance.rs:358-401 done once at upload
// instead of per read), CRC-32/IEEE (tskv/src/tsm/page.rs:58-76).
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../include/tskv_gpu.h"
#include "kinds.h"

namespace tskv {

// CRC-32/IEEE (crc32fast semantics), slicing-by-8.
uint32_t crc32_ieee(const uint8_t *data, size_t len);
// The slicing-by-8 tables ([8][256]), for the device-side verifier.
const uint32_t *crc32_tables();

struct PageHeader {
  uint32_t bitset_len;
  uint64_t n_rows;
  uint32_t crc;
  const uint8_t *bitset;
  const uint8_t *data;
  uint64_t data_len;
};

// Parses `u32be bitset_len | u64be rows | u32be crc | bitset | data` (page.rs:334-345).
// Returns false when the page is shorter than its own framing.
bool parse_page(const uint8_t *page, uint64_t size, PageHeader *out);

// Decode kind (DK_* of cursors.cuh) of a page of physical type `phys_type`; error kinds encode the
// reference's decode errors so that they surface when (and only when) the page is read.
uint8_t classify_page(const PageHeader &h, uint8_t phys_type);

}  // namespace tskv
