// oracle/page.cc — CPU restatement of the reference page framing + CRC. TEST INFRASTRUCTURE ONLY.
#include <cstring>
#include <vector>

#include "tskv_oracle.h"

namespace {
uint32_t crc_table[8][256];
bool crc_init_done = false;
void crc_init() {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    crc_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; i++)
    for (int t = 1; t < 8; t++) crc_table[t][i] = (crc_table[t - 1][i] >> 8) ^ crc_table[0][crc_table[t - 1][i] & 0xff];
  crc_init_done = true;
}
inline uint32_t be32(const uint8_t *p) {
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}
inline uint64_t be64(const uint8_t *p) {
  uint64_t v = 0;
  for (int i = 0; i < 8; i++) v = (v << 8) | p[i];
  return v;
}
}  // namespace

extern "C" {

// crc32fast 1.4.2 Hasher::new()/update/finalize == CRC-32/IEEE (reflected, init/xorout 0xffffffff).
// Slicing-by-8 so that the CPU arm is not handicapped (crc32fast itself uses SIMD folding).
uint32_t orc_crc32(const uint8_t *data, uint64_t len) {
  if (!crc_init_done) crc_init();
  uint32_t c = 0xffffffffu;
  uint64_t i = 0;
  for (; i + 8 <= len; i += 8) {
    uint32_t lo, hi;
    memcpy(&lo, data + i, 4);
    memcpy(&hi, data + i + 4, 4);
    lo ^= c;
    c = crc_table[7][lo & 0xff] ^ crc_table[6][(lo >> 8) & 0xff] ^ crc_table[5][(lo >> 16) & 0xff] ^ crc_table[4][lo >> 24] ^
        crc_table[3][hi & 0xff] ^ crc_table[2][(hi >> 8) & 0xff] ^ crc_table[1][(hi >> 16) & 0xff] ^ crc_table[0][hi >> 24];
  }
  for (; i < len; i++) c = crc_table[0][(c ^ data[i]) & 0xff] ^ (c >> 8);
  return c ^ 0xffffffffu;
}

// tskv/src/tsm/page.rs:334-345 (and :577-586): u32be bitset_len | u64be rows | u32be crc32(data) |
// bitset | data.
uint64_t orc_page_build(const uint8_t *bitset, uint32_t bitset_len, uint64_t n_rows,
                        const uint8_t *data, uint64_t data_len, uint8_t *out) {
  uint32_t crc = orc_crc32(data, data_len);
  out[0] = (uint8_t)(bitset_len >> 24);
  out[1] = (uint8_t)(bitset_len >> 16);
  out[2] = (uint8_t)(bitset_len >> 8);
  out[3] = (uint8_t)bitset_len;
  for (int i = 0; i < 8; i++) out[4 + i] = (uint8_t)(n_rows >> (56 - 8 * i));
  out[12] = (uint8_t)(crc >> 24);
  out[13] = (uint8_t)(crc >> 16);
  out[14] = (uint8_t)(crc >> 8);
  out[15] = (uint8_t)crc;
  if (bitset_len) memcpy(out + 16, bitset, bitset_len);
  if (data_len) memcpy(out + 16 + bitset_len, data, data_len);
  return 16 + (uint64_t)bitset_len + data_len;
}

// tskv/src/tsm/page.rs:58-94 (crc_validation, null_bitset, data_buffer) then
// tsm/reader.rs:658-731 (data_buf_to_arrow_array).
tskv_status orc_page_decode(uint32_t phys_type, const uint8_t *page, uint64_t size, int verify_crc,
                            uint64_t *out_vals, uint8_t *out_valid, uint64_t cap_rows,
                            uint64_t *out_n_rows) {
  if (size < 16) return TSKV_ERR_PAGE_FORMAT;
  uint64_t bitset_len = be32(page);
  uint64_t n_rows = be64(page + 4);
  uint32_t crc = be32(page + 12);
  if (16 + bitset_len > size) return TSKV_ERR_PAGE_FORMAT;
  if (bitset_len * 8 < n_rows) return TSKV_ERR_PAGE_FORMAT;  // append_packed_range would panic
  const uint8_t *bitset = page + 16;
  const uint8_t *data = page + 16 + bitset_len;
  uint64_t data_len = size - 16 - bitset_len;
  if (verify_crc && orc_crc32(data, data_len) != crc) return TSKV_ERR_CRC_MISMATCH;
  if (out_n_rows) *out_n_rows = n_rows;
  if (n_rows > cap_rows) return TSKV_ERR_INVALID_ARG;
  return orc_decode_column(phys_type, data, data_len, bitset, n_rows, out_vals, out_valid);
}

// Same output layout as tskvgpu_decode_pages (include/tskv_gpu.h).
tskv_status orc_decode_pages(const uint8_t *arena, uint64_t arena_len, const tskv_page_desc *descs,
                             uint64_t n_descs, uint64_t first_page, uint64_t n_pages,
                             int verify_crc, uint64_t *out_values, uint8_t *out_validity) {
  if (first_page + n_pages > n_descs) return TSKV_ERR_INVALID_ARG;
  uint64_t row_off = 0, val_off = 0;
  std::vector<uint8_t> valid;
  for (uint64_t p = first_page; p < first_page + n_pages; p++) {
    const tskv_page_desc &d = descs[p];
    if (d.offset + d.size > arena_len) return TSKV_ERR_INVALID_ARG;
    valid.assign(d.num_values ? d.num_values : 1, 0);
    uint64_t n_rows = 0;
    tskv_status st = orc_page_decode(d.phys_type, arena + d.offset, d.size, verify_crc,
                                     out_values + row_off, valid.data(), d.num_values, &n_rows);
    if (st != TSKV_OK) return st;
    if (n_rows != d.num_values) return TSKV_ERR_PAGE_FORMAT;
    uint64_t bm_bytes = ((uint64_t)d.num_values + 63) / 64 * 8;
    memset(out_validity + val_off, 0, bm_bytes);
    for (uint64_t r = 0; r < n_rows; r++)
      if (valid[r]) out_validity[val_off + (r >> 3)] |= (uint8_t)(1u << (r & 7));
    row_off += d.num_values;
    val_off += bm_bytes;
  }
  return TSKV_OK;
}

}  // extern "C"
