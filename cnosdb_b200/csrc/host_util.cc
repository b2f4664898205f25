// host_util.cc — see host_util.h.
#include "host_util.h"

#include <algorithm>
#include <cmath>
#include <vector>
#include <mutex>

namespace tskv {

namespace {
uint32_t g_tab[8][256];
std::once_flag g_once;
void init_tables() {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1)));
    g_tab[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; i++)
    for (int t = 1; t < 8; t++) g_tab[t][i] = (g_tab[t - 1][i] >> 8) ^ g_tab[0][g_tab[t - 1][i] & 0xff];
}
inline uint32_t rd32be(const uint8_t *p) {
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
}
}  // namespace

const uint32_t *crc32_tables() {
  std::call_once(g_once, init_tables);
  return &g_tab[0][0];
}

uint32_t crc32_ieee(const uint8_t *p, size_t len) {
  std::call_once(g_once, init_tables);
  uint32_t crc = ~0u;
  while (len >= 8) {
    uint32_t lo = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    uint32_t hi = (uint32_t)p[4] | ((uint32_t)p[5] << 8) | ((uint32_t)p[6] << 16) | ((uint32_t)p[7] << 24);
    lo ^= crc;
    crc = g_tab[7][lo & 0xff] ^ g_tab[6][(lo >> 8) & 0xff] ^ g_tab[5][(lo >> 16) & 0xff] ^
          g_tab[4][lo >> 24] ^ g_tab[3][hi & 0xff] ^ g_tab[2][(hi >> 8) & 0xff] ^
          g_tab[1][(hi >> 16) & 0xff] ^ g_tab[0][hi >> 24];
    p += 8;
    len -= 8;
  }
  while (len--) crc = (crc >> 8) ^ g_tab[0][(crc ^ *p++) & 0xff];
  return ~crc;
}

double plan_selected_fraction(const uint32_t *arena_series, uint64_t n_arena, const uint32_t *sel, uint64_t n_sel) {
  if (!sel || n_arena == 0) return 1.0;
  const uint32_t *ie = sel + n_sel;
  const uint64_t in_range = (uint64_t)(std::upper_bound(sel, ie, arena_series[n_arena - 1]) - std::lower_bound(sel, ie, arena_series[0]));
  return std::min(1.0, (double)in_range / (double)n_arena);
}

bool plan_use_cooperative(double est_selected_pages, int sm_count, int min_blocks_per_sm, int threads_per_block) {
  return est_selected_pages < 0.25 * (double)sm_count * min_blocks_per_sm * threads_per_block;
}

uint32_t plan_gorilla_group(double est_gorilla_pages, double resident_warps) {
  uint32_t g = 1;
  while (g < 32 && est_gorilla_pages / g > 0.75 * resident_warps) g *= 2;
  return g;
}

void plan_serial_grids(int n_bins, const double *chunks, const double *t_chunk, const int *occ, int sm_count,
                       int warps_per_block, int *grid_out) {
  std::vector<double> cand;
  for (int b = 0; b < n_bins; b++) {
    grid_out[b] = 0;
    if (chunks[b] <= 0) continue;
    for (int k = 1; k <= 64; k++) cand.push_back(k * t_chunk[b]);
  }
  if (cand.empty()) return;
  std::sort(cand.begin(), cand.end());
  auto blocks_for = [&](int b, double T) -> double {
    const double rounds = std::floor(T / t_chunk[b] + 1e-9);
    if (rounds < 1) return -1;
    return std::ceil(chunks[b] / (rounds * warps_per_block));
  };
  double best = cand.back();
  for (double T : cand) {
    double sm_used = 0;
    bool ok = true;
    for (int b = 0; b < n_bins && ok; b++) {
      if (chunks[b] <= 0) continue;
      const double n = blocks_for(b, T);
      if (n < 0) ok = false;
      else sm_used += n / std::max(1, occ[b]);
    }
    if (ok && sm_used <= sm_count) { best = T; break; }
  }
  for (int b = 0; b < n_bins; b++)
    if (chunks[b] > 0) grid_out[b] = std::max(1, (int)std::max(1.0, blocks_for(b, best)));
}

bool parse_page(const uint8_t *page, uint64_t size, PageHeader *h) {
  if (size < 16) return false;
  h->bitset_len = rd32be(page);
  h->n_rows = ((uint64_t)rd32be(page + 4) << 32) | rd32be(page + 8);
  h->crc = rd32be(page + 12);
  if (16 + (uint64_t)h->bitset_len > size) return false;
  if ((uint64_t)h->bitset_len * 8 < h->n_rows) return false;
  h->bitset = page + 16;
  h->data = page + 16 + h->bitset_len;
  h->data_len = size - 16 - h->bitset_len;
  return true;
}

uint8_t classify_page(const PageHeader &h, uint8_t phys_type) {
  if (h.data_len == 0) return DK_ALLNULL;  // every codec: empty buffer => all-null array
  const uint8_t *d = h.data;
  unsigned enc = d[0];
  if (enc == TSKV_ENC_QUANTILE) return DK_UNSUPPORTED;
  if (enc == TSKV_ENC_NULL) return ((h.data_len - 1) & 7) ? DK_BAD_LENGTH : DK_RAWBE;
  bool ts_family;
  switch (phys_type) {
    case TSKV_PT_TIME: ts_family = enc != TSKV_ENC_DELTA; break;      // get_ts_codec
    case TSKV_PT_I64: ts_family = enc == TSKV_ENC_DELTA_TS; break;    // get_i64_codec
    case TSKV_PT_U64: ts_family = false; break;                       // get_u64_codec
    case TSKV_PT_F64:                                                 // get_f64_codec => gorilla
      return h.data_len < 10 ? DK_SHORT : DK_GORILLA;
    default: return DK_UNSUPPORTED;
  }
  if (h.data_len < 2) return DK_SHORT;  // src[0] on an empty slice panics in the reference
  unsigned sub = d[1] >> 4;
  if (sub > 2) return DK_BAD_ENCODING;
  if (sub == 0) {
    uint64_t l = h.data_len - 2;
    if (l == 0 || (l & 7)) return DK_BAD_LENGTH;
    return ts_family ? DK_RAW_SC : DK_RAW_ZZ;
  }
  if (h.data_len < 10) return DK_SHORT;
  if (sub == 2) return ts_family ? DK_RLE_SC : DK_RLE_ZZ;
  if ((h.data_len - 10) & 7) return DK_SHORT;  // simple8b::decode would slice out of bounds
  return ts_family ? DK_S8B_SC : DK_S8B_ZZ;
}

}  // namespace tskv
