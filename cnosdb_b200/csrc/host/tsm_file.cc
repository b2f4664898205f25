// tsm_file.cc — TSM FILE -> page arena + descriptor table: the host-side loader that lets real
// `.tsm` files, not only synthetic arenas, feed the engine, and the matching writer the tests build files with.
//
// Layout of a TSM file (all citations relative to the reference tree):
//   magic 0x012CDA16 (u32 BE, tskv/src/tsm/writer.rs:38,148-156) | pages, back to back (writer.rs:316-350) |
//   META = chunks | chunk groups | chunk-group meta (writer.rs:165-231,497-520) | footer (FOOTER_SIZE = 131140 bytes,
//   tskv/src/tsm/mod.rs:18).
// Reading order is TsmReader::open's (tsm/reader.rs:120-168,399-474): footer -> META = [footer.series.chunk_offset,
// len - FOOTER_SIZE) -> ChunkGroupMeta at footer.table.{chunk_group_offset,size} (offsets INSIDE META) -> every
// table's ChunkGroup -> every series' Chunk -> its ColumnGroups -> PageWriteSpec{offset (absolute), size, meta}.
// Every struct is serialised with bincode 1.3.3's default options: little-endian fixed-width integers, u64 lengths
// for String / Vec / maps, u8 tag for Option, u32 variant index for enums, struct fields in declaration order:
//   Footer{version: TsmVersion, time_range, table: TableMeta{chunk_group_offset, chunk_group_size},
//          series: SeriesMeta{bloom_filter: BloomFilter{b: Vec<u8>, mask: u64}, chunk_offset, chunk_size}}  footer.rs:19-127
//   ChunkGroupMeta{tables: BTreeMap<String, ChunkGroupWriteSpec{table_schema: TskvTableSchema, chunk_group_offset,
//          chunk_group_size, time_range, count: usize}>}                                          chunk_group.rs:58-106
//   TskvTableSchema (hand-written Serialize: tenant, db, name, schema_version u64, next_column_id u32,
//          columns: Vec<TableColumn>, columns_index: HashMap<String, usize>)   common/models/src/schema/tskv_table_schema.rs:36-64
//   TableColumn{id u32, name, column_type: ColumnType{Tag | Time(TimeUnit) | Field(ValueType)}, encoding: Encoding} :532-537,761-765
//   ChunkGroup{chunks: Vec<ChunkWriteSpec{series_id u32, chunk_offset, chunk_size, statics{time_range}}>}  chunk_group.rs:14-17, chunk.rs:150-190
//   Chunk{time_range, table_name, series_id, series_key: SeriesKey{tags: Vec<Tag{key, value: Vec<u8>}>, table},
//          next_column_group_id u64, column_groups: BTreeMap<u64, ColumnGroup>}                  chunk.rs:16-25
//   ColumnGroup{column_group_id u64, pages_offset, size, time_range, pages: Vec<PageWriteSpec>}  column_group.rs:9-17
//   PageWriteSpec{offset u64, size u64, meta: PageMeta{num_values u32, column: TableColumn,
//          statistics: PageStatistics{Bool|F64|I64|U64|Bytes}(ValueStatistics{min, max: Option<T>,
//          distinct_count: Option<u64>, null_count u64})}}                                      page.rs:599-620, statistics/mod.rs:4-9
// TsmVersion::V2 files carry META through the string codec (writer.rs:507-518, reader.rs:133-144):
//   [Encoding id][0x10 for snappy] payload; Snappy (raw format, one length-prefixed string) and Zlib / Gzip (zlib) are
//   decoded here; Zstd / Bzip META returns TSKV_ERR_UNSUPPORTED (no such library in this image).
// Parity status: UNPINNED beyond structure - the reference tree holds no .tsm fixture and cannot be run here; what IS
// pinned is the bincode layout rule set above against the reference's FOOTER_SIZE constant (4 + 16 + 16 + (8 + 131072 +
// 8) + 16 = 131140) and the reference reader's field order. The engine's page bytes themselves are covered by the codec
// goldens.
// Pages in a TSM file are packed back to back; the engine wants every page 16-byte aligned, so the loader REPACKS the
// selected pages into a fresh arena (one memcpy per page) and reports each column group's time_range() for statistics
// pruning (tskvgpu_pages_set_time_bounds).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <zlib.h>

#include "../../../include/tskv_tsm.h"

namespace {

constexpr uint64_t FOOTER_SIZE = 131140;  // tskv/src/tsm/mod.rs:18
constexpr uint32_t TSM_MAGIC = 0x012CDA16;
constexpr uint64_t BLOOM_BYTES = 1024 * 1024 / 8;  // BLOOM_FILTER_BITS, tsm/mod.rs:17

struct Rd {  // bounds-checked bincode reader
  const uint8_t *p;
  uint64_t n, i = 0;
  bool ok = true;
  bool need(uint64_t k) {
    if (!ok || k > n - i) { ok = false; return false; }
    return true;
  }
  uint8_t u8() { if (!need(1)) return 0; return p[i++]; }
  uint32_t u32() { if (!need(4)) return 0; uint32_t v; memcpy(&v, p + i, 4); i += 4; return v; }
  uint64_t u64() { if (!need(8)) return 0; uint64_t v; memcpy(&v, p + i, 8); i += 8; return v; }
  int64_t i64() { return (int64_t)u64(); }
  uint16_t u16() { if (!need(2)) return 0; uint16_t v; memcpy(&v, p + i, 2); i += 2; return v; }
  std::string str() {
    uint64_t len = u64();
    if (!need(len)) return {};
    std::string s((const char *)p + i, (size_t)len);
    i += len;
    return s;
  }
  void skip(uint64_t k) { if (need(k)) i += k; }
  void skip_bytes_vec() { skip(u64()); }
};

struct Wr {  // bincode writer (tests build files with it)
  std::vector<uint8_t> b;
  void u8(uint8_t v) { b.push_back(v); }
  void u16(uint16_t v) { raw(&v, 2); }
  void u32(uint32_t v) { raw(&v, 4); }
  void u64(uint64_t v) { raw(&v, 8); }
  void i64(int64_t v) { raw(&v, 8); }
  void raw(const void *q, size_t k) { const uint8_t *c = (const uint8_t *)q; b.insert(b.end(), c, c + k); }
  void str(const std::string &s) { u64(s.size()); raw(s.data(), s.size()); }
};

// ColumnType / ValueType (tskv_table_schema.rs:761-765, value_type.rs:8-16) -> the engine's physical type, or -1
struct ColType {
  uint32_t kind = 0;   // 0 Tag, 1 Time, 2 Field
  uint32_t sub = 0;    // TimeUnit / ValueType variant
};
ColType read_column_type(Rd &r) {
  ColType c;
  c.kind = r.u32();
  if (c.kind == 1) c.sub = r.u32();          // TimeUnit: Second, Millisecond, Microsecond, Nanosecond
  else if (c.kind == 2) {
    c.sub = r.u32();                          // ValueType: Unknown, Float, Integer, Unsigned, Boolean, String, Geometry
    if (c.sub == 6) { r.u32(); r.u16(); }     // Geometry{sub_type: GeometryType, srid: i16}
  } else if (c.kind != 0) r.ok = false;
  return c;
}
int phys_type_of(const ColType &c) {
  if (c.kind == 1) return TSKV_PT_TIME;
  if (c.kind == 2) {
    if (c.sub == 1) return TSKV_PT_F64;
    if (c.sub == 2) return TSKV_PT_I64;
    if (c.sub == 3) return TSKV_PT_U64;
    if (c.sub == 4) return TSKV_PT_BOOL;
  }
  return -1;  // tags, string, geometry: not on this engine's path
}
struct Column {
  uint32_t id = 0;
  std::string name;
  ColType type;
  uint32_t encoding = 0;
};
Column read_table_column(Rd &r) {
  Column c;
  c.id = r.u32();
  c.name = r.str();
  c.type = read_column_type(r);
  c.encoding = r.u32();
  return c;
}
// PageStatistics{Bool|F64|I64|U64|Bytes}(ValueStatistics{min, max: Option<T>, distinct_count: Option<u64>, null_count})
// (page.rs:607-613, statistics/mod.rs:4-9). min / max of the numeric / boolean variants -> tskv_value_stats.
tskv_value_stats read_value_statistics(Rd &r, uint32_t variant) {
  tskv_value_stats st{};
  bool have[2] = {false, false};
  uint64_t v[2] = {0, 0};
  for (int k = 0; k < 2; k++) {
    if (!r.u8()) continue;  // None
    switch (variant) {
      case 0: v[k] = r.u8(); have[k] = true; break;             // bool
      case 1: case 2: case 3: v[k] = r.u64(); have[k] = true; break;  // f64 bits / i64 / u64
      case 4: r.skip_bytes_vec(); break;                         // Vec<u8> (strings): not on this engine's path
      default: r.ok = false;
    }
  }
  if (r.u8()) r.u64();  // distinct_count
  r.u64();              // null_count
  if (have[0] && have[1]) { st.min = v[0]; st.max = v[1]; st.flags = TSKV_STATS_MINMAX; }
  return st;
}
void skip_table_schema(Rd &r) {
  r.str(); r.str(); r.str();  // tenant, db, name
  r.u64();                    // schema_version
  r.u32();                    // next_column_id
  uint64_t n = r.u64();
  for (uint64_t i = 0; i < n && r.ok; i++) read_table_column(r);
  n = r.u64();                // columns_index: HashMap<String, usize>
  for (uint64_t i = 0; i < n && r.ok; i++) { r.str(); r.u64(); }
}

// ---- snappy raw format (format_description.txt of google/snappy): uvarint length, then literal / copy elements
bool snappy_decompress(const uint8_t *s, uint64_t n, std::vector<uint8_t> &out) {
  uint64_t i = 0, len = 0;
  for (unsigned shift = 0;; shift += 7) {
    if (i >= n || shift > 35) return false;
    uint8_t b = s[i++];
    len |= (uint64_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) break;
  }
  out.clear();
  out.reserve(len);
  while (i < n) {
    const uint8_t tag = s[i++];
    if ((tag & 3) == 0) {  // literal
      uint64_t l = tag >> 2;
      if (l >= 60) {
        const unsigned extra = (unsigned)l - 59;
        if (i + extra > n) return false;
        l = 0;
        for (unsigned k = 0; k < extra; k++) l |= (uint64_t)s[i + k] << (8 * k);
        i += extra;
      }
      l += 1;
      if (i + l > n) return false;
      out.insert(out.end(), s + i, s + i + l);
      i += l;
    } else {
      uint64_t l, off;
      if ((tag & 3) == 1) {
        if (i + 1 > n) return false;
        l = 4 + ((tag >> 2) & 7);
        off = ((uint64_t)(tag >> 5) << 8) | s[i];
        i += 1;
      } else if ((tag & 3) == 2) {
        if (i + 2 > n) return false;
        l = 1 + (tag >> 2);
        off = s[i] | ((uint64_t)s[i + 1] << 8);
        i += 2;
      } else {
        if (i + 4 > n) return false;
        l = 1 + (tag >> 2);
        off = s[i] | ((uint64_t)s[i + 1] << 8) | ((uint64_t)s[i + 2] << 16) | ((uint64_t)s[i + 3] << 24);
        i += 4;
      }
      if (off == 0 || off > out.size()) return false;
      for (uint64_t k = 0; k < l; k++) out.push_back(out[out.size() - off]);
    }
  }
  return out.size() == len;
}
// literals only: a valid (uncompressed) snappy stream, enough for the writer the tests use
void snappy_store(const uint8_t *s, uint64_t n, std::vector<uint8_t> &out) {
  uint64_t v = n;
  do { uint8_t b = v & 0x7f; v >>= 7; out.push_back(b | (v ? 0x80 : 0)); } while (v);
  for (uint64_t i = 0; i < n;) {
    const uint64_t l = std::min<uint64_t>(n - i, 1u << 16);
    if (l <= 60) out.push_back((uint8_t)((l - 1) << 2));
    else { out.push_back((uint8_t)(61 << 2)); out.push_back((uint8_t)((l - 1) & 0xff)); out.push_back((uint8_t)((l - 1) >> 8)); }
    out.insert(out.end(), s + i, s + i + l);
    i += l;
  }
}
bool zlib_inflate(const uint8_t *s, uint64_t n, bool gzip, std::vector<uint8_t> &out) {
  z_stream z{};
  if (inflateInit2(&z, gzip ? 15 + 16 : 15) != Z_OK) return false;
  z.next_in = const_cast<uint8_t *>(s);
  z.avail_in = (uInt)n;
  out.clear();
  uint8_t buf[1 << 16];
  int rc;
  do {
    z.next_out = buf;
    z.avail_out = sizeof(buf);
    rc = inflate(&z, Z_NO_FLUSH);
    if (rc != Z_OK && rc != Z_STREAM_END) { inflateEnd(&z); return false; }
    out.insert(out.end(), buf, buf + (sizeof(buf) - z.avail_out));
  } while (rc != Z_STREAM_END);
  inflateEnd(&z);
  return true;
}

// The string codec's inverse for ONE string (codec/string.rs:170-330): snappy holds uvarint(len) | bytes, the
// others u64-BE(len) | bytes.
tskv_status decode_meta_v2(const uint8_t *s, uint64_t n, std::vector<uint8_t> &meta, std::string &err) {
  if (n < 2) { err = "TSM V2: metadata block too short"; return TSKV_ERR_PAGE_FORMAT; }
  std::vector<uint8_t> raw;
  const uint8_t enc = s[0];
  if (enc == TSKV_ENC_GORILLA + 1 /* Snappy = 7 */) {
    if ((s[1] >> 4) != 1 || !snappy_decompress(s + 2, n - 2, raw)) { err = "TSM V2: snappy metadata does not decode"; return TSKV_ERR_PAGE_FORMAT; }
    uint64_t i = 0, len = 0;
    for (unsigned shift = 0;; shift += 7) {
      if (i >= raw.size() || shift > 63) { err = "TSM V2: bad string length"; return TSKV_ERR_PAGE_FORMAT; }
      uint8_t b = raw[i++];
      len |= (uint64_t)(b & 0x7f) << shift;
      if (!(b & 0x80)) break;
    }
    if (len > raw.size() - i) { err = "TSM V2: bad string length"; return TSKV_ERR_PAGE_FORMAT; }
    meta.assign(raw.begin() + i, raw.begin() + i + len);
    return TSKV_OK;
  }
  if (enc == 9 /* Zlib */ || enc == 4 /* Gzip */) {
    if (!zlib_inflate(s + 1, n - 1, enc == 4, raw) || raw.size() < 8) { err = "TSM V2: zlib metadata does not decode"; return TSKV_ERR_PAGE_FORMAT; }
    uint64_t len = 0;
    for (int k = 0; k < 8; k++) len = (len << 8) | raw[k];
    if (len > raw.size() - 8) { err = "TSM V2: bad string length"; return TSKV_ERR_PAGE_FORMAT; }
    meta.assign(raw.begin() + 8, raw.begin() + 8 + len);
    return TSKV_OK;
  }
  err = "TSM V2: metadata compressed with an encoding this build cannot decode (zstd / bzip)";
  return TSKV_ERR_UNSUPPORTED;
}

thread_local std::string g_err;

}  // namespace

extern "C" {


const char *tskvtsm_last_error(void) { return g_err.c_str(); }

void tskvtsm_free(tskvtsm_result *r) {
  if (!r) return;
  free(r->arena);
  free(r->descs);
  free(r->cg_bounds);
  free(r->value_stats);
  memset(r, 0, sizeof(*r));
}

// TsmReader::open (tsm/reader.rs:120-168): footer -> META -> chunk group meta -> chunk groups -> chunks; then the page
// specs of every column group -> descriptors + repacked arena. `table` (may be NULL / empty) restricts to one table.
tskv_status tskvtsm_load(const uint8_t *file, uint64_t len, const char *table, tskvtsm_result *out) {
  g_err.clear();
  if (!file || !out) return TSKV_ERR_INVALID_ARG;
  memset(out, 0, sizeof(*out));
  if (len < FOOTER_SIZE + 4) { g_err = "file is too small"; return TSKV_ERR_PAGE_FORMAT; }  // reader.rs:400-404
  uint32_t magic_be = 0;
  memcpy(&magic_be, file, 4);
  if (__builtin_bswap32(magic_be) != TSM_MAGIC) { g_err = "not a TSM file (magic)"; return TSKV_ERR_PAGE_FORMAT; }
  // ---- footer
  Rd f{file + (len - FOOTER_SIZE), FOOTER_SIZE};
  const uint32_t version_idx = f.u32();  // TsmVersion: V1 = variant 0, V2 = variant 1
  out->min_ts = f.i64();
  out->max_ts = f.i64();
  const uint64_t cgm_off = f.u64(), cgm_size = f.u64();
  const uint64_t bloom_len = f.u64();
  f.skip(bloom_len);
  f.u64();  // mask
  const uint64_t chunk_off = f.u64();
  f.u64();  // chunk_size
  if (!f.ok || version_idx > 1 || bloom_len != BLOOM_BYTES || chunk_off > len - FOOTER_SIZE) {
    g_err = "footer does not parse";
    return TSKV_ERR_PAGE_FORMAT;
  }
  out->version = version_idx + 1;
  // ---- META
  const uint8_t *meta = file + chunk_off;
  uint64_t meta_len = len - FOOTER_SIZE - chunk_off;
  std::vector<uint8_t> inflated;
  if (version_idx == 1) {
    tskv_status st = decode_meta_v2(meta, meta_len, inflated, g_err);
    if (st != TSKV_OK) return st;
    meta = inflated.data();
    meta_len = inflated.size();
  }
  if (cgm_size > meta_len || cgm_off > meta_len - cgm_size) { g_err = "chunk group meta out of bounds"; return TSKV_ERR_PAGE_FORMAT; }
  // ---- ChunkGroupMeta -> (table, chunk group offset / size)
  struct Table { std::string name; uint64_t off, size; };
  std::vector<Table> tables;
  {
    Rd r{meta + cgm_off, cgm_size};
    const uint64_t n = r.u64();
    for (uint64_t i = 0; i < n && r.ok; i++) {
      Table t;
      t.name = r.str();
      skip_table_schema(r);
      t.off = r.u64();
      t.size = r.u64();
      r.i64(); r.i64();  // time_range
      r.u64();           // count
      tables.push_back(t);
    }
    if (!r.ok) { g_err = "chunk group meta does not parse"; return TSKV_ERR_PAGE_FORMAT; }
  }
  std::vector<tskv_page_desc> descs;
  std::vector<tskv_value_stats> vstats;  // per descriptor
  std::vector<tskv_time_range> bounds;
  std::vector<uint8_t> arena;
  uint64_t skipped = 0;
  for (const Table &t : tables) {
    if (table && table[0] && t.name != table) continue;
    if (t.size > meta_len || t.off > meta_len - t.size) { g_err = "chunk group out of bounds"; return TSKV_ERR_PAGE_FORMAT; }
    Rd g{meta + t.off, t.size};
    const uint64_t n_chunks = g.u64();
    for (uint64_t c = 0; c < n_chunks && g.ok; c++) {
      g.u32();  // series_id (repeated inside the chunk)
      const uint64_t coff = g.u64(), csize = g.u64();
      g.i64(); g.i64();
      if (!g.ok) break;
      if (csize > meta_len || coff > meta_len - csize) { g_err = "chunk out of bounds"; return TSKV_ERR_PAGE_FORMAT; }
      Rd r{meta + coff, csize};
      r.i64(); r.i64();                 // time_range
      r.str();                          // table_name
      const uint32_t series_id = r.u32();
      const uint64_t n_tags = r.u64();  // series_key.tags
      for (uint64_t k = 0; k < n_tags && r.ok; k++) { r.skip_bytes_vec(); r.skip_bytes_vec(); }
      r.str();                          // series_key.table
      r.u64();                          // next_column_group_id
      const uint64_t n_cg = r.u64();
      for (uint64_t k = 0; k < n_cg && r.ok; k++) {
        r.u64();  // map key
        r.u64();  // column_group_id
        r.u64();  // pages_offset
        r.u64();  // size
        tskv_time_range tr{r.i64(), r.i64()};
        const uint64_t n_pages = r.u64();
        // the engine's descriptor order: TIME page first, then the field pages by ascending column id
        std::vector<tskv_page_desc> field;
        std::vector<tskv_value_stats> field_stats;
        tskv_page_desc time_desc{};
        bool have_time = false;
        std::vector<std::pair<uint64_t, uint64_t>> src;  // (file offset, size) in `field` order; time first
        uint64_t time_src = 0;
        for (uint64_t p = 0; p < n_pages && r.ok; p++) {
          const uint64_t off = r.u64(), size = r.u64();
          const uint32_t num_values = r.u32();
          const Column col = read_table_column(r);
          const tskv_value_stats vst = read_value_statistics(r, r.u32());
          if (!r.ok) break;
          if (size > len || off > len - size || size >= (1ull << 32)) { g_err = "page out of bounds"; return TSKV_ERR_PAGE_FORMAT; }
          const int pt = phys_type_of(col.type);
          if (pt < 0) { skipped++; continue; }
          tskv_page_desc d{};
          d.offset = off;  // rewritten below
          d.size = (uint32_t)size;
          d.num_values = num_values;
          d.series_id = series_id;
          d.column_id = (uint16_t)col.id;
          d.phys_type = (uint8_t)pt;
          if (pt == TSKV_PT_TIME) { time_desc = d; have_time = true; time_src = off; }
          else { field.push_back(d); field_stats.push_back(vst); }
        }
        if (!r.ok) break;
        if (!have_time) { g_err = "column group without a time page (column_group.rs:67-79)"; return TSKV_ERR_PAGE_FORMAT; }
        std::vector<size_t> forder(field.size());
        for (size_t k = 0; k < forder.size(); k++) forder[k] = k;
        std::stable_sort(forder.begin(), forder.end(), [&](size_t a, size_t b) { return field[a].column_id < field[b].column_id; });
        auto put = [&](tskv_page_desc d, uint64_t file_off, const tskv_value_stats &vs) {
          const uint64_t pad = (16 - (arena.size() & 15)) & 15;
          arena.insert(arena.end(), pad, 0);
          d.offset = arena.size();
          arena.insert(arena.end(), file + file_off, file + file_off + d.size);
          descs.push_back(d);
          vstats.push_back(vs);
        };
        put(time_desc, time_src, tskv_value_stats{});
        for (size_t k : forder) put(field[k], field[k].offset, field_stats[k]);
        bounds.push_back(tr);
      }
      if (!r.ok) { g_err = "chunk does not parse"; return TSKV_ERR_PAGE_FORMAT; }
    }
    if (!g.ok) { g_err = "chunk group does not parse"; return TSKV_ERR_PAGE_FORMAT; }
  }
  out->arena_len = arena.size();
  out->arena = (uint8_t *)malloc(std::max<size_t>(arena.size(), 1));
  out->descs = (tskv_page_desc *)malloc(std::max<size_t>(descs.size(), 1) * sizeof(tskv_page_desc));
  out->cg_bounds = (tskv_time_range *)malloc(std::max<size_t>(bounds.size(), 1) * sizeof(tskv_time_range));
  out->value_stats = (tskv_value_stats *)malloc(std::max<size_t>(vstats.size(), 1) * sizeof(tskv_value_stats));
  if (!out->arena || !out->descs || !out->cg_bounds || !out->value_stats) { tskvtsm_free(out); return TSKV_ERR_OOM; }
  if (!vstats.empty()) memcpy(out->value_stats, vstats.data(), vstats.size() * sizeof(tskv_value_stats));
  if (!arena.empty()) memcpy(out->arena, arena.data(), arena.size());
  if (!descs.empty()) memcpy(out->descs, descs.data(), descs.size() * sizeof(tskv_page_desc));
  if (!bounds.empty()) memcpy(out->cg_bounds, bounds.data(), bounds.size() * sizeof(tskv_time_range));
  out->n_descs = descs.size();
  out->n_column_groups = bounds.size();
  out->n_skipped_pages = skipped;
  return TSKV_OK;
}

// ---- writer (tests): one table, one chunk per series, column groups in descriptor order. Follows TsmWriter::write_pages
// + finish (writer.rs:316-350,497-520). meta_encoding: 1 = Encoding::Null -> TsmVersion::V1, 7 = Snappy -> V2.
// column_names may be NULL (names "c<id>"). Returns the file size, or 0 (and sets the error) on bad input.
uint64_t tskvtsm_write(const uint8_t *arena, const tskv_page_desc *descs, uint64_t n_descs, const tskv_time_range *cg_bounds,
                       uint64_t n_cg, const char *table_name, uint32_t meta_encoding, uint8_t *out, uint64_t cap) {
  return tskvtsm_write_stats(arena, descs, n_descs, cg_bounds, n_cg, table_name, meta_encoding, nullptr, out, cap);
}

uint64_t tskvtsm_write_stats(const uint8_t *arena, const tskv_page_desc *descs, uint64_t n_descs, const tskv_time_range *cg_bounds,
                             uint64_t n_cg, const char *table_name, uint32_t meta_encoding, const tskv_value_stats *value_stats,
                             uint8_t *out, uint64_t cap) {
  g_err.clear();
  if (!table_name || (meta_encoding != 1 && meta_encoding != 7)) { g_err = "writer: table name / encoding"; return 0; }
  std::vector<uint8_t> file;
  const uint32_t magic = __builtin_bswap32(TSM_MAGIC);
  file.insert(file.end(), (const uint8_t *)&magic, (const uint8_t *)&magic + 4);
  struct Cg { uint64_t first, n; uint32_t series; std::vector<uint64_t> off; };
  std::vector<Cg> cgs;
  for (uint64_t i = 0; i < n_descs;) {
    if (descs[i].phys_type != TSKV_PT_TIME) { g_err = "writer: column group does not start with a time page"; return 0; }
    Cg cg{i, 1, descs[i].series_id, {}};
    while (i + cg.n < n_descs && descs[i + cg.n].phys_type != TSKV_PT_TIME) cg.n++;
    for (uint64_t k = 0; k < cg.n; k++) {  // pages back to back, like TsmWriter::write_pages
      cg.off.push_back(file.size());
      file.insert(file.end(), arena + descs[i + k].offset, arena + descs[i + k].offset + descs[i + k].size);
    }
    cgs.push_back(cg);
    i += cg.n;
  }
  if (cgs.size() != n_cg) { g_err = "writer: cg_bounds count"; return 0; }
  auto column = [&](Wr &w, const tskv_page_desc &d) {
    w.u32(d.column_id);
    w.str(d.phys_type == TSKV_PT_TIME ? "time" : "c" + std::to_string(d.column_id));
    if (d.phys_type == TSKV_PT_TIME) { w.u32(1); w.u32(3); }  // Time(Nanosecond)
    else { w.u32(2); w.u32(d.phys_type == TSKV_PT_F64 ? 1 : d.phys_type == TSKV_PT_I64 ? 2 : d.phys_type == TSKV_PT_BOOL ? 4 : 3); }
    w.u32(0);  // Encoding::Default
  };
  // chunks (one per series, series ascending like the BTreeMap), then the chunk group, then the chunk group meta
  std::map<uint32_t, std::vector<size_t>> by_series;
  for (size_t k = 0; k < cgs.size(); k++) by_series[cgs[k].series].push_back(k);
  Wr meta;
  struct Spec { uint32_t series; uint64_t off, size; tskv_time_range tr; };
  std::vector<Spec> specs;
  tskv_time_range all{INT64_MAX, INT64_MIN};
  for (auto &kv : by_series) {
    Wr c;
    tskv_time_range tr{INT64_MAX, INT64_MIN};
    for (size_t k : kv.second) { tr.min_ts = std::min(tr.min_ts, cg_bounds[k].min_ts); tr.max_ts = std::max(tr.max_ts, cg_bounds[k].max_ts); }
    all.min_ts = std::min(all.min_ts, tr.min_ts);
    all.max_ts = std::max(all.max_ts, tr.max_ts);
    c.i64(tr.min_ts); c.i64(tr.max_ts);
    c.str(table_name);
    c.u32(kv.first);
    c.u64(0);            // series_key.tags
    c.str(table_name);   // series_key.table
    c.u64(kv.second.size());  // next_column_group_id
    c.u64(kv.second.size());  // column_groups
    uint64_t id = 0;
    for (size_t k : kv.second) {
      const Cg &cg = cgs[k];
      uint64_t size = 0;
      for (uint64_t p = 0; p < cg.n; p++) size += descs[cg.first + p].size;
      c.u64(id); c.u64(id);
      c.u64(cg.off[0]); c.u64(size);
      c.i64(cg_bounds[k].min_ts); c.i64(cg_bounds[k].max_ts);
      c.u64(cg.n);
      for (uint64_t p = 0; p < cg.n; p++) {
        const tskv_page_desc &d = descs[cg.first + p];
        c.u64(cg.off[p]); c.u64(d.size);
        c.u32(d.num_values);
        column(c, d);
        c.u32(d.phys_type == TSKV_PT_F64 ? 1 : d.phys_type == TSKV_PT_U64 ? 3 : d.phys_type == TSKV_PT_BOOL ? 0 : 2);  // PageStatistics variant (time: I64)
        const tskv_value_stats *vs = value_stats ? &value_stats[cg.first + p] : nullptr;
        if (vs && (vs->flags & TSKV_STATS_MINMAX) && d.phys_type != TSKV_PT_TIME) {  // Some(min), Some(max)
          for (int k = 0; k < 2; k++) {
            c.u8(1);
            const uint64_t v = k == 0 ? vs->min : vs->max;
            if (d.phys_type == TSKV_PT_BOOL) c.u8((uint8_t)(v != 0));
            else c.u64(v);
          }
        } else {
          c.u8(0); c.u8(0);                   // min None, max None
        }
        c.u8(0); c.u64(0);                    // distinct None, null_count
      }
      id++;
    }
    specs.push_back({kv.first, meta.b.size(), c.b.size(), tr});
    meta.raw(c.b.data(), c.b.size());
  }
  const uint64_t chunk_size = meta.b.size();
  const uint64_t group_off = meta.b.size();
  {
    Wr g;
    g.u64(specs.size());
    for (const Spec &s : specs) { g.u32(s.series); g.u64(s.off); g.u64(s.size); g.i64(s.tr.min_ts); g.i64(s.tr.max_ts); }
    meta.raw(g.b.data(), g.b.size());
  }
  const uint64_t group_size = meta.b.size() - group_off;
  const uint64_t cgm_off = meta.b.size();
  {
    Wr m;
    m.u64(1);
    m.str(table_name);
    m.str("cnosdb"); m.str("public"); m.str(table_name);  // TskvTableSchema: tenant, db, name
    m.u64(0);   // schema_version
    m.u32(0);   // next_column_id
    m.u64(0);   // columns
    m.u64(0);   // columns_index
    m.u64(group_off); m.u64(group_size);
    m.i64(all.min_ts); m.i64(all.max_ts);
    m.u64(0);   // count
    meta.raw(m.b.data(), m.b.size());
  }
  const uint64_t cgm_size = meta.b.size() - cgm_off;
  const uint64_t chunk_offset = file.size();
  if (meta_encoding == 1) {
    file.insert(file.end(), meta.b.begin(), meta.b.end());
  } else {  // string codec, snappy: [7][1 << 4] snappy(uvarint(len) | bytes)
    std::vector<uint8_t> plain;
    uint64_t v = meta.b.size();
    do { uint8_t b = v & 0x7f; v >>= 7; plain.push_back(b | (v ? 0x80 : 0)); } while (v);
    plain.insert(plain.end(), meta.b.begin(), meta.b.end());
    file.push_back(7);
    file.push_back(1 << 4);
    snappy_store(plain.data(), plain.size(), file);
  }
  Wr f;
  f.u32(meta_encoding == 1 ? 0 : 1);  // TsmVersion variant
  f.i64(all.min_ts); f.i64(all.max_ts);
  f.u64(cgm_off); f.u64(cgm_size);
  f.u64(BLOOM_BYTES);
  f.b.insert(f.b.end(), BLOOM_BYTES, 0);  // (the series bloom filter is only used by TsmReader::statistics)
  f.u64(1024 * 1024 - 1);  // mask
  f.u64(chunk_offset); f.u64(chunk_size);
  if (f.b.size() != FOOTER_SIZE) { g_err = "writer: footer size"; return 0; }
  file.insert(file.end(), f.b.begin(), f.b.end());
  if (out && file.size() <= cap) memcpy(out, file.data(), file.size());
  return file.size();
}

}  // extern "C"
