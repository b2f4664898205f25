// batch_reader.cc — see batch_reader.h. Pure host-side marshalling over the C ABI; no compute here.
#include "batch_reader.h"

#include <algorithm>
#include <functional>
#include <map>

namespace tskv {
namespace reader {

namespace {
// dense results above this many buckets are refused (2^24 buckets x 8 bytes per output column)
constexpr uint64_t kMaxBuckets = 1ull << 24;
const char *agg_name(AggregateKind k) {
  switch (k) {
    case AggregateKind::Count: return "count";
    case AggregateKind::Sum: return "sum";
    case AggregateKind::Min: return "min";
    case AggregateKind::Max: return "max";
    case AggregateKind::Mean: return "mean";
    case AggregateKind::First: return "first";
    case AggregateKind::Last: return "last";
  }
  return "?";
}
// `sliding_window(t, w, w, origin, 0)` start (time_window.rs:184-198), used to size the dense bucket range.
int64_t window_start(int64_t t, int64_t w, int64_t origin) {
  int64_t o = origin % w;
  int64_t dividend = (int64_t)((uint64_t)t - (uint64_t)o + (uint64_t)w);
  return (int64_t)((uint64_t)t - (uint64_t)(dividend % w));
}
}  // namespace

TskvResult<std::shared_ptr<GpuEngine>> GpuEngine::create(int device) {
  TskvResult<std::shared_ptr<GpuEngine>> r;
  tskv_ctx *ctx = nullptr;
  tskv_status st = tskvgpu_ctx_create(device, &ctx);
  if (st != TSKV_OK) {
    r.error = {st, "tskvgpu_ctx_create failed: no usable CUDA device", -1};
    return r;
  }
  r.value = std::shared_ptr<GpuEngine>(new GpuEngine(ctx));
  return r;
}
GpuEngine::~GpuEngine() { tskvgpu_ctx_destroy(ctx_); }
TskvError GpuEngine::last_error(tskv_status st) const {
  return {st, tskvgpu_last_error(ctx_), tskvgpu_last_error_page(ctx_)};
}

GpuAggregateBatchReader::GpuAggregateBatchReader(std::shared_ptr<GpuEngine> engine, const uint8_t *arena,
                                                 uint64_t arena_len, std::vector<ColumnGroup> column_groups,
                                                 QueryOption option, bool verify_crc)
    : engine_(std::move(engine)), arena_(arena), arena_len_(arena_len), column_groups_(std::move(column_groups)),
      option_(std::move(option)), verify_crc_(verify_crc) {}

void GpuAggregateBatchReader::fmt_as(std::ostream &f) const {
  f << "GpuAggregateBatchReader: column_groups=" << column_groups_.size() << ", aggregates=[";
  for (size_t i = 0; i < option_.aggregates.size(); i++)
    f << (i ? ", " : "") << agg_name(option_.aggregates[i].kind) << "(" << option_.aggregates[i].column << ")";
  f << "]";
  if (option_.bucket) f << ", bucket=" << option_.bucket->width;
}

TskvResult<SendableTskvRecordBatchStream> GpuAggregateBatchReader::process() {
  TskvResult<SendableTskvRecordBatchStream> out;
  // ---- descriptor table: column group by column group, time page first (column_group.rs:9-17) --------
  std::vector<tskv_page_desc> descs;
  std::vector<uint64_t> cg_files;  // file id of every column group handed to the engine
  TimeRange data_range = TimeRange::none();
  pruned_column_groups_ = 0;
  for (const ColumnGroup &cg : column_groups_) {
    // filter_column_groups (reader/chunk.rs:12-50) for the time predicate: a column group whose time range
    // (PageMeta statistics) misses every query range is never read
    if (!option_.time_ranges.empty() &&
        std::none_of(option_.time_ranges.begin(), option_.time_ranges.end(), [&](const TimeRange &r) { return cg.time_range().overlaps(r); })) {
      pruned_column_groups_++;
      continue;
    }
    data_range.merge(cg.time_range());
    cg_files.push_back(cg.file_id());
    for (const PageWriteSpec &p : cg.pages()) {
      tskv_page_desc d{};
      d.offset = p.offset;
      d.size = (uint32_t)p.size;
      d.num_values = p.meta.num_values;
      d.series_id = cg.series_id();
      d.column_id = p.meta.column.id;
      d.phys_type = (uint8_t)p.meta.column.column_type;
      descs.push_back(d);
    }
  }
  if (descs.empty()) {  // nothing left to read: an empty stream, like a reader tree without chunks
    counters_ = tskv_counters{};
    return out;
  }
  // ---- query: one tskv_agg_column per referenced column, aggregates OR-ed into its mask ---------------
  std::vector<tskv_agg_column> cols;
  for (const PushedAggregateFunction &a : option_.aggregates) {
    auto it = std::find_if(cols.begin(), cols.end(), [&](const tskv_agg_column &c) { return c.column_id == a.column; });
    if (it == cols.end()) {
      auto tc = std::find_if(option_.table_columns.begin(), option_.table_columns.end(),
                             [&](const TableColumn &c) { return c.id == a.column; });
      if (tc == option_.table_columns.end()) {
        out.error = {TSKV_ERR_INVALID_ARG, "aggregate references a column missing from table_columns", -1};
        return out;
      }
      cols.push_back(tskv_agg_column{a.column, (uint8_t)tc->column_type, 0});
      it = cols.end() - 1;
    }
    it->agg_mask |= (uint8_t)a.kind;
  }
  std::vector<tskv_time_range> ranges;
  TimeRange scan_range = data_range;  // PageMeta statistics bound the rows that can appear
  if (!option_.time_ranges.empty()) {
    TimeRange q = TimeRange::none();
    for (const TimeRange &r : option_.time_ranges) {
      ranges.push_back({r.min_ts, r.max_ts});
      q.merge(r);
    }
    scan_range.min_ts = std::max(scan_range.min_ts, q.min_ts);
    scan_range.max_ts = std::min(scan_range.max_ts, q.max_ts);
  }
  tskv_query q{};
  if (option_.series_ids) {
    q.series_ids = option_.series_ids->data();
    q.n_series = (uint32_t)option_.series_ids->size();
  }
  q.time_ranges = ranges.data();
  q.n_time_ranges = (uint32_t)ranges.size();
  q.n_buckets = 1;
  if (scan_range.min_ts > scan_range.max_ts) return out;  // the query ranges miss the data: an empty stream, like the pruned path
  if (option_.bucket && option_.bucket->width > 0) {
    q.origin = option_.bucket->origin;
    q.width = option_.bucket->width;
    q.first_bucket_start = window_start(scan_range.min_ts, q.width, q.origin);
    const int64_t last = window_start(scan_range.max_ts, q.width, q.origin);
    // bucket count in 128-bit arithmetic: (last - first) can exceed the i64 range, the count the u32 of the ABI
    const __int128 n = ((__int128)last - (__int128)q.first_bucket_start) / q.width + 1;
    if (n < 1 || n > (__int128)kMaxBuckets) {
      out.error = {TSKV_ERR_INVALID_ARG, "bucket expression yields too many buckets for a dense result", -1};
      return out;
    }
    q.n_buckets = (uint32_t)n;
  }
  q.group_by_series = option_.group_by_series ? 1 : 0;
  q.columns = cols.data();
  q.n_columns = (uint32_t)cols.size();

  // ---- C ABI: upload (read_adjacent_pages + crc_validation), scan, results ---------------------------
  tskv_ctx *ctx = engine_->ctx();
  tskv_pages *pages = nullptr;
  tskv_status st = tskvgpu_upload_pages(ctx, arena_, arena_len_, descs.data(), descs.size(),
                                        verify_crc_ ? TSKV_UPLOAD_VERIFY_CRC : 0, &pages);
  if (st != TSKV_OK) {
    out.error = engine_->last_error(st);
    return out;
  }
  if (!tombstones_.empty()) {
    st = tskvgpu_pages_set_tombstones(ctx, pages, tombstones_.data(), tombstones_.size());
    if (st != TSKV_OK) {
      out.error = engine_->last_error(st);
      tskvgpu_pages_destroy(ctx, pages);
      return out;
    }
  }
  // chunks of several files (build_series_reader, reader/iterator.rs:463-560): overlapping ones are merged on the device
  if (std::adjacent_find(cg_files.begin(), cg_files.end(), std::not_equal_to<uint64_t>()) != cg_files.end()) {
    st = tskvgpu_pages_set_chunk_files(ctx, pages, cg_files.data(), cg_files.size());
    if (st != TSKV_OK) {
      out.error = engine_->last_error(st);
      tskvgpu_pages_destroy(ctx, pages);
      return out;
    }
  }
  tskv_output_layout L{};
  st = tskvgpu_query_output_layout(pages, &q, &L);
  std::vector<uint64_t> values(L.n_out * L.n_cells);
  std::vector<uint8_t> validity(L.validity_bytes);
  if (st == TSKV_OK) st = tskvgpu_scan_aggregate(ctx, pages, &q, values.data(), validity.data());
  if (st != TSKV_OK) {
    out.error = st == TSKV_ERR_INVALID_ARG && std::string(tskvgpu_last_error(ctx)).empty()
                    ? TskvError{st, "invalid query", -1}
                    : engine_->last_error(st);
    tskvgpu_pages_destroy(ctx, pages);
    return out;
  }
  tskvgpu_get_counters(ctx, &counters_);
  tskvgpu_pages_destroy(ctx, pages);

  // ---- wrap as one RecordBatch -------------------------------------------------------------------------
  RecordBatch batch;
  batch.num_rows = (size_t)L.n_cells;
  const size_t bm = (size_t)(L.n_cells + 7) / 8;
  if (option_.group_by_series) {
    ArrayData slot{"series_slot", PhysicalDType::Unsigned, {}, std::vector<uint8_t>(bm, 0xff)};
    for (uint64_t c = 0; c < L.n_cells; c++) slot.values.push_back(c / q.n_buckets);
    batch.columns.push_back(std::move(slot));
  }
  if (q.width > 0) {
    ArrayData t{"time", PhysicalDType::Time, {}, std::vector<uint8_t>(bm, 0xff)};
    for (uint64_t c = 0; c < L.n_cells; c++) t.values.push_back((uint64_t)(q.first_bucket_start + (int64_t)(c % q.n_buckets) * q.width));
    batch.columns.push_back(std::move(t));
  }
  // library order: query columns in order, inside a column the set agg bits ascending
  std::map<std::pair<ColumnId, uint8_t>, uint64_t> out_index;
  uint64_t j = 0;
  for (const tskv_agg_column &c : cols)
    for (unsigned bit = 0; bit < 7; bit++)
      if (c.agg_mask & (1u << bit)) out_index[{c.column_id, (uint8_t)(1u << bit)}] = j++;
  for (const PushedAggregateFunction &a : option_.aggregates) {
    const uint64_t k = out_index[{a.column, (uint8_t)a.kind}];
    auto tc = std::find_if(option_.table_columns.begin(), option_.table_columns.end(),
                           [&](const TableColumn &c) { return c.id == a.column; });
    ArrayData col;
    col.name = std::string(agg_name(a.kind)) + "(" + tc->name + ")";
    col.type = a.kind == AggregateKind::Mean ? PhysicalDType::Float
               : a.kind == AggregateKind::Count ? PhysicalDType::Unsigned : tc->column_type;
    col.values.assign(values.begin() + k * L.n_cells, values.begin() + (k + 1) * L.n_cells);
    col.validity.assign(validity.begin() + k * L.bitmap_stride, validity.begin() + k * L.bitmap_stride + bm);
    batch.columns.push_back(std::move(col));
  }
  out.value.push_back(std::move(batch));
  return out;
}

}  // namespace reader
}  // namespace tskv
