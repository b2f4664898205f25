// batch_reader.h — C++ host-side mirror of the reference's reader interface for the scan path, on top of
// the C ABI (include/tskv_gpu.h). The reference is Rust; this image has no Rust toolchain, so the host
// side is written in C++ with the reference's names, argument meaning and error behaviour
// (paths relative to the reference tree):
//   BatchReader / BatchReaderRef            tskv/src/reader/mod.rs:159-164
//   QueryOption, PushedAggregateFunction    tskv/src/reader/iterator.rs:713-741, models predicate/domain.rs:1840-1843
//   TimeRange (closed interval)             common/models/src/predicate/domain.rs:35-98
//   ColumnGroup, PageWriteSpec, PageMeta    tskv/src/tsm/column_group.rs:9-80, tskv/src/tsm/page.rs:599-620
//   TskvError::Decode & friends             tskv/src/error.rs:293-299
// A scan is one vnode-level call: the factory (SeriesGroupBatchReaderFactory::create, iterator.rs:123-264)
// would return a GpuAggregateBatchReader instead of the per-series reader tree.
#pragma once
#include <cstdint>
#include <memory>
#include <optional>
#include <ostream>
#include <string>
#include <vector>

#include "../../../include/tskv_gpu.h"

namespace tskv {
namespace reader {

using SeriesId = uint32_t;   // common/models/src/lib.rs:40
using ColumnId = uint16_t;

struct TimeRange {  // closed interval
  int64_t min_ts, max_ts;
  static TimeRange all() { return {INT64_MIN, INT64_MAX}; }
  static TimeRange none() { return {INT64_MAX, INT64_MIN}; }
  bool contains(int64_t t) const { return t >= min_ts && t <= max_ts; }
  bool overlaps(const TimeRange &o) const { return !(min_ts > o.max_ts || max_ts < o.min_ts); }
  void merge(const TimeRange &o) {
    min_ts = min_ts < o.min_ts ? min_ts : o.min_ts;
    max_ts = max_ts > o.max_ts ? max_ts : o.max_ts;
  }
};

enum class PhysicalDType : uint8_t { Time = TSKV_PT_TIME, Integer = TSKV_PT_I64, Unsigned = TSKV_PT_U64, Float = TSKV_PT_F64, Boolean = TSKV_PT_BOOL };

struct TableColumn {
  ColumnId id;
  std::string name;
  PhysicalDType column_type;
};
struct PageMeta {
  uint32_t num_values;
  TableColumn column;
};
struct PageWriteSpec {
  uint64_t offset;  // into the arena the column group was read into
  uint64_t size;
  PageMeta meta;
};

class ColumnGroup {
 public:
  ColumnGroup(uint64_t id, SeriesId series) : column_group_id_(id), series_id_(series), time_range_(TimeRange::none()) {}
  uint64_t column_group_id() const { return column_group_id_; }
  SeriesId series_id() const { return series_id_; }
  const TimeRange &time_range() const { return time_range_; }
  void time_range_merge(const TimeRange &tr) { time_range_.merge(tr); }
  const std::vector<PageWriteSpec> &pages() const { return pages_; }
  void push(PageWriteSpec page) { pages_.push_back(std::move(page)); }  // time page first, then fields by column id
  size_t row_len() const { return pages_.empty() ? 0 : pages_.front().meta.num_values; }
  // ColumnFile::file_id() of the file (or memcache) this group's chunk lives in. Column groups of one series with
  // different file ids are chunks of different files: overlapping ones are merged (DataMerger, reader/merge.rs).
  uint64_t file_id() const { return file_id_; }
  void set_file_id(uint64_t id) { file_id_ = id; }

 private:
  uint64_t file_id_ = 0;
  uint64_t column_group_id_;
  SeriesId series_id_;
  TimeRange time_range_;
  std::vector<PageWriteSpec> pages_;
};

// Extends PushedAggregateFunction (today only Count(col)) with the aggregates DataFusion runs above the scan.
enum class AggregateKind : uint8_t { Count = TSKV_AGG_COUNT, Sum = TSKV_AGG_SUM, Min = TSKV_AGG_MIN, Max = TSKV_AGG_MAX,
                                     Mean = TSKV_AGG_MEAN, First = TSKV_AGG_FIRST, Last = TSKV_AGG_LAST };
struct PushedAggregateFunction {
  AggregateKind kind;
  ColumnId column;
};
// time_window(time, width) / date_bin(width, time, origin) pushed below the aggregate.
struct TimeBucket {
  int64_t origin;
  int64_t width;
};

struct QueryOption {
  std::vector<TimeRange> time_ranges;             // split.time_ranges(); empty => all
  std::optional<std::vector<SeriesId>> series_ids;  // get_series_id_by_filter result (sorted); nullopt => all
  std::vector<PushedAggregateFunction> aggregates;
  std::optional<TimeBucket> bucket;               // nullopt => ungrouped in time
  bool group_by_series = false;
  std::vector<TableColumn> table_columns;         // the fields referenced by `aggregates`
};

// Error: mirrors TskvError (status = tskv_status; Decode errors carry the page index).
struct TskvError {
  tskv_status status = TSKV_OK;
  std::string reason;
  int64_t page = -1;
  bool ok() const { return status == TSKV_OK; }
};
template <typename T>
struct TskvResult {
  T value{};
  TskvError error;
  bool ok() const { return error.ok(); }
};

// Minimal Arrow-shaped batch: 8-byte value buffers + LSB-first validity bitmaps, zero-copy wrappable.
struct ArrayData {
  std::string name;        // "time" | "<agg>(<column>)" | "series_slot"
  PhysicalDType type;      // Float for mean
  std::vector<uint64_t> values;
  std::vector<uint8_t> validity;  // ceil(rows / 8) bytes
  bool is_valid(size_t i) const { return (validity[i >> 3] >> (i & 7)) & 1; }
};
struct RecordBatch {
  size_t num_rows = 0;
  std::vector<ArrayData> columns;
};
using SendableTskvRecordBatchStream = std::vector<RecordBatch>;  // the stream is materialised: one batch per scan

class BatchReader;
using BatchReaderRef = std::shared_ptr<BatchReader>;
class BatchReader {
 public:
  virtual ~BatchReader() = default;
  virtual TskvResult<SendableTskvRecordBatchStream> process() = 0;
  virtual void fmt_as(std::ostream &f) const = 0;
  virtual std::vector<BatchReaderRef> children() const = 0;
};

// One CUDA device + stream (tskv_ctx). Thread-safe; share one per device.
class GpuEngine {
 public:
  static TskvResult<std::shared_ptr<GpuEngine>> create(int device);
  ~GpuEngine();
  tskv_ctx *ctx() const { return ctx_; }
  TskvError last_error(tskv_status st) const;

 private:
  explicit GpuEngine(tskv_ctx *c) : ctx_(c) {}
  tskv_ctx *ctx_;
};

// Replaces the ColumnGroupReader / DataFilter / SeriesReader tree + the DataFusion aggregate above it.
// `arena` holds the raw page bytes of `column_groups` (what read_adjacent_pages would return), page offsets
// 16-byte aligned. Output: one RecordBatch with a "time" column (bucket start; omitted when ungrouped in time),
// "series_slot" when group_by_series, then one column per pushed aggregate in request order.
class GpuAggregateBatchReader : public BatchReader {
 public:
  GpuAggregateBatchReader(std::shared_ptr<GpuEngine> engine, const uint8_t *arena, uint64_t arena_len,
                          std::vector<ColumnGroup> column_groups, QueryOption option, bool verify_crc = true);
  TskvResult<SendableTskvRecordBatchStream> process() override;
  void fmt_as(std::ostream &f) const override;
  std::vector<BatchReaderRef> children() const override { return {}; }
  const tskv_counters &metrics() const { return counters_; }  // page_read_count/bytes, elapsed_* (column_group/mod.rs:141-193)
  // The TsmTombstone of the file(s) behind the arena (ColumnGroupReader carries `tomb`, column_group/mod.rs:25-45;
  // applied by decode_pages, tsm/reader.rs:507-551). Entries as in include/tskv_gpu.h.
  void set_tombstones(std::vector<tskv_tombstone> tombs) { tombstones_ = std::move(tombs); }
  // Column groups skipped by the statistics pruning of the last process() (filter_column_groups, reader/chunk.rs:12-50).
  uint64_t pruned_column_groups() const { return pruned_column_groups_; }

 private:
  std::shared_ptr<GpuEngine> engine_;
  const uint8_t *arena_;
  uint64_t arena_len_;
  std::vector<ColumnGroup> column_groups_;
  QueryOption option_;
  bool verify_crc_;
  std::vector<tskv_tombstone> tombstones_;
  tskv_counters counters_{};
  uint64_t pruned_column_groups_ = 0;
};

}  // namespace reader
}  // namespace tskv
