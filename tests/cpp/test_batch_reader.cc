// C++ test of the host-side BatchReader mirror (cnosdb_b200/csrc/host/batch_reader.h) over the C ABI, checked
// against the CPU oracle. Reads like the reference's reader tests (tskv/src/reader/column_group/mod.rs:266-368):
// build column groups, run the reader, compare the RecordBatch.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <vector>

#include "../../cnosdb_b200/csrc/host/batch_reader.h"
#include "../../cnosdb_b200/csrc/host/tsm_writer.h"
#include "../../oracle/tskv_oracle.h"

using namespace tskv::reader;

#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main() {
  // ---- 40 series x (time + i64 field 1 + f64 field 2), 500 rows, 10 s spacing ----------------------
  const int n_series = 40, n = 500;
  const int64_t t0 = 1640995200000000000ll, step = 10000000000ll, w = 60000000000ll;
  tskv::Bytes arena;
  std::vector<ColumnGroup> cgs;
  std::vector<tskv_page_desc> descs;
  uint64_t seed = 7;
  auto rnd = [&]() { seed = seed * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(seed >> 33); };
  for (int s = 0; s < n_series; s++) {
    std::vector<int64_t> ts(n), iv(n);
    std::vector<double> fv(n);
    int64_t v = rnd() % 100;
    for (int i = 0; i < n; i++) {
      ts[i] = t0 + i * step + (s % 3 == 0 ? (int64_t)(rnd() % 1000) : 0);
      v += (int64_t)(rnd() % 7) - 3;
      iv[i] = v;
      fv[i] = v * 0.5 + (rnd() % 1000) / 1024.0;
    }
    ColumnGroup cg(s, (SeriesId)(100 + s));
    cg.time_range_merge(TimeRange{ts.front(), ts.back()});
    auto add = [&](const tskv::Bytes &data, ColumnId id, const char *name, PhysicalDType pt) {
      while (arena.size() & 15) arena.push_back(0);
      uint64_t off = arena.size();
      tskv::append_page(nullptr, n, data, arena);
      cg.push(PageWriteSpec{off, arena.size() - off, PageMeta{(uint32_t)n, TableColumn{id, name, pt}}});
      descs.push_back(tskv_page_desc{off, (uint32_t)(arena.size() - off), (uint32_t)n, (uint32_t)(100 + s), id, (uint8_t)pt, 0});
    };
    tskv::Bytes d;
    tskv::encode_timestamps(ts.data(), n, d); add(d, 0, "time", PhysicalDType::Time);
    d.clear(); tskv::encode_integers(iv.data(), n, d); add(d, 1, "usage_user", PhysicalDType::Integer);
    d.clear(); tskv::encode_floats(fv.data(), n, d); add(d, 2, "usage_idle", PhysicalDType::Float);
    cgs.push_back(cg);
  }
  // ---- the pushed-down scan: WHERE time BETWEEN .. AND series IN (..) GROUP BY time_window(1m) -------
  QueryOption opt;
  opt.time_ranges = {TimeRange{t0 + 95 * step, t0 + 400 * step}};
  std::vector<SeriesId> sel;
  for (int s = 0; s < n_series; s += 2) sel.push_back(100 + s);
  opt.series_ids = sel;
  opt.bucket = TimeBucket{0, w};
  opt.table_columns = {TableColumn{1, "usage_user", PhysicalDType::Integer}, TableColumn{2, "usage_idle", PhysicalDType::Float}};
  opt.aggregates = {{AggregateKind::Mean, 1}, {AggregateKind::Max, 1}, {AggregateKind::Count, 2}, {AggregateKind::Sum, 2},
                    {AggregateKind::First, 2}, {AggregateKind::Last, 1}};
  auto eng = GpuEngine::create(0);
  CHECK(eng.ok());
  GpuAggregateBatchReader reader(eng.value, arena.data(), arena.size(), cgs, opt);
  std::ostringstream desc;
  reader.fmt_as(desc);
  CHECK(desc.str().find("GpuAggregateBatchReader") == 0 && reader.children().empty());
  auto res = reader.process();
  CHECK(res.ok());
  CHECK(res.value.size() == 1);
  const RecordBatch &b = res.value[0];
  CHECK(b.columns.size() == 1 + opt.aggregates.size());
  CHECK(b.columns[0].name == "time" && b.columns[1].name == "mean(usage_user)" && b.columns[5].name == "first(usage_idle)");
  CHECK(reader.metrics().page_read_count == (uint64_t)(n_series / 2) * 3);

  // ---- oracle on the same arena / equivalent tskv_query ------------------------------------------------
  tskv_agg_column qcols[2] = {{1, TSKV_PT_I64, (uint8_t)(TSKV_AGG_MEAN | TSKV_AGG_MAX | TSKV_AGG_LAST)},
                              {2, TSKV_PT_F64, (uint8_t)(TSKV_AGG_COUNT | TSKV_AGG_SUM | TSKV_AGG_FIRST)}};
  tskv_time_range tr{opt.time_ranges[0].min_ts, opt.time_ranges[0].max_ts};
  tskv_query q{};
  q.series_ids = sel.data(); q.n_series = (uint32_t)sel.size();
  q.time_ranges = &tr; q.n_time_ranges = 1;
  q.origin = 0; q.width = w;
  q.first_bucket_start = (int64_t)b.columns[0].values[0];
  q.n_buckets = (uint32_t)b.num_rows;
  q.columns = qcols; q.n_columns = 2;
  tskv_output_layout L{};
  CHECK(orc_query_output_layout(descs.data(), descs.size(), &q, &L) == TSKV_OK);
  std::vector<uint64_t> ov(L.n_out * L.n_cells);
  std::vector<uint8_t> ob(L.validity_bytes);
  CHECK(orc_scan_aggregate(arena.data(), arena.size(), descs.data(), descs.size(), &q, 1, 1, ov.data(), ob.data(), nullptr) == TSKV_OK);
  // oracle output order: col1 {max, mean, last}, col2 {count, sum, first}
  struct Map { int batch_col; int oracle_col; bool approx; } maps[] = {{1, 1, true}, {2, 0, false}, {3, 3, false}, {4, 4, true}, {5, 5, false}, {6, 2, false}};
  for (const Map &m : maps) {
    const ArrayData &c = b.columns[m.batch_col];
    for (size_t i = 0; i < b.num_rows; i++) {
      bool ev = (ob[m.oracle_col * L.bitmap_stride + (i >> 3)] >> (i & 7)) & 1;
      CHECK(c.is_valid(i) == ev);
      if (!ev) continue;
      uint64_t e = ov[m.oracle_col * L.n_cells + i], g = c.values[i];
      if (m.approx) {
        double ed, gd;
        memcpy(&ed, &e, 8); memcpy(&gd, &g, 8);
        CHECK(std::fabs(ed - gd) <= 1e-6 * std::fabs(ed));
      } else {
        CHECK(e == g);
      }
    }
  }
  // ---- with a TsmTombstone attached (ColumnGroupReader's `tomb`, decode_pages reader.rs:507-551) -----------
  std::vector<tskv_tombstone> tombs = {{100, 1, t0 + 100 * step, t0 + 180 * step},             // (series 100, usage_user) reads NULL
                                       {102, TSKV_TOMB_ALL, t0 + 200 * step, t0 + 260 * step},  // rows of series 102 dropped
                                       {TSKV_TOMB_ALL, TSKV_TOMB_ALL, t0 + 300 * step, t0 + 310 * step}};
  GpuAggregateBatchReader tomb_reader(eng.value, arena.data(), arena.size(), cgs, opt);
  tomb_reader.set_tombstones(tombs);
  auto tres = tomb_reader.process();
  CHECK(tres.ok() && tres.value.size() == 1 && tres.value[0].num_rows == b.num_rows);
  std::vector<uint64_t> tv(L.n_out * L.n_cells);
  std::vector<uint8_t> tb(L.validity_bytes);
  CHECK(orc_scan_aggregate_tomb(arena.data(), arena.size(), descs.data(), descs.size(), &q, tombs.data(), tombs.size(), 1, 1,
                                tv.data(), tb.data(), nullptr) == TSKV_OK);
  bool differs = false;
  for (const Map &m : maps) {
    const ArrayData &c = tres.value[0].columns[m.batch_col];
    for (size_t i = 0; i < b.num_rows; i++) {
      bool ev = (tb[m.oracle_col * L.bitmap_stride + (i >> 3)] >> (i & 7)) & 1;
      CHECK(c.is_valid(i) == ev);
      if (!ev) continue;
      uint64_t e = tv[m.oracle_col * L.n_cells + i], g = c.values[i];
      differs |= e != ov[m.oracle_col * L.n_cells + i];
      if (m.approx) {
        double ed, gd;
        memcpy(&ed, &e, 8); memcpy(&gd, &g, 8);
        CHECK(std::fabs(ed - gd) <= 1e-6 * std::fabs(ed));
      } else {
        CHECK(e == g);
      }
    }
  }
  CHECK(differs);  // the tombstones did change the answer
  // ---- statistics pruning (filter_column_groups, reader/chunk.rs:12-50): a query range that misses a column group's
  //      time range keeps it from being read at all -----------------------------------------------------------------
  {
    QueryOption late = opt;
    late.time_ranges = {TimeRange{t0 + 600 * step, t0 + 900 * step}};  // after every row
    GpuAggregateBatchReader none_reader(eng.value, arena.data(), arena.size(), cgs, late);
    auto nres = none_reader.process();
    CHECK(nres.ok() && nres.value.empty());  // an empty stream; the library is not even called
    CHECK(none_reader.pruned_column_groups() == (uint64_t)n_series && none_reader.metrics().page_read_count == 0);
    CHECK(reader.pruned_column_groups() == 0);
  }
  // ---- overlapping chunks (DataMerger, reader/merge.rs): a "delta file" (file id 2) rewrites rows 100..199 of the first
  //      selected series and adds 50 rows after its end; the merged scan must equal the oracle's merge ------------------
  {
    tskv::Bytes arena2 = arena;
    std::vector<ColumnGroup> cgs2 = cgs;
    std::vector<tskv_page_desc> descs2 = descs;
    std::vector<uint64_t> files(cgs.size(), 1);
    const int m = 150;
    std::vector<int64_t> ts(m), iv(m);
    for (int i = 0; i < m; i++) {
      ts[i] = t0 + (i < 100 ? 100 + i : 500 + (i - 100)) * step;  // rows 100..199 get series 100's exact (jittered) times below
      iv[i] = 1000000 + i;
    }
    // series 100 is s = 0, which has jittered timestamps: take the exact timestamps of its rows 100..199 from the page we wrote
    {
      uint64_t sd = 7;
      auto r2 = [&]() { sd = sd * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(sd >> 33); };
      (void)r2();  // v
      for (int i = 0; i < n; i++) {
        const int64_t t = t0 + i * step + (int64_t)(r2() % 1000);
        (void)r2(); (void)r2();
        if (i >= 100 && i < 200) ts[i - 100] = t;
      }
    }
    ColumnGroup cg((uint64_t)cgs.size(), (SeriesId)100);
    cg.set_file_id(2);
    cg.time_range_merge(TimeRange{ts.front(), ts.back()});
    auto add2 = [&](const tskv::Bytes &data, ColumnId id, const char *name, PhysicalDType pt) {
      while (arena2.size() & 15) arena2.push_back(0);
      uint64_t off = arena2.size();
      tskv::append_page(nullptr, m, data, arena2);
      cg.push(PageWriteSpec{off, arena2.size() - off, PageMeta{(uint32_t)m, TableColumn{id, name, pt}}});
      descs2.push_back(tskv_page_desc{off, (uint32_t)(arena2.size() - off), (uint32_t)m, 100u, id, (uint8_t)pt, 0});
    };
    tskv::Bytes d;
    tskv::encode_timestamps(ts.data(), m, d); add2(d, 0, "time", PhysicalDType::Time);
    d.clear(); tskv::encode_integers(iv.data(), m, d); add2(d, 1, "usage_user", PhysicalDType::Integer);
    for (ColumnGroup &c : cgs2) c.set_file_id(1);
    cgs2.push_back(cg);
    files.push_back(2);
    QueryOption mo = opt;
    mo.time_ranges.clear();
    mo.aggregates = {{AggregateKind::Count, 1}, {AggregateKind::Sum, 1}, {AggregateKind::Max, 1}, {AggregateKind::Count, 2}};
    GpuAggregateBatchReader mreader(eng.value, arena2.data(), arena2.size(), cgs2, mo);
    auto mres = mreader.process();
    CHECK(mres.ok() && mres.value.size() == 1);
    const RecordBatch &mb = mres.value[0];
    tskv_agg_column mcols[2] = {{1, TSKV_PT_I64, (uint8_t)(TSKV_AGG_COUNT | TSKV_AGG_SUM | TSKV_AGG_MAX)}, {2, TSKV_PT_F64, (uint8_t)TSKV_AGG_COUNT}};
    tskv_query mq{};
    mq.series_ids = sel.data(); mq.n_series = (uint32_t)sel.size();
    mq.origin = 0; mq.width = w;
    mq.first_bucket_start = (int64_t)mb.columns[0].values[0];
    mq.n_buckets = (uint32_t)mb.num_rows;
    mq.columns = mcols; mq.n_columns = 2;
    tskv_output_layout ML{};
    CHECK(orc_query_output_layout(descs2.data(), descs2.size(), &mq, &ML) == TSKV_OK);
    std::vector<uint64_t> mv(ML.n_out * ML.n_cells);
    std::vector<uint8_t> mbm(ML.validity_bytes);
    orc_handle *H = nullptr;
    CHECK(orc_open(arena2.data(), arena2.size(), descs2.data(), descs2.size(), 1, &H) == TSKV_OK);
    CHECK(orc_set_chunk_files(H, files.data(), files.size()) == TSKV_OK);
    CHECK(orc_scan(H, &mq, nullptr, 0, 1, mv.data(), mbm.data(), nullptr) == TSKV_OK);
    orc_close(H);
    // oracle output order: col1 {count, sum, max}, col2 {count}; batch: time, count(1), sum(1), max(1), count(2)
    uint64_t total_rows = 0;
    for (int k = 0; k < 4; k++)
      for (uint64_t i = 0; i < ML.n_cells; i++) {
        CHECK(mb.columns[1 + k].is_valid(i) == (bool)((mbm[k * ML.bitmap_stride + (i >> 3)] >> (i & 7)) & 1));
        if (mb.columns[1 + k].is_valid(i)) CHECK(mb.columns[1 + k].values[i] == mv[k * ML.n_cells + i]);
        if (k == 0) total_rows += mb.columns[1].values[i];
      }
    CHECK(total_rows == (uint64_t)(n_series / 2) * n + 50);  // 100 rewritten rows collapse, 50 are new
  }
  // ---- error behaviour: a corrupted page surfaces TsmPageFileHashCheckFailed, not a crash ----------------
  tskv::Bytes bad = arena;
  bad[cgs[0].pages()[1].offset + cgs[0].pages()[1].size - 1] ^= 0x10;
  GpuAggregateBatchReader bad_reader(eng.value, bad.data(), bad.size(), cgs, opt);
  auto bad_res = bad_reader.process();
  CHECK(!bad_res.ok() && bad_res.error.status == TSKV_ERR_CRC_MISMATCH && bad_res.error.page == 1);
  printf("test_batch_reader ok: %zu buckets, %llu pages read\n", b.num_rows, (unsigned long long)reader.metrics().page_read_count);
  return 0;
}
