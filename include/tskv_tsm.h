/* tskv_tsm.h — C ABI of the TSM file loader: a `.tsm` file image -> the page arena,
 * descriptor table and per-column-group time bounds that tskvgpu_upload_pages / tskvgpu_pages_set_time_bounds take.
 * Implemented by cnosdb_b200/csrc/host/tsm_file.cc (libtskv_hostgen.so; pure host code, no CUDA).
 *
 * Replaces, on the reference side (paths relative to the CnosDB tree):
 *   TsmReader::open                 tskv/src/tsm/reader.rs:120-168   footer -> metadata -> chunk groups -> chunks
 *   read_footer / read_chunk_*      tskv/src/tsm/reader.rs:399-474
 *   ColumnGroup::time_range()       tskv/src/tsm/column_group.rs:9-17 (reported per column group for statistics pruning)
 *   PageMeta.statistics             tskv/src/tsm/page.rs:599-613 (reported per page for value-statistics pruning)
 * A host that already has a TsmReader (the Rust shim of INTEGRATION.md) does not need this: it hands the engine the
 * PageWriteSpec list it holds. The loader is for hosts that start from the file.
 */
#ifndef TSKV_TSM_H
#define TSKV_TSM_H

#include "tskv_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tskvtsm_result {
  uint8_t *arena;             /* repacked pages, each 16-byte aligned (malloc-ed; release with tskvtsm_free) */
  uint64_t arena_len;
  tskv_page_desc *descs;      /* column group by column group, TIME page first */
  uint64_t n_descs;
  tskv_time_range *cg_bounds; /* ColumnGroup::time_range() per column group, in descriptor order */
  uint64_t n_column_groups;
  tskv_value_stats *value_stats; /* PageMeta.statistics per descriptor (min / max of I64 / U64 / F64 / Bool pages, when the
                                    file holds them): what tskvgpu_pages_set_value_stats takes */
  uint64_t n_skipped_pages;   /* pages of column types outside this engine's path (tag / bool / string / geometry) */
  int64_t min_ts, max_ts;     /* Footer.time_range */
  uint32_t version;           /* TsmVersion: 1 | 2 */
  uint32_t reserved;
} tskvtsm_result;

/* Parses a whole file image. `table` (NULL or "" = every table) restricts the result to one table's chunk group.
 * TSKV_ERR_PAGE_FORMAT for a truncated / inconsistent file (what the reference reports as TsmFileBroken /
 * ReadTsm errors, tsm/reader.rs:400-404), TSKV_ERR_UNSUPPORTED for V2 metadata in a codec this build cannot decode
 * (zstd / bzip). On error `out` is zeroed and tskvtsm_last_error() describes the failure. */
tskv_status tskvtsm_load(const uint8_t *file, uint64_t len, const char *table, tskvtsm_result *out);

void tskvtsm_free(tskvtsm_result *r);

/* Thread-local description of the last failure of this module. */
const char *tskvtsm_last_error(void);

/* The inverse, for tests and tooling: one table, one chunk per series, column groups in descriptor order (follows
 * TsmWriter::write_pages + finish, tskv/src/tsm/writer.rs:316-350,497-520). meta_encoding: 1 = Encoding::Null ->
 * TsmVersion::V1, 7 = Snappy -> TsmVersion::V2. Returns the file size (call with out = NULL / cap = 0 to size the
 * buffer), or 0 on bad input. */
uint64_t tskvtsm_write(const uint8_t *arena, const tskv_page_desc *descs, uint64_t n_descs,
                       const tskv_time_range *cg_bounds, uint64_t n_cg, const char *table_name, uint32_t meta_encoding,
                       uint8_t *out, uint64_t cap);
/* The same with PageMeta.statistics: value_stats[i] (or NULL = min / max None everywhere) goes into page i's metadata. */
uint64_t tskvtsm_write_stats(const uint8_t *arena, const tskv_page_desc *descs, uint64_t n_descs,
                             const tskv_time_range *cg_bounds, uint64_t n_cg, const char *table_name, uint32_t meta_encoding,
                             const tskv_value_stats *value_stats, uint8_t *out, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* TSKV_TSM_H */
