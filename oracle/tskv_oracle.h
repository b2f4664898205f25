/*
 * tskv_oracle.h — CPU restatement ("oracle") of the reference's tskv decode -> filter ->
 * bucket-aggregate path.  TEST INFRASTRUCTURE ONLY: nothing under cnosdb_b200/ may include, link
 * or call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and only as the checker / the CPU arm.
 *
 * Parity status: PINNED against the reference's own golden vectors (tests/golden/ JSON files, extracted
 * from /root/reference by tests/golden/make_golden.py): the two byte-exact InfluxDB vectors
 * (integer.rs:438-483), the zig-zag table (integer.rs:270-280), simple8b encoded lengths
 * (simple8b.rs:232-252), every round-trip corpus of timestamp.rs / integer.rs / simple8b.rs /
 * float.rs, the time_window known-answer tuples (time_window.rs:318-368) and the SQL goldens of
 * sqllogicaltests/cases/function/common/ .slt files; the overlap merge against the three tables of
 * reader/sort_merge.rs:449-539 and the grouping table of reader/utils.rs:330-353 (tests/test_oracle_merge.py); the
 * boolean codec against the vectors of tsm/codec/boolean.rs:150-260 (tests/test_oracle_bool.py).
 * The reference itself (Rust) cannot be compiled in
 * this image (no rustc/cargo), so there is no oracle/_ref build.
 *
 * All citations are relative to /root/reference.
 */
#ifndef TSKV_ORACLE_H_
#define TSKV_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#include "../include/tskv_gpu.h" /* shared descriptor / query structs and status codes */

#ifdef __cplusplus
extern "C" {
#endif

/* ---- codecs (tskv/src/tsm/codec/). Encoders return bytes written, or a negative status. */
int64_t orc_simple8b_encode(const uint64_t *src, uint64_t n, uint8_t *dst, uint64_t cap);
int64_t orc_simple8b_decode(const uint8_t *src, uint64_t len, uint64_t *dst, uint64_t cap);
uint64_t orc_zigzag_encode(int64_t v);
int64_t orc_zigzag_decode(uint64_t v);
/* Full column encoders incl. the leading Encoding id byte. */
int64_t orc_ts_encode(const int64_t *src, uint64_t n, uint8_t *dst, uint64_t cap);   /* DeltaTs=11 */
int64_t orc_i64_encode(const int64_t *src, uint64_t n, uint8_t *dst, uint64_t cap);  /* Delta=2 */
int64_t orc_f64_encode(const double *src, uint64_t n, uint8_t *dst, uint64_t cap);   /* Gorilla=6 */
int64_t orc_raw_encode(const uint64_t *src, uint64_t n, uint8_t *dst, uint64_t cap); /* Null=1 */
int64_t orc_bool_encode(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap);     /* BitPack=10 (boolean.rs:24-64) */
int64_t orc_bool_raw_encode(const uint8_t *src, uint64_t n, uint8_t *dst, uint64_t cap); /* Null=1 (boolean.rs:66-76) */
/* Column decode with codec dispatch (tsm/reader.rs:658-731 + codec/instance.rs:358-401).
 * out_vals: n_rows 8-byte cells (0 for null rows); out_valid: n_rows bytes (0/1). */
tskv_status orc_decode_column(uint32_t phys_type, const uint8_t *data, uint64_t data_len,
                              const uint8_t *bitset, uint64_t n_rows, uint64_t *out_vals,
                              uint8_t *out_valid);

/* ---- page framing (tskv/src/tsm/page.rs:58-94, 334-345). */
uint32_t orc_crc32(const uint8_t *data, uint64_t len);
/* Builds header|bitset|data ; returns total size (16 + bitset_len + data_len). */
uint64_t orc_page_build(const uint8_t *bitset, uint32_t bitset_len, uint64_t n_rows,
                        const uint8_t *data, uint64_t data_len, uint8_t *out);
tskv_status orc_page_decode(uint32_t phys_type, const uint8_t *page, uint64_t size, int verify_crc,
                            uint64_t *out_vals, uint8_t *out_valid, uint64_t cap_rows,
                            uint64_t *out_n_rows);

/* ---- bucket expression (time_window.rs:184-198 == transform_time_window.rs:251-296). */
void orc_sliding_window(int64_t t, int64_t window, int64_t slide, int64_t start_time, int64_t i,
                        int64_t *out_start, int64_t *out_end);
void orc_ceil_sliding_window(int64_t t, int64_t window, int64_t slide, int64_t start_time,
                             int64_t *out_start, int64_t *out_end);
void orc_floor_sliding_window(int64_t t, int64_t window, int64_t slide, int64_t start_time,
                              int64_t *out_start, int64_t *out_end);

/* ---- whole path: same inputs / outputs as tskvgpu_decode_pages / tskvgpu_scan_aggregate. */
tskv_status orc_decode_pages(const uint8_t *arena, uint64_t arena_len, const tskv_page_desc *descs,
                             uint64_t n_descs, uint64_t first_page, uint64_t n_pages,
                             int verify_crc, uint64_t *out_values, uint8_t *out_validity);
tskv_status orc_query_output_layout(const tskv_page_desc *descs, uint64_t n_descs,
                                    const tskv_query *q, tskv_output_layout *out);
/* n_threads <= 1: single-threaded, rows in (series slot, page, row) order — the arrival order of
 * a single-partition reference scan.  n_threads > 1: contiguous series chunks of size
 * (n + ncpu) / ncpu like tskv/src/reader/iterator.rs:232-235, partial tables merged in chunk order.
 * verify_crc mirrors Page::crc_validation on every page read (tsm/reader.rs:259,492). */
tskv_status orc_scan_aggregate(const uint8_t *arena, uint64_t arena_len,
                               const tskv_page_desc *descs, uint64_t n_descs, const tskv_query *q,
                               int verify_crc, int n_threads, uint64_t *out_values,
                               uint8_t *out_validity, uint64_t *out_points /* may be NULL */);
/* Same with a TsmTombstone attached (tsm/reader.rs:507-551): entries as in include/tskv_gpu.h. Rows are located
 * the reference's way, by binary search over the page's time values (reader.rs:634-656). */
tskv_status orc_scan_aggregate_tomb(const uint8_t *arena, uint64_t arena_len,
                                    const tskv_page_desc *descs, uint64_t n_descs, const tskv_query *q,
                                    const tskv_tombstone *tombs, uint64_t n_tombs,
                                    int verify_crc, int n_threads, uint64_t *out_values,
                                    uint8_t *out_validity, uint64_t *out_points /* may be NULL */);
/* Handle API used by bench.py's CPU arms: `orc_open` builds the series index once (the reference keeps the
 * TsmReader's chunk metadata cached after open: tsm/reader.rs:120-168, tsfamily/version.rs:158-172) and starts a
 * persistent pool of n_threads workers; `orc_scan` is one query over the opened pages (same semantics and results
 * as orc_scan_aggregate_tomb, which is now open + scan + close). The caller keeps arena / descs alive. */
typedef struct orc_handle orc_handle;
tskv_status orc_open(const uint8_t *arena, uint64_t arena_len, const tskv_page_desc *descs, uint64_t n_descs,
                     int n_threads, orc_handle **out);
tskv_status orc_scan(orc_handle *h, const tskv_query *q, const tskv_tombstone *tombs, uint64_t n_tombs,
                     int verify_crc, uint64_t *out_values, uint8_t *out_validity, uint64_t *out_points);
/* Overlapping chunks (reader/iterator.rs:463-560, reader/sort_merge.rs, reader/batch_builder.rs): file id of every
 * column group in descriptor-table order; later scans merge + de-duplicate the overlapping chunks of a series. */
tskv_status orc_set_chunk_files(orc_handle *h, const uint64_t *cg_file_id, uint64_t n_cg);
void orc_close(orc_handle *h);
const char *orc_last_error(void);
/* Tests only: scans with field predicates normally skip the column groups whose page min / max rule a predicate out
 * (filter_column_groups, reader/chunk.rs:12-50); 0 turns that off so that both ways can be compared. */
void orc_set_value_stats_pruning(int on);

#ifdef __cplusplus
}
#endif
#endif
