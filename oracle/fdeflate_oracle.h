/*
 * fdeflate_oracle.h -- CPU ORACLE for the fdeflate hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the reference crate's algorithms (image-rs/fdeflate
 * 0.4.0-dev, files cited per function in fdeflate_oracle.c).  It exists so that the CUDA
 * path can be checked bit-for-bit on a machine with no Rust toolchain.  Nothing in the
 * product (fdeflate_b200/, include/) may link, import or call it; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity pin status (see DESIGN.md "Oracle"):
 *   inflate      : pinned by the reference's golden vectors (tests/ .zz files, fuzz corpus, fixed-table KAT,
 *                  huffman.rs KATs, unit-test behaviours) + differential against system zlib.
 *   table build  : pinned by FIXED_LITLEN_TABLE / FIXED_DIST_TABLE known answers.
 *   adler32      : standard RFC 1950 (third-party simd-adler32 ^0.3.4 in the reference, not vendored);
 *                  pinned by decompress.rs:1351 (adler32(example1 output) == 751299) and zlib.adler32.
 *   ultra-fast   : PARITY UNPINNED BY REFERENCE TESTS (the reference has no byte-level golden for it);
 *                  pinned only by line-by-line restatement + HEADER constant + zlib round trips.
 */
#ifndef FDEFLATE_ORACLE_H
#define FDEFLATE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* DecompressionError in declaration order (decompress.rs:13-48), 0 = Ok,
 * 17 = BoundedDecompressionError::OutputTooLarge (decompress.rs:1098). */
enum fdo_status {
    FDO_OK = 0,
    FDO_BAD_ZLIB_HEADER = 1,
    FDO_INSUFFICIENT_INPUT = 2,
    FDO_INVALID_BLOCK_TYPE = 3,
    FDO_INVALID_UNCOMPRESSED_BLOCK_LENGTH = 4,
    FDO_INVALID_HLIT = 5,
    FDO_INVALID_HDIST = 6,
    FDO_INVALID_CODE_LENGTH_REPEAT = 7,
    FDO_BAD_CODE_LENGTH_HUFFMAN_TREE = 8,
    FDO_BAD_LITERAL_LENGTH_HUFFMAN_TREE = 9,
    FDO_BAD_DISTANCE_HUFFMAN_TREE = 10,
    FDO_INVALID_LITERAL_LENGTH_CODE = 11,
    FDO_INVALID_DISTANCE_CODE = 12,
    FDO_INPUT_STARTS_WITH_RUN = 13,
    FDO_DISTANCE_TOO_FAR_BACK = 14,
    FDO_WRONG_CHECKSUM = 15,
    FDO_EXTRA_INPUT = 16,
    FDO_OUTPUT_TOO_LARGE = 17
};

#define FDO_FLAG_IGNORE_ADLER32 1u

/* ---- adler32 (RFC 1950; stands in for simd_adler32::Adler32) ---- */
uint32_t fdo_adler32(uint32_t adler, const uint8_t *data, size_t len);

/* ---- huffman.rs:18-184 build_table ---- */
/* secondary_table must have room for FDO_SECONDARY_CAP u16 entries. Returns 1/0 like the reference. */
#define FDO_SECONDARY_CAP 4096
int fdo_build_table(const uint8_t *lengths, size_t n_lengths, const uint32_t *entries, size_t n_entries,
                    uint16_t *codes, uint32_t *primary_table, size_t primary_size,
                    uint16_t *secondary_table, size_t *secondary_len, int is_distance_table,
                    int double_literal);
/* decompress.rs:561-606 CompressedBlock::build_tables on a 320-entry length array.
 * litlen_table[4096], dist_table[512]; returns an fdo_status. */
int fdo_build_tables(size_t hlit, const uint8_t *code_lengths, uint32_t *litlen_table, uint32_t *dist_table,
                     uint16_t *secondary, size_t *secondary_len, uint16_t *dist_secondary,
                     size_t *dist_secondary_len);

/* ---- decompress.rs:96-342 streaming Decompressor ---- */
typedef struct fdo_decompressor fdo_decompressor;
fdo_decompressor *fdo_decompressor_new(void);
void fdo_decompressor_free(fdo_decompressor *d);
void fdo_decompressor_ignore_adler32(fdo_decompressor *d);
int fdo_decompressor_is_done(const fdo_decompressor *d);
/* Decompressor::read. Returns fdo_status; on FDO_OK *consumed / *produced are set. */
int fdo_decompressor_read(fdo_decompressor *d, const uint8_t *input, size_t input_len, uint8_t *output,
                          size_t output_len, size_t output_position, size_t *consumed, size_t *produced);

/* ---- decompress.rs:1111-1144 decompress_to_vec_bounded, writing into a caller buffer of
 * capacity maxlen (same 1024 / +32 KiB growth schedule for the visible output slice).
 * Returns fdo_status (FDO_OUTPUT_TOO_LARGE => out[0..*out_len) is the partial output). */
int fdo_inflate_into(const uint8_t *input, size_t input_len, uint8_t *out, size_t maxlen, uint32_t flags,
                     size_t *out_len, size_t *consumed);
/* decompress_to_vec (unbounded): mallocs *out (caller frees with fdo_free). */
int fdo_decompress_to_vec(const uint8_t *input, size_t input_len, uint32_t flags, uint8_t **out,
                          size_t *out_len);
void fdo_free(void *p);

/* ---- compress/ultrafast.rs + compress/mod.rs:313-317 compress_to_vec_ultra_fast ---- */
size_t fdo_ultrafast_bound(size_t n);
/* returns compressed length (always succeeds if out_cap >= fdo_ultrafast_bound(n)); 0 on overflow. */
size_t fdo_compress_ultra_fast(const uint8_t *data, size_t n, uint8_t *out, size_t out_cap);

/* UltraFastCompressor call by call (ultrafast.rs:70-181): new(writer) -> write_data(..)* -> finish.  The output
 * depends on how the input is cut into write_data calls (the zero-run state and the 8-byte chunking restart with
 * every call, :97-99).  `out` plays the writer; finish returns the stream length (0 = out_cap too small) and frees. */
typedef struct fdo_ultrafast fdo_ultrafast;
fdo_ultrafast *fdo_ultrafast_new(uint8_t *out, size_t out_cap);
void fdo_ultrafast_write_data(fdo_ultrafast *c, const uint8_t *data, size_t n);
size_t fdo_ultrafast_finish(fdo_ultrafast *c);

/* ---- compress/mod.rs:69-101,126-156,194-214,241-268: Compressor::new(w, 0, true) +
 * one write_data(whole input) + finish ---- */
size_t fdo_stored_bound(size_t n);
size_t fdo_compress_stored(const uint8_t *data, size_t n, uint8_t *out, size_t out_cap);

/* ---- constant tables exposed for tests ---- */
const uint8_t *fdo_huffman_lengths(void); /* 286 */
const uint16_t *fdo_huffman_codes(void);  /* 286, lib.rs:103-127 compute_codes */
const uint8_t *fdo_ultrafast_header(void); /* 54 bytes, ultrafast.rs:82-86 */
const uint32_t *fdo_litlen_table_entries(void);   /* 288, tables.rs:99-122 */
const uint32_t *fdo_distance_table_entries(void); /* 32, tables.rs:130-140 */
const uint16_t *fdo_length_to_symbol(void);       /* 256 */
const uint8_t *fdo_length_to_len_extra(void);     /* 256 */

/* ---- multi-threaded batch drivers (CPU baseline; one stream per task) ---- */
/* Each returns wall seconds spent inside the worker threads' region. */
double fdo_inflate_batch(const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                         uint8_t *out_base, const uint64_t *out_off, const uint64_t *out_cap,
                         uint64_t *out_len, int32_t *status, size_t n, uint32_t flags, int nthreads);
double fdo_compress_ultra_fast_batch(const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                                     uint8_t *out_base, const uint64_t *out_off, const uint64_t *out_cap,
                                     uint64_t *out_len, size_t n, int nthreads);
int fdo_hardware_threads(void);

/* ---- synthetic PNG-filtered RGBA tiles (the benchmark input, same bytes as the product library's
 * fdb_synth_tiles_host): n_tiles tiles of height * (1 + 4 * width) bytes back to back, generated on nthreads ---- */
void fdo_synth_tiles(uint8_t *out, uint64_t first_tile, uint64_t n_tiles, uint32_t width, uint32_t height,
                     uint64_t seed, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
