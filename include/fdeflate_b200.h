/*
 * fdeflate_b200.h -- C ABI of the B200-native batch zlib codec (libfdeflate_b200.so).
 *
 * Drop-in boundary for the hot path of image-rs/fdeflate (reference citations are file:line in the
 * reference crate, version 0.4.0-dev).  The reference has no FFI of its own (it is a pure-Rust crate,
 * `#![forbid(unsafe_code)]`, src/lib.rs:21); its boundary is the public Rust API re-exported at
 * src/lib.rs:29-36.  Each entry point below names the Rust item it replaces; INTEGRATION.md shows the
 * `extern "C"` block and the thin wrappers a maintainer adds on the Rust side.
 *
 * Conventions
 *   - plain pointers and sizes only; no allocation crosses the ABI; the caller owns every buffer.
 *   - a batch is n independent zlib streams: stream i reads in_base[in_off[i] .. +in_len[i]) and writes
 *     out_base[out_off[i] .. +out_cap[i]).  Slots must not overlap.  Offsets may have any alignment;
 *     16-byte aligned offsets take the vectorised paths.
 *   - function return value: 0 = the batch ran; non-zero = batch-level failure (CUDA error or bad
 *     argument; text via fdb_last_error).  Per-stream results are in status[]; one bad stream never
 *     aborts the batch.
 *   - *_device entry points take DEVICE pointers (including the offset/length/result arrays) and
 *     enqueue on `cuda_stream` without synchronising; the others take HOST pointers, stage through
 *     device memory owned by the context, and return when the results are in the caller's buffers.
 *   - device input buffers must be readable up to the next 16-byte boundary past the last stream
 *     (true for any cudaMalloc'ed buffer).
 *   - slots may come in any order.  The host-buffer entry points move runs of adjacent slots (less than
 *     256 bytes apart) and evenly spaced slots in one copy each, and a batch of more than 64 widely and
 *     unevenly spaced slots per pipeline chunk as one range: bytes of out_base BETWEEN slots may therefore be
 *     overwritten.  Nothing before the first slot or behind the end of the last one is ever touched, and
 *     nothing outside in_base's [min in_off, max in_off + in_len) is read.
 *   - a context is bound to one GPU and is not re-entrant; use one context per host thread.  The
 *     *_device entry points share the context's work counters and scratch: at most ONE device-pointer
 *     call per context may be in flight on the GPU at a time (enqueue the next one on the same stream, or
 *     after the previous one has finished; use one context per concurrent stream).
 *   - there is no CPU fallback: without a CUDA device fdb_create fails.
 */
#ifndef FDEFLATE_B200_H
#define FDEFLATE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-stream status.  1..16 = DecompressionError variants in declaration order
 * (src/decompress.rs:13-48); 17 = BoundedDecompressionError::OutputTooLarge (src/decompress.rs:1098);
 * 18 = compress side only, the caller's slot is smaller than the encoded stream. */
enum fdb_status {
    FDB_OK = 0,
    FDB_BAD_ZLIB_HEADER = 1,
    FDB_INSUFFICIENT_INPUT = 2,
    FDB_INVALID_BLOCK_TYPE = 3,
    FDB_INVALID_UNCOMPRESSED_BLOCK_LENGTH = 4,
    FDB_INVALID_HLIT = 5,
    FDB_INVALID_HDIST = 6,
    FDB_INVALID_CODE_LENGTH_REPEAT = 7,
    FDB_BAD_CODE_LENGTH_HUFFMAN_TREE = 8,
    FDB_BAD_LITERAL_LENGTH_HUFFMAN_TREE = 9,
    FDB_BAD_DISTANCE_HUFFMAN_TREE = 10,
    FDB_INVALID_LITERAL_LENGTH_CODE = 11,
    FDB_INVALID_DISTANCE_CODE = 12,
    FDB_INPUT_STARTS_WITH_RUN = 13,
    FDB_DISTANCE_TOO_FAR_BACK = 14,
    FDB_WRONG_CHECKSUM = 15,
    FDB_EXTRA_INPUT = 16,
    FDB_OUTPUT_TOO_LARGE = 17,
    FDB_OUTPUT_BUFFER_TOO_SMALL = 18
};

/* flags for the inflate entry points */
#define FDB_FLAG_IGNORE_ADLER32 1u /* Decompressor::ignore_adler32, src/decompress.rs:154-156 */
#define FDB_FLAG_GENERAL_ONLY 2u   /* skip the ultra-fast-format fast path (testing / profiling) */
#define FDB_FLAG_SPLIT_LARGE 4u    /* device-pointer calls: decode ultra-fast-format streams of >= 128 KiB with many
                                      warps each (three extra small launches per batch).  The host-buffer calls
                                      turn it on by themselves when a batch holds such a stream. */

typedef struct fdb_ctx fdb_ctx;

/* Context life cycle.  device = CUDA device ordinal. */
int fdb_create(int device, fdb_ctx** ctx);
void fdb_destroy(fdb_ctx* ctx);
const char* fdb_last_error(const fdb_ctx* ctx);
const char* fdb_version(void);

/* ---- inflate ---------------------------------------------------------------------------------
 * Replaces, per stream, decompress_to_vec_bounded(input, maxlen = out_cap[i])
 * (src/decompress.rs:1111-1144), i.e. the Decompressor::read state machine (src/decompress.rs:179-337)
 * run to completion.  decompress_to_vec (src/decompress.rs:1079) is the same call with a slot known
 * to be large enough.  On FDB_OK out_len[i] bytes are valid and equal the reference's Vec; on
 * FDB_OUTPUT_TOO_LARGE out_len[i] == out_cap[i] and the slot holds the reference's partial_output;
 * on any other status the slot content is unspecified.  consumed (may be NULL) receives the number
 * of input bytes read up to and including the adler32 trailer (trailing bytes are ignored, as in the
 * reference).  adler32 is verified on the device unless FDB_FLAG_IGNORE_ADLER32. */
int fdb_inflate_batch_device(fdb_ctx* ctx, const void* d_in_base, const uint64_t* d_in_off,
                             const uint64_t* d_in_len, void* d_out_base, const uint64_t* d_out_off,
                             const uint64_t* d_out_cap, uint64_t* d_out_len, uint64_t* d_consumed,
                             int32_t* d_status, size_t n, uint32_t flags, void* cuda_stream);
int fdb_inflate_batch(fdb_ctx* ctx, const uint8_t* in_base, const uint64_t* in_off, const uint64_t* in_len,
                      uint8_t* out_base, const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len,
                      uint64_t* consumed, int32_t* status, size_t n, uint32_t flags);

/* ---- ultra-fast deflate ------------------------------------------------------------------------
 * Replaces, per stream, compress_to_vec_ultra_fast(input) (src/compress/mod.rs:313-317) =
 * UltraFastCompressor::new + one write_data(whole input) + finish (src/compress/ultrafast.rs:70-181).
 * Output bytes are identical to the reference's.  A slot of fdb_deflate_ultrafast_bound(in_len)
 * bytes always suffices. */
size_t fdb_deflate_ultrafast_bound(size_t in_len);
int fdb_deflate_ultrafast_batch_device(fdb_ctx* ctx, const void* d_in_base, const uint64_t* d_in_off,
                                       const uint64_t* d_in_len, void* d_out_base, const uint64_t* d_out_off,
                                       const uint64_t* d_out_cap, uint64_t* d_out_len, int32_t* d_status, size_t n,
                                       void* cuda_stream);
int fdb_deflate_ultrafast_batch(fdb_ctx* ctx, const uint8_t* in_base, const uint64_t* in_off,
                                const uint64_t* in_len, uint8_t* out_base, const uint64_t* out_off,
                                const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t n);

/* ---- stored ("level 0") deflate ------------------------------------------------------------------
 * Replaces Compressor::new(w, 0, true) + one write_data(whole input) + finish()
 * (src/compress/mod.rs:69-101, :126-156, :194-214, :241-268) -- the north star's StoredOnlyCompressor. */
size_t fdb_deflate_stored_bound(size_t in_len);
int fdb_deflate_stored_batch_device(fdb_ctx* ctx, const void* d_in_base, const uint64_t* d_in_off,
                                    const uint64_t* d_in_len, void* d_out_base, const uint64_t* d_out_off,
                                    const uint64_t* d_out_cap, uint64_t* d_out_len, int32_t* d_status, size_t n,
                                    void* cuda_stream);
int fdb_deflate_stored_batch(fdb_ctx* ctx, const uint8_t* in_base, const uint64_t* in_off, const uint64_t* in_len,
                             uint8_t* out_base, const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len,
                             int32_t* status, size_t n);

/* ---- PNG row filters (the step either side of the zlib path in a PNG codec) ----
 * Not part of image-rs/fdeflate: the `png` crate applies them around its fdeflate calls.  Semantics are the PNG
 * specification's section 9.  A FILTERED image is `height` rows of (1 filter-type byte + stride bytes) -- exactly
 * what inflating a non-interlaced IDAT stream yields -- a RAW image is `height` rows of stride bytes; bpp = bytes
 * per complete pixel (1..8).  All pointers are device pointers; image i sits at base + off[i].
 *   unfilter: filtered -> raw.  status 19 = a row's filter type is not 0..4, 20 = bpp out of range.
 *   filter  : raw -> filtered.  mode 0..4 = None/Sub/Up/Average/Paeth on every row, 5 = per row the type with the
 *             smallest sum of absolute signed differences (PNG 12.8), lowest type on ties. */
int fdb_png_unfilter_batch_device(fdb_ctx* ctx, const void* d_filtered_base, const uint64_t* d_filtered_off,
                                  void* d_raw_base, const uint64_t* d_raw_off, const uint32_t* d_height,
                                  const uint32_t* d_stride, const uint32_t* d_bpp, int32_t* d_status, size_t n,
                                  void* cuda_stream);
int fdb_png_filter_batch_device(fdb_ctx* ctx, const void* d_raw_base, const uint64_t* d_raw_off, void* d_filtered_base,
                                const uint64_t* d_filtered_off, const uint32_t* d_height, const uint32_t* d_stride,
                                const uint32_t* d_bpp, uint32_t mode, int32_t* d_status, size_t n, void* cuda_stream);

/* filter + ultra-fast deflate of raw images in ONE kernel, device pointers: the encoder computes the filtered bytes of
 * image i (mode 0..4, one type on every row) from the raw rows while it stages them, so the filtered image is never
 * stored.  Output as fdb_deflate_ultrafast_batch_device (slot i >= fdb_deflate_ultrafast_bound(height * (1 + stride)));
 * d_filter_status[i] = 0, or 20 for bpp outside 1..8, mode > 4 (the adaptive mode 5 takes the two calls
 * fdb_png_filter_batch_device + fdb_deflate_ultrafast_batch_device) or an image of 4 GiB and more -- such an image
 * produces no stream (d_out_len[i] = 0).  The raw buffer must be readable up to the end of the 32-bit word that holds
 * its last byte.  Measured 1.6-2.7x SLOWER than the two calls on 4096 RGBA tiles (the raw bytes are fetched word by word
 * through L1 instead of in coalesced 16-byte vectors): use it when the filtered copy's device memory is what matters.
 * fdb_png_encode_batch / fdb_png_encode_files_batch take this path when the context was created with FDB_PNG_FUSED=1. */
int fdb_png_encode_batch_device(fdb_ctx* ctx, const void* d_raw_base, const uint64_t* d_raw_off, const uint32_t* d_height,
                                const uint32_t* d_stride, const uint32_t* d_bpp, uint32_t mode, void* d_out_base,
                                const uint64_t* d_out_off, const uint64_t* d_out_cap, uint64_t* d_out_len,
                                int32_t* d_filter_status, int32_t* d_status, size_t n, void* cuda_stream);

/* the same with host buffers (staged through the context; every array is a host array) */
int fdb_png_unfilter_batch(fdb_ctx* ctx, const uint8_t* filtered_base, const uint64_t* filtered_off, uint8_t* raw_base,
                           const uint64_t* raw_off, const uint32_t* height, const uint32_t* stride, const uint32_t* bpp,
                           int32_t* status, size_t n);
int fdb_png_filter_batch(fdb_ctx* ctx, const uint8_t* raw_base, const uint64_t* raw_off, uint8_t* filtered_base,
                         const uint64_t* filtered_off, const uint32_t* height, const uint32_t* stride, const uint32_t* bpp,
                         uint32_t mode, int32_t* status, size_t n);

/* PNG image data in one call, host buffers, the filtered image stays on the device:
 *   decode: n zlib streams (the concatenated IDAT payload of one non-interlaced image each) -> raw pixels at
 *           raw_base + raw_off[i] (height[i] * stride[i] bytes).  status[i] = the inflate status if not Ok (a stream
 *           that ends before height * (1 + stride) bytes is InsufficientInput, one that goes on is OutputTooLarge),
 *           else the unfilter status.
 *   encode: raw pixels -> filter (mode as above) -> ultra-fast deflate; slot i needs
 *           fdb_deflate_ultrafast_bound(height[i] * (1 + stride[i])) bytes. */
int fdb_png_decode_batch(fdb_ctx* ctx, const uint8_t* idat_base, const uint64_t* idat_off, const uint64_t* idat_len,
                         uint8_t* raw_base, const uint64_t* raw_off, const uint32_t* height, const uint32_t* stride,
                         const uint32_t* bpp, int32_t* status, size_t n);
int fdb_png_encode_batch(fdb_ctx* ctx, const uint8_t* raw_base, const uint64_t* raw_off, const uint32_t* height,
                         const uint32_t* stride, const uint32_t* bpp, uint32_t mode, uint8_t* out_base,
                         const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len, int32_t* status, size_t n);

/* PNG FILES in one call.  The chunk structure is walked on the host (signature, IHDR, chunk order, lengths: PNG 5.2-5.6,
 * 11.2.2; only chunk headers are read); chunk CRCs, the gathering of several IDAT payloads into one zlib stream,
 * inflate and unfilter run on the device.  status: 0, an inflate / unfilter status, or 21 = not a PNG / broken chunk
 * structure / invalid IHDR, 22 = a chunk's CRC-32 does not match, 23 = valid but not decoded here (interlaced).
 *   probe : host only, no context: geometry of every file (stride = bytes per raw row; pixels need height * stride).
 *   decode: raw pixels of file i to raw_base + raw_off[i], a slot of raw_cap[i] bytes.  The sizes come from the
 *           (untrusted) file: an image that needs more than raw_cap[i] bytes gets status 17 (OutputTooLarge), nothing
 *           of it is written and the rest of the batch is unaffected.  Images whose slots follow each other with
 *           less than 16 bytes of padding are copied back in one piece, padding included. */
int fdb_png_probe_batch(const uint8_t* file_base, const uint64_t* file_off, const uint64_t* file_len, uint32_t* width,
                        uint32_t* height, uint32_t* bit_depth, uint32_t* color_type, uint32_t* stride, int32_t* status,
                        size_t n);
int fdb_png_decode_files_batch(fdb_ctx* ctx, const uint8_t* file_base, const uint64_t* file_off, const uint64_t* file_len,
                               uint8_t* raw_base, const uint64_t* raw_off, const uint64_t* raw_cap, int32_t* status, size_t n);

/* ... and the other way: raw pixels (8- or 16-bit gray, gray + alpha, RGB, RGBA; 16-bit samples big-endian as in the
 * file) -> complete PNG files (signature, IHDR, one IDAT chunk holding an ultra-fast zlib stream, IEND) at
 * file_base + file_off[i]; file_cap[i] >= fdb_png_file_bound(...).  Filter `mode` as in fdb_png_filter_batch.  The row
 * filter, the deflate and the IDAT CRC run on the device, on the pipeline of the host-buffer deflate call. */
size_t fdb_png_file_bound(uint32_t width, uint32_t height, uint32_t bit_depth, uint32_t color_type);
int fdb_png_encode_files_batch(fdb_ctx* ctx, const uint8_t* raw_base, const uint64_t* raw_off, const uint32_t* width,
                               const uint32_t* height, const uint32_t* bit_depth, const uint32_t* color_type, uint32_t mode,
                               uint8_t* file_base, const uint64_t* file_off, const uint64_t* file_cap, uint64_t* file_len,
                               int32_t* status, size_t n);

/* ---- CRC-32 of a batch of byte ranges (PNG chunk CRCs; polynomial 0xEDB88320, the value zlib's crc32() gives) ----
 * crc[i] = CRC-32 of base[off[i] .. off[i] + len[i]) continued from `seed` = the CRC of whatever precedes every
 * range (0 = nothing; e.g. crc32("IDAT") for the payloads of IDAT chunks).  One warp per range. */
int fdb_crc32_batch_device(fdb_ctx* ctx, const void* d_base, const uint64_t* d_off, const uint64_t* d_len, uint32_t seed,
                           uint32_t* d_crc, size_t n, void* cuda_stream);
int fdb_crc32_batch(fdb_ctx* ctx, const uint8_t* base, const uint64_t* off, const uint64_t* len, uint32_t seed,
                    uint32_t* crc, size_t n);

/* ---- synthetic PNG-filtered RGBA tiles (benchmark / test input) -----------------
 * Tile t is width x height RGBA8, row 0 Sub-filtered, other rows Paeth-filtered; each row is
 * 1 filter-type byte + 4*width residual bytes, so a tile is height*(1+4*width) bytes, laid out
 * back to back.  Integer-only and counter-based, so the host and device versions agree exactly. */
size_t fdb_synth_tile_bytes(uint32_t width, uint32_t height);
int fdb_synth_tiles_host(uint8_t* out, uint64_t first_tile, uint64_t n_tiles, uint32_t width, uint32_t height,
                         uint64_t seed);
int fdb_synth_tiles_device(fdb_ctx* ctx, void* d_out, uint64_t first_tile, uint64_t n_tiles, uint32_t width,
                           uint32_t height, uint64_t seed, void* cuda_stream);

/* Host-buffer batches (fdb_inflate_batch, fdb_deflate_*_batch) are pipelined: the batch is cut into
 * chunks of about `bytes` of slot span and chunk k+1 is copied to the device while chunk k is in the
 * kernels and chunk k-1 is copied back.  0 restores the default (128 MiB).  A tuning knob, not part of
 * the reference's interface; results do not depend on it. */
int fdb_set_pipeline_chunk(fdb_ctx* ctx, size_t bytes);

/* number of kernels this library has launched through ctx since creation (bench bookkeeping) */
uint64_t fdb_launch_count(const fdb_ctx* ctx);
/* how many streams of the most recent inflate batch on this context were declined by the
 * ultra-fast-format fast path and decoded by the general kernel.  Synchronises `cuda_stream`. */
int64_t fdb_last_general_count(fdb_ctx* ctx, void* cuda_stream);
/* Long streams, many warps each.  An ultra-fast-format stream of >= 128 KiB is inflated span by span and an
 * input of >= 256 KiB is ultra-fast-deflated segment by segment (64 KiB units, three passes),
 * so that a batch of few or very uneven streams still fills the GPU.  A deflate batch with more inputs than the device
 * runs warps only cuts an input that holds an eighth of the batch's bytes (the work counter balances the rest).
 * Results are identical either way.  The host-buffer calls turn this on by themselves for a batch
 * (chunk) that holds a stream of >= 128 KiB (inflate) or an input of >= 1 MiB (deflate); for the device-pointer
 * calls, which do not see the sizes, it is off unless fdb_set_split_large(ctx, 1) (costs a few extra small
 * launches per batch).  fdb_set_split_threshold changes the two sizes (0 = default). */
int fdb_set_split_large(fdb_ctx* ctx, int on);
int fdb_set_split_threshold(fdb_ctx* ctx, size_t inflate_stream_bytes, size_t deflate_input_bytes);
/* Span-by-span inflate keeps what its first pass learns about every span (one word per lane and segment: 1/8 of the
 * compressed bytes) so that the pass that writes does not count again.  The host-buffer calls size that scratch from
 * the batch; a device-pointer call, which does not see the sizes, uses a pool of `bytes` (default 256 MiB = 2 GiB of
 * compressed input per call; spans beyond the pool are simply counted twice; 0 = no pool).  Device memory, allocated
 * on the first span-by-span call. */
int fdb_set_split_scratch(fdb_ctx* ctx, size_t bytes);
/* how many spans the long streams of the most recent inflate batch on this context were cut into
 * (0 = every stream was decoded by one warp).  Synchronises `cuda_stream`. */
int64_t fdb_last_split_spans(fdb_ctx* ctx, void* cuda_stream);

/* ---- streaming decoders: Decompressor::read with its state kept on the device (src/decompress.rs:96-113, :158-219) ----
 * A decoder is opened, fed any number of times with whatever input has arrived and whatever room the caller has, and
 * closed.  Every call of fdb_stream_read_batch advances all the decoders it names in ONE launch.  For decoder ids[i]
 * the call TAKES all of in_base[in_off[i] .. + in_len[i]) (what cannot be parsed yet -- the input may end inside a
 * token or a block header -- is kept by the context), writes at most out_room[i] bytes to out_base + out_off[i] and
 * reports in produced[i] how many; the bytes come out in stream order, each exactly once, whatever the chunking.
 * status[i]: FDB_OK = the stream is complete (checksum verified unless FDB_FLAG_IGNORE_ADLER32; later calls produce
 * nothing, :185-187), FDB_STREAM_NEED_INPUT = everything given so far is decoded, FDB_STREAM_OUTPUT_FULL = the room is
 * used up and more output is pending (call again, in_len may be 0), or a DecompressionError (1..16), after which the
 * decoder stays in that state.  The work of a call is proportional to the bytes of that call, not to the length of
 * the stream so far: the decoder resumes at a token boundary with its tables, its 32 KiB window, a match cut short by
 * a full output (the reference's QueuedOutput, :1066-1070) and its running adler32 kept in device memory. */
#define FDB_STREAM_NEED_INPUT (-2)
#define FDB_STREAM_OUTPUT_FULL (-3)
int fdb_stream_open_batch(fdb_ctx* ctx, uint32_t* ids, size_t n);
int fdb_stream_read_batch(fdb_ctx* ctx, const uint32_t* ids, const uint8_t* in_base, const uint64_t* in_off,
                          const uint64_t* in_len, uint8_t* out_base, const uint64_t* out_off, const uint64_t* out_room,
                          uint64_t* produced, int32_t* status, size_t n, uint32_t flags);
int fdb_stream_close_batch(fdb_ctx* ctx, const uint32_t* ids, size_t n);

/* ---- several GPUs of one box behind one handle (SURVEY 8b `device set`, 8e) -----------------------------
 * A batch shards by stream with no exchange step: fdb_multi_* partitions the streams over the devices of the
 * set, byte-balanced on in_len + out_cap (resp. in_len) -- contiguous runs of streams when that balances within 3 %
 * of the longest-processing-time greedy partition (it does with many streams per device, and keeps every device's
 * slots one range of the caller's buffers), the greedy partition itself otherwise; a stream is never split across
 * GPUs -- and drives one host thread and one context per device: each runs the ordinary pipelined host-buffer
 * call on its share, reading and writing the caller's slots in place.  The call returns when every share is done.
 * With interleaved shares every device copies back exactly its own slots (nothing between them is written).
 * Arguments and per-stream results are exactly those of the single-device calls; results do not depend on the
 * device set.  devices[] holds CUDA ordinals (a device may appear more than once: that many contexts on it).
 * fdb_multi_last_error returns the text of the first device that failed. */
typedef struct fdb_multi fdb_multi;
int fdb_multi_create(const int* devices, int n_devices, fdb_multi** out);
void fdb_multi_destroy(fdb_multi* m);
int fdb_multi_device_count(const fdb_multi* m);
const char* fdb_multi_last_error(const fdb_multi* m);
int fdb_multi_inflate_batch(fdb_multi* m, const uint8_t* in_base, const uint64_t* in_off, const uint64_t* in_len,
                            uint8_t* out_base, const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len,
                            uint64_t* consumed, int32_t* status, size_t n, uint32_t flags);
int fdb_multi_deflate_ultrafast_batch(fdb_multi* m, const uint8_t* in_base, const uint64_t* in_off, const uint64_t* in_len,
                                      uint8_t* out_base, const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len,
                                      int32_t* status, size_t n);
int fdb_multi_deflate_stored_batch(fdb_multi* m, const uint8_t* in_base, const uint64_t* in_off, const uint64_t* in_len,
                                   uint8_t* out_base, const uint64_t* out_off, const uint64_t* out_cap, uint64_t* out_len,
                                   int32_t* status, size_t n);
/* which device (index into devices[]) stream i of the most recent fdb_multi_* call ran on; owner[n] */
int fdb_multi_last_partition(const fdb_multi* m, uint32_t* owner, size_t n);

#ifdef __cplusplus
}
#endif
#endif
