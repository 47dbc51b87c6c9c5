//! Raw binding of `include/fdeflate_b200.h`.  One item per C declaration, same order as the header; see the header
//! for the contract of every call (which reference item it replaces, ownership, per-stream status codes).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct fdb_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct fdb_multi {
    _private: [u8; 0],
}

// per-stream status: 1..=16 = DecompressionError variants in declaration order (src/decompress.rs:13-48)
pub const FDB_OK: i32 = 0;
pub const FDB_INSUFFICIENT_INPUT: i32 = 2;
pub const FDB_WRONG_CHECKSUM: i32 = 15;
pub const FDB_OUTPUT_TOO_LARGE: i32 = 17;
pub const FDB_OUTPUT_BUFFER_TOO_SMALL: i32 = 18;
pub const FDB_STREAM_NEED_INPUT: i32 = -2;
pub const FDB_STREAM_OUTPUT_FULL: i32 = -3;

pub const FDB_FLAG_IGNORE_ADLER32: u32 = 1;
pub const FDB_FLAG_GENERAL_ONLY: u32 = 2;
pub const FDB_FLAG_SPLIT_LARGE: u32 = 4;

extern "C" {
    pub fn fdb_create(device: c_int, ctx: *mut *mut fdb_ctx) -> c_int;
    pub fn fdb_destroy(ctx: *mut fdb_ctx);
    pub fn fdb_last_error(ctx: *const fdb_ctx) -> *const c_char;
    pub fn fdb_version() -> *const c_char;

    // ---- inflate: decompress_to_vec_bounded per stream (src/decompress.rs:1111-1144) ----
    pub fn fdb_inflate_batch_device(
        ctx: *mut fdb_ctx, d_in_base: *const c_void, d_in_off: *const u64, d_in_len: *const u64, d_out_base: *mut c_void,
        d_out_off: *const u64, d_out_cap: *const u64, d_out_len: *mut u64, d_consumed: *mut u64, d_status: *mut i32, n: usize,
        flags: u32, cuda_stream: *mut c_void,
    ) -> c_int;
    pub fn fdb_inflate_batch(
        ctx: *mut fdb_ctx, in_base: *const u8, in_off: *const u64, in_len: *const u64, out_base: *mut u8, out_off: *const u64,
        out_cap: *const u64, out_len: *mut u64, consumed: *mut u64, status: *mut i32, n: usize, flags: u32,
    ) -> c_int;

    // ---- ultra-fast deflate: compress_to_vec_ultra_fast per stream (src/compress/mod.rs:313-317) ----
    pub fn fdb_deflate_ultrafast_bound(in_len: usize) -> usize;
    pub fn fdb_deflate_ultrafast_batch_device(
        ctx: *mut fdb_ctx, d_in_base: *const c_void, d_in_off: *const u64, d_in_len: *const u64, d_out_base: *mut c_void,
        d_out_off: *const u64, d_out_cap: *const u64, d_out_len: *mut u64, d_status: *mut i32, n: usize, cuda_stream: *mut c_void,
    ) -> c_int;
    pub fn fdb_deflate_ultrafast_batch(
        ctx: *mut fdb_ctx, in_base: *const u8, in_off: *const u64, in_len: *const u64, out_base: *mut u8, out_off: *const u64,
        out_cap: *const u64, out_len: *mut u64, status: *mut i32, n: usize,
    ) -> c_int;

    // ---- stored ("level 0") deflate: Compressor::new(w, 0, true) (src/compress/mod.rs:69-101, :241-268) ----
    pub fn fdb_deflate_stored_bound(in_len: usize) -> usize;
    pub fn fdb_deflate_stored_batch_device(
        ctx: *mut fdb_ctx, d_in_base: *const c_void, d_in_off: *const u64, d_in_len: *const u64, d_out_base: *mut c_void,
        d_out_off: *const u64, d_out_cap: *const u64, d_out_len: *mut u64, d_status: *mut i32, n: usize, cuda_stream: *mut c_void,
    ) -> c_int;
    pub fn fdb_deflate_stored_batch(
        ctx: *mut fdb_ctx, in_base: *const u8, in_off: *const u64, in_len: *const u64, out_base: *mut u8, out_off: *const u64,
        out_cap: *const u64, out_len: *mut u64, status: *mut i32, n: usize,
    ) -> c_int;

    // ---- PNG rows / image data / files (the `png` crate's side of the path) ----
    pub fn fdb_png_unfilter_batch(
        ctx: *mut fdb_ctx, filtered_base: *const u8, filtered_off: *const u64, raw_base: *mut u8, raw_off: *const u64,
        height: *const u32, stride: *const u32, bpp: *const u32, status: *mut i32, n: usize,
    ) -> c_int;
    pub fn fdb_png_filter_batch(
        ctx: *mut fdb_ctx, raw_base: *const u8, raw_off: *const u64, filtered_base: *mut u8, filtered_off: *const u64,
        height: *const u32, stride: *const u32, bpp: *const u32, mode: u32, status: *mut i32, n: usize,
    ) -> c_int;
    pub fn fdb_png_decode_batch(
        ctx: *mut fdb_ctx, idat_base: *const u8, idat_off: *const u64, idat_len: *const u64, raw_base: *mut u8, raw_off: *const u64,
        height: *const u32, stride: *const u32, bpp: *const u32, status: *mut i32, n: usize,
    ) -> c_int;
    pub fn fdb_png_encode_batch(
        ctx: *mut fdb_ctx, raw_base: *const u8, raw_off: *const u64, height: *const u32, stride: *const u32, bpp: *const u32,
        mode: u32, out_base: *mut u8, out_off: *const u64, out_cap: *const u64, out_len: *mut u64, status: *mut i32, n: usize,
    ) -> c_int;
    // filter (mode 0..4) + ultra-fast deflate in one kernel, device pointers: the filtered image is never stored
    pub fn fdb_png_encode_batch_device(
        ctx: *mut fdb_ctx, d_raw_base: *const c_void, d_raw_off: *const u64, d_height: *const u32, d_stride: *const u32,
        d_bpp: *const u32, mode: u32, d_out_base: *mut c_void, d_out_off: *const u64, d_out_cap: *const u64, d_out_len: *mut u64,
        d_filter_status: *mut i32, d_status: *mut i32, n: usize, cuda_stream: *mut c_void,
    ) -> c_int;
    pub fn fdb_png_probe_batch(
        file_base: *const u8, file_off: *const u64, file_len: *const u64, width: *mut u32, height: *mut u32, bit_depth: *mut u32,
        color_type: *mut u32, stride: *mut u32, status: *mut i32, n: usize,
    ) -> c_int;
    pub fn fdb_png_decode_files_batch(
        ctx: *mut fdb_ctx, file_base: *const u8, file_off: *const u64, file_len: *const u64, raw_base: *mut u8, raw_off: *const u64,
        raw_cap: *const u64, status: *mut i32, n: usize,
    ) -> c_int;
    pub fn fdb_png_file_bound(width: u32, height: u32, bit_depth: u32, color_type: u32) -> usize;
    pub fn fdb_png_encode_files_batch(
        ctx: *mut fdb_ctx, raw_base: *const u8, raw_off: *const u64, width: *const u32, height: *const u32, bit_depth: *const u32,
        color_type: *const u32, mode: u32, file_base: *mut u8, file_off: *const u64, file_cap: *const u64, file_len: *mut u64,
        status: *mut i32, n: usize,
    ) -> c_int;
    pub fn fdb_crc32_batch(
        ctx: *mut fdb_ctx, base: *const u8, off: *const u64, len: *const u64, seed: u32, crc: *mut u32, n: usize,
    ) -> c_int;

    // ---- tuning / bookkeeping ----
    pub fn fdb_set_pipeline_chunk(ctx: *mut fdb_ctx, bytes: usize) -> c_int;
    pub fn fdb_launch_count(ctx: *const fdb_ctx) -> u64;
    pub fn fdb_last_general_count(ctx: *mut fdb_ctx, cuda_stream: *mut c_void) -> i64;
    pub fn fdb_set_split_large(ctx: *mut fdb_ctx, on: c_int) -> c_int;
    pub fn fdb_set_split_threshold(ctx: *mut fdb_ctx, inflate_stream_bytes: usize, deflate_input_bytes: usize) -> c_int;
    pub fn fdb_set_split_scratch(ctx: *mut fdb_ctx, bytes: usize) -> c_int;
    pub fn fdb_last_split_spans(ctx: *mut fdb_ctx, cuda_stream: *mut c_void) -> i64;

    // ---- streaming decoders: Decompressor::read with its state on the device (src/decompress.rs:158-219) ----
    pub fn fdb_stream_open_batch(ctx: *mut fdb_ctx, ids: *mut u32, n: usize) -> c_int;
    pub fn fdb_stream_read_batch(
        ctx: *mut fdb_ctx, ids: *const u32, in_base: *const u8, in_off: *const u64, in_len: *const u64, out_base: *mut u8,
        out_off: *const u64, out_room: *const u64, produced: *mut u64, status: *mut i32, n: usize, flags: u32,
    ) -> c_int;
    pub fn fdb_stream_close_batch(ctx: *mut fdb_ctx, ids: *const u32, n: usize) -> c_int;

    // ---- several GPUs behind one handle ----
    pub fn fdb_multi_create(devices: *const c_int, n_devices: c_int, out: *mut *mut fdb_multi) -> c_int;
    pub fn fdb_multi_destroy(m: *mut fdb_multi);
    pub fn fdb_multi_device_count(m: *const fdb_multi) -> c_int;
    pub fn fdb_multi_last_error(m: *const fdb_multi) -> *const c_char;
    pub fn fdb_multi_inflate_batch(
        m: *mut fdb_multi, in_base: *const u8, in_off: *const u64, in_len: *const u64, out_base: *mut u8, out_off: *const u64,
        out_cap: *const u64, out_len: *mut u64, consumed: *mut u64, status: *mut i32, n: usize, flags: u32,
    ) -> c_int;
    pub fn fdb_multi_deflate_ultrafast_batch(
        m: *mut fdb_multi, in_base: *const u8, in_off: *const u64, in_len: *const u64, out_base: *mut u8, out_off: *const u64,
        out_cap: *const u64, out_len: *mut u64, status: *mut i32, n: usize,
    ) -> c_int;
    pub fn fdb_multi_deflate_stored_batch(
        m: *mut fdb_multi, in_base: *const u8, in_off: *const u64, in_len: *const u64, out_base: *mut u8, out_off: *const u64,
        out_cap: *const u64, out_len: *mut u64, status: *mut i32, n: usize,
    ) -> c_int;
    pub fn fdb_multi_last_partition(m: *const fdb_multi, owner: *mut u32, n: usize) -> c_int;
}
