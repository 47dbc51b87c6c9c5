// fdb_common.h -- definitions shared by host code and kernels.
#pragma once
#include <stdint.h>
#include <stddef.h>

namespace fdb {

// Per-stream status = the reference's DecompressionError in declaration order
// (reference src/decompress.rs:13-48); 17 = BoundedDecompressionError::OutputTooLarge (:1098).
enum Status : int32_t {
    ST_OK = 0,
    ST_BAD_ZLIB_HEADER = 1,
    ST_INSUFFICIENT_INPUT = 2,
    ST_INVALID_BLOCK_TYPE = 3,
    ST_INVALID_UNCOMPRESSED_BLOCK_LENGTH = 4,
    ST_INVALID_HLIT = 5,
    ST_INVALID_HDIST = 6,
    ST_INVALID_CODE_LENGTH_REPEAT = 7,
    ST_BAD_CODE_LENGTH_HUFFMAN_TREE = 8,
    ST_BAD_LITERAL_LENGTH_HUFFMAN_TREE = 9,
    ST_BAD_DISTANCE_HUFFMAN_TREE = 10,
    ST_INVALID_LITERAL_LENGTH_CODE = 11,
    ST_INVALID_DISTANCE_CODE = 12,
    ST_INPUT_STARTS_WITH_RUN = 13,
    ST_DISTANCE_TOO_FAR_BACK = 14,
    ST_WRONG_CHECKSUM = 15,
    ST_EXTRA_INPUT = 16,
    ST_OUTPUT_TOO_LARGE = 17,
    // compress side only: caller's slot is smaller than the encoded stream (the reference writes
    // into a growing Vec and cannot hit this)
    ST_OUTPUT_BUFFER_TOO_SMALL = 18,
    // PNG row filters (png_filter.cuh): a row's filter-type byte is not 0..4; bpp outside 1..8 or an unknown mode
    ST_PNG_BAD_FILTER_TYPE = 19,
    ST_PNG_BAD_GEOMETRY = 20,
    // PNG files (fdb_png_probe_batch / fdb_png_decode_files_batch): not a PNG / broken chunk structure / invalid IHDR,
    // a chunk whose CRC-32 does not match, a valid file of a kind that is not decoded here (interlaced)
    ST_PNG_BAD_FILE = 21,
    ST_PNG_BAD_CRC = 22,
    ST_PNG_UNSUPPORTED = 23,
    // internal: ultra-fast-format fast path declined the stream; the general kernel redoes it
    ST_PENDING_GENERAL = -1,
    // streaming decoders (fdb_stream_read_batch): the call stopped at a token boundary because the input given so
    // far ends inside the next item / the caller's room is full; both are "call again", not errors
    ST_STREAM_NEED_INPUT = -2,
    ST_STREAM_OUTPUT_FULL = -3,
};

enum : uint32_t { FLAG_IGNORE_ADLER32 = 1u };

// One batch of independent zlib streams, all pointers are DEVICE pointers.
struct InflateBatch {
    const uint8_t* in_base;
    const uint64_t* in_off;   // [n] byte offset of stream i in in_base
    const uint64_t* in_len;   // [n]
    uint8_t* out_base;
    const uint64_t* out_off;  // [n] byte offset of slot i in out_base
    const uint64_t* out_cap;  // [n] slot capacity (= maxlen of decompress_to_vec_bounded)
    uint64_t* out_len;        // [n] bytes produced
    uint64_t* consumed;       // [n] input bytes consumed (may be null)
    int32_t* status;          // [n]
    uint32_t n;
    uint32_t flags;
    uint32_t* general_out = nullptr;  // if set: receives how many streams the fast path handed to the general kernel
    uint32_t* split_out = nullptr;    // if set (and FDB_FLAG_SPLIT_LARGE): receives how many spans long streams were cut into
};

#if defined(__CUDACC__) && !defined(FDB_EMUL)
#define FDB_HD __host__ __device__ __forceinline__
#else
#define FDB_HD static inline
#endif

struct DeflateBatch {
    const uint8_t* in_base;
    const uint64_t* in_off;
    const uint64_t* in_len;
    uint8_t* out_base;
    const uint64_t* out_off;
    const uint64_t* out_cap;
    uint64_t* out_len;
    int32_t* status;
    uint32_t n;
};

// ---- decode-table entry formats (ours; only the decoded bytes have to match the reference) ----
// litlen entry:
//   [3:0] nbits  code bits consumed by the entry (both literals for a pair; the length CODE only)
//   [5:4] number of literals (1 or 2; 0 = not a literal entry)   [6] LEN  [7] EOB
//         none of LIT/LEN/EOB: code longer than the table
//   literal : [15:8] sym1  [23:16] sym2 (0 for a single)  [27:24] code bits of the first literal
//             [29:28] number of literals again, so that `e >> 28` is the byte count (bits 30,31 = 0)
//   length  : [10:8] extra-bit count  [24:16] base length      (so `e >> 28` == 0)
//   [31] QUIRK: fixed-code symbols 286/287, which the reference treats as end-of-block
//        (reference tables.rs:99-122 + decompress.rs:743-748)
enum : uint32_t {
    LL_LIT = 3u << 4,   // mask: non-zero for literal entries
    LL_LIT1 = 1u << 4,  // one literal
    LL_LIT2 = 2u << 4,  // two literals
    LL_LEN = 1u << 6,
    LL_EOB = 1u << 7,
    LL_QUIRK = 1u << 31,
};
// dist entry: [3:0] code bits  [7:4] extra-bit count  [8] VALID  [9] LONG (code > 9 bits)  [31:16] base
enum : uint32_t { DS_VALID = 1u << 8, DS_LONG = 1u << 9 };

// RFC 1951 length / distance symbol parameters (reference tables.rs:68-88 holds them as arrays)
FDB_HD uint32_t len_sym_extra(uint32_t sym) {  // sym in 257..285
    return (sym < 265 || sym == 285) ? 0u : ((sym - 261u) >> 2);
}
FDB_HD uint32_t len_sym_base(uint32_t sym) {
    if (sym < 265) return sym - 254u;
    if (sym == 285) return 258u;
    uint32_t e = (sym - 261u) >> 2;
    return ((4u + ((sym - 261u) & 3u)) << e) + 3u;
}
FDB_HD uint32_t dist_sym_extra(uint32_t d) { return d < 4 ? 0u : (d >> 1) - 1u; }  // d in 0..29
FDB_HD uint32_t dist_sym_base(uint32_t d) {
    if (d < 4) return d + 1u;
    uint32_t e = (d >> 1) - 1u;
    return ((2u + (d & 1u)) << e) + 1u;
}
FDB_HD uint32_t make_litlen_entry(uint32_t sym, uint32_t nbits) {
    if (sym < 256) return nbits | LL_LIT1 | (sym << 8) | (nbits << 24) | (1u << 28);
    if (sym == 256) return nbits | LL_EOB;
    if (sym < 286) return nbits | LL_LEN | (len_sym_extra(sym) << 8) | (len_sym_base(sym) << 16);
    return nbits | LL_EOB | LL_QUIRK;
}
// single-literal entry e1 (first literal, l1 bits) + a second literal sym2 of l2 bits -> pair entry
FDB_HD uint32_t make_litlen_pair(uint32_t e1, uint32_t sym2, uint32_t l1, uint32_t l2) {
    return (l1 + l2) | LL_LIT2 | (e1 & 0xff00u) | (sym2 << 16) | (l1 << 24) | (2u << 28);
}
FDB_HD uint32_t make_dist_entry(uint32_t sym, uint32_t nbits) {
    if (sym < 30) return nbits | DS_VALID | (dist_sym_extra(sym) << 4) | (dist_sym_base(sym) << 16);
    return nbits;  // symbols 30/31: invalid distance code
}

// ---- constant decode tables of the ultra-fast format (K4 only; index = the next 12 stream bits) ----
// The layouts put every field where the decode loops can use the entry AS IT IS: the bit count sits in the low five bits
// (a funnel shift takes its amount mod 32, so `bits >> n` needs no mask), and a count entry can be ADDED to the lane's
// position / byte accumulator in one instruction.
// UW "write table", u32:
//   literal  [4:0] bits consumed (1..12)   [12:5] [20:13] [28:21] up to three literal bytes in stream order (unused = 0)
//            [31:30] bytes produced (1..3): the leading literals whose codes fit in the 12 bits together
//   special  [31:30] == 0 and [4:0] == 0 (a special entry read as a second entry consumes and produces nothing):
//            [8:5] code bits  [11:9] extra-bit count  [12] end of block  [21:13] base length
//            - end of block, or a length token (every run: the write loop checks the byte before it there)
// UC "count table", u32 (0 = special: end of block, a length token longer than 12 bits or with distance bit 1):
//   [4:0] bits consumed by the 1..6 leading literals / the one short run    [13:10] bytes they produce
//   [27:24] bits of the first token alone   [28] RUN: the entry is a short-run token
//   [29] ENDNZ: its last byte is non-zero   [30] FIRSTNZ: its first token is a non-zero literal
//   A lane keeps  acc = position in its row + K4_BIAS | bytes << 10  and adds whole entries to it; what the flag bits
//   add up to above bit 24 is never read.
// UB "boundary table", u16: for a window whose count entry is a group of literals, bit k-1 is set when one of those
//   literals ends k bits into the window (the top set bit is the entry's bit count); 0 for every other window.  With it
//   the first token boundary at or after a given bit of an entry is two shifts and a find-first-set, and the bytes up to
//   there a population count: the count and warm-up walks cross their limit with whole entries and step back.
// All tables are stored BIT-REVERSED: the entry for index x lives at slot uf_slot(x) = the 12 index
// bits in reverse order.  The low index bits are the first code of the window and are far from uniform
// (39 % of the bench bytes are the 2-bit code 00), so a table in natural order sends most lanes to a
// few banks (measured 4.1-4.9 wavefronts per lookup); reversed, the bank is chosen by index bits 7..11.
// On the device this is BREV + one shift, cheaper than masking the index.
enum : uint32_t { UC_CNT_SHIFT = 10, UC_FIRST_SHIFT = 24, UC_RUN = 1u << 28, UC_ENDNZ = 1u << 29, UC_FIRSTNZ = 1u << 30 };
enum : uint32_t { UW_SPECIAL_SHIFT = 5, UW_EOB = 1u << 7 /* of the special fields, i.e. entry >> UW_SPECIAL_SHIFT */, UW_LITERAL_MIN = 1u << 30 };
FDB_HD uint32_t uf_slot(uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < 12; i++) r |= ((bits >> i) & 1u) << (11u - i);
    return r;
}

static const uint32_t ADLER_MOD = 65521u;

// Which inputs of a DEFLATE batch are worth cutting into segments: those that would be the tail of the batch (the count
// pass costs 60 % more work on the inputs it is spent on).
//  * A batch with fewer inputs than the device runs warps gives every input a warp at once and ends with its longest
//    input: everything above the configured minimum is cut (2048 inputs of 64 KiB..16 MiB: 3.3x faster, 256: 15x).
//  * A batch with more inputs than warps is balanced by the work counter (measured, 8192 such inputs: one warp per input
//    409-433 GB/s with the longest first, 368-384 GB/s in any order, against 343-360 GB/s with everything cut), so only
//    an input that holds an eighth of the whole batch is cut.
// (Inflate always cuts above its minimum: one warp decodes a long stream so slowly -- 16 MiB take 105 ms -- that the tail
// dominates even with 8192 streams: 205 GB/s on one warp each against 300-324 GB/s span by span.)
FDB_HD uint64_t deflate_split_threshold(uint64_t min_bytes, uint64_t total, uint32_t n_inputs, uint32_t slots) {
    const uint64_t t = n_inputs >= slots ? total / 8 : 0;
    return t > min_bytes ? t : min_bytes;
}

}  // namespace fdb
