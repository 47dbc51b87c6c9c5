// deflate_uf.cuh -- K1/K2 fused: ultra-fast deflate, ONE WARP PER STREAM, single pass.
//
// Replaces the reference's UltraFastCompressor (src/compress/ultrafast.rs:16-181) as driven by
// compress_to_vec_ultra_fast (src/compress/mod.rs:313-317): constant 53-byte + 5-bit header,
// literals coded with the fixed PNG-trained Huffman code, zero runs coded as
// `lit0 ; floor((R-1)/258) x (sym285, dist 1) ; tail`, end-of-block, adler32.
// Output is byte-identical to the reference for one write_data() call over the whole input.
//
// The reference walks the input 8 bytes at a time with a sequential `run` counter.  Here each lane
// owns 16 bytes (two of the reference's chunks) of a 512-byte warp step and the sequential state
// becomes three warp scans:
//   1. run carry     x -> (all-zero ? x + len : trailing zeros)      (function composition scan)
//   2. bit offsets   sum of token bit lengths                        (prefix sum)
//   3. partial word  x -> (completed a word ? my tail bits : x | my bits)   (function composition)
// Every output word is written exactly once, by the lane that completes it, into a shared-memory
// staging window that is flushed with coalesced 128-byte stores: no atomics, no pre-zeroed output.
// adler32 rides along as position-weighted sums of the bytes already in registers (adler.cuh).
//
// Which zero bytes are "run bytes" (reference ultrafast.rs:97-157, restated per 8-byte chunk):
// with x = pending run length entering the chunk, an all-zero chunk extends the run; otherwise the
// leading zero bytes belong to the run iff x > 0, and the trailing zero bytes start a new run.
// Bytes past the last whole 8-byte chunk are always literals.
#pragma once
#include "simt.h"
#include "fdb_common.h"
#include <type_traits>

#include "adler.cuh"

namespace fdb {

struct UfEncTables {
    uint2 lit[512];          // {code, nbits} of every literal (reference tables.rs:7-25); 256..511 = {0, 0}
    uint32_t tail_tok[258];  // run tail r = (R-1) mod 258: bits | nbits << 24   (ultrafast.rs:54-64)
    uint32_t header[14];     // the 54 header bytes (ultrafast.rs:82-86), little-endian words
};

static const uint32_t UF_CODE285_DIST1 = 343u;  // sym 285 (9 bits) followed by the 1-bit distance code 0
static const uint32_t UF_EOB_CODE = 2303u;      // sym 256, 12 bits
static const uint32_t UF_HEADER_BITS = 53u * 8u + 5u;

static const int DEFLATE_WARPS = 8;
#ifndef DEFLATE_MIN_CTAS
#define DEFLATE_MIN_CTAS 4
#endif
static const uint32_t STG_WORDS = 320;  // >= 512 bytes * 18 bits / 32 + slack

struct DeflateSmem {
    uint2 lit[512];
    uint32_t tail_tok[258];
    uint32_t header[14];
    uint32_t stg[DEFLATE_WARPS][STG_WORDS];
};

// high bit of every byte that is non-zero
FDB_DEVICE uint64_t nonzero_bytes(uint64_t c) {
    return (((c & 0x7f7f7f7f7f7f7f7full) + 0x7f7f7f7f7f7f7f7full) | c) & 0x8080808080808080ull;
}
FDB_DEVICE uint32_t ctz64(uint64_t v) {  // v != 0
    uint32_t lo = (uint32_t)v;
    return lo ? simt::ffs(lo) - 1u : 32u + simt::ffs((uint32_t)(v >> 32)) - 1u;
}
FDB_DEVICE uint32_t clz64(uint64_t v) {  // v != 0
    uint32_t hi = (uint32_t)(v >> 32);
    return hi ? simt::clz(hi) : 32u + simt::clz((uint32_t)v);
}

// Packs (value, nbits <= 32) groups LSB-first into the warp's staging window.  Every word a lane
// completes is stored by that lane; the first one may lack the bits earlier lanes put below this
// lane's start, which the caller ORs in afterwards (patch_first).
#ifndef DF_WIDE_MUL
#define DF_WIDE_MUL 0
#endif
// DF_FMA_ADD: the packer's three additions (bit count, the word-full step back, the staging address) as multiply-adds
// x * 1 + y, the 1 read from constant memory so that the compiler keeps the multiply.  IMAD runs on the FMA pipe, which
// idles at 19 % while the integer ALU pipe is the encoder's binding unit (74 %): 1.266 -> 1.245 ms.  What did NOT help
// (profiles/r04_variants.txt): the same for the pair widths of word_pairs, for the constant additions of chunk_plan, for
// the head's bit count, for only the predicated two or only the sum, and for the row pointer step of K4's lane reader.
#ifndef DF_FMA_ADD
#define DF_FMA_ADD 1
#endif
#if DF_FMA_ADD && !defined(FDB_EMUL)
__constant__ uint32_t df_one = 1u;
FDB_DEVICE uint32_t df_add(uint32_t a, uint32_t b) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(df_one), "r"(b));
    return r;
}
#else
template <class T>
FDB_DEVICE T df_add(T a, uint32_t b) { return a + b; }  // (a shared-window address is wider than 32 bits in the emulator)
#endif
struct BitPacker {
    uint32_t lo, hi;     // accumulator: bits [accn) of lo|hi<<32 are valid
    uint32_t accn;       // < 32 between emits
    simt::saddr wa;      // staging address of the word `lo` maps to
    FDB_MEMBER void emit(uint32_t v, uint32_t n) {  // v < 2^n, n <= 32
#if DF_WIDE_MUL && !defined(FDB_EMUL)
        // the 64-bit shift as one wide multiply on the FMA pipe (the ALU pipe is the binding one); the factor is made
        // opaque so that the compiler does not turn the multiply back into two funnel shifts
        uint32_t m = 1u << accn;
        asm volatile("" : "+r"(m));
        const uint64_t sh = (uint64_t)v * m;
#else
        const uint64_t sh = (uint64_t)v << accn;
#endif
        lo |= (uint32_t)sh;
        hi = (uint32_t)(sh >> 32);  // (hi carries nothing between emits: accn < 32)
        accn = df_add(accn, n);
        if (accn >= 32) {
            simt::sts32(wa, lo);
            wa = df_add(wa, 4u);
            lo = hi;
            accn = df_add(accn, 0xffffffe0u);  // - 32 (mod 2^32)
        }
    }
};

// What one 8-byte chunk owes besides its plain literals, in stream order:  head | literals | tail.
//   keep    = bit j set: byte j is coded as a literal (clear: swallowed by a run, or past the end)
//   head    = tokens that close a run entering the chunk:   [sym285+dist]? [run tail]   (<= 28 bits)
//             or, for an all-zero chunk, every token the chunk owes: [lit0]? [sym285+dist]? [run tail]?
//   tail    = tokens that open a run on the trailing zeros:  [lit0] [run tail if the run ends here]
struct ChunkPlan {
    uint32_t keep;
    uint32_t head_v, head_n;
    uint32_t tail_v, tail_n;
};

// kind 2: whole chunk inside the run-logic prefix (reference ultrafast.rs:98-153).  The pending run
// entering the chunk is given as (r, pend) = (length mod 258, length > 0) -- all the encoder ever needs
// of it -- and is replaced by the run leaving the chunk; cont = the byte after the chunk is a zero that
// extends the run.  kind 1: the final partial chunk, `rem` literal bytes (ultrafast.rs:159-164).
// kind 0: nothing.
//
// A run of R bytes is `lit0 ; floor((R-1)/258) x (sym285, dist 1) ; tail((R-1) mod 258)`, and those
// tokens are owned by run offsets 0, 258*j and R-1.  With z = the zero bytes at the start of this chunk
// that belong to the run (run offsets x .. x+z-1), the chunk owes sym285 iff a positive multiple of 258
// lies in that range, and the tail iff the run ends inside the chunk.
FDB_DEVICE void chunk_plan(ChunkPlan& t, uint64_t nz /* the chunk, little-endian */, uint32_t kind, uint32_t rem, uint32_t& r, uint32_t& pend,
                           bool cont, const uint32_t* tail_tok) {
    // straight-line code: nearly every warp step has lanes in every case, so branches only add replays
    const bool k2 = kind == 2;
    const bool allz = nz == 0;
    const uint32_t lead = allz ? 8u : (ctz64(nz) >> 3);   // leading / trailing zero BYTES of the chunk
    const uint32_t trail = allz ? 0u : (clz64(nz) >> 3);
    const bool pending = pend != 0;
    const uint32_t z = (k2 && (allz || pending)) ? lead : 0u;  // zero bytes at the start that belong to the run
    const uint32_t q = r + z - 1u;                             // run offset of the last of them, before the wrap
    const bool wrap = z != 0 && q >= 258u;
    // sym 285 + distance 1 (:49-52): a positive multiple of 258 among the run offsets x .. x+z-1
    const bool cross = z != 0 && pending && (r == 0 || wrap);
    // the run ends inside this chunk (:54-64, :105-108)
    const bool ends = z != 0 && !(allz && cont);
    const uint32_t n0 = (k2 && allz && !pending) ? 2u : 0u;  // lit 0 opens a run inside an all-zero chunk (:46)
    const uint32_t tt = tail_tok[ends ? (wrap ? q - 258u : q) : 0u];  // (tail_tok[0] == 0: nothing owed)
    const uint32_t n1 = cross ? n0 + 10u : n0;
    t.head_v = (cross ? (UF_CODE285_DIST1 << n0) : 0u) | ((tt & 0xffffffu) << n1);
    t.head_n = n1 + (tt >> 24);
    // trailing zeros open a new run (:111, :130): lit 0, then the run's tail if it ends right here
    const bool has_tail = k2 && trail != 0;
    const uint32_t tt2 = tail_tok[(has_tail && !cont) ? trail - 1u : 0u];
    t.tail_v = (tt2 & 0xffffffu) << 2;
    t.tail_n = (has_tail ? 2u : 0u) + (tt2 >> 24);
    // one bit per byte that stays a literal: positions [z, 8 - trail); the final partial chunk keeps `rem`
    const uint32_t keep2 = allz ? 0u : ((0xffu << z) & (0xffu >> trail));
    const uint32_t keep1 = kind == 1 ? (1u << rem) - 1u : 0u;
    t.keep = k2 ? keep2 : keep1;
    const uint32_t r8 = r + 8u >= 258u ? r + 8u - 258u : r + 8u;
    r = k2 ? (allz ? r8 : trail) : 0u;
    pend = (k2 && (allz || trail != 0)) ? 1u : 0u;
}

// Literal pairs of one 32-bit input word: two (value, nbits) groups of <= 24 bits.  A byte that is not
// kept (swallowed by a run, or past the end) indexes the second half of the table, whose entries are
// {0, 0}: it vanishes from the stream without any per-token masking.  keep4 = keep bits of the 4 bytes.
struct PairTok {
    uint32_t v, n;
};
FDB_DEVICE void word_pairs(PairTok& a, PairTok& b, uint32_t w, uint32_t keep4, const uint2* lit) {
    // one 0/1 byte per input byte: 1 = not kept
    const uint32_t nk = ((~keep4 & 0xfu) * 0x00204081u) & 0x01010101u;
    // index = byte k of w | (byte k of nk) << 8      (selector nibbles 12..15: sign of an nk byte = 0)
    const uint2 e0 = lit[simt::prmt(w, nk, 0xcc40u)];
    const uint2 e1 = lit[simt::prmt(w, nk, 0xdd51u)];
    const uint2 e2 = lit[simt::prmt(w, nk, 0xee62u)];
    const uint2 e3 = lit[simt::prmt(w, nk, 0xff73u)];
    a.v = e0.x | (e1.x << e0.y);
    a.n = e0.y + e1.y;
    b.v = e2.x | (e3.x << e2.y);
    b.n = e2.y + e3.y;
}

FDB_DEVICE uint4 load16_guarded(const uint8_t* in, uint64_t g, uint64_t n, bool aligned) {
    if (g + 16 <= n && aligned) return simt::ldg128((const uint4*)(in + g));
    uint32_t w[4] = {0, 0, 0, 0};
    for (uint32_t j = 0; j < 16; j++)
        if (g + j < n) w[j >> 2] |= (uint32_t)simt::ldg8(in + g + j) << (8u * (j & 3u));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- segments: a long input encoded by several warps ------------------------------------------
// An input of many 64 KiB SEGMENTS is encoded in three passes.  COUNT: every segment finds the run that
// is pending where it starts (the zero bytes just before it: the reference's `run` counter at a chunk
// boundary is exactly their number), walks its warp steps without packing anything and reports how many
// bits it will emit, plus its adler32 partial sums.  A prefix sum over the segments of a stream turns the
// bit counts into bit offsets (split_scan_kernel, which also knows the final length and checksum then).
// WRITE: every segment encodes again, at its offset; the one word it shares with the segment before and
// the one it shares with the segment after are merged with atomicOr into words the scan has zeroed, every
// other word is stored exactly once as in the one-warp encoder.  DF_WHOLE is that one-warp encoder.
enum : int { DF_WHOLE = 0, DF_COUNT = 1, DF_WRITE = 2 };
static const uint64_t DF_SEG_BYTES = 64u << 10;              // multiple of the 512-byte warp step
static const uint64_t DF_SPLIT_MIN_BYTES = 4 * DF_SEG_BYTES;   // default: inputs of >= 256 KiB are split (never < 2 segments)
static const uint64_t DF_AUTO_SPLIT_BYTES = 16 * DF_SEG_BYTES;  // host-buffer calls turn the segment path on for a chunk that holds an input this long
static const uint32_t DF_NO_ITEM = 0xffffffffu;

struct DfSpan {
    uint64_t begin, end;  // input bytes of the segment: begin is a multiple of 512, end = n for the last one
    uint64_t bit_off;     // DF_WRITE: stream bit (0 = first bit of the zlib header) where the segment's tokens start
    uint32_t run_in;      // pending run entering the segment, in bytes
    uint32_t first, last; // the stream's first / last segment (header / end of block + checksum)
    uint32_t adler;       // DF_WRITE, last: the stream's checksum
};
struct DfSpanOut {
    uint64_t bits;        // DF_COUNT: bits of the segment's tokens
    uint64_t s1, s2;      // DF_COUNT: adler32 partial sums of the segment's bytes (s2 mod 65521)
};

// DF_WHOLE: one stream, one warp; returns the encoded length, or 0 with *status != ST_OK.
template <int MODE>
FDB_DEVICE uint64_t deflate_uf_run(const uint2* lit, const uint32_t* tail_tok, const uint32_t* header,
                                   uint32_t* stg, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap,
                                   int32_t* status, const DfSpan* sp, DfSpanOut* so) {
    const unsigned lane = simt::lane_id();
    const simt::saddr stg_s = simt::smem_addr(stg);
    const uint32_t oab = (uint32_t)((uintptr_t)out & 3u);  // out's offset inside its aligned word
    uint32_t* const obase = (uint32_t*)(out - oab);
    const bool in_aligned = ((uintptr_t)in & 15u) == 0;
    const uint64_t n8 = n & ~(uint64_t)7;
    const uint32_t rem = (uint32_t)(n - n8);
    bool overflow = false;
    const uint64_t cap_words = (cap + oab) >> 2;  // virtual words that end inside the caller's slot

    // guarded store of virtual word k (bytes 4k-oab .. 4k-oab+3 of the stream)
    auto store_word = [&](uint64_t k, uint32_t v) {
        uint64_t end_byte = 4 * (k + 1) - oab;  // stream byte index one past this word
        if (end_byte > cap) {
            overflow = true;
            return;
        }
        if (k == 0 && oab) {
            for (uint32_t j = oab; j < 4; j++) out[j - oab] = (uint8_t)(v >> (8u * j));
        } else {
            obase[k] = v;
        }
    };

    // ---- header (ultrafast.rs:81-91): 53 bytes + the low 5 bits of byte 53 ----
    uint64_t vbit = 8ull * oab + UF_HEADER_BITS;  // virtual bit cursor (bit 0 = bit 0 of obase[0])
    uint32_t wcarry = 0;                          // bits of the incomplete word below the cursor
    // DF_WRITE: the first word this segment completes also holds the last bits of the segment before
    bool merge_first = false;
    if (MODE == DF_WRITE && !sp->first) {
        vbit = 8ull * oab + sp->bit_off;
        merge_first = true;
    }
    if (MODE == DF_WHOLE || (MODE == DF_WRITE && sp->first)) {
        uint32_t hw = 0;  // virtual word `lane` of the header
        if (lane < 16) {
            for (uint32_t j = 0; j < 4; j++) {
                int32_t sb = (int32_t)(4 * lane + j) - (int32_t)oab;
                if (sb >= 0 && sb < 54) hw |= ((header[sb >> 2] >> (8u * (sb & 3))) & 0xffu) << (8u * j);
            }
        }
        uint32_t full = (uint32_t)(vbit >> 5);
        if (lane < full) store_word(lane, hw);
        wcarry = simt::shfl(hw, full) & ((1u << (vbit & 31)) - 1u);
    }

    // ---- data (ultrafast.rs:94-167) ----
    AdlerAcc ad = {0, 0};
    uint32_t run_carry = MODE == DF_WHOLE ? 0u : sp->run_in;
    const uint64_t iters = (n + 511) >> 9;                                   // warp steps of the whole stream
    const uint64_t it_begin = MODE == DF_WHOLE ? 0 : sp->begin >> 9;         // ... and of this call
    const uint64_t it_end = MODE == DF_WHOLE ? iters : (sp->end + 511) >> 9;
    uint4 nxt = make_uint4(0, 0, 0, 0);
    if (it_begin < it_end) nxt = load16_guarded(in, (it_begin << 9) + (uint64_t)lane * 16, n, in_aligned);
    // "is the first byte of the next step zero": needed by lane 31 long before that step's data, so it
    // is fetched one step earlier than the data itself
    uint32_t nfb_next = it_begin + 1 < iters ? simt::ldg8(in + ((it_begin + 1) << 9)) : 1u;  // (the raw byte: compared when it is used)
    // One 512-byte warp step.  FULL: every chunk of the step is a whole chunk inside the run-logic prefix
    // (all steps of a stream but the last one or two), which removes every end-of-input test.
    auto step = [&](uint64_t it, auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        const uint64_t base = it << 9;
        const uint64_t g = base + (uint64_t)lane * 16;
        uint4 q = nxt;
        if (it + 1 < iters) nxt = load16_guarded(in, g + 512, n, in_aligned);
        const uint32_t nfb = nfb_next;
        if (it + 2 < iters) nfb_next = simt::ldg8(in + base + 1024);

        // adler partial sums
        if (MODE == DF_WRITE) {
        } else if (FULL || g + 16 <= n) {
            adler_add16(ad, q, g);
            if ((it & 63) == 63) adler_fold(ad);
        } else if (g < n) {
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
            for (uint32_t j = 0; g + j < n; j++) adler_add1(ad, (w[j >> 2] >> (8u * (j & 3u))) & 0xffu, g + j);
        }

        const uint64_t c0 = ((uint64_t)q.y << 32) | q.x, c1 = ((uint64_t)q.w << 32) | q.z;
        const uint64_t nz0 = c0, nz1 = c1;  // (only "== 0", ctz and clz are taken of these)
        // chunk kinds: 2 = whole chunk in the run-logic prefix, 1 = the final partial chunk, 0 = past the end
        uint32_t k0 = 2u, k1 = 2u;
        if (!FULL) {  // only the last warp step(s) of a stream
            k0 = (g + 8 <= n8) ? 2u : (g == n8 && rem) ? 1u : 0u;
            k1 = (g + 16 <= n8) ? 2u : (g + 8 == n8 && rem) ? 1u : 0u;
        }

        // 1. pending run entering this lane.  Per lane the run length maps as f(x) = a ? x + 16 : b
        //    (a: both chunks all-zero; b: zeros the lane leaves pending), so the value entering lane i is
        //    b of the nearest lane below that is not all-zero, plus 16 per all-zero lane in between.
        uint32_t fa, fb;
        {
            uint32_t a0 = 0, b0 = 0, a1 = 0, b1 = 0;
            if (k0 == 2) {
                if (nz0 == 0) { a0 = 1; b0 = 8; } else { b0 = clz64(nz0) >> 3; }
            }
            if (k1 == 2) {
                if (nz1 == 0) { a1 = 1; b1 = 8; } else { b1 = clz64(nz1) >> 3; }
            }
            fa = a0 & a1;
            fb = a1 ? b0 + b1 : b1;
        }
        uint32_t x0;
        {
            const uint32_t zmask = simt::ballot(fa != 0);
            const uint32_t below = ~zmask & simt::lanemask_lt();  // lanes below me that reset the run
            const uint32_t j = below ? 31u - simt::clz(below) : 0u;
            const uint32_t bj = simt::shfl(fb, j);
            x0 = below ? bj + 16u * (lane - 1u - j) : run_carry + 16u * lane;
            const uint32_t out31 = simt::shfl(fa ? x0 + 16u : fb, 31);
            run_carry = out31;
        }

        // does the byte after each chunk continue a run?
        uint32_t first_byte_zero = (k0 == 2 && (q.x & 0xffu) == 0) ? 1u : 0u;
        uint32_t next_lane_first = simt::shfl_down(first_byte_zero, 1);
        if (lane == 31) next_lane_first = (it + 1 < iters && base + 512 + 8 <= n8 && nfb == 0) ? 1u : 0u;
        const bool cont0 = (k1 == 2) && ((q.z & 0xffu) == 0);
        const bool cont1 = next_lane_first != 0;

        // 2. tokens (looked up once, kept in registers as literal pairs), bit lengths and offsets
        ChunkPlan t0, t1;
        uint32_t rr = x0 % 258u, pend = x0 > 0 ? 1u : 0u;
        chunk_plan(t0, nz0, k0, rem, rr, pend, cont0, tail_tok);
        chunk_plan(t1, nz1, k1, rem, rr, pend, cont1, tail_tok);
        PairTok p[8];
        word_pairs(p[0], p[1], q.x, t0.keep, lit);
        word_pairs(p[2], p[3], q.y, t0.keep >> 4, lit);
        word_pairs(p[4], p[5], q.z, t1.keep, lit);
        word_pairs(p[6], p[7], q.w, t1.keep >> 4, lit);
        const uint32_t my_bits = ((p[0].n + p[1].n + p[2].n) + (p[3].n + p[4].n + p[5].n)) +
                                 ((p[6].n + p[7].n + t0.head_n) + (t0.tail_n + t1.head_n + t1.tail_n));
        const uint32_t incl_bits = simt::scan_incl_add(my_bits);
        const uint32_t total_bits = simt::shfl(incl_bits, 31);
        if (MODE == DF_COUNT) {
            vbit += total_bits;
            return;
        }
        const uint64_t o = vbit + (incl_bits - my_bits);
        const uint64_t wbase = vbit >> 5;

        // 3. pack
        BitPacker bp;
        bp.lo = bp.hi = 0;
        bp.accn = (uint32_t)(o & 31);
        const uint32_t w_first = (uint32_t)((o >> 5) - wbase);
        const simt::saddr wa_first = stg_s + 4u * w_first;
        bp.wa = wa_first;
        bp.emit(t0.head_v, t0.head_n);
        bp.emit(p[0].v, p[0].n);
        bp.emit(p[1].v, p[1].n);
        bp.emit(p[2].v, p[2].n);
        bp.emit(p[3].v, p[3].n);
        bp.emit(t0.tail_v | (t1.head_v << t0.tail_n), t0.tail_n + t1.head_n);  // together <= 30 bits
        bp.emit(p[4].v, p[4].n);
        bp.emit(p[5].v, p[5].n);
        bp.emit(p[6].v, p[6].n);
        bp.emit(p[7].v, p[7].n);
        bp.emit(t1.tail_v, t1.tail_n);

        // partial-word carry: g(x) = m ? x | v : v, with m = "this lane did not complete its first word".
        // Usually every lane completes a word, and the carry into a lane is just its neighbour's tail bits.
        const uint32_t gm = (bp.wa == wa_first) ? 1u : 0u, gv = bp.lo;
        uint32_t carry_in;
        if (!simt::any(gm != 0)) {
            carry_in = simt::shfl_up(gv, 1);
            if (lane == 0) carry_in = wcarry;
            wcarry = simt::shfl(gv, 31);
        } else {
            uint32_t im = gm, iv = gv;
#pragma unroll
            for (unsigned d = 1; d < 32; d <<= 1) {
                uint32_t pm = simt::shfl_up(im, d), pv = simt::shfl_up(iv, d);
                if (lane >= d) {
                    iv = im ? (pv | iv) : iv;
                    im = im & pm;
                }
            }
            uint32_t em = simt::shfl_up(im, 1), ev = simt::shfl_up(iv, 1);
            if (lane == 0) { em = 1; ev = 0; }
            carry_in = em ? (wcarry | ev) : ev;
            uint32_t lm = simt::shfl(im, 31), lv = simt::shfl(iv, 31);
            wcarry = lm ? (wcarry | lv) : lv;
        }
        // the first word this lane stored lacks the bits below its start
        if (!gm) simt::sts32(wa_first, simt::lds32(wa_first) | carry_in);
        simt::syncwarp();

        // 4. flush the completed words, coalesced
        vbit += total_bits;
        const uint32_t nwords = (uint32_t)((vbit >> 5) - wbase);
        // (the header is 53 bytes, so these are never the stream's first, possibly partial, word)
        {
            const uint32_t fit = cap_words > wbase ? (uint32_t)(cap_words - wbase < nwords ? cap_words - wbase : nwords) : 0u;
            uint32_t* const dst = obase + wbase;
            if (fit < nwords) overflow = true;
            // a step normally completes ~55 words: two straight-line rounds, then a loop for the rest
            if (MODE == DF_WRITE && merge_first && fit > 0) {
                if (lane == 0) simt::atomic_or(dst, simt::lds32(stg_s));
                else if (lane < fit) dst[lane] = simt::lds32(stg_s + 4u * lane);
                merge_first = false;
            } else if (lane < fit) dst[lane] = simt::lds32(stg_s + 4u * lane);
            if (lane + 32 < fit) dst[lane + 32] = simt::lds32(stg_s + 4u * lane + 128u);
#pragma unroll 1
            for (uint32_t k = lane + 64; k < fit; k += 32) dst[k] = simt::lds32(stg_s + 4u * k);
        }
        simt::syncwarp();
    };
    for (uint64_t it = it_begin; it < it_end; it++) {
        if ((it << 9) + 512 <= n8)
            step(it, std::true_type{});
        else
            step(it, std::false_type{});
    }
    if (MODE == DF_COUNT) {
        so->bits = vbit - (8ull * oab + UF_HEADER_BITS);
        so->s1 = simt::reduce_add(ad.s1);
        so->s2 = simt::reduce_add(ad.s2 % ADLER_MOD) % ADLER_MOD;
        return 0;
    }
    if (MODE == DF_WRITE && !sp->last) {
        // the word under the cursor is shared with the next segment: merge my bits into it
        if (lane == 0 && (vbit & 31) && ((vbit >> 5) + 1) * 4 - oab <= cap) simt::atomic_or(obase + (vbit >> 5), wcarry);
        return 0;
    }

    // ---- finish (ultrafast.rs:170-181): EOB, pad to a byte, adler32 big-endian ----
    const uint32_t adler = MODE == DF_WRITE ? sp->adler : adler_finish_warp(ad, n);
    overflow = simt::any(overflow);
    uint64_t total_len = 0;
    {
        uint64_t acc = wcarry;
        uint32_t accn = (uint32_t)(vbit & 31);
        acc |= (uint64_t)UF_EOB_CODE << accn;
        accn += 12;
        accn = (accn + 7u) & ~7u;
        uint64_t vb = (vbit >> 5) * 4;           // virtual byte index of acc's byte 0
        uint32_t nbytes = accn >> 3;             // <= 6
        uint8_t tail[12];
        for (uint32_t j = 0; j < nbytes; j++) tail[j] = (uint8_t)(acc >> (8u * j));
        tail[nbytes + 0] = (uint8_t)(adler >> 24);
        tail[nbytes + 1] = (uint8_t)(adler >> 16);
        tail[nbytes + 2] = (uint8_t)(adler >> 8);
        tail[nbytes + 3] = (uint8_t)adler;
        nbytes += 4;
        total_len = vb + nbytes - oab;
        if (total_len > cap) overflow = true;
        if (!overflow && lane == 0) {
            for (uint32_t j = 0; j < nbytes; j++)
                if (vb + j >= oab) out[vb + j - oab] = tail[j];
        }
    }
    *status = overflow ? ST_OUTPUT_BUFFER_TOO_SMALL : ST_OK;
    return overflow ? 0 : total_len;
}

FDB_DEVICE uint64_t deflate_uf_stream(const uint2* lit, const uint32_t* tail_tok, const uint32_t* header,
                                      uint32_t* stg, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap,
                                      int32_t* status) {
    return deflate_uf_run<DF_WHOLE>(lit, tail_tok, header, stg, in, n, out, cap, status, nullptr, nullptr);
}

// ---- work order: longest inputs first ------------------------------------------------------------
// The persistent kernel hands inputs to warps as they become free.  With ragged inputs (BASELINE configs[4]: 64 KiB ..
// 16 MiB) a long input that is picked up late keeps one warp busy after every other has finished; in longest-first
// order the counter balances them (measured on 8192 such inputs: 368-384 GB/s in caller order, 409-433 GB/s longest
// first).  Two tiny kernels sort the batch by size class (eighths of an octave, descending): a histogram and a scatter.
static const uint32_t DF_ORDER_CLASSES = 512;
FDB_DEVICE uint32_t df_size_class(uint64_t n) {  // 0 = the longest
    if (n < 8) return DF_ORDER_CLASSES - 1u - (uint32_t)n;
    const uint32_t lg = 63u - clz64(n);                           // >= 3
    const uint32_t frac = (uint32_t)(n >> (lg - 3)) & 7u;        // the three bits below the leading one
    return DF_ORDER_CLASSES - 1u - (8u * lg + frac);             // 8 * 63 + 7 = 511
}
FDB_GLOBAL void order_hist_kernel(const uint64_t* len, uint32_t n, uint32_t* hist) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        simt::atomic_add(&hist[df_size_class(len[i])], 1u);
}
FDB_GLOBAL void order_scatter_kernel(const uint64_t* len, uint32_t n, const uint32_t* hist, uint32_t* cursor, uint32_t* order) {
    FDB_SHARED uint32_t base[DF_ORDER_CLASSES];
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (uint32_t c = 0; c < DF_ORDER_CLASSES; c++) {
            base[c] = acc;
            acc += hist[c];
        }
    }
    simt::syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t c = df_size_class(len[i]);
        order[base[c] + simt::atomic_add(&cursor[c], 1u)] = i;
    }
}

FDB_GLOBAL void FDB_LAUNCH_BOUNDS(DEFLATE_WARPS * 32, DEFLATE_MIN_CTAS)
    deflate_uf_kernel(DeflateBatch b, const UfEncTables* tables, uint32_t* next, const uint32_t* split_item0,
                      const uint32_t* order) {
    FDB_SHARED DeflateSmem s;
    for (uint32_t i = threadIdx.x; i < 512; i += blockDim.x) s.lit[i] = tables->lit[i];
    for (uint32_t i = threadIdx.x; i < 258; i += blockDim.x) s.tail_tok[i] = tables->tail_tok[i];
    for (uint32_t i = threadIdx.x; i < 14; i += blockDim.x) s.header[i] = tables->header[i];
    simt::syncthreads();
    const unsigned lane = simt::lane_id();
    uint32_t* stg = s.stg[simt::warp_in_block()];
    // every warp's first input is fixed (input c + grid * w for warp w of CTA c): a batch of about one input per
    // resident warp then loads every SM alike (see inflate_uf_kernel); further inputs come from the counter
    const uint32_t slots = gridDim.x * DEFLATE_WARPS;
    bool first = true;
    for (;;) {
        uint32_t i = blockIdx.x + gridDim.x * simt::warp_in_block();
        if (!first) {
            if (lane == 0) i = slots + simt::atomic_add(next, 1u);
            i = simt::shfl(i, 0);
        }
        first = false;
        if (i >= b.n) break;
        if (order) i = order[i];
        if (split_item0 && split_item0[i] != DF_NO_ITEM) continue;  // encoded segment by segment (below)
        int32_t st = ST_OK;
        uint64_t len = deflate_uf_stream(s.lit, s.tail_tok, s.header, stg, b.in_base + b.in_off[i], b.in_len[i],
                                         b.out_base + b.out_off[i], b.out_cap[i], &st);
        if (lane == 0) {
            b.out_len[i] = len;
            b.status[i] = st;
        }
    }
}

// ---- segment path ------------------------------------------------------------------------------
struct DfItem {
    uint32_t stream, j;
    uint32_t run_in;   // count pass
    uint32_t skip;     // scan: the stream does not fit its slot, nothing is written
    uint64_t bits;     // count pass
    uint64_t s1, s2;   // count pass
    uint64_t bit_off;  // scan
};
struct DfSplit {
    DfItem* items;
    uint32_t item_cap;
    uint32_t* n_items;
    uint32_t* item0;    // [n] first item of stream i, or DF_NO_ITEM
    uint32_t* nseg;     // [n]
    uint32_t* adler;    // [n] scan: checksum of the whole input
    uint32_t* next_count;
    uint32_t* next_scan;
    uint32_t* next_write;
    uint64_t min_bytes;  // inputs at least this long are split ...
    const uint64_t* total;  // ... if they are also long against the batch: sum of the batch's input lengths
    uint32_t slots;         // warps the device runs at once
};

FDB_GLOBAL void deflate_uf_plan_kernel(DeflateBatch b, DfSplit sp) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    sp.item0[i] = DF_NO_ITEM;
    sp.nseg[i] = 0;
    const uint64_t n = b.in_len[i];
    if (n < deflate_split_threshold(sp.min_bytes, *sp.total, b.n, sp.slots) || n < 2 * DF_SEG_BYTES) return;
    const uint64_t S = n / DF_SEG_BYTES;  // the last segment also takes the remainder
    if (S > 0x7fffffffull) return;
    const uint32_t base = simt::atomic_add(sp.n_items, (uint32_t)S);
    if ((uint64_t)base + S > sp.item_cap) return;  // scratch exhausted: one warp encodes this stream
    sp.item0[i] = base;
    sp.nseg[i] = (uint32_t)S;
    for (uint32_t j = 0; j < (uint32_t)S; j++) {
        sp.items[base + j].stream = i;
        sp.items[base + j].j = j;
        sp.items[base + j].skip = 1;
    }
}
FDB_DEVICE uint32_t df_item_count(const DfSplit& sp) {
    const uint32_t c = *sp.n_items;
    return c < sp.item_cap ? c : sp.item_cap;
}
FDB_DEVICE bool df_item_live(const DfSplit& sp, uint32_t idx, const DfItem& it, uint32_t n) {
    return it.stream < n && sp.item0[it.stream] != DF_NO_ITEM && sp.item0[it.stream] + it.j == idx;
}
FDB_DEVICE void df_span_of(DfSpan& span, uint32_t j, uint32_t S, uint64_t n) {
    span.begin = (uint64_t)j * DF_SEG_BYTES;
    span.end = j + 1 == S ? n : span.begin + DF_SEG_BYTES;
    span.first = j == 0;
    span.last = j + 1 == S;
    span.bit_off = 0;
    span.run_in = 0;
    span.adler = 0;
}

// The run pending at byte `pos` (a multiple of 512): the zero bytes just before it.  The reference's
// counter (ultrafast.rs:97-131) is `run + 8` over an all-zero chunk and the chunk's trailing zeros
// otherwise, i.e. at a chunk boundary it is the length of the zero suffix of everything before.
FDB_DEVICE uint32_t pending_run_before(const uint8_t* in, uint64_t pos, bool in_aligned) {
    const unsigned lane = simt::lane_id();
    uint64_t x = 0;
    while (pos > 0) {  // (pos stays a multiple of 512)
        const uint64_t g = pos - 16u * (lane + 1u);  // lane 0 holds the 16 bytes just before pos
        const uint4 q = in_aligned ? simt::ldg128((const uint4*)(in + g)) : load16_guarded(in, g, pos, false);
        const bool nz = (q.x | q.y | q.z | q.w) != 0;
        const uint32_t m = simt::ballot(nz);
        if (m) {
            const uint32_t l = simt::ffs(m) - 1u;
            const uint32_t w3 = simt::shfl(q.w, l), w2 = simt::shfl(q.z, l), w1 = simt::shfl(q.y, l), w0 = simt::shfl(q.x, l);
            const uint32_t tz = w3 ? simt::clz(w3) >> 3 : w2 ? 4u + (simt::clz(w2) >> 3) : w1 ? 8u + (simt::clz(w1) >> 3) : 12u + (simt::clz(w0) >> 3);
            x += 16u * l + tz;
            break;
        }
        x += 512;
        pos -= 512;
    }
    // all the encoder ever uses of a run length is "mod 258" and "> 0" (chunk_plan): keep those, stay small
    return (uint32_t)(x % 258u) + (x >= 258u ? 258u : 0u);
}

FDB_GLOBAL void FDB_LAUNCH_BOUNDS(DEFLATE_WARPS * 32, DEFLATE_MIN_CTAS)
    deflate_uf_split_count_kernel(DeflateBatch b, const UfEncTables* tables, DfSplit sp) {
    FDB_SHARED DeflateSmem s;
    const uint32_t count = df_item_count(sp);
    if (count == 0) return;
    for (uint32_t i = threadIdx.x; i < 512; i += blockDim.x) s.lit[i] = tables->lit[i];
    for (uint32_t i = threadIdx.x; i < 258; i += blockDim.x) s.tail_tok[i] = tables->tail_tok[i];
    for (uint32_t i = threadIdx.x; i < 14; i += blockDim.x) s.header[i] = tables->header[i];
    simt::syncthreads();
    const unsigned lane = simt::lane_id();
    uint32_t* stg = s.stg[simt::warp_in_block()];
    for (;;) {
        uint32_t idx = 0;
        if (lane == 0) idx = simt::atomic_add(sp.next_count, 1u);
        idx = simt::shfl(idx, 0);
        if (idx >= count) break;
        DfItem& it = sp.items[idx];
        if (!df_item_live(sp, idx, it, b.n)) continue;
        const uint32_t i = it.stream;
        const uint8_t* in = b.in_base + b.in_off[i];
        const uint64_t n = b.in_len[i];
        DfSpan span;
        df_span_of(span, it.j, sp.nseg[i], n);
        span.run_in = pending_run_before(in, span.begin, ((uintptr_t)in & 15u) == 0);
        DfSpanOut so;
        int32_t st = ST_OK;
        deflate_uf_run<DF_COUNT>(s.lit, s.tail_tok, s.header, stg, in, n, b.out_base + b.out_off[i], b.out_cap[i], &st, &span, &so);
        if (lane == 0) {
            it.run_in = span.run_in;
            it.bits = so.bits;
            it.s1 = so.s1;
            it.s2 = so.s2;
        }
        simt::syncwarp();
    }
}

// One warp per split stream: bit offsets of its segments, final length, checksum, status; zeroes the
// words two segments share.
FDB_GLOBAL void deflate_uf_split_scan_kernel(DeflateBatch b, DfSplit sp) {
    const unsigned lane = simt::lane_id();
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = simt::atomic_add(sp.next_scan, 1u);
        i = simt::shfl(i, 0);
        if (i >= b.n) break;
        const uint32_t base = sp.item0[i];
        if (base == DF_NO_ITEM) continue;
        const uint32_t S = sp.nseg[i];
        const uint64_t n = b.in_len[i];
        uint8_t* out = b.out_base + b.out_off[i];
        const uint32_t oab = (uint32_t)((uintptr_t)out & 3u);
        uint64_t carry = UF_HEADER_BITS, t1 = 0, t2 = 0;
        for (uint32_t g = 0; g < S; g += 32) {
            const uint32_t j = g + lane;
            const bool have = j < S;
            const uint64_t bits = have ? sp.items[base + j].bits : 0ull;
            if (have) {
                t1 += sp.items[base + j].s1;
                t2 += sp.items[base + j].s2;
            }
            const uint64_t incl = simt::scan_incl_add(bits);
            if (have) sp.items[base + j].bit_off = carry + incl - bits;
            carry += simt::shfl(incl, 31);
        }
        // end of block (12 bits), pad to a byte, adler32 (ultrafast.rs:170-181)
        const uint64_t total_len = (carry + 12 + 7) / 8 + 4;
        const bool fits = total_len <= b.out_cap[i];
        t1 = simt::reduce_add(t1);
        t2 = simt::reduce_add(t2 % ADLER_MOD) % ADLER_MOD;
        const uint32_t s1m = (uint32_t)(t1 % ADLER_MOD), s2m = (uint32_t)t2, nm = (uint32_t)(n % ADLER_MOD);
        const uint32_t A = (1u + s1m) % ADLER_MOD;
        const uint32_t B = (uint32_t)(((uint64_t)nm + (uint64_t)nm * s1m % ADLER_MOD + ADLER_MOD - s2m) % ADLER_MOD);
        for (uint32_t j = lane; j < S; j += 32) {
            sp.items[base + j].skip = fits ? 0u : 1u;
            if (fits && j > 0) {  // the word segments j-1 and j share
                const uint64_t k = (8ull * oab + sp.items[base + j].bit_off) >> 5;
                *(uint32_t*)(out - oab + 4 * k) = 0u;
            }
        }
        if (lane == 0) {
            sp.adler[i] = (B << 16) | A;
            b.out_len[i] = fits ? total_len : 0;
            b.status[i] = fits ? ST_OK : ST_OUTPUT_BUFFER_TOO_SMALL;
        }
        simt::syncwarp();
    }
}

FDB_GLOBAL void FDB_LAUNCH_BOUNDS(DEFLATE_WARPS * 32, DEFLATE_MIN_CTAS)
    deflate_uf_split_write_kernel(DeflateBatch b, const UfEncTables* tables, DfSplit sp) {
    FDB_SHARED DeflateSmem s;
    const uint32_t count = df_item_count(sp);
    if (count == 0) return;
    for (uint32_t i = threadIdx.x; i < 512; i += blockDim.x) s.lit[i] = tables->lit[i];
    for (uint32_t i = threadIdx.x; i < 258; i += blockDim.x) s.tail_tok[i] = tables->tail_tok[i];
    for (uint32_t i = threadIdx.x; i < 14; i += blockDim.x) s.header[i] = tables->header[i];
    simt::syncthreads();
    const unsigned lane = simt::lane_id();
    uint32_t* stg = s.stg[simt::warp_in_block()];
    for (;;) {
        uint32_t idx = 0;
        if (lane == 0) idx = simt::atomic_add(sp.next_write, 1u);
        idx = simt::shfl(idx, 0);
        if (idx >= count) break;
        const DfItem& it = sp.items[idx];
        if (!df_item_live(sp, idx, it, b.n) || it.skip) continue;
        const uint32_t i = it.stream;
        const uint8_t* in = b.in_base + b.in_off[i];
        const uint64_t n = b.in_len[i];
        DfSpan span;
        df_span_of(span, it.j, sp.nseg[i], n);
        span.run_in = it.run_in;
        span.bit_off = it.bit_off;
        span.adler = sp.adler[i];
        int32_t st = ST_OK;
        deflate_uf_run<DF_WRITE>(s.lit, s.tail_tok, s.header, stg, in, n, b.out_base + b.out_off[i], b.out_cap[i], &st, &span, nullptr);
        simt::syncwarp();
    }
}

}  // namespace fdb
