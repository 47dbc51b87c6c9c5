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
struct BitPacker {
    uint32_t lo, hi;     // accumulator: bits [accn) of lo|hi<<32 are valid
    uint32_t accn;       // < 32 between emits
    simt::saddr wa;      // staging address of the word `lo` maps to
    FDB_MEMBER void emit(uint32_t v, uint32_t n) {  // v < 2^n, n <= 32
        const uint64_t sh = (uint64_t)v << accn;
        lo |= (uint32_t)sh;
        hi = (uint32_t)(sh >> 32);  // (hi carries nothing between emits: accn < 32)
        accn += n;
        if (accn >= 32) {
            simt::sts32(wa, lo);
            wa += 4;
            lo = hi;
            accn -= 32;
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

// One stream, one warp.  Returns the encoded length, or 0 with *status != ST_OK.
FDB_DEVICE uint64_t deflate_uf_stream(const uint2* lit, const uint32_t* tail_tok, const uint32_t* header,
                                      uint32_t* stg, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap,
                                      int32_t* status) {
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
    {
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
    uint32_t run_carry = 0;
    const uint64_t iters = (n + 511) >> 9;
    uint4 nxt = make_uint4(0, 0, 0, 0);
    if (iters > 0) nxt = load16_guarded(in, (uint64_t)lane * 16, n, in_aligned);
    // "is the first byte of the next step zero": needed by lane 31 long before that step's data, so it
    // is fetched one step earlier than the data itself
    uint32_t nfb_next = iters > 1 ? simt::ldg8(in + 512) : 1u;  // (the raw byte: compared when it is used)
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
        if (FULL || g + 16 <= n) {
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
            if (lane < fit) dst[lane] = simt::lds32(stg_s + 4u * lane);
            if (lane + 32 < fit) dst[lane + 32] = simt::lds32(stg_s + 4u * lane + 128u);
#pragma unroll 1
            for (uint32_t k = lane + 64; k < fit; k += 32) dst[k] = simt::lds32(stg_s + 4u * k);
        }
        simt::syncwarp();
    };
    for (uint64_t it = 0; it < iters; it++) {
        if ((it << 9) + 512 <= n8)
            step(it, std::true_type{});
        else
            step(it, std::false_type{});
    }

    // ---- finish (ultrafast.rs:170-181): EOB, pad to a byte, adler32 big-endian ----
    const uint32_t adler = adler_finish_warp(ad, n);
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

FDB_GLOBAL void FDB_LAUNCH_BOUNDS(DEFLATE_WARPS * 32, DEFLATE_MIN_CTAS)
    deflate_uf_kernel(DeflateBatch b, const UfEncTables* tables, uint32_t* next) {
    FDB_SHARED DeflateSmem s;
    for (uint32_t i = threadIdx.x; i < 512; i += blockDim.x) s.lit[i] = tables->lit[i];
    for (uint32_t i = threadIdx.x; i < 258; i += blockDim.x) s.tail_tok[i] = tables->tail_tok[i];
    for (uint32_t i = threadIdx.x; i < 14; i += blockDim.x) s.header[i] = tables->header[i];
    simt::syncthreads();
    const unsigned lane = simt::lane_id();
    uint32_t* stg = s.stg[simt::warp_in_block()];
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = simt::atomic_add(next, 1u);
        i = simt::shfl(i, 0);
        if (i >= b.n) break;
        int32_t st = ST_OK;
        uint64_t len = deflate_uf_stream(s.lit, s.tail_tok, s.header, stg, b.in_base + b.in_off[i], b.in_len[i],
                                         b.out_base + b.out_off[i], b.out_cap[i], &st);
        if (lane == 0) {
            b.out_len[i] = len;
            b.status[i] = st;
        }
    }
}

}  // namespace fdb
