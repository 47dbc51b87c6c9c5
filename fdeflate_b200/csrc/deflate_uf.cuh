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
#include "adler.cuh"

namespace fdb {

struct UfEncTables {
    uint32_t lit_tok[256];   // code | nbits << 16          (reference tables.rs:7-25)
    uint32_t tail_tok[258];  // run tail r = (R-1) mod 258: bits | nbits << 24   (ultrafast.rs:54-64)
    uint32_t header[14];     // the 54 header bytes (ultrafast.rs:82-86), little-endian words
};

static const uint32_t UF_CODE285_DIST1 = 343u;  // sym 285 (9 bits) followed by the 1-bit distance code 0
static const uint32_t UF_EOB_CODE = 2303u;      // sym 256, 12 bits
static const uint32_t UF_HEADER_BITS = 53u * 8u + 5u;

static const int DEFLATE_WARPS = 8;
#ifndef DEFLATE_MIN_CTAS
#define DEFLATE_MIN_CTAS 3
#endif
static const uint32_t STG_WORDS = 320;  // >= 512 bytes * 18 bits / 32 + slack

struct DeflateSmem {
    uint32_t lit_tok[256];
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

// Packs (value, nbits <= 32) groups LSB-first into the warp's staging window.  The first word a lane
// touches may be shared with earlier lanes, so it is kept back (first_word) until the partial-word
// carry scan has run; every later word is complete and owned by this lane.
struct BitPacker {
    uint64_t acc;
    uint32_t accn;
    uint32_t w;        // staging word the accumulator's low half maps to
    uint32_t w_first;  // first word this lane touches (possibly shared with earlier lanes)
    uint32_t first_word;
    uint32_t* stg;
    FDB_MEMBER void emit(uint32_t v, uint32_t n) {
        acc |= (uint64_t)v << accn;
        accn += n;
        if (accn >= 32) {
            uint32_t word = (uint32_t)acc;
            if (w == w_first)
                first_word = word;
            else
                stg[w] = word;
            w++;
            acc >>= 32;
            accn -= 32;
        }
    }
};

// The tokens of one 8-byte chunk, in stream order:  head group | 8 per-byte literal tokens | tail group.
//   lit[j]  = code | nbits << 16 of byte j, or 0 when byte j is swallowed by a run / is past the end
//   head    = tokens that close a run entering the chunk:   [sym285+dist]? [run tail]   (<= 28 bits)
//             or, for an all-zero chunk, every token the chunk owes: [lit0]? [sym285+dist]? [run tail]?
//   tail    = tokens that open a run on the trailing zeros:  [lit0] [run tail if the run ends here]
struct ChunkTokens {
    uint32_t lit[8];
    uint32_t head_v, head_n;
    uint32_t tail_v, tail_n;
};

FDB_DEVICE uint32_t lit_bits_sum(const ChunkTokens& t) {
    // codes are <= 12 bits, so eight of them cannot carry into the nbits field at bit 16
    uint32_t s = (t.lit[0] + t.lit[1] + t.lit[2]) + (t.lit[3] + t.lit[4] + t.lit[5]) + (t.lit[6] + t.lit[7]);
    return (s >> 16) + t.head_n + t.tail_n;
}

// kind 2: whole chunk inside the run-logic prefix (reference ultrafast.rs:98-153); x = pending run
// length entering the chunk; cont = the byte after the chunk is a zero that extends the run.
// kind 1: the final partial chunk, `rem` literal bytes (ultrafast.rs:159-164).  kind 0: nothing.
// Returns the pending run length leaving the chunk.
FDB_DEVICE uint32_t chunk_tokens(ChunkTokens& t, uint64_t c, uint64_t nz, uint32_t kind, uint32_t rem, uint32_t x,
                                 bool cont, const uint32_t* lit_tok, const uint32_t* tail_tok) {
    t.head_v = t.head_n = t.tail_v = t.tail_n = 0;
    const uint32_t lo = (uint32_t)c, hi = (uint32_t)(c >> 32);
    t.lit[0] = lit_tok[lo & 0xffu];
    t.lit[1] = lit_tok[(lo >> 8) & 0xffu];
    t.lit[2] = lit_tok[(lo >> 16) & 0xffu];
    t.lit[3] = lit_tok[lo >> 24];
    t.lit[4] = lit_tok[hi & 0xffu];
    t.lit[5] = lit_tok[(hi >> 8) & 0xffu];
    t.lit[6] = lit_tok[(hi >> 16) & 0xffu];
    t.lit[7] = lit_tok[hi >> 24];
    if (kind != 2) {
        const uint32_t keep = kind == 1 ? rem : 0u;
#pragma unroll
        for (uint32_t j = 0; j < 8; j++)
            if (j >= keep) t.lit[j] = 0;
        return 0;
    }
    if (nz == 0x8080808080808080ull && x == 0) return 0;  // no zero byte, no run pending: eight plain literals
    if (nz == 0) {  // all-zero chunk: extends (or opens) a run
        uint32_t v = 0, n = 0;
        if (x == 0) n = 2;  // lit 0 opens the run (ultrafast.rs:46); its code is 00
        const uint32_t r0 = x % 258u;
        if (r0 == 0 ? (x > 0) : (258u - r0 <= 7u)) {  // a multiple of 258 falls inside: sym 285 + distance 1 (:49-52)
            v |= UF_CODE285_DIST1 << n;
            n += 10;
        }
        if (!cont) {  // the run ends with this chunk (:54-64)
            const uint32_t tt = tail_tok[(x + 7u) % 258u];
            v |= (tt & 0xffffffu) << n;
            n += tt >> 24;
        }
        t.head_v = v;
        t.head_n = n;
#pragma unroll
        for (uint32_t j = 0; j < 8; j++) t.lit[j] = 0;
        return x + 8u;
    }
    const uint32_t lead = ctz64(nz) >> 3, trail = clz64(nz) >> 3;
    uint32_t first_lit = 0;
    if (x > 0 && lead > 0) {  // the run entering the chunk ends on its leading zeros (:105-108)
        uint32_t v = 0, n = 0;
        const uint32_t r0 = x % 258u;
        if (r0 == 0 || 258u - r0 <= lead - 1u) {
            v = UF_CODE285_DIST1;
            n = 10;
        }
        const uint32_t tt = tail_tok[(x + lead - 1u) % 258u];
        v |= (tt & 0xffffffu) << n;
        n += tt >> 24;
        t.head_v = v;
        t.head_n = n;
        first_lit = lead;
    }
    if (trail > 0) {  // trailing zeros open a new run (:111, :130)
        uint32_t n = 2;  // lit 0
        uint32_t v = 0;
        if (!cont) {
            const uint32_t tt = tail_tok[trail - 1u];
            v = (tt & 0xffffffu) << 2;
            n += tt >> 24;
        }
        t.tail_v = v;
        t.tail_n = n;
    }
    if (first_lit | trail) {
        // one bit per byte that stays a literal: positions [first_lit, 8 - trail)
        const uint32_t keep = (0xffu << first_lit) & (0xffu >> trail);
#pragma unroll
        for (uint32_t j = 0; j < 8; j++)
            if (!(keep & (1u << j))) t.lit[j] = 0;
    }
    return trail;
}

FDB_DEVICE void emit_chunk(BitPacker& bp, const ChunkTokens& t) {
    bp.emit(t.head_v, t.head_n);
#pragma unroll
    for (uint32_t k = 0; k < 8; k += 2) {
        const uint32_t a = t.lit[k], b = t.lit[k + 1];
        bp.emit((a & 0xffffu) | ((b & 0xffffu) << (a >> 16)), (a >> 16) + (b >> 16));
    }
    bp.emit(t.tail_v, t.tail_n);
}

FDB_DEVICE uint4 load16_guarded(const uint8_t* in, uint64_t g, uint64_t n, bool aligned) {
    if (g + 16 <= n && aligned) return simt::ldg128((const uint4*)(in + g));
    uint32_t w[4] = {0, 0, 0, 0};
    for (uint32_t j = 0; j < 16; j++)
        if (g + j < n) w[j >> 2] |= (uint32_t)simt::ldg8(in + g + j) << (8u * (j & 3u));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// One stream, one warp.  Returns the encoded length, or 0 with *status != ST_OK.
FDB_DEVICE uint64_t deflate_uf_stream(const uint32_t* lit_tok, const uint32_t* tail_tok, const uint32_t* header,
                                      uint32_t* stg, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap,
                                      int32_t* status) {
    const unsigned lane = simt::lane_id();
    const uint32_t oab = (uint32_t)((uintptr_t)out & 3u);  // out's offset inside its aligned word
    uint32_t* const obase = (uint32_t*)(out - oab);
    const bool in_aligned = ((uintptr_t)in & 15u) == 0;
    const uint64_t n8 = n & ~(uint64_t)7;
    const uint32_t rem = (uint32_t)(n - n8);
    bool overflow = false;

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
    for (uint64_t it = 0; it < iters; it++) {
        const uint64_t base = it << 9;
        const uint64_t g = base + (uint64_t)lane * 16;
        uint4 q = nxt;
        if (it + 1 < iters) nxt = load16_guarded(in, g + 512, n, in_aligned);

        // adler partial sums
        if (g + 16 <= n) {
            adler_add16(ad, q, g);
            if ((it & 63) == 63) adler_fold(ad);
        } else if (g < n) {
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
            for (uint32_t j = 0; g + j < n; j++) adler_add1(ad, (w[j >> 2] >> (8u * (j & 3u))) & 0xffu, g + j);
        }

        const uint64_t c0 = ((uint64_t)q.y << 32) | q.x, c1 = ((uint64_t)q.w << 32) | q.z;
        const uint64_t nz0 = nonzero_bytes(c0), nz1 = nonzero_bytes(c1);
        // chunk kinds: 2 = whole chunk in the run-logic prefix, 1 = the final partial chunk, 0 = past the end
        uint32_t k0 = 2u, k1 = 2u;
        if (base + 512 > n8) {  // only the last warp step of a stream
            k0 = (g + 8 <= n8) ? 2u : (g == n8 && rem) ? 1u : 0u;
            k1 = (g + 16 <= n8) ? 2u : (g + 8 == n8 && rem) ? 1u : 0u;
        }

        // 1. run-carry scan.  f(x) = a ? x + b : b
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
        uint32_t ia = fa, ib = fb;  // inclusive scan of the composition
#pragma unroll
        for (unsigned d = 1; d < 32; d <<= 1) {
            uint32_t pa = simt::shfl_up(ia, d), pb = simt::shfl_up(ib, d);
            if (lane >= d) {
                ib = ia ? pb + ib : ib;
                ia = ia & pa;
            }
        }
        uint32_t ea = simt::shfl_up(ia, 1), eb = simt::shfl_up(ib, 1);
        if (lane == 0) { ea = 1; eb = 0; }
        const uint32_t x0 = ea ? run_carry + eb : eb;  // pending run entering this lane
        {
            uint32_t la = simt::shfl(ia, 31), lb = simt::shfl(ib, 31);
            run_carry = la ? run_carry + lb : lb;
        }

        // does the byte after each chunk continue a run?
        uint32_t first_byte_zero = (k0 == 2 && (q.x & 0xffu) == 0) ? 1u : 0u;
        uint32_t next_lane_first = simt::shfl_down(first_byte_zero, 1);
        uint32_t next_iter_first = simt::shfl((uint32_t)(((nxt.x & 0xffu) == 0) ? 1u : 0u), 0);
        if (lane == 31) next_lane_first = (it + 1 < iters && base + 512 + 8 <= n8) ? next_iter_first : 0u;
        const bool cont0 = (k1 == 2) && ((q.z & 0xffu) == 0);
        const bool cont1 = next_lane_first != 0;

        // 2. tokens (looked up once, kept in registers), bit lengths and offsets
        ChunkTokens t0, t1;
        const uint32_t x1 = chunk_tokens(t0, c0, nz0, k0, rem, x0, cont0, lit_tok, tail_tok);
        chunk_tokens(t1, c1, nz1, k1, rem, x1, cont1, lit_tok, tail_tok);
        const uint32_t my_bits = lit_bits_sum(t0) + lit_bits_sum(t1);
        const uint32_t incl_bits = simt::scan_incl_add(my_bits);
        const uint32_t total_bits = simt::shfl(incl_bits, 31);
        const uint64_t o = vbit + (incl_bits - my_bits);
        const uint64_t wbase = vbit >> 5;

        // 3. pack
        BitPacker bp;
        bp.acc = 0;
        bp.accn = (uint32_t)(o & 31);
        bp.w = bp.w_first = (uint32_t)((o >> 5) - wbase);
        bp.first_word = 0;
        bp.stg = stg;
        emit_chunk(bp, t0);
        emit_chunk(bp, t1);

        // partial-word carry: g(x) = m ? x | v : v, with m = "this lane did not complete its first word"
        uint32_t gm = (bp.w == bp.w_first) ? 1u : 0u, gv = (uint32_t)bp.acc;
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
        const uint32_t carry_in = em ? (wcarry | ev) : ev;
        if (!gm) stg[bp.w_first] = bp.first_word | carry_in;
        {
            uint32_t lm = simt::shfl(im, 31), lv = simt::shfl(iv, 31);
            wcarry = lm ? (wcarry | lv) : lv;
        }
        simt::syncwarp();

        // 4. flush the completed words, coalesced
        vbit += total_bits;
        const uint32_t nwords = (uint32_t)((vbit >> 5) - wbase);
        for (uint32_t k = lane; k < nwords; k += 32) store_word(wbase + k, stg[k]);
        simt::syncwarp();
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
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s.lit_tok[i] = tables->lit_tok[i];
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
        uint64_t len = deflate_uf_stream(s.lit_tok, s.tail_tok, s.header, stg, b.in_base + b.in_off[i], b.in_len[i],
                                         b.out_base + b.out_off[i], b.out_cap[i], &st);
        if (lane == 0) {
            b.out_len[i] = len;
            b.status[i] = st;
        }
    }
}

}  // namespace fdb
