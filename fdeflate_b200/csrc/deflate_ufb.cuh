// deflate_ufb.cuh -- K1/K2, second generation: ultra-fast deflate with 64-BYTE LANE BLOCKS.
//
// Same output as deflate_uf.cuh (byte-identical to the reference's UltraFastCompressor for one write_data() call,
// src/compress/ultrafast.rs:16-181), same per-chunk logic (chunk_plan, word_pairs, BitPacker), different division of
// labour.  In deflate_uf.cuh a lane owns 16 bytes of a 512-byte warp step, and half of the kernel's instructions are
// what it costs to put 32 lanes back together after every 16 bytes: the prefix sum of the bit counts, the run carry,
// the partial-word carry, the flush, the loop.  Here a lane owns 64 consecutive bytes (eight of the reference's 8-byte
// chunks) of a 2048-byte warp step and encodes them the way the reference does -- sequentially, the run state in
// registers -- into a PRIVATE bit string that starts at bit 0 of its own shared-memory row.  Only then do the lanes
// meet: prefix sum of the string lengths, and every lane moves its string to its bit offset in the warp's output
// window with one funnel shift per word.  The per-step costs are paid once per 64 bytes instead of once per 16.
//
// Per warp step of 2048 bytes:
//   1. stage   : coalesced 16-byte loads -> shared-memory rows of 64 bytes (row stride 17 words: a lane walking its own
//                row and 32 lanes storing one vector each are both conflict-free); adler32 on the vectors in registers
//   2. carry   : the run pending at the start of every lane's block, from the blocks below (all-zero blocks extend it,
//                any other block resets it to its zero suffix): one ballot + one shuffle, as in deflate_uf.cuh
//   3. encode  : eight chunks per lane, sequentially: plan, literal lookups, pack into the private row
//   4. scan    : prefix sum of the string lengths -> bit offsets
//   5. move    : word j of a lane's string lands at (offset >> 5) + j, shifted by offset & 31; the word a lane only
//                starts travels on as the partial-word carry (one shuffle; a composition scan when some lane's whole
//                string ends inside one word), so every window word is stored exactly once -- no atomics, no zeroing
//   6. flush   : completed words leave with coalesced stores
// The window IS the staging rows (dead once the lanes have encoded): 6.4 KB of shared memory per warp, 32 warps per SM.
#pragma once
#include "deflate_uf.cuh"

namespace fdb {

static const uint32_t UB_BLOCK = 64;                   // bytes per lane and warp step
static const uint32_t UB_STEP = 32 * UB_BLOCK;         // 2048
static const uint32_t UB_IN_STRIDE = 17;               // words per staging row
static const uint32_t UB_ROW_WORDS = 33;               // private bit string: 8 chunks x <= 124 bits = 31 words, + the zero word behind it
static const uint32_t UB_WIN_WORDS = 32 * UB_IN_STRIDE;  // the warp's output window = the staging rows (544 words: a step of
                                                         // up to 8.5 bits per input byte leaves in one round, else in several)
#ifndef UB_WARPS_PER_CTA
#define UB_WARPS_PER_CTA 8
#endif
static const int UB_WARPS = UB_WARPS_PER_CTA;
#ifndef UB_UNROLL
#define UB_UNROLL 2
#endif
static const int UB_UNROLL_N = UB_UNROLL;  // chunks of the encode loop unrolled together
#ifndef UB_MIN_CTAS
#define UB_MIN_CTAS 4
#endif

// Bit packer of a lane's private row.  `pos` is the bit position in the row and never wraps: the funnel shifts take it
// mod 32, a completed word shows as a flip of bit 5, and the row's length is pos itself -- one instruction less per emit
// than an accumulator count that is reduced by 32 after every store (deflate_uf.cuh's BitPacker), five emits per chunk.
#ifndef UB_POS_PACKER
#define UB_POS_PACKER 0
#endif
struct RowPacker {
    uint32_t lo;       // the word under construction: its bits below pos
    uint32_t pos;      // bits emitted so far
    simt::saddr wa;    // row address of that word
    FDB_MEMBER void emit(uint32_t v, uint32_t n) {  // v < 2^n, n <= 32
        const uint32_t nlo = lo | simt::funnel_l(0u, v, pos);  // v << (pos mod 32)
        const uint32_t nhi = simt::funnel_l(v, 0u, pos);       // what does not fit (0 when pos mod 32 == 0)
        const uint32_t npos = pos + n;
#if !defined(FDB_EMUL) && UB_POS_PACKER == 2
        asm volatile(
            "{\n\t.reg .pred q;\n\t.reg .b32 x;\n\t"
            "xor.b32 x, %2, %3;\n\tand.b32 x, x, 32;\n\tsetp.ne.u32 q, x, 0;\n\t"
            "@q st.shared.u32 [%1], %4;\n\t@q add.u32 %1, %1, 4;\n\tselp.b32 %0, %5, %4, q;\n\t}"
            : "=r"(lo), "+r"(wa)
            : "r"(npos), "r"(pos), "r"(nlo), "r"(nhi)
            : "memory");
#else
        const bool full = ((npos ^ pos) & 32u) != 0;
        if (full) {
            simt::sts32(wa, nlo);
            wa += 4;
        }
        lo = full ? nhi : nlo;
#endif
        pos = npos;
    }
};

struct UbWarp {
    uint32_t win[UB_WIN_WORDS + 1];  // input staging rows, then the output window (+1: chunk 7 of lane 31 looks one word ahead)
    uint32_t rows[32 * UB_ROW_WORDS];
};

struct UbSmem {
    uint2 lit[512];
    uint32_t tail_tok[258];
    uint32_t header[14];
    UbWarp warp[UB_WARPS];
};

// Where the bytes to encode come from.  UbPlainSrc: n bytes in device memory.  deflate_png.cuh adds a source that
// computes the bytes of a PNG-filtered image from the raw pixels on the fly (the filter fused into this prologue).
//   load16(g): bytes [g, g + 16) of the stream (g a multiple of 16; zeros past the end)   byte(g): one byte, g < n
struct UbPlainSrc {
    const uint8_t* in;
    uint64_t n;
    bool aligned;
    FDB_MEMBER uint4 load16(uint64_t g) const { return load16_guarded(in, g, n, aligned); }
    FDB_MEMBER uint32_t byte(uint64_t g) const { return simt::ldg8(in + g); }
    FDB_MEMBER void prefetch(uint64_t g) const { simt::prefetch_l2(in + g); }
};

// One stream, one warp; returns the encoded length, or 0 with *status != ST_OK.
template <class SRC>
FDB_DEVICE uint64_t deflate_ufb_stream(const uint2* lit, const uint32_t* tail_tok, const uint32_t* header, UbWarp& ws,
                                       const SRC& src, uint64_t n, uint8_t* out, uint64_t cap, int32_t* status) {
    const unsigned lane = simt::lane_id();
    const simt::saddr win_s = simt::smem_addr(ws.win);
    const simt::saddr my_in = win_s + 4u * UB_IN_STRIDE * lane;                     // my staging row
    const simt::saddr my_row = simt::smem_addr(ws.rows) + 4u * UB_ROW_WORDS * lane;  // my bit string
    const uint32_t oab = (uint32_t)((uintptr_t)out & 3u);  // out's offset inside its aligned word
    uint32_t* const obase = (uint32_t*)(out - oab);
    const uint64_t n8 = n & ~(uint64_t)7;
    const uint32_t rem = (uint32_t)(n - n8);
    bool overflow = false;
    // (cap_words = (cap + oab) >> 2, the virtual words that end inside the caller's slot, is recomputed where it is used:
    // one value less to keep across the encode loop, which runs at the register limit)

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
        const uint32_t full = (uint32_t)(vbit >> 5);
        if (lane < full) {
            const uint64_t end_byte = 4ull * (lane + 1) - oab;
            if (end_byte > cap) {
                overflow = true;
            } else if (lane == 0 && oab) {
                for (uint32_t j = oab; j < 4; j++) out[j - oab] = (uint8_t)(hw >> (8u * j));
            } else {
                obase[lane] = hw;
            }
        }
        wcarry = simt::shfl(hw, full) & ((1u << (vbit & 31)) - 1u);
    }

    // ---- data (ultrafast.rs:94-167) ----
    AdlerAcc ad = {0, 0};
    uint32_t run_carry = 0;
    const uint64_t iters = (n + UB_STEP - 1) / UB_STEP;
    // "is the first byte of the next step zero" (lane 31's last chunk needs it)
    uint32_t nfb_next = iters > 1 ? src.byte(UB_STEP) : 1u;

    auto step = [&](uint64_t it, auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        const uint64_t base = it * UB_STEP;
        const uint32_t nfb = nfb_next;
        if (it + 2 < iters) nfb_next = src.byte(base + 2 * UB_STEP);

        // the next step's 16 lines are needed in a few thousand cycles: start them towards L2 now
        if (lane < 16 && base + UB_STEP + 128ull * lane < n) src.prefetch(base + UB_STEP + 128ull * lane);
        // 1. stage: vector v = lane + 32 k holds bytes [16 v, 16 v + 16) of the step = words 4 (v & 3) .. of row v >> 2
        simt::syncwarp();
        // (all four loads first: the stores below are ordered asm statements, and a load issued behind the store of
        // the vector before it waits out its own trip to L2 -- four latencies per step instead of one)
        uint4 q4[4];
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            const uint64_t g = base + 16ull * (lane + 32u * k);
            q4[k] = make_uint4(0, 0, 0, 0);
            if (FULL || g < n) q4[k] = src.load16(g);
        }
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t v = lane + 32u * k;
            const uint64_t g = base + 16ull * v;
            const uint4 q = q4[k];
            if (FULL || g + 16 <= n) {
                adler_add16(ad, q, g);
            } else if (g < n) {
                uint32_t w[4] = {q.x, q.y, q.z, q.w};
                for (uint32_t j = 0; g + j < n; j++) adler_add1(ad, (w[j >> 2] >> (8u * (j & 3u))) & 0xffu, g + j);
            }
            const simt::saddr a = win_s + 4u * ((v >> 2) * UB_IN_STRIDE + 4u * (v & 3u));
            simt::sts32(a, q.x);
            simt::sts32(a + 4u, q.y);
            simt::sts32(a + 8u, q.z);
            simt::sts32(a + 12u, q.w);
        }
        if ((it & 15) == 15) adler_fold(ad);
        simt::syncwarp();

        const uint64_t g0 = base + (uint64_t)lane * UB_BLOCK;  // my first byte
        // chunk kinds: 2 = whole chunk in the run-logic prefix, 1 = the final partial chunk, 0 = past the end
        auto kind_of = [&](uint32_t c) -> uint32_t {
            if (FULL) return 2u;
            const uint64_t g = g0 + 8ull * c;
            return (g + 8 <= n8) ? 2u : (g == n8 && rem) ? 1u : 0u;
        };

        // 2. what my block does to a pending run: f(x) = fa ? x + 64 : fb  (fa: all eight chunks are whole and zero;
        //    fb: the zero bytes my block leaves pending = its zero suffix, as long as the chunks are whole)
        uint32_t fa = 1, fb = 0;
        {
            bool open = true;  // still inside the zero suffix
            for (int c = 7; c >= 0 && simt::any(open); c--) {
                const uint32_t lo = simt::lds32(my_in + 8u * (uint32_t)c), hi = simt::lds32(my_in + 8u * (uint32_t)c + 4u);
                if (open) {
                    if (kind_of((uint32_t)c) != 2u) {
                        fa = 0;
                        fb = 0;
                        open = false;
                    } else if ((lo | hi) == 0) {
                        fb += 8u;
                    } else {
                        fb += hi ? simt::clz(hi) >> 3 : 4u + (simt::clz(lo) >> 3);
                        fa = 0;
                        open = false;
                    }
                }
            }
        }
        uint32_t x0;
        {
            const uint32_t zmask = simt::ballot(fa != 0);
            const uint32_t below = ~zmask & simt::lanemask_lt();  // lanes below me that reset the run
            const uint32_t j = below ? 31u - simt::clz(below) : 0u;
            const uint32_t bj = simt::shfl(fb, j);
            x0 = below ? bj + UB_BLOCK * (lane - 1u - j) : run_carry + UB_BLOCK * lane;
            run_carry = simt::shfl(fa ? x0 + UB_BLOCK : fb, 31);
        }
        // does the byte after my block continue a run?  (the first byte of the next lane's block, if that chunk is whole)
        const uint32_t first_word = simt::lds32(my_in);
        const uint32_t my_first_zero = (kind_of(0) == 2u && (first_word & 0xffu) == 0) ? 1u : 0u;
        uint32_t next_first_zero = simt::shfl_down(my_first_zero, 1);
        if (lane == 31) next_first_zero = (it + 1 < iters && base + UB_STEP + 8 <= n8 && nfb == 0) ? 1u : 0u;

        // 3. encode my eight chunks into my row
#if UB_POS_PACKER
        RowPacker bp;
        bp.lo = 0;
        bp.pos = 0;
        bp.wa = my_row;
#else
        BitPacker bp;
        bp.lo = bp.hi = 0;
        bp.accn = 0;
        bp.wa = my_row;
#endif
        uint32_t rr = x0 % 258u, pend = x0 > 0 ? 1u : 0u;
        uint32_t hold_v = 0, hold_n = 0;  // the tail tokens of the chunk before, emitted together with the next head
        uint32_t lo = first_word, hi = simt::lds32(my_in + 4u);
#pragma unroll UB_UNROLL_N
        for (uint32_t c = 0; c < 8; c++) {
            // (the row has 17 words: the loads for c == 7 read the pad word and the next row's first word, unused)
            const uint32_t nlo = simt::lds32(my_in + 8u * c + 8u), nhi = simt::lds32(my_in + 8u * c + 12u);
            const uint32_t kind = kind_of(c);
            const bool cont = c < 7 ? (kind_of(c + 1) == 2u && (nlo & 0xffu) == 0) : next_first_zero != 0;
            ChunkPlan t;
            chunk_plan(t, ((uint64_t)hi << 32) | lo, kind, rem, rr, pend, cont, tail_tok);
            PairTok p0, p1, p2, p3;
            word_pairs(p0, p1, lo, t.keep, lit);
            word_pairs(p2, p3, hi, t.keep >> 4, lit);
            bp.emit(hold_v | (t.head_v << hold_n), hold_n + t.head_n);  // together <= 30 bits
            bp.emit(p0.v, p0.n);
            bp.emit(p1.v, p1.n);
            bp.emit(p2.v, p2.n);
            bp.emit(p3.v, p3.n);
            hold_v = t.tail_v;
            hold_n = t.tail_n;
            lo = nlo;
            hi = nhi;
        }
        bp.emit(hold_v, hold_n);
        simt::sts32(bp.wa, bp.lo);   // the incomplete last word (zeros above the last bit)
        simt::sts32(bp.wa + 4u, 0u);  // ... and a zero word behind it (the move reads one word past the end)
#if UB_POS_PACKER
        const uint32_t my_bits = bp.pos;
#else
        const uint32_t my_bits = 8u * (uint32_t)(bp.wa - my_row) + bp.accn;
#endif

        // 4. bit offsets
        const uint32_t incl_bits = simt::scan_incl_add(my_bits);
        const uint32_t total_bits = simt::shfl(incl_bits, 31);
        const uint64_t o = vbit + (incl_bits - my_bits);
        const uint64_t wbase = vbit >> 5;
        const uint32_t s = (uint32_t)(o & 31);
        const uint32_t w0 = (uint32_t)((o >> 5) - wbase);
        const uint32_t nc = (uint32_t)(((o + my_bits) >> 5) - (o >> 5));  // words I complete
        simt::syncwarp();  // every lane is done with the staging rows: the window may overwrite them

        // 5. move.  Partial-word carry: g(x) = gm ? x | gv : gv with gm = "I complete no word", gv = my bits in the
        //    word I leave incomplete
        const uint32_t r_last = nc ? simt::lds32(my_row + 4u * (nc - 1u)) : 0u;
        const uint32_t r_next = simt::lds32(my_row + 4u * nc);
        const uint32_t gv = simt::funnel_l(r_last, r_next, s);
        const uint32_t gm = nc == 0 ? 1u : 0u;
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
        // ... and 6. flush the completed words, coalesced; in rounds of one window (normally one)
        vbit += total_bits;
        const uint32_t nwords = (uint32_t)((vbit >> 5) - wbase);
        // (the header is 53 bytes, so these are never the stream's first, possibly partial, word)
        const uint64_t cap_words = (cap + oab) >> 2;
        const uint32_t fit = cap_words > wbase ? (uint32_t)(cap_words - wbase < nwords ? cap_words - wbase : nwords) : 0u;
        uint32_t* const dst = obase + wbase;
        if (fit < nwords) overflow = true;
        for (uint32_t r0 = 0; r0 < nwords; r0 += UB_WIN_WORDS) {
            if (r0) simt::syncwarp();  // the round before has left the window
            // my word j goes to window word w0 + j - r0
            const uint32_t jlo = r0 > w0 ? r0 - w0 : 0u;
            const uint32_t jhi = r0 + UB_WIN_WORDS > w0 ? (nc < r0 + UB_WIN_WORDS - w0 ? nc : r0 + UB_WIN_WORDS - w0) : 0u;
            if (jlo < jhi) {
                uint32_t prev = jlo ? simt::lds32(my_row + 4u * (jlo - 1u)) : 0u;
                simt::saddr wp = win_s + 4u * (w0 + jlo - r0);
                for (uint32_t j = jlo; j < jhi; j++) {
                    const uint32_t cur = simt::lds32(my_row + 4u * j);
                    uint32_t v = simt::funnel_l(prev, cur, s);
                    if (j == 0) v |= carry_in;
                    simt::sts32(wp, v);
                    wp += 4u;
                    prev = cur;
                }
            }
            simt::syncwarp();
            const uint32_t hi_w = r0 + UB_WIN_WORDS < fit ? r0 + UB_WIN_WORDS : fit;
#pragma unroll 4
            for (uint32_t k = r0 + lane; k < hi_w; k += 32) dst[k] = simt::lds32(win_s + 4u * (k - r0));
        }
    };
    for (uint64_t it = 0; it < iters; it++) {
        if (it * UB_STEP + UB_STEP <= n8)
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

FDB_GLOBAL void FDB_LAUNCH_BOUNDS(UB_WARPS * 32, UB_MIN_CTAS)
    deflate_ufb_kernel(DeflateBatch b, const UfEncTables* tables, uint32_t* next, const uint32_t* split_item0,
                       const uint32_t* order) {
    FDB_DYN_SMEM(smem_raw);
    UbSmem& s = *reinterpret_cast<UbSmem*>(smem_raw);
    for (uint32_t i = threadIdx.x; i < 512; i += blockDim.x) s.lit[i] = tables->lit[i];
    for (uint32_t i = threadIdx.x; i < 258; i += blockDim.x) s.tail_tok[i] = tables->tail_tok[i];
    for (uint32_t i = threadIdx.x; i < 14; i += blockDim.x) s.header[i] = tables->header[i];
    simt::syncthreads();
    const unsigned lane = simt::lane_id();
    UbWarp& ws = s.warp[simt::warp_in_block()];
    // every warp's first input is fixed (input c + grid * w for warp w of CTA c), further inputs come from the counter
    // (see deflate_uf_kernel)
    const uint32_t slots = gridDim.x * UB_WARPS;
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
        if (split_item0 && split_item0[i] != DF_NO_ITEM) continue;  // encoded segment by segment (deflate_uf.cuh)
        int32_t st = ST_OK;
        const UbPlainSrc src = {b.in_base + b.in_off[i], b.in_len[i], ((uintptr_t)(b.in_base + b.in_off[i]) & 15u) == 0};
        uint64_t len = deflate_ufb_stream(s.lit, s.tail_tok, s.header, ws, src, src.n, b.out_base + b.out_off[i], b.out_cap[i], &st);
        if (lane == 0) {
            b.out_len[i] = len;
            b.status[i] = st;
        }
    }
}

}  // namespace fdb
