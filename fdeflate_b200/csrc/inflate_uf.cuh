// inflate_uf.cuh -- K4: inflate of ultra-fast-format streams, ONE WARP PER STREAM with all 32
// lanes decoding different parts of the SAME stream.
//
// A stream qualifies when its first 53 bytes + 5 bits equal the constant header the reference's
// UltraFastCompressor writes (src/compress/ultrafast.rs:81-91): one dynamic block whose litlen code
// is HUFFMAN_LENGTHS (src/tables.rs:7-20) and whose only distance code is "distance 1".  For such
// streams the decode table is a constant (built once on the host, shared by the whole CTA) and every
// match replicates the previous byte, so output positions are a pure prefix sum of per-token byte
// counts.  That is what makes intra-stream parallelism possible:
//
//   per segment of 32 x SUBW words of compressed bits
//   1. stage   : coalesced 16-byte loads -> padded shared memory (conflict-free per-lane reads)
//   2. count   : lane i starts WARM words BEFORE its sub-sequence at a guessed bit position, decodes
//                single tokens until it crosses its boundary (Huffman codes self-synchronise within
//                a few tokens), then counts the bytes of its own sub-sequence.  Lane 0 starts at the
//                known true position.
//   3. verify  : lane i's start must equal lane i-1's end; since lane 0 is exact this proves every
//                lane.  A lane that had not synchronised is re-run from its predecessor's end
//                (loop until consistent; normally zero rounds).
//   4. scan    : exclusive prefix sum of byte counts -> output offsets; last-literal propagation
//                gives every lane the byte its leading match replicates.
//   5. write   : lanes decode again (two-literal table entries) and drop literals into a
//                zero-initialised shared-memory window; zero runs are skipped, not written.  Full
//                windows leave with 16-byte coalesced stores and feed adler32 on the way out.
//
// Anything irregular (foreign header, distance bit 1, truncation, output larger than the slot,
// match at position 0) is not diagnosed here: the stream is appended to a work list and the general
// kernel (K3) redoes it from scratch, which yields exactly the reference's status.
//
// Semantics replaced: reference src/decompress.rs:611-1018 (decode loop) + :306-326 (checksum) for
// this stream class; results are identical to K3's and to the oracle's.
#pragma once
#include "simt.h"
#include "fdb_common.h"
#include "adler.cuh"

namespace fdb {

static const int K4_WARPS = 16;        // warps per CTA (share one 16 KiB decode table)
static const uint32_t K4_SUBW = 8;     // 32-bit words of compressed data per lane per segment
static const uint32_t K4_WARM = 4;     // warm-up words before a lane's sub-sequence
static const uint32_t K4_TAILW = 3;    // words after it (token overrun + two-word look-ahead)
static const uint32_t K4_ROWW = K4_WARM + K4_SUBW + K4_TAILW;  // 15 words seen by one lane
static const uint32_t K4_ROWS_ALLOC = K4_ROWW + 1;             // +1: the look-ahead may touch one more
static const uint32_t K4_SEG_WORDS = K4_WARM + 32 * K4_SUBW + 4;  // words staged per segment (vectors of 4)
static const uint32_t K4_LIM_LO = 32u * K4_WARM;               // a lane's sub-sequence, in bits of its row
static const uint32_t K4_LIM_HI = 32u * (K4_WARM + K4_SUBW);
static const uint32_t K4_WIN = 4096;   // output window bytes (a segment normally expands to ~2.5 KiB)
static const uint32_t K4_INVALID = 0xffffffffu;

// Staging is TRANSPOSED and PRIVATE per lane: row i holds the words lane i can ever touch
// (WARM words before its sub-sequence, the sub-sequence, TAILW after), word c of row i at
// stg[c * 32 + i].  Every lane therefore reads bank == lane (never a conflict) and walks its row
// with a constant +128-byte pointer step.  Overlapping words are simply stored twice.
struct K4Warp {
    uint32_t stg[K4_ROWS_ALLOC * 32];
    uint8_t win[K4_WIN + 16];  // +16: one overhang byte for a literal pair straddling the window end
};

struct K4Smem {
    uint32_t table[4096];
    K4Warp warp[K4_WARPS];
};

struct UfDecTables {
    uint32_t table[4096];  // litlen entries (fdb_common.h format) for HUFFMAN_LENGTHS, two-literal entries included
    uint32_t header[14];   // the constant 54 header bytes
};

// lane-private LSB-first bit reader over the lane's staging row: a 32-bit window is one funnel
// shift of (w0, w1); w2 is fetched one word ahead so the shared-memory latency stays off the
// decode dependency chain.
struct LaneBits {
    uint32_t w0, w1, w2;
    uint32_t rp;         // bit position relative to the start of the row
    const uint32_t* nx;  // next word to fetch
};
FDB_DEVICE void lb_start(LaneBits& b, const uint32_t* row, uint32_t rp) {
    b.rp = rp;
    const uint32_t* p = row + (rp >> 5) * 32;
    b.w0 = p[0];
    b.w1 = p[32];
    b.w2 = p[64];
    b.nx = p + 96;
}
FDB_DEVICE uint32_t lb_peek(const LaneBits& b) { return simt::funnel_r(b.w0, b.w1, b.rp); }  // shift is mod 32
FDB_DEVICE void lb_advance(LaneBits& b, uint32_t n) {  // n < 32
    uint32_t nrp = b.rp + n;
    if ((nrp ^ b.rp) & 32u) {
        b.w0 = b.w1;
        b.w1 = b.w2;
        b.w2 = *b.nx;
        b.nx += 32;
    }
    b.rp = nrp;
}

struct LaneCount {
    uint32_t end;     // bit position (row-relative) where the lane stopped: first token boundary >= LIM_HI
    uint32_t cnt;     // bytes produced in [start, end)
    uint32_t flags;   // CF_*
    uint32_t lastlit; // last literal value (valid with CF_HASLIT)
};
enum : uint32_t { CF_EOB = 1, CF_BAD = 2, CF_HASLIT = 4, CF_MATCH_FIRST = 8 };

// Count the bytes of the tokens in [start, LIM_HI); stop at the first token boundary >= LIM_HI or at
// EOB (then end = position of the EOB code).  No early exits inside the loops, so lanes that still
// iterate stay converged and the others wait at the loop exit.
FDB_DEVICE LaneCount count_tokens(const uint32_t* tab, const uint32_t* row, uint32_t start, uint32_t active) {
    LaneCount c = {K4_INVALID, 0, 0, 0};
    LaneBits b;
    lb_start(b, row, active ? start : 0u);
    uint32_t laste = 0, cnt = 0, flags = 0;
    uint32_t stop = active ? 0u : 1u;
    // main loop: no entry can reach LIM_HI from here (an entry consumes at most 12 + 5 + 1 bits), so
    // pairs are taken blindly
    while (!stop && b.rp < K4_LIM_HI - 18) {
        uint32_t bits = lb_peek(b);
        uint32_t e = tab[bits & 0xfffu];
        uint32_t n = e & 15u;
        cnt += e >> 28;
        if (e & LL_LIT) {
            laste = e;
        } else if (e & LL_LEN) {
            uint32_t xb = (e >> 8) & 7u;
            uint32_t v = bits >> n;
            if (cnt == 0) flags |= CF_MATCH_FIRST;
            cnt += ((e >> 16) & 0x1ffu) + (v & ((1u << xb) - 1u));
            if ((v >> xb) & 1u) flags |= CF_BAD;  // distance code "1" is not in the ultra-fast code
            n += xb + 1u;
        } else {  // end of block
            flags |= CF_EOB;
            n = 0;
            stop = 1;
        }
        lb_advance(b, n);
    }
    // tail: the pair that would cross LIM_HI gives its second literal to the next lane
    while (!stop && b.rp < K4_LIM_HI) {
        uint32_t bits = lb_peek(b);
        uint32_t e = tab[bits & 0xfffu];
        uint32_t n = e & 15u;
        if (e & LL_LIT) {
            uint32_t l1 = (e >> 24) & 15u;
            if ((e & LL_LIT2) && b.rp + l1 >= K4_LIM_HI) {
                n = l1;
                e = (e & ~(LL_LIT | (3u << 28))) | LL_LIT1 | (1u << 28);
            }
            cnt += e >> 28;
            laste = e;
        } else if (e & LL_LEN) {
            uint32_t xb = (e >> 8) & 7u;
            uint32_t v = bits >> n;
            if (cnt == 0) flags |= CF_MATCH_FIRST;
            cnt += ((e >> 16) & 0x1ffu) + (v & ((1u << xb) - 1u));
            if ((v >> xb) & 1u) flags |= CF_BAD;
            n += xb + 1u;
        } else {
            flags |= CF_EOB;
            n = 0;
            stop = 1;
        }
        lb_advance(b, n);
    }
    if (active) {
        c.end = b.rp;
        c.cnt = cnt;
        c.flags = flags;
        if (laste) {
            c.flags |= CF_HASLIT;
            c.lastlit = (laste & LL_LIT2) ? ((laste >> 16) & 0xffu) : ((laste >> 8) & 0xffu);
        }
    }
    return c;
}

// warm-up: single tokens from a guessed start until the first boundary >= LIM_LO
FDB_DEVICE uint32_t warm_up(const uint32_t* tab, const uint32_t* row, uint32_t active) {
    LaneBits b;
    lb_start(b, row, 0u);
    uint32_t stop = active ? 0u : 1u, dead = 0;
    while (!stop && b.rp < K4_LIM_LO) {
        uint32_t e = tab[lb_peek(b) & 0xfffu];
        uint32_t n = e & 15u;
        if (e & LL_LIT) {
            uint32_t l1 = (e >> 24) & 15u;
            if ((e & LL_LIT2) && b.rp + l1 >= K4_LIM_LO) n = l1;
        } else if (e & LL_LEN) {
            n += ((e >> 8) & 7u) + 1u;
        } else {
            dead = 1;  // speculative EOB: this lane has no valid guess
            stop = 1;
            n = 0;
        }
        lb_advance(b, n);
    }
    return (active && !dead) ? b.rp : K4_INVALID;
}

struct K4Stream {
    const uint8_t* in;
    uint64_t n;
    uint8_t* out;
    uint64_t cap;
};

// Returns ST_OK / ST_WRONG_CHECKSUM, or ST_PENDING_GENERAL when the stream must go to K3.
FDB_DEVICE int32_t inflate_uf_stream(const uint32_t* tab, const uint32_t* hdr, K4Warp& ws, const K4Stream& s,
                                     uint32_t flags, uint64_t* out_len, uint64_t* consumed) {
    const unsigned lane = simt::lane_id();
    uint32_t* stg = ws.stg;
    const uint32_t* row = ws.stg + lane;
    uint8_t* win = ws.win;
    *out_len = 0;
    *consumed = 0;

    // ---- header must be the ultra-fast constant (ultrafast.rs:82-91) ----
    if (s.n < 54 + 2 + 4) return ST_PENDING_GENERAL;
    {
        bool ok = true;
        for (uint32_t j = lane; j < 54; j += 32) {
            uint32_t want = (hdr[j >> 2] >> (8u * (j & 3u))) & 0xffu;
            uint32_t got = simt::ldg8(s.in + j);
            if (j == 53) got &= 0x1fu;
            ok = ok && (got == want);
        }
        if (!simt::all(ok)) return ST_PENDING_GENERAL;
    }

    const uint8_t* abase = (const uint8_t*)((uintptr_t)s.in & ~(uintptr_t)15);
    const uint64_t first_byte = (uint64_t)((uintptr_t)s.in & 15u);
    const uint64_t end_byte = first_byte + s.n;
    const uint64_t vstart = first_byte * 8 + 53 * 8 + 5;  // first data bit, virtual (bit 0 = bit 0 of abase)
    const uint64_t vend = end_byte * 8;
    const uint32_t oalign = (uint32_t)((uintptr_t)s.out & 15u);
    uint8_t* const obase = s.out - oalign;  // virtual output position vo = oalign + stream position

    uint64_t seg_word = ((vstart >> 5) >> 2) << 2;  // first word of lane 0's sub-sequence, 16-byte aligned
    uint64_t p0 = vstart;                           // true bit position where lane 0 starts
    uint64_t o0 = 0;                                // bytes produced so far
    uint32_t prev_byte = 0;
    bool have_prev = false;
    uint64_t win_vo = 0;  // virtual output position of win[0] (multiple of 16)
    AdlerAcc ad = {0, 0};

    // zero the output window
    for (uint32_t v = lane; v < (K4_WIN + 16) / 16; v += 32) ((uint4*)win)[v] = make_uint4(0, 0, 0, 0);
    simt::syncwarp();

    // Store the finished vectors win[0 .. 16*nvec) at virtual position win_vo, feed adler32, and zero
    // them again.  Bytes outside [oalign, stream_end_vo) (first / last vector of the stream) are masked.
    auto flush_vectors = [&](uint32_t nvec, uint64_t stream_end_vo) {
        for (uint32_t v = lane; v < nvec; v += 32) {
            uint4 q = ((const uint4*)win)[v];
            ((uint4*)win)[v] = make_uint4(0, 0, 0, 0);
            uint64_t vo = win_vo + 16ull * v;
            bool head_cut = vo < oalign;
            bool tail_cut = vo + 16 > stream_end_vo;
            if (!head_cut && !tail_cut) {
                simt::stcs128((uint4*)(obase + vo), q);
                adler_add16(ad, q, vo - oalign);
            } else {
                uint32_t w[4] = {q.x, q.y, q.z, q.w};
                for (uint32_t j = 0; j < 16; j++) {
                    uint64_t bpos = vo + j;
                    if (bpos >= oalign && bpos < stream_end_vo) {
                        uint32_t byte = (w[j >> 2] >> (8u * (j & 3u))) & 0xffu;
                        obase[bpos] = (uint8_t)byte;
                        adler_add1(ad, byte, bpos - oalign);
                    }
                }
            }
        }
    };

    for (;;) {
        if ((seg_word << 5) >= vend) return ST_PENDING_GENERAL;  // ran off the input without an EOB
        const uint64_t s0 = seg_word - K4_WARM;                  // first staged word (virtual word index)

        // ---- 1. stage: coalesced 16-byte loads, every word scattered to the row(s) that can see it ----
        simt::syncwarp();
        for (uint32_t v = lane; v < K4_SEG_WORDS / 4; v += 32) {
            uint64_t byte0 = (s0 << 2) + 16ull * v;  // relative to abase
            uint4 q = make_uint4(0, 0, 0, 0);
            if (byte0 + 16 > first_byte && byte0 < end_byte) {
                q = simt::ldg128((const uint4*)(abase + byte0));
                if (byte0 < first_byte || byte0 + 16 > end_byte) {
                    uint32_t w[4] = {q.x, q.y, q.z, q.w};
                    for (uint32_t j = 0; j < 16; j++) {
                        uint64_t bpos = byte0 + j;
                        if (bpos < first_byte || bpos >= end_byte) w[j >> 2] &= ~(0xffu << (8u * (j & 3u)));
                    }
                    q = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            // staged word t = 4v + j is word (t & 7) of row (t >> 3) and word (t & 7) + 8 of the row before
            const uint32_t r1 = v >> 1, c1 = (v & 1) * 4;
            const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (uint32_t j = 0; j < 4; j++) {
                if (r1 < 32) stg[(c1 + j) * 32 + r1] = w4[j];
                if (r1 >= 1 && c1 + j + 8 < K4_ROWW) stg[(c1 + j + 8) * 32 + (r1 - 1)] = w4[j];
            }
        }
        simt::syncwarp();

        // ---- 2. count ----
        uint32_t start = warm_up(tab, row, lane != 0);
        if (lane == 0) start = (uint32_t)(p0 - (s0 << 5));
        LaneCount c = count_tokens(tab, row, start, start != K4_INVALID);

        // ---- 3. verify the chain: my start must be my predecessor's end (rows are 32*SUBW bits apart) ----
        uint32_t eob_lane = 32;
        for (;;) {
            uint32_t prev_end = simt::shfl_up(c.end, 1);
            uint32_t prev_flags = simt::shfl_up(c.flags, 1);
            uint32_t want = prev_end - 32u * K4_SUBW;
            bool mismatch = (lane > 0) && (prev_end == K4_INVALID || start != want || (prev_flags & CF_EOB));
            uint32_t mm = simt::ballot(mismatch);
            uint32_t em = simt::ballot((c.flags & CF_EOB) != 0 && start != K4_INVALID);
            uint32_t first_mis = mm ? simt::ffs(mm) - 1 : 32;
            uint32_t first_eob = em ? simt::ffs(em) - 1 : 32;
            if (first_eob < first_mis) {  // the stream ends inside a verified lane
                eob_lane = first_eob;
                break;
            }
            if (first_mis == 32) break;  // every lane verified, no EOB in this segment
            // re-run the lanes whose start disagrees with their predecessor's end
            bool redo = mismatch && !(prev_flags & CF_EOB) && prev_end != K4_INVALID;
            if (mismatch) start = redo ? want : K4_INVALID;
            LaneCount c2 = count_tokens(tab, row, start, redo);
            if (mismatch) c = c2;  // (unresolved lanes get end = INVALID, cnt = 0, flags = 0)
        }
        if (lane > eob_lane) {
            c.cnt = 0;
            c.flags = 0;
        }
        if (simt::any((c.flags & CF_BAD) != 0)) return ST_PENDING_GENERAL;

        // ---- 4. scan ----
        const uint32_t incl = simt::scan_incl_add(c.cnt);
        const uint64_t seg_bytes = simt::shfl(incl, 31);
        if (o0 + seg_bytes > s.cap) return ST_PENDING_GENERAL;  // K3 reports OutputTooLarge
        uint64_t op = oalign + o0 + (incl - c.cnt);             // my virtual output position
        const uint64_t my_end_vo = op + c.cnt;
        uint32_t lit_mask = simt::ballot((c.flags & CF_HASLIT) != 0);
        uint32_t below = lit_mask & simt::lanemask_lt();
        uint32_t src = below ? 31u - simt::clz(below) : 0u;
        uint32_t from_lane = simt::shfl(c.lastlit, src);
        uint32_t fv = below ? from_lane : prev_byte;
        bool fv_known = below ? true : have_prev;
        if (simt::any((c.flags & CF_MATCH_FIRST) && !fv_known)) return ST_PENDING_GENERAL;  // match at position 0
        if (lit_mask) {
            prev_byte = simt::shfl(c.lastlit, 31u - simt::clz(lit_mask));
            have_prev = true;
        }

        // ---- 5. write ----
        // The window base slides with the output: it is the 16-byte vector holding the first byte of
        // this segment, so a segment whose output fits in K4_WIN is written by all lanes at once.
        const uint64_t seg_end_vo = oalign + o0 + seg_bytes;
        LaneBits b;
        lb_start(b, row, start != K4_INVALID ? start : 0u);
        uint32_t fin = (start == K4_INVALID || lane > eob_lane) ? 1u : 0u;  // no more tokens to decode
        uint32_t pend = 0;  // bytes of a non-zero fill still owed to later windows
        for (;;) {
            const uint64_t wend = win_vo + K4_WIN;
            const bool mine = !(fin && pend == 0) && op < wend;
            uint32_t wp = mine ? (uint32_t)(op - win_vo) : 0u;
            if (mine && pend == 0 && my_end_vo <= wend) {
                // fast path: everything this lane still has to write fits in the window
                uint8_t* wptr = win + wp;
                uint32_t laste = 0;
                while (!fin && b.rp < K4_LIM_HI - 18) {
                    uint32_t bits = lb_peek(b);
                    uint32_t e = tab[bits & 0xfffu];
                    uint32_t n = e & 15u;
                    if (e & LL_LIT) {
                        // both bytes always: a single's second byte is 0 and the next token of this lane
                        // overwrites it (or it is a zero the stream would have produced anyway)
                        wptr[0] = (uint8_t)(e >> 8);
                        wptr[1] = (uint8_t)(e >> 16);
                        wptr += e >> 28;
                        laste = e;
                    } else if (e & LL_LEN) {
                        uint32_t xb = (e >> 8) & 7u;
                        uint32_t len = ((e >> 16) & 0x1ffu) + ((bits >> n) & ((1u << xb) - 1u));
                        n += xb + 1u;
                        if (laste) fv = (laste >> (8u * (laste >> 28))) & 0xffu;  // last literal written
                        if (fv != 0)
                            for (uint32_t k = 0; k < len; k++) wptr[k] = (uint8_t)fv;
                        wptr += len;  // zero runs: the window is zero-initialised, nothing to write
                    } else {
                        fin = 1;
                        n = 0;
                    }
                    lb_advance(b, n);
                }
                if (laste) fv = (laste >> (8u * (laste >> 28))) & 0xffu;
                while (!fin && b.rp < K4_LIM_HI) {
                    uint32_t bits = lb_peek(b);
                    uint32_t e = tab[bits & 0xfffu];
                    uint32_t n = e & 15u;
                    if (e & LL_LIT) {
                        uint32_t two = (e >> 5) & 1u, l1 = (e >> 24) & 15u;
                        if (two && b.rp + l1 >= K4_LIM_HI) {
                            two = 0;
                            n = l1;
                        }
                        wptr[0] = (uint8_t)(e >> 8);
                        fv = (e >> 8) & 0xffu;
                        if (two) {
                            wptr[1] = (uint8_t)(e >> 16);
                            fv = (e >> 16) & 0xffu;
                        }
                        wptr += 1u + two;
                    } else if (e & LL_LEN) {
                        uint32_t xb = (e >> 8) & 7u;
                        uint32_t len = ((e >> 16) & 0x1ffu) + ((bits >> n) & ((1u << xb) - 1u));
                        n += xb + 1u;
                        if (fv != 0)
                            for (uint32_t k = 0; k < len; k++) wptr[k] = (uint8_t)fv;
                        wptr += len;
                    } else {
                        fin = 1;
                        n = 0;
                    }
                    lb_advance(b, n);
                }
                fin = 1;
                wp = (uint32_t)(wptr - win);
            } else if (mine) {
                // careful path: this lane's output crosses the window end (long runs) or it still owes
                // bytes of a non-zero fill; stop at the window end and resume after the flush
                while (wp < K4_WIN && !(fin && pend == 0)) {
                    if (pend) {
                        while (pend && wp < K4_WIN) {
                            win[wp++] = (uint8_t)fv;
                            pend--;
                        }
                    } else {
                        uint32_t bits = lb_peek(b);
                        uint32_t e = tab[bits & 0xfffu];
                        uint32_t n = e & 15u;
                        if (e & LL_LIT) {
                            uint32_t two = (e >> 5) & 1u, l1 = (e >> 24) & 15u;
                            if (two && b.rp + l1 >= K4_LIM_HI) {
                                two = 0;
                                n = l1;
                            }
                            win[wp] = (uint8_t)(e >> 8);
                            fv = (e >> 8) & 0xffu;
                            if (two) {
                                win[wp + 1] = (uint8_t)(e >> 16);  // may be the overhang byte win[K4_WIN]
                                fv = (e >> 16) & 0xffu;
                            }
                            wp += 1u + two;
                        } else if (e & LL_LEN) {
                            uint32_t xb = (e >> 8) & 7u;
                            uint32_t len = ((e >> 16) & 0x1ffu) + ((bits >> n) & ((1u << xb) - 1u));
                            n += xb + 1u;
                            if (fv == 0) {
                                wp += len;
                            } else {
                                while (len && wp < K4_WIN) {
                                    win[wp++] = (uint8_t)fv;
                                    len--;
                                }
                                pend = len;
                            }
                        } else {
                            fin = 1;
                            n = 0;
                        }
                        lb_advance(b, n);
                        if (b.rp >= K4_LIM_HI) fin = 1;
                    }
                }
            }
            if (mine) op = win_vo + wp;
            simt::syncwarp();
            if (seg_end_vo < wend) break;  // the rest of this segment fits: leave it in the window
            // the window is complete: flush all of it and slide by K4_WIN
            uint32_t over = win[K4_WIN];
            simt::syncwarp();
            flush_vectors(K4_WIN / 16, ~0ull);
            if (lane == 0) {
                win[K4_WIN] = 0;
                win[0] = (uint8_t)over;
            }
            simt::syncwarp();
            win_vo = wend;
        }

        // ---- next segment or finish ----
        adler_fold(ad);
        o0 += seg_bytes;
        if (eob_lane < 32) {
            const uint32_t eob_rel = simt::shfl(c.end, eob_lane);
            const uint64_t eob_end = ((s0 + (uint64_t)K4_SUBW * eob_lane) << 5) + eob_rel + 12;  // EOB = 12 bits
            const uint64_t trailer_byte = (eob_end + 7) >> 3;   // relative to abase
            if (trailer_byte + 4 > end_byte) return ST_PENDING_GENERAL;  // truncated: K3 reports it
            uint32_t left = (uint32_t)(oalign + o0 - win_vo);
            flush_vectors((left + 15) / 16, oalign + o0);
            simt::syncwarp();
            const uint8_t* t = abase + trailer_byte;
            uint32_t stored = ((uint32_t)simt::ldg8(t) << 24) | ((uint32_t)simt::ldg8(t + 1) << 16) |
                              ((uint32_t)simt::ldg8(t + 2) << 8) | (uint32_t)simt::ldg8(t + 3);
            uint32_t got = adler_finish_warp(ad, o0);
            *out_len = o0;
            *consumed = trailer_byte + 4 - first_byte;
            if (!(flags & FLAG_IGNORE_ADLER32) && got != stored) return ST_WRONG_CHECKSUM;
            return ST_OK;
        }
        // flush the finished vectors of this segment and slide the window base to the vector that
        // holds the next output byte (its already-written bytes move to win[0..16))
        {
            const uint32_t nvec = (uint32_t)((seg_end_vo - win_vo) >> 4);
            uint4 tail = ((const uint4*)win)[nvec];
            simt::syncwarp();
            flush_vectors(nvec, ~0ull);
            simt::syncwarp();
            if (nvec > 0 && lane == 0) {
                ((uint4*)win)[nvec] = make_uint4(0, 0, 0, 0);
                ((uint4*)win)[0] = tail;
            }
            simt::syncwarp();
            win_vo += 16ull * nvec;
        }
        p0 = ((s0 + (uint64_t)K4_SUBW * 31) << 5) + simt::shfl(c.end, 31);
        seg_word += 32 * K4_SUBW;
    }
}

// Persistent kernel.  Streams the fast path declines are appended to worklist[] (count in *work_count)
// with status ST_PENDING_GENERAL; the host launches K3 over that list next, on the same stream.
FDB_GLOBAL void FDB_LAUNCH_BOUNDS(K4_WARPS * 32, 2)
    inflate_uf_kernel(InflateBatch b, const UfDecTables* tables, uint32_t* next, uint32_t* worklist,
                      uint32_t* work_count) {
    FDB_DYN_SMEM(smem_raw);
    K4Smem& sm = *reinterpret_cast<K4Smem*>(smem_raw);
    FDB_SHARED uint32_t hdr[14];
    for (uint32_t i = threadIdx.x; i < 4096; i += blockDim.x) sm.table[i] = tables->table[i];
    if (threadIdx.x < 14) hdr[threadIdx.x] = tables->header[threadIdx.x];
    simt::syncthreads();
    const unsigned lane = simt::lane_id();
    K4Warp& ws = sm.warp[simt::warp_in_block()];
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = simt::atomic_add(next, 1u);
        i = simt::shfl(i, 0);
        if (i >= b.n) break;
        K4Stream s = {b.in_base + b.in_off[i], b.in_len[i], b.out_base + b.out_off[i], b.out_cap[i]};
        uint64_t out_len = 0, consumed = 0;
        int32_t st = inflate_uf_stream(sm.table, hdr, ws, s, b.flags, &out_len, &consumed);
        if (lane == 0) {
            b.status[i] = st;
            b.out_len[i] = out_len;
            if (b.consumed) b.consumed[i] = consumed;
            if (st == ST_PENDING_GENERAL) worklist[simt::atomic_add(work_count, 1u)] = i;
        }
        simt::syncwarp();
    }
}

}  // namespace fdb
