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
static const uint32_t K4_SUBW = 16;    // 32-bit words of compressed data per lane per segment
static const uint32_t K4_WARM = 4;     // warm-up words before a lane's sub-sequence
static const uint32_t K4_TAILW = 4;    // slack after the segment (token overrun + refill look-ahead)
static const uint32_t K4_STG_WORDS = K4_WARM + 32 * K4_SUBW + K4_TAILW;     // 520
static const uint32_t K4_STG_PADDED = K4_STG_WORDS + (K4_STG_WORDS >> 4) + 1;
static const uint32_t K4_WIN = 2048;   // output window bytes
static const uint32_t K4_INVALID = 0xffffffffu;

struct K4Warp {
    uint32_t stg[K4_STG_PADDED];
    uint32_t pad_[(4 - (K4_STG_PADDED & 3)) & 3];
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

// lane-private LSB-first bit reader over the padded staging buffer
struct LaneBits {
    uint64_t bb;
    uint32_t nb;
    uint32_t gw;  // next staging word
    uint32_t rp;  // bit position relative to staging word 0
};
FDB_DEVICE uint32_t stg_word(const uint32_t* stg, uint32_t g) { return stg[g + (g >> 4)]; }
FDB_DEVICE void lb_start(LaneBits& b, const uint32_t* stg, uint32_t rp) {
    b.rp = rp;
    b.gw = rp >> 5;
    uint32_t sh = rp & 31;
    b.bb = (uint64_t)(stg_word(stg, b.gw) >> sh);
    b.nb = 32 - sh;
    b.gw++;
}
FDB_DEVICE void lb_refill(LaneBits& b, const uint32_t* stg) {  // afterwards nb >= 33
    if (b.nb <= 32) {
        b.bb |= (uint64_t)stg_word(stg, b.gw) << b.nb;
        b.nb += 32;
        b.gw++;
    }
}
FDB_DEVICE void lb_consume(LaneBits& b, uint32_t n) {
    b.bb >>= n;
    b.nb -= n;
    b.rp += n;
}

struct LaneCount {
    uint32_t end;     // bit position (relative) where the lane stopped: first token boundary >= its limit
    uint32_t cnt;     // bytes produced in [start, end)
    uint32_t flags;   // CF_*
    uint32_t lastlit; // last literal value (valid with CF_HASLIT)
};
enum : uint32_t { CF_EOB = 1, CF_BAD = 2, CF_HASLIT = 4, CF_MATCH_FIRST = 8 };

// count the bytes of the tokens in [start, limit); stop at the first token boundary >= limit or at EOB
FDB_DEVICE LaneCount count_tokens(const uint32_t* tab, const uint32_t* stg, uint32_t start, uint32_t limit) {
    LaneCount c = {0, 0, 0, 0};
    LaneBits b;
    lb_start(b, stg, start);
    bool first = true;
    while (b.rp < limit) {
        lb_refill(b, stg);
        uint32_t e = tab[(uint32_t)b.bb & 0xfffu];
        uint32_t n = e & 15u;
        if (e & LL_LIT) {
            uint32_t two = (e >> 5) & 1u, l1 = (e >> 24) & 15u;
            if (two && b.rp + l1 >= limit) {  // the pair's second literal belongs to the next lane
                two = 0;
                n = l1;
            }
            c.cnt += 1u + two;
            c.lastlit = two ? ((e >> 16) & 0xffu) : ((e >> 8) & 0xffu);
            c.flags |= CF_HASLIT;
        } else if (e & LL_LEN) {
            uint32_t xb = (e >> 8) & 7u;
            uint32_t v = (uint32_t)(b.bb >> n);
            c.cnt += ((e >> 16) & 0x1ffu) + (v & ((1u << xb) - 1u));
            if ((v >> xb) & 1u) c.flags |= CF_BAD;  // distance code "1" is not in the ultra-fast code
            if (first) c.flags |= CF_MATCH_FIRST;
            n += xb + 1u;
        } else {  // end of block
            c.flags |= CF_EOB;
            c.end = b.rp;  // position of the EOB code itself
            return c;
        }
        first = false;
        lb_consume(b, n);
    }
    c.end = b.rp;
    return c;
}

// warm-up: single tokens from a guessed start until the first boundary >= limit
FDB_DEVICE uint32_t warm_up(const uint32_t* tab, const uint32_t* stg, uint32_t start, uint32_t limit) {
    LaneBits b;
    lb_start(b, stg, start);
    while (b.rp < limit) {
        lb_refill(b, stg);
        uint32_t e = tab[(uint32_t)b.bb & 0xfffu];
        uint32_t n = e & 15u;
        if (e & LL_LIT) {
            uint32_t l1 = (e >> 24) & 15u;
            if ((e & LL_LIT2) && b.rp + l1 >= limit) n = l1;
        } else if (e & LL_LEN) {
            n += ((e >> 8) & 7u) + 1u;
        } else {
            return K4_INVALID;  // speculative EOB: this lane has no valid guess
        }
        lb_consume(b, n);
    }
    return b.rp;
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

    uint64_t seg_word = ((vstart >> 5) >> 2) << 2;  // first segment's word, 16-byte aligned
    uint64_t p0 = vstart;                           // true bit position where lane 0 starts
    uint64_t o0 = 0;                                // bytes produced so far
    uint32_t prev_byte = 0;
    bool have_prev = false;
    uint64_t win_vo = 0;  // virtual output position of win[0] (multiple of K4_WIN)
    AdlerAcc ad = {0, 0};

    // zero the output window
    for (uint32_t v = lane; v < (K4_WIN + 16) / 16; v += 32) ((uint4*)win)[v] = make_uint4(0, 0, 0, 0);
    simt::syncwarp();

    // flush win[0..nbytes) to virtual position win_vo (nbytes multiple of 16 except at stream end)
    auto flush_window = [&](uint32_t nbytes, uint64_t stream_end_vo) {
        for (uint32_t v = lane; v < (nbytes + 15) / 16; v += 32) {
            uint4 q = ((const uint4*)win)[v];
            uint64_t vo = win_vo + 16ull * v;
            bool head_cut = vo < oalign;                 // first vector of the stream, bytes before out[0]
            bool tail_cut = vo + 16 > stream_end_vo;     // last vector, bytes after the stream end
            if (!head_cut && !tail_cut) {
                simt::stcs128((uint4*)(obase + vo), q);
                adler_add16(ad, q, vo - oalign);
            } else {
                uint32_t w[4] = {q.x, q.y, q.z, q.w};
                for (uint32_t j = 0; j < 16; j++) {
                    uint64_t b = vo + j;
                    if (b >= oalign && b < stream_end_vo) {
                        uint32_t byte = (w[j >> 2] >> (8u * (j & 3u))) & 0xffu;
                        obase[b] = (uint8_t)byte;
                        adler_add1(ad, byte, b - oalign);
                    }
                }
            }
        }
    };

    for (;;) {
        if ((seg_word << 5) >= vend) return ST_PENDING_GENERAL;  // ran off the input without an EOB
        const uint64_t s0 = seg_word - K4_WARM;                  // staging word 0 (virtual word index)

        // ---- 1. stage ----
        simt::syncwarp();
        for (uint32_t v = lane; v < K4_STG_WORDS / 4; v += 32) {
            uint64_t byte0 = (s0 << 2) + 16ull * v;  // relative to abase
            uint4 q = make_uint4(0, 0, 0, 0);
            if (byte0 + 16 > first_byte && byte0 < end_byte) {
                q = simt::ldg128((const uint4*)(abase + byte0));
                if (byte0 < first_byte || byte0 + 16 > end_byte) {
                    uint32_t w[4] = {q.x, q.y, q.z, q.w};
                    for (uint32_t j = 0; j < 16; j++) {
                        uint64_t b = byte0 + j;
                        if (b < first_byte || b >= end_byte) w[j >> 2] &= ~(0xffu << (8u * (j & 3u)));
                    }
                    q = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            uint32_t g = 4 * v;
            uint32_t p = g + (g >> 4);
            stg[p] = q.x; stg[p + 1] = q.y; stg[p + 2] = q.z; stg[p + 3] = q.w;
        }
        simt::syncwarp();

        // ---- 2. count ----
        const uint32_t my_limit_lo = 32u * (K4_WARM + lane * K4_SUBW);  // my boundary B_i
        const uint32_t my_limit_hi = my_limit_lo + 32u * K4_SUBW;       // B_{i+1}
        uint32_t start;
        if (lane == 0) start = (uint32_t)(p0 - (s0 << 5));
        else start = warm_up(tab, stg, my_limit_lo - 32u * K4_WARM, my_limit_lo);
        LaneCount c = {K4_INVALID, 0, 0, 0};
        if (start != K4_INVALID) c = count_tokens(tab, stg, start, my_limit_hi);

        // ---- 3. verify the chain ----
        uint32_t eob_lane = 32;
        for (;;) {
            uint32_t prev_end = simt::shfl_up(c.end, 1);
            uint32_t prev_flags = simt::shfl_up(c.flags, 1);
            bool mismatch = (lane > 0) && (start != prev_end || (prev_flags & CF_EOB));
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
            if (redo) {
                start = prev_end;
                c = count_tokens(tab, stg, start, my_limit_hi);
            } else if (mismatch) {
                start = K4_INVALID;  // predecessor is itself unresolved or ended the stream
                c.end = K4_INVALID;
                c.flags = 0;
                c.cnt = 0;
            }
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
        const uint64_t seg_end_vo = oalign + o0 + seg_bytes;
        LaneBits b;
        lb_start(b, stg, start != K4_INVALID ? start : 0u);
        bool done = (start == K4_INVALID) || (lane > eob_lane);
        uint32_t pend = 0;  // bytes of a non-zero fill still owed to later windows
        for (;;) {
            const uint64_t wend = win_vo + K4_WIN;
            if (!done && op < wend) {
                uint32_t wp = (uint32_t)(op - win_vo);
                while (pend && wp < K4_WIN) {
                    win[wp++] = (uint8_t)fv;
                    pend--;
                }
                while (!pend && wp < K4_WIN) {
                    if (b.rp >= my_limit_hi) { done = true; break; }
                    lb_refill(b, stg);
                    uint32_t e = tab[(uint32_t)b.bb & 0xfffu];
                    uint32_t n = e & 15u;
                    if (e & LL_LIT) {
                        uint32_t two = (e >> 5) & 1u, l1 = (e >> 24) & 15u;
                        if (two && b.rp + l1 >= my_limit_hi) {
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
                        uint32_t len = ((e >> 16) & 0x1ffu) + ((uint32_t)(b.bb >> n) & ((1u << xb) - 1u));
                        n += xb + 1u;
                        if (fv == 0) {
                            wp += len;  // the window is zero-initialised: nothing to write
                        } else {
                            while (len && wp < K4_WIN) {
                                win[wp++] = (uint8_t)fv;
                                len--;
                            }
                            pend = len;
                        }
                    } else {
                        done = true;
                        break;
                    }
                    lb_consume(b, n);
                }
                op = win_vo + wp;
            }
            simt::syncwarp();
            if (seg_end_vo < wend) break;  // window not complete yet: keep it for the next segment
            flush_window(K4_WIN, ~0ull);
            simt::syncwarp();
            uint32_t over = win[K4_WIN];
            simt::syncwarp();
            for (uint32_t v = lane; v < (K4_WIN + 16) / 16; v += 32) ((uint4*)win)[v] = make_uint4(0, 0, 0, 0);
            simt::syncwarp();
            if (lane == 0) win[0] = (uint8_t)over;
            simt::syncwarp();
            win_vo = wend;
        }

        // ---- next segment or finish ----
        o0 += seg_bytes;
        if (eob_lane < 32) {
            const uint32_t eob_rel = simt::shfl(c.end, eob_lane);
            const uint64_t eob_end = (s0 << 5) + eob_rel + 12;  // EOB code is 12 bits (sym 256)
            const uint64_t trailer_byte = (eob_end + 7) >> 3;   // relative to abase
            if (trailer_byte + 4 > end_byte) return ST_PENDING_GENERAL;  // truncated: K3 reports it
            // last partial window
            uint32_t left = (uint32_t)(oalign + o0 - win_vo);
            flush_window(left, oalign + o0);
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
        p0 = (s0 << 5) + simt::shfl(c.end, 31);
        seg_word += 32 * K4_SUBW;
    }
}

// Persistent kernel.  Streams the fast path declines are appended to worklist[] (count in *work_count)
// with status ST_PENDING_GENERAL; the host launches K3 over that list next, on the same stream.
FDB_GLOBAL void FDB_LAUNCH_BOUNDS(K4_WARPS * 32, 1)
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
