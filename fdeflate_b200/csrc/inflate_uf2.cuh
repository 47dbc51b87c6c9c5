// inflate_uf2.cuh -- K4, whole streams, TWO UNITS PER LANE in the count phase.
//
// inflate_uf.cuh decodes a segment of 32 sub-sequences ("units", 256 bits each) at a time: every lane warms up on the
// 128 bits before its unit, counts its unit, the chain is verified, offsets are scanned, and the lanes decode again to
// write.  The warm-up walks bits another lane walks too; it is 16 % of the kernel's instructions, and the per-segment
// chores (verify, scan, ballots, the tails of three loops) come on top.  Here a warp stages TWO segments (64 units) and
// lane i counts units 2i and 2i+1 one after the other -- one warm-up per 512 bits instead of one per 256 -- then the 64
// units are written in two phases of 32 lanes each (phase p, lane j: unit 32p + j, whose start and byte count come from
// lane 16p + j/2 by shuffle), through the same window and the same write loops as before.  Staging for two segments
// costs 2 KB more per warp; the kernel runs 28 warps per SM with a 3 KB window (a segment of PNG-filtered data expands
// to ~2.3 KB; larger ones take the careful path as before).
//
// Semantics are those of inflate_uf_run<K4_WHOLE> (same tables, same loops, same fallbacks to K3).
#pragma once
#include "inflate_uf.cuh"

namespace fdb {

#if K4_PAIR_UNITS

// staging: unit u of the pair of segments lives in plane u & 1, column u >> 1: word c of its row at
// stg[((u & 1) * K4_ROWS_ALLOC + c) * 32 + (u >> 1)] -- lane i reads its two units from bank i
FDB_DEVICE uint32_t k4_unit_word(uint32_t unit, uint32_t c) { return ((unit & 1u) * K4_ROWS_ALLOC + c) * 32u + (unit >> 1); }

FDB_DEVICE int32_t inflate_uf_whole2(const UfTabs& t, const uint32_t* hdr, K4Warp& ws, const K4Stream& s, uint32_t flags,
                                     uint64_t* out_len, uint64_t* consumed) {
    const unsigned lane = simt::lane_id();
    uint32_t* stg = ws.stg;
    const simt::saddr stg_s = simt::smem_addr(ws.stg);
    const simt::saddr row_a = stg_s + 4u * lane;                             // unit 2 * lane
    const simt::saddr row_b = row_a + 4u * K4_ROWS_ALLOC * 32u;              // unit 2 * lane + 1
    uint8_t* win = ws.win;
    const simt::saddr win_s = simt::smem_addr(ws.win);
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

    uint64_t seg_word = ((vstart >> 5) >> 2) << 2;  // first word of unit 0, 16-byte aligned
    uint64_t p0 = vstart;                           // true bit position where unit 0 starts
    uint64_t o0 = 0;                                // bytes produced so far
    uint32_t prev_nz = 1;                           // "the previous byte is non-zero, or there is none"
    uint64_t win_vo = 0;                            // virtual output position of win[0] (multiple of 16)
    AdlerAcc ad = {0, 0};
    const uint64_t lo_vo = oalign;

    for (uint32_t v = lane; v < (K4_WIN + 16) / 16; v += 32) ((uint4*)win)[v] = make_uint4(0, 0, 0, 0);
    simt::syncwarp();

    auto flush_vectors = [&](uint32_t nvec, uint64_t stream_end_vo) {
        if (win_vo >= lo_vo && win_vo + 16ull * nvec <= stream_end_vo) {
            uint8_t* const dst = obase + win_vo;
            const uint64_t pos0 = win_vo - oalign;
            for (uint32_t v = lane; v < nvec; v += 32) {
                uint4 q = ((const uint4*)win)[v];
                ((uint4*)win)[v] = make_uint4(0, 0, 0, 0);
                simt::stcs128((uint4*)(dst + 16u * v), q);
                adler_add16(ad, q, pos0 + 16u * v);
            }
            return;
        }
        for (uint32_t v = lane; v < nvec; v += 32) {
            uint4 q = ((const uint4*)win)[v];
            ((uint4*)win)[v] = make_uint4(0, 0, 0, 0);
            uint64_t vo = win_vo + 16ull * v;
            bool head_cut = vo < lo_vo;
            bool tail_cut = vo + 16 > stream_end_vo;
            if (!head_cut && !tail_cut) {
                simt::stcs128((uint4*)(obase + vo), q);
                adler_add16(ad, q, vo - oalign);
            } else {
                uint32_t w[4] = {q.x, q.y, q.z, q.w};
                for (uint32_t j = 0; j < 16; j++) {
                    uint64_t bpos = vo + j;
                    if (bpos >= lo_vo && bpos < stream_end_vo) {
                        uint32_t byte = (w[j >> 2] >> (8u * (j & 3u))) & 0xffu;
                        obase[bpos] = (uint8_t)byte;
                        adler_add1(ad, byte, bpos - oalign);
                    }
                }
            }
        }
    };

    const uint32_t PAIR_WORDS = 64u * K4_SUBW;                  // words of the two segments
    const uint32_t STAGED = K4_WARM + PAIR_WORDS + 4u;          // words staged (vectors of 4)
    for (;;) {
        if ((seg_word << 5) >= vend) return ST_PENDING_GENERAL;  // ran off the input without an EOB
        const uint64_t s0 = seg_word - K4_WARM;                  // first staged word (virtual word index)

        // ---- 1. stage 64 units: lane l takes the vectors 4l .. 4l+3 (the eight words of units 2l and 2l+1: bank l
        //         in every store) and lanes 0, 1 the two vectors behind unit 63 ----
        simt::syncwarp();
        const bool seg_inside = (s0 << 2) >= first_byte && (s0 << 2) + 4ull * STAGED <= end_byte;
        if (seg_inside) {
            const uint64_t pf = (s0 << 2) + 4u * PAIR_WORDS + 128u * lane;
            if (lane < 17 && pf < end_byte) simt::prefetch_l2(abase + pf);
        }
        for (uint32_t it = 0; it < 5; it++) {
            const uint32_t v = it < 4 ? 4u * lane + it : 128u + lane;
            if (v >= STAGED / 4) continue;
            uint64_t byte0 = (s0 << 2) + 16ull * v;  // relative to abase
            uint4 q = make_uint4(0, 0, 0, 0);
            if (seg_inside) {
                q = simt::ldg128((const uint4*)(abase + byte0));
            } else if (byte0 + 16 > first_byte && byte0 < end_byte) {
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
            // staged word 4v + j is word (4v + j) & 7 of unit (4v + j) >> 3 and word that + 8 of the unit before
            const uint32_t r1 = v >> 1, c1 = (v & 1) * 4;
            const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (uint32_t j = 0; j < 4; j++) {
                const uint32_t wr = simt::brev(w4[j]);  // MSB-first (LaneBits)
                if (r1 < 64) stg[k4_unit_word(r1, c1 + j)] = wr;
                if (r1 >= 1 && c1 + j + 8 < K4_ROWW) stg[k4_unit_word(r1 - 1, c1 + j + 8)] = wr;
            }
        }
        simt::syncwarp();

        // ---- 2. count: lane i walks units 2i and 2i+1 (one warm-up for both) ----
        uint32_t start = warm_up(t, row_a, lane != 0);
        if (lane == 0) start = (uint32_t)(p0 - (s0 << 5));
        LaneCount c1 = count_tokens(t, row_a, start, start != K4_INVALID);
        auto second = [&](const LaneCount& a) {
            const bool go = a.end != K4_INVALID && !(a.flags & CF_EOB);
            return count_tokens(t, row_b, go ? a.end - 32u * K4_SUBW : 0u, go);
        };
        LaneCount c2 = second(c1);

        // ---- 3. verify the chain: my start must be my predecessor's end (lanes are two units apart) ----
        uint32_t eob_lane = 32;
        for (;;) {
            const uint32_t my_eob = (start != K4_INVALID && ((c1.flags | c2.flags) & CF_EOB)) ? 1u : 0u;
            const uint32_t prev_end = simt::shfl_up(c2.end, 1);  // (relative to the predecessor's second row)
            const uint32_t prev_eob = simt::shfl_up(my_eob, 1);
            const uint32_t want = prev_end - 32u * K4_SUBW;
            const bool mismatch = (lane > 0) && (prev_end == K4_INVALID || start != want || prev_eob);
            const uint32_t mm = simt::ballot(mismatch);
            const uint32_t em = simt::ballot(my_eob != 0);
            const uint32_t first_mis = mm ? simt::ffs(mm) - 1 : 32;
            const uint32_t first_eob = em ? simt::ffs(em) - 1 : 32;
            if (first_eob < first_mis) {  // the stream ends inside a verified lane
                eob_lane = first_eob;
                break;
            }
            if (first_mis == 32) break;  // every lane verified, no EOB in these segments
            const bool redo = mismatch && !prev_eob && prev_end != K4_INVALID;
            if (mismatch) start = redo ? want : K4_INVALID;
            const LaneCount d1 = count_tokens(t, row_a, start, redo);
            LaneCount d2 = {K4_INVALID, 0, 0};
            {
                const bool go = redo && d1.end != K4_INVALID && !(d1.flags & CF_EOB);
                const LaneCount e2 = count_tokens(t, row_b, go ? d1.end - 32u * K4_SUBW : 0u, go);
                if (go) d2 = e2;
            }
            if (mismatch) {
                c1 = d1;  // (unresolved lanes get end = INVALID, cnt = 0, flags = 0)
                c2 = d2;
            }
        }
        if (lane > eob_lane) {
            c1.cnt = c2.cnt = 0;
            c1.flags = c2.flags = 0;
        }
        const uint32_t eob_in_first = (lane == eob_lane && (c1.flags & CF_EOB)) ? 1u : 0u;
        if (eob_in_first) {
            c2.cnt = 0;
            c2.flags = 0;
        }
        // the unit that holds the end-of-block code (64 = none)
        const uint32_t eob_unit = eob_lane < 32 ? 2u * eob_lane + (simt::shfl(eob_in_first, eob_lane) ? 0u : 1u) : 64u;
        if (simt::any(((c1.flags | c2.flags) & CF_BAD) != 0)) return ST_PENDING_GENERAL;

        // ---- 4. scan ----
        const uint32_t lane_cnt = c1.cnt + c2.cnt;
        const uint32_t incl = simt::scan_incl_add(lane_cnt);
        const uint64_t pair_bytes = simt::shfl(incl, 31);
        if (o0 + pair_bytes > s.cap) return ST_PENDING_GENERAL;  // K3 reports OutputTooLarge
        const uint32_t lane_base = incl - lane_cnt;               // bytes of these segments before my first unit
        {
            // a unit that opens with a run needs a zero byte before it.  Inside a lane: unit 2i+1 after unit 2i; across
            // lanes: the last byte of the nearest lane below that produced anything, else of the segments before
            const uint32_t first_run = c1.cnt ? (c1.flags & CF_FIRSTRUN) : (c2.flags & CF_FIRSTRUN);
            const uint32_t last_nz = c2.cnt ? (c2.flags & CF_LASTNZ) : (c1.flags & CF_LASTNZ);
            const bool inner = c1.cnt != 0 && c2.cnt != 0 && (c2.flags & CF_FIRSTRUN) != 0 && (c1.flags & CF_LASTNZ) != 0;
            const uint32_t has_mask = simt::ballot(lane_cnt != 0);
            const uint32_t nz_mask = simt::ballot(lane_cnt != 0 && last_nz != 0);
            const uint32_t below = has_mask & simt::lanemask_lt();
            const uint32_t pred_nz = below ? ((nz_mask >> (31u - simt::clz(below))) & 1u) : prev_nz;
            if (simt::any(inner || (lane_cnt != 0 && first_run != 0 && pred_nz != 0))) return ST_PENDING_GENERAL;
            if (has_mask) prev_nz = (nz_mask >> (31u - simt::clz(has_mask))) & 1u;
        }

        // ---- 5. write, 32 units at a time ----
        const uint32_t c1_valid_end = (c1.end != K4_INVALID && !(c1.flags & CF_EOB)) ? c1.end - 32u * K4_SUBW : K4_INVALID;
        for (uint32_t p = 0; p < 2; p++) {
            // my unit: 32p + lane, counted by lane `owner` as its first (h = 0) or second (h = 1) unit
            const uint32_t unit = 32u * p + lane;
            const uint32_t owner = 16u * p + (lane >> 1), h = lane & 1u;
            const uint32_t o_start = simt::shfl(start, owner), o_start2 = simt::shfl(c1_valid_end, owner);
            const uint32_t o_cnt1 = simt::shfl(c1.cnt, owner), o_cnt2 = simt::shfl(c2.cnt, owner);
            const uint32_t o_base = simt::shfl(lane_base, owner);
            const uint32_t ustart = h ? o_start2 : o_start;
            const uint32_t ucnt = h ? o_cnt2 : o_cnt1;
            // bytes of the pair before this phase, and of this phase
            const uint32_t before = p ? simt::shfl(incl, 15) : 0u;
            const uint32_t phase_bytes = p ? (uint32_t)pair_bytes - before : simt::shfl(incl, 15);
            uint64_t op = oalign + o0 + o_base + (h ? o_cnt1 : 0u);  // my virtual output position
            const uint64_t my_end_vo = op + ucnt;
            const uint64_t seg_end_vo = oalign + o0 + before + phase_bytes;
            const simt::saddr wrow = stg_s + 4u * k4_unit_word(unit, 0);
            LaneBits b;
            lb_start(b, wrow, ustart != K4_INVALID ? ustart : 0u);
            uint32_t fin = (ustart == K4_INVALID || unit > eob_unit) ? 1u : 0u;  // no more tokens to decode
            uint32_t run_bad = 0;   // some run of this unit follows a non-zero byte (see inflate_uf.cuh)
            uint32_t gone_nz = 0;   // my last byte has left the window and was non-zero
            for (;;) {
                // A round: the units whose output ends inside the window are written whole (fast path).  The first unit
                // that does not fit stops the round -- the units behind it wait, the window is flushed up to where that
                // unit starts and slides there -- and only a unit that does not fit an empty window either (long runs)
                // is written piecewise, a window at a time (careful path).
                const uint64_t wend = win_vo + K4_WIN;
                const uint32_t stuck_mask = simt::ballot(!fin && my_end_vo > wend);
                const uint32_t first_stuck = stuck_mask ? simt::ffs(stuck_mask) - 1u : 32u;
                const bool mine = !fin && lane <= first_stuck && (lane < first_stuck || op - win_vo < 16);
                uint32_t wp = mine ? (uint32_t)(op - win_vo) : 0u;
                if (mine && lane < first_stuck) {
                    // fast path: everything this unit still has to write fits in the window
                    simt::saddr wptr = win_s + wp;
                    const simt::saddr wfirst = wptr;
                    auto run_check = [&]() { run_bad |= wptr != wfirst ? simt::lds8(wptr - 1u) : gone_nz; };
                    while (!fin && b.rp <= K4_LIM_HI - K4_PAIR) {
                        const uint32_t bits = lb_peek(b);
                        const uint32_t e1 = wt_at(t, bits);
                        uint32_t n;
                        if (e1 < UW_LITERAL_MIN) {
                            const uint32_t w = e1 >> UW_SPECIAL_SHIFT;
                            if (w & UW_EOB) {
                                fin = 1;
                                n = 0;
                            } else {
                                uint32_t len, bd;
                                uf_long_run(w, bits, n, len, bd);
                                run_check();
                                wptr += len;
                            }
                        } else {
                            simt::sts8(wptr, e1 >> 5);
                            simt::sts8_if(wptr + 1, e1 >> 13, (int32_t)e1 < 0);
                            simt::sts8_if(wptr + 2, e1 >> 21, e1 >= (3u << 30));
                            wptr += e1 >> 30;
                            const uint32_t e2 = wt_at(t, lb_skip(bits, e1));
                            simt::sts8_if(wptr, e2 >> 5, e2 >= (1u << 30));
                            simt::sts8_if(wptr + 1, e2 >> 13, (int32_t)e2 < 0);
                            simt::sts8_if(wptr + 2, e2 >> 21, e2 >= (3u << 30));
                            wptr += e2 >> 30;
                            n = (e1 & 31u) + (e2 & 31u);
                        }
                        lb_advance(b, n);
                    }
                    while (!fin && b.rp <= K4_LIM_HI - 12u) {
                        const uint32_t bits = lb_peek(b);
                        const uint32_t e1 = wt_at(t, bits);
                        uint32_t n;
                        if (e1 < UW_LITERAL_MIN) {
                            const uint32_t w = e1 >> UW_SPECIAL_SHIFT;
                            if (w & UW_EOB) {
                                fin = 1;
                                n = 0;
                            } else {
                                uint32_t len, bd;
                                uf_long_run(w, bits, n, len, bd);
                                run_check();
                                wptr += len;
                            }
                        } else {
                            simt::sts8(wptr, e1 >> 5);
                            simt::sts8_if(wptr + 1, e1 >> 13, (int32_t)e1 < 0);
                            simt::sts8_if(wptr + 2, e1 >> 21, e1 >= (3u << 30));
                            wptr += e1 >> 30;
                            n = e1 & 31u;
                        }
                        lb_advance(b, n);
                    }
                    while (!fin && b.rp < K4_LIM_HI) {
                        const uint32_t bits = lb_peek(b);
                        const uint32_t e = wt_at(t, bits);
                        uint32_t n;
                        if (e < UW_LITERAL_MIN) {
                            const uint32_t w = e >> UW_SPECIAL_SHIFT;
                            if (w & UW_EOB) {
                                fin = 1;
                                n = 0;
                            } else {
                                uint32_t len, bd;
                                uf_long_run(w, bits, n, len, bd);
                                run_check();
                                wptr += len;
                            }
                        } else {
                            simt::sts8(wptr, e >> 5);
                            wptr += 1u;
                            n = (ct_at(t, bits) >> UC_FIRST_SHIFT) & 15u;
                        }
                        lb_advance(b, n);
                    }
                    fin = 1;
                    wp = (uint32_t)(wptr - win_s);
                } else if (mine) {
                    // careful path: this unit's output crosses the window end (long runs)
                    while (wp < K4_WIN && !fin) {
                        const uint32_t bits = lb_peek(b);
                        const uint32_t e = wt_at(t, bits);
                        uint32_t n;
                        if (e < UW_LITERAL_MIN) {
                            const uint32_t w = e >> UW_SPECIAL_SHIFT;
                            if (w & UW_EOB) {
                                fin = 1;
                                n = 0;
                            } else {
                                uint32_t len, bd;
                                uf_long_run(w, bits, n, len, bd);
                                run_bad |= gone_nz;
                                gone_nz = 0;
                                wp += len;
                            }
                        } else {
                            const uint32_t byte = (e >> 5) & 0xffu;
                            simt::sts8(win_s + wp, byte);
                            gone_nz = byte;
                            wp += 1u;
                            n = (ct_at(t, bits) >> UC_FIRST_SHIFT) & 15u;
                        }
                        lb_advance(b, n);
                        if (b.rp >= K4_LIM_HI) fin = 1;
                    }
                }
                if (mine) op = win_vo + wp;
                simt::syncwarp();
                if (!stuck_mask) break;  // the rest of this phase is in the window: leave it there
                // flush up to the first unit that is not done (the whole window if the careful path has filled it) and
                // slide the window base to the vector that holds that unit's next byte
                const uint64_t stuck_op = simt::shfl(op, first_stuck);
                const uint64_t ahead = (stuck_op - win_vo) >> 4;  // (a long run can reach beyond the window: zeros)
                const uint32_t nvec = ahead < K4_WIN / 16 ? (uint32_t)ahead : K4_WIN / 16;
                const uint4 tail = ((const uint4*)win)[nvec];
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
            if (simt::any(run_bad != 0)) return ST_PENDING_GENERAL;  // K3 replicates the byte
            adler_fold(ad);
            if (eob_unit < 32u * (p + 1u)) break;  // the stream ends in this phase
            // flush the finished vectors of this phase and slide the window base to the vector that holds the next
            // output byte (its already-written bytes move to win[0..16))
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

        // ---- next pair of segments or finish ----
        o0 += pair_bytes;
        if (eob_unit < 64) {
            const uint32_t eob_rel = simt::shfl(eob_in_first ? c1.end : c2.end, eob_lane);  // position of the code in that unit's row
            const uint64_t eob_end = ((s0 + (uint64_t)K4_SUBW * eob_unit) << 5) + eob_rel + 12;  // EOB = 12 bits
            const uint64_t trailer_byte = (eob_end + 7) >> 3;   // relative to abase
            if (trailer_byte + 4 > end_byte) return ST_PENDING_GENERAL;  // truncated: K3 reports it
            uint32_t left = (uint32_t)(oalign + o0 - win_vo);
            flush_vectors((left + 15) / 16, oalign + o0);
            simt::syncwarp();
            const uint8_t* tr = abase + trailer_byte;
            uint32_t stored = ((uint32_t)simt::ldg8(tr) << 24) | ((uint32_t)simt::ldg8(tr + 1) << 16) |
                              ((uint32_t)simt::ldg8(tr + 2) << 8) | (uint32_t)simt::ldg8(tr + 3);
            uint32_t got = adler_finish_warp(ad, o0);
            *out_len = o0;
            *consumed = trailer_byte + 4 - first_byte;
            if (!(flags & FLAG_IGNORE_ADLER32) && got != stored) return ST_WRONG_CHECKSUM;
            return ST_OK;
        }
        p0 = ((s0 + (uint64_t)K4_SUBW * 63) << 5) + simt::shfl(c2.end, 31);
        seg_word += 64 * K4_SUBW;
    }
}

#endif  // K4_PAIR_UNITS

}  // namespace fdb
