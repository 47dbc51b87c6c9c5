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
//   1. stage   : coalesced 16-byte loads -> transposed per-lane rows in shared memory (conflict-free reads), every word
//                bit-reversed once so that the lane readers are MSB-first and a window's top 12 bits index the tables
//   2. count   : lane i starts WARM words BEFORE its sub-sequence at a guessed bit position and walks table entries
//                (up to six literals or one short run each, two entries per 32-bit window) until it has crossed its
//                boundary -- Huffman codes self-synchronise within a few tokens -- and is taken back to the first TOKEN
//                boundary behind it by the boundary table (uf_land); then it counts the bytes of its own sub-sequence
//                the same way.  Lane 0 starts at the known true position.
//   3. verify  : lane i's start must equal lane i-1's end; since lane 0 is exact this proves every
//                lane.  A lane that had not synchronised is re-run from its predecessor's end
//                (loop until consistent; normally zero rounds).
//   4. scan    : exclusive prefix sum of byte counts -> output offsets; ballots check that no lane opens with a run
//                behind a non-zero byte (the decoder only knows runs of zeros).
//   5. write   : lanes decode again (entries of up to three literals), bounded by their byte counts, gather the literals
//                in a register and store 32-bit words into a zero-initialised shared-memory window; zero runs are
//                skipped, not written.  Finished 16-byte vectors leave with coalesced stores and feed adler32 on the way.
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

// K4_PAIR_UNITS: whole streams are decoded two segments at a time, every lane counting two units behind one warm-up
// (inflate_uf2.cuh).  Staging for two segments costs 2 KB more per warp: 28 warps per SM and a 3 KB window instead of
// 32 warps and 4 KB.
#ifndef K4_PAIR_UNITS
#define K4_PAIR_UNITS 0
#endif
// 28 warps, not 32: 72 registers per thread instead of 64 (room for the hoisted staging loads, K4_STAGE_HOIST), and
// 148 x 28 = 4144 warps still take the bench's 4096 streams in one round.  Measured (profiles/r04_variants.txt): 32 warps
// 1.272 ms, 32 + hoist 1.248, 30 + hoist 1.248, 28 1.268, 28 + hoist 1.232 ms.
#ifndef K4_WARPS_PER_CTA
#define K4_WARPS_PER_CTA 28
#endif
static const int K4_WARPS = K4_WARPS_PER_CTA;  // warps per CTA (one CTA per SM; they share the 40 KiB of tables)
static const uint32_t K4_SUBW = 8;     // 32-bit words of compressed data per lane per segment
static const uint32_t K4_WARM = 4;     // warm-up words before a lane's sub-sequence
static const uint32_t K4_TAILW = 3;    // words after it (token overrun + two-word look-ahead)
static const uint32_t K4_ROWW = K4_WARM + K4_SUBW + K4_TAILW;  // 15 words seen by one lane
static const uint32_t K4_ROWS_ALLOC = K4_ROWW + 1;             // +1: the look-ahead may touch one more
static const uint32_t K4_SEG_WORDS = K4_WARM + 32 * K4_SUBW + 4;  // words staged per segment (vectors of 4)
static const uint32_t K4_LIM_LO = 32u * K4_WARM;               // a lane's sub-sequence, in bits of its row
static const uint32_t K4_LIM_HI = 32u * (K4_WARM + K4_SUBW);
static const uint32_t K4_PAIR = 24;    // two table entries consume at most 2 x 12 bits
#ifndef K4_WIN_BYTES
#define K4_WIN_BYTES (K4_PAIR_UNITS ? 3008 : 3904)
#endif
static const uint32_t K4_WIN = K4_WIN_BYTES;  // output window bytes (a segment normally expands to ~2.3 KiB)
static const uint32_t K4_INVALID = 0xffffffffu;

// Staging is TRANSPOSED and PRIVATE per lane: row i holds the words lane i can ever touch
// (WARM words before its sub-sequence, the sub-sequence, TAILW after), word c of row i at
// stg[c * 32 + i].  Every lane therefore reads bank == lane (never a conflict) and walks its row
// with a constant +128-byte pointer step.  Overlapping words are simply stored twice.
struct K4Warp {
    uint32_t stg[(K4_PAIR_UNITS ? 2 : 1) * K4_ROWS_ALLOC * 32];
    uint8_t win[K4_WIN + 16];
};

struct K4Smem {
    uint32_t wt[4096];
    uint32_t ct[4096];
    uint16_t bt[4096];
    K4Warp warp[K4_WARPS];
};

struct UfDecTables {
    uint32_t wt[4096];     // UW write table (fdb_common.h) for HUFFMAN_LENGTHS
    uint32_t ct[4096];     // UC count table
    uint16_t bt[4096];     // UB boundary table
    uint32_t header[14];   // the constant 54 header bytes
};

// shared-window addresses of the two tables
struct UfTabs {
    simt::saddr wt, ct, bt;
};
// `bits` is an MSB-FIRST window (lb_peek): its top twelve bits are the next twelve stream bits, first bit on top --
// which is the bit-reversed slot the tables are stored at, so a lookup is mask, shift-and-add, load.
// K4_FMA_PIPE: field extraction on the FMA pipe.  The integer ALU pipe (shifts, logic, selects, compares: one warp
// instruction per two cycles) is the kernel's binding unit (75 % busy, the FMA pipe 15 %), and a shift by a constant is a
// multiply: x >> k = hi32(x * 2^(32-k)), x << k = x * 2^k.  The factors are read from constant memory so that the
// compiler cannot turn the multiplies back into shifts; IMAD / IMAD.HI take a constant-bank operand directly.
// Measured (profiles/r03_k4_variants.txt): with the table lookups done this way the kernel is 8.5 % SLOWER (IMAD.HI sits
// in the decode dependency chain and is slower than the two ALU instructions it replaces); K4_FMA_PIPE == 2 moves only
// the literal extraction of the write loop, which is off that chain.
// K4_STAGE_HOIST: the three staging loads of a segment are issued together, ahead of the stores of the first (one trip
// to L2 per segment instead of three; profiles/r04_variants.txt)
#ifndef K4_STAGE_HOIST
#define K4_STAGE_HOIST 1
#endif
#ifndef K4_FMA_PIPE
#define K4_FMA_PIPE 0
#endif
#if K4_FMA_PIPE && !defined(FDB_EMUL)
__constant__ uint32_t k4_mul[4] = {4u, 4096u, 1u << 25, 0u};
#define K4_M4 k4_mul[0]
#define K4_M4096 k4_mul[1]
#define K4_M2P25 k4_mul[2]
FDB_DEVICE uint32_t k4_mulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
// base + 4 * uf_slot(bits): two multiply-adds, no ALU instruction
#endif
#if K4_FMA_PIPE == 1 && !defined(FDB_EMUL)
FDB_DEVICE simt::saddr uf_entry_addr(simt::saddr base, uint32_t bits) { return base + k4_mulhi(bits, K4_M4096) * K4_M4; }
#elif K4_FMA_PIPE == 3 && !defined(FDB_EMUL)
// shift on the ALU pipe, scale-and-add on the FMA pipe
FDB_DEVICE simt::saddr uf_entry_addr(simt::saddr base, uint32_t bits) { return base + (bits >> 20) * K4_M4; }
#else
// mask, then shift-and-add in one LEA.HI
FDB_DEVICE simt::saddr uf_entry_addr(simt::saddr base, uint32_t bits) {
#if !defined(FDB_EMUL)
    uint32_t a;
    asm("mad.hi.u32 %0, %1, 16384, %2;" : "=r"(a) : "r"(bits & 0xfff00000u), "r"(base));
    return a;
#else
    return base + ((bits & 0xfff00000u) >> 18);
#endif
}
#endif
FDB_DEVICE uint32_t wt_at(const UfTabs& t, uint32_t bits) { return simt::lds32_ro(uf_entry_addr(t.wt, bits)); }
FDB_DEVICE uint32_t ct_at(const UfTabs& t, uint32_t bits) { return simt::lds32_ro(uf_entry_addr(t.ct, bits)); }
FDB_DEVICE uint32_t bt_at(const UfTabs& t, uint32_t bits) { return simt::lds16_ro(t.bt + ((bits >> 20) << 1)); }
// the window behind the first n bits of `bits` (n = the low five bits of a table entry: the shift is taken mod 32)
FDB_DEVICE uint32_t lb_skip(uint32_t bits, uint32_t n) { return simt::funnel_l(0u, bits, n); }

// lane-private bit reader over the lane's staging row.  The staged words are BIT-REVERSED once, when they are staged
// (stream bit k of a word sits at bit 31 - k), so the reader is MSB-first: a 32-bit window is one funnel shift of
// (w0, w1) with the next stream bit on top, and its top twelve bits index the tables directly -- no BREV per lookup
// (it was 7 % of the kernel's instructions, on the slow XU pipe, in the middle of the decode dependency chain).
// w2 is fetched one word ahead so the shared-memory latency stays off that chain.
struct LaneBits {
    uint32_t w0, w1, w2;
    uint32_t rp;      // bit position relative to the start of the row
    simt::saddr nx;   // next word to fetch
};
FDB_DEVICE void lb_start(LaneBits& b, simt::saddr row, uint32_t rp) {
    b.rp = rp;
    const simt::saddr p = row + (rp >> 5) * 128u;
    b.w0 = simt::lds32(p);
    b.w1 = simt::lds32(p + 128u);
    b.w2 = simt::lds32(p + 256u);
    b.nx = p + 384u;
}
FDB_DEVICE uint32_t lb_peek(const LaneBits& b) { return simt::funnel_l(b.w1, b.w0, b.rp); }  // ((w0:w1) << rp mod 32) >> 32
// move to position nrp (b.rp <= nrp < b.rp + 32; only bit 5 of the two is compared, so callers may keep other fields
// above bit 9 of the position word: count_tokens does)
FDB_DEVICE void lb_advance_to(LaneBits& b, uint32_t nrp) {
#if !defined(FDB_EMUL)
    // the same four predicated instructions, spelled out (the compiler's version shuffles the three words through
    // temporaries: ten instructions per advance in the round-1 SASS)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 x;\n\t"
        "xor.b32 x, %4, %5;\n\tand.b32 x, x, 32;\n\tsetp.ne.u32 p, x, 0;\n\t"
        "@p mov.u32 %0, %1;\n\t@p mov.u32 %1, %2;\n\t@p ld.shared.u32 %2, [%3];\n\t@p add.u32 %3, %3, 128;\n\t}"
        : "+r"(b.w0), "+r"(b.w1), "+r"(b.w2), "+r"(b.nx)
        : "r"(nrp), "r"(b.rp)
        : "memory");
#else
    if ((nrp ^ b.rp) & 32u) {
        b.w0 = b.w1;
        b.w1 = b.w2;
        b.w2 = simt::lds32(b.nx);
        b.nx += 128u;
    }
#endif
    b.rp = nrp;
}
FDB_DEVICE void lb_advance(LaneBits& b, uint32_t n) { lb_advance_to(b, b.rp + n); }  // n < 32

// A special write-table entry that is not the end of block: a length token read from the 32-bit
// window `bits`.  w = the entry's special fields (entry >> UW_SPECIAL_SHIFT).  Sets the bits the token occupies
// (code + extra + distance bit), its length, and whether its distance bit is 1 (a distance the ultra-fast code does
// not have).
FDB_DEVICE void uf_long_run(uint32_t w, uint32_t bits, uint32_t& n, uint32_t& len, uint32_t& bad_dist) {
    const uint32_t nc = w & 15u, xb = (w >> 4) & 7u, v = simt::brev(bits) >> nc;  // (extra bits are LSB-first values)
    len = ((w >> 8) & 0x1ffu) + (v & ((1u << xb) - 1u));
    bad_dist = (v >> xb) & 1u;
    n = nc + xb + 1u;
}

// K4_WORD_STORES: the write loop gathers a lane's literals in a register and stores whole 32-bit words.
// A lane's bytes are consecutive, so it keeps the word under construction in `acc` (its low wptr & 3 bytes are filled) and
// stores it when an entry crosses into the next word: one store per four bytes instead of one per byte (the byte stores
// were a third of the kernel's shared-memory wavefronts, conflicting 2.5-3.2 ways each).  Words two lanes share: the lane
// that fills the word's LAST byte stores it in the loop, seeded with what the window already held below its first byte
// (bytes of earlier segments and rounds); every lane's unfinished last word is OR-ed in behind a warp barrier, after all
// the plain stores (the window is zero wherever nothing has been written).
#ifndef K4_WORD_STORES
#define K4_WORD_STORES 1
#endif
struct WinWriter {
    simt::saddr wptr;  // shared-window address of the next byte
    uint32_t acc;      // the bytes of the word at wptr & ~3 below wptr
};
// what the window holds below byte `at` in at's word (bytes of earlier segments and rounds): the seed of a lane's first word
FDB_DEVICE uint32_t ww_seed(simt::saddr at) {
    const uint32_t s = ((uint32_t)at & 3u) * 8u;
    return simt::lds32(at & ~(simt::saddr)3) & ~(0xffffffffu << s);
}
FDB_DEVICE void ww_start(WinWriter& w, simt::saddr at, uint32_t seed) {
    w.wptr = at;
    w.acc = seed;
}
// `take` literals (0..3), in the low bytes of `lits` with zeros above them
FDB_DEVICE void ww_put_lits(WinWriter& w, uint32_t lits, uint32_t take) {
    const uint32_t s = (uint32_t)w.wptr << 3;  // funnel shifts take it mod 32: 8 * (wptr & 3)
    const uint32_t lo = w.acc | simt::funnel_l(0u, lits, s);
    const uint32_t hi = simt::funnel_l(lits, 0u, s);  // what does not fit in this word (0 when s == 0)
    const simt::saddr nw = w.wptr + take;
#if !defined(FDB_EMUL)
    // one predicate for the store and the select (the compiler's version turns the bool into a register and back)
    asm volatile(
        "{\n\t.reg .pred q;\n\t.reg .b32 x;\n\t"
        "xor.b32 x, %1, %2;\n\tand.b32 x, x, 4;\n\tsetp.ne.u32 q, x, 0;\n\t"
        "and.b32 x, %2, 0xfffffffc;\n\t@q st.shared.u32 [x], %3;\n\tselp.b32 %0, %4, %3, q;\n\t}"
        : "=r"(w.acc)
        : "r"(nw), "r"(w.wptr), "r"(lo), "r"(hi)
        : "memory");
#else
    const bool cross = ((nw ^ w.wptr) & 4u) != 0;
    simt::sts32_if(w.wptr & ~(simt::saddr)3, lo, cross);
    w.acc = cross ? hi : lo;
#endif
    w.wptr = nw;
}
// the 0..3 literals of write-table entry e (0: none)
FDB_DEVICE void ww_put(WinWriter& w, uint32_t e) {
#if (K4_FMA_PIPE == 1 || K4_FMA_PIPE == 2) && !defined(FDB_EMUL)
    ww_put_lits(w, k4_mulhi(e * K4_M4, K4_M2P25), k4_mulhi(e, K4_M4));  // (e << 2) >> 7: bits 28..5; e >> 30
#else
    ww_put_lits(w, (e >> 5) & 0xffffffu, e >> 30);
#endif
}
// the leading literals of literal entry e (window `bits`), at most rem >= 1 of them; returns the bits they occupy when
// that matters to the caller (EXACT), else the bits of the whole entry
template <bool EXACT>
FDB_DEVICE uint32_t ww_put_upto(WinWriter& w, const UfTabs& t, uint32_t e, uint32_t bits, uint32_t rem) {
    const uint32_t cnt = e >> 30;
    if (rem >= cnt) {
        ww_put_lits(w, (e >> 5) & 0xffffffu, cnt);
        return e & 31u;
    }
    ww_put_lits(w, (e >> 5) & ~(0xffffffffu << (8u * rem)), rem);  // rem is 1 or 2
    if (!EXACT) return e & 31u;
    const uint32_t m = bt_at(t, bits);  // the same leading literals (fdb_common.h)
    return simt::ffs(rem == 1u ? m : (m & (m - 1u)));
}
// before a run: the unfinished word goes out as it is (the run's zeros complete it: a run is at least 3 bytes long)
FDB_DEVICE void ww_run(WinWriter& w) {
    if ((uint32_t)w.wptr & 3u) simt::sts32(w.wptr & ~(simt::saddr)3, w.acc);
    w.acc = 0;
}

struct LaneCount {
    uint32_t end;    // bit position (row-relative) where the lane stopped: first token boundary >= LIM_HI
    uint32_t cnt;    // bytes produced in [start, end)
    uint32_t flags;  // CF_*
};
// CF_BAD: a token the fast path does not decode (distance bit 1).  CF_FIRSTRUN / CF_LASTNZ let the warp check that no
// lane opens with a run behind a non-zero byte; a run behind a non-zero byte INSIDE a lane is found by the write loop.
enum : uint32_t { CF_EOB = 1, CF_BAD = 2, CF_FIRSTRUN = 4, CF_LASTNZ = 8 };

// The count and warm-up loops keep ONE word per lane:  acc = position in the row + K4_BIAS | bytes << 10 | (junk above
// bit 24), and add whole count-table entries to it (fdb_common.h).  The bias is a multiple of 32 (funnel shifts and the
// word-crossing test see the position mod 32 / its bit 5) chosen so that "position >= limit" is one bit test.
static const uint32_t K4_BIAS = 128;
static const uint32_t K4_POS_MASK = 0x3ffu;
static_assert(K4_LIM_HI + K4_PAIR + 20 + K4_BIAS < 1024 && K4_BIAS % 32 == 0, "position field of acc");
static_assert(K4_LIM_HI + K4_BIAS == 512 && K4_LIM_LO + K4_BIAS == 256, "loop bounds as bit tests");
static_assert(K4_LIM_HI + K4_PAIR + 32 <= 32 * K4_ROWW, "the walk may cross its limit by a pair of entries");
FDB_DEVICE uint32_t acc_pos(uint32_t acc) { return (acc & K4_POS_MASK) - K4_BIAS; }

// Where a walk that crossed `lim` (position + K4_BIAS) with its last pair of entries (e1, e2: literal groups or one short
// run each; e2 may be 0) should have stopped: at the first TOKEN boundary at or after the limit.  prev = acc before that
// trip (position < lim), bits = its window.  Returns acc at that boundary (bytes counted up to it) and whether the byte
// before it is non-zero.  The boundary table gives the answer without walking the tokens (fdb_common.h): this replaces
// the token-at-a-time tails, which ran at half the lanes and were a fifth of the kernel's instructions.
FDB_DEVICE uint32_t uf_land(const UfTabs& t, uint32_t prev, uint32_t bits, uint32_t e1, uint32_t e2, uint32_t lim, uint32_t& last_nz) {
    const uint32_t d = lim - (prev & K4_POS_MASK);          // 1 .. 24: bits from prev to the limit
    const uint32_t n1 = e1 & 31u;
    const bool in1 = d <= n1;                               // the limit falls into the first entry
    const uint32_t e = in1 ? e1 : e2;
    const uint32_t base = in1 ? prev : prev + e1;
    const uint32_t dd = in1 ? d : d - n1;                   // 1 .. 12: bits from the entry's start to the limit
    if (e & UC_RUN) {                                       // one token: it ends at or after the limit
        last_nz = 0;
        return base + e;
    }
    const uint32_t m = bt_at(t, in1 ? bits : lb_skip(bits, e1));
    const uint32_t off = dd - 1u + simt::ffs(m >> (dd - 1u));            // first boundary >= dd (the entry's end is one)
    const uint32_t k = simt::popc(m & ~(0xffffffffu << off));          // literals up to it
    last_nz = ((((m << 1) | 1u) >> (off - 2u)) & 1u) ^ 1u;             // a literal of two bits is the byte 0, no other is
    return base + off + (k << UC_CNT_SHIFT);
}

// Count the bytes of the tokens in [start, LIM_HI); stop at the first token boundary >= LIM_HI or at
// EOB (then end = position of the EOB code).  No early exits inside the loop, so lanes that still
// iterate stay converged and the others wait at the loop exit.  A lane that is done (end of block, or not taking
// part) has all position bits of acc set, which ends the loop without a separate flag.
FDB_DEVICE LaneCount count_tokens(const UfTabs& t, simt::saddr row, uint32_t start, uint32_t active) {
    LaneCount c = {K4_INVALID, 0, 0};
    LaneBits b;
    lb_start(b, row, active ? start : 0u);
    b.rp += K4_BIAS;  // b.rp is `acc` from here on
    uint32_t flags = 0;
    uint32_t bad = 0;      // some run had distance bit 1
    uint32_t eob_acc = 0;  // acc when the end-of-block code was read
    {  // does the lane open with a run token?
        const uint32_t bits = lb_peek(b);
        const uint32_t c1 = ct_at(t, bits);
        uint32_t fr = c1 & UC_RUN;
        if (c1 == 0) fr = ((wt_at(t, bits) >> UW_SPECIAL_SHIFT) & UW_EOB) ? 0u : 1u;
        if (fr) flags |= CF_FIRSTRUN;
    }
    if (!active) b.rp |= K4_POS_MASK;
    // two entries per 32-bit window (together <= 24 bits) until the position is at or behind LIM_HI; the last trip is
    // then taken back to the first token boundary (uf_land).  A special second entry is 0 ("0 bits, 0 bytes") and comes
    // back as a first entry; a special first entry is one token, which may cross the limit as it is.
    uint32_t l1 = 0, l2 = 0;       // the entries of the last trip (0, 0: a special token, or no trip)
    uint32_t pacc = b.rp, pbits = 0;  // acc and window before the last trip
    while (!(b.rp & 0x200u)) {  // position < LIM_HI
        const uint32_t bits = lb_peek(b);
        const uint32_t c1 = ct_at(t, bits);
        uint32_t nacc;
        pacc = b.rp;
        pbits = bits;
        if (c1 == 0) {
            const uint32_t w = wt_at(t, bits) >> UW_SPECIAL_SHIFT;
            if (w & UW_EOB) {  // (l1, l2 stay: what the byte before the end of block is)
                flags |= CF_EOB;
                eob_acc = b.rp;
                nacc = b.rp | K4_POS_MASK;
            } else {
                uint32_t n, len, bd;
                uf_long_run(w, bits, n, len, bd);
                nacc = b.rp + n + (len << UC_CNT_SHIFT);
                bad |= bd;
                l1 = 0;
                l2 = 0;
            }
        } else {
            const uint32_t c2 = ct_at(t, lb_skip(bits, c1));  // the shift is taken mod 32
            nacc = b.rp + c1 + c2;
            l1 = c1;
            l2 = c2;
        }
        lb_advance_to(b, nacc);
    }
    if (active) {
        uint32_t acc = b.rp, last_nz = 0;
        if (flags & CF_EOB) {
            acc = eob_acc;
            last_nz = ((l2 ? l2 : l1) & UC_ENDNZ) ? 1u : 0u;
        } else if (l1) {
            acc = uf_land(t, pacc, pbits, l1, l2, K4_LIM_HI + K4_BIAS, last_nz);
        }
        c.end = acc_pos(acc);
        c.cnt = (acc >> UC_CNT_SHIFT) & 0x3fffu;
        c.flags = flags | (bad ? CF_BAD : 0u) | (last_nz ? CF_LASTNZ : 0u);
    }
    return c;
}

// warm-up: tokens from a guessed start until the first boundary >= LIM_LO
FDB_DEVICE uint32_t warm_up(const UfTabs& t, simt::saddr row, uint32_t active) {
    LaneBits b;
    lb_start(b, row, 0u);
    b.rp = active ? K4_BIAS : K4_POS_MASK;  // `acc` as in count_tokens (the byte field is never read here)
    uint32_t dead = 0;
    uint32_t l1 = 0, l2 = 0, pacc = b.rp, pbits = 0;
    while (!(b.rp & 0x300u)) {  // position < LIM_LO
        const uint32_t bits = lb_peek(b);
        const uint32_t c1 = ct_at(t, bits);
        uint32_t nacc;
        pacc = b.rp;
        pbits = bits;
        if (c1 == 0) {
            const uint32_t w = wt_at(t, bits) >> UW_SPECIAL_SHIFT;
            l1 = 0;
            l2 = 0;
            nacc = b.rp + (w & 15u) + ((w >> 4) & 7u) + 1u;
            if (w & UW_EOB) {  // speculative EOB: this lane has no valid guess
                dead = 1;
                nacc = b.rp | K4_POS_MASK;
            }
        } else {
            const uint32_t c2 = ct_at(t, lb_skip(bits, c1));
            nacc = b.rp + c1 + c2;
            l1 = c1;
            l2 = c2;
        }
        lb_advance_to(b, nacc);
    }
    if (!active || dead) return K4_INVALID;
    uint32_t last_nz;
    const uint32_t acc = l1 ? uf_land(t, pacc, pbits, l1, l2, K4_LIM_LO + K4_BIAS, last_nz) : b.rp;
    return acc_pos(acc);
}

struct K4Stream {
    const uint8_t* in;
    uint64_t n;
    uint8_t* out;
    uint64_t cap;
};

// ---- spans: a long stream decoded by several warps --------------------------------------------
// A stream whose compressed size is many segments long is cut into SPANS of K4_SPAN_SEGS segments on
// the stream's own segment grid.  Where a span starts is not known in advance, so it is found the way a
// lane finds its sub-sequence: a COUNT pass starts one whole segment early at a guessed bit, drops
// that segment (by its end the walk sits on true token boundaries, Huffman self-synchronisation over
// 8192 bits), then counts the bytes of its own segments; the position where it started counting must
// equal the position where the span before it stopped, and span 0 starts at the exact first data bit, so
// equality along the chain proves every span (split_scan_kernel; anything else sends the stream to K3).
// A prefix sum of the byte counts gives every span its output offset, and a WRITE pass decodes the spans
// again, each into its own part of the slot.  K4_WHOLE is the ordinary one-warp-per-stream decode.
enum : int { K4_WHOLE = 0, K4_COUNT = 1, K4_WRITE = 2 };
static const uint32_t K4_SPAN_SEGS = 64;                            // segments per span (64 KiB of compressed data)
static const uint64_t K4_SPAN_WORDS = (uint64_t)K4_SPAN_SEGS * 32 * K4_SUBW;
enum : uint32_t { SP_EOB = 1, SP_FIRSTRUN = 2, SP_LASTNZ = 4, SP_DEAD = 8, SP_SKIP = 16, SP_SYNCFAIL = 32 };
static const uint32_t K4_NO_ITEM = 0xffffffffu;
static const uint64_t K4_SPLIT_MIN_BYTES = 2ull * K4_SPAN_WORDS * 4;  // default: streams of >= 2 spans (128 KiB) are split (with lane records a cut costs one extra segment per span)

// LANE RECORDS.  The count pass knows, for every lane of every segment of its span, where the lane's sub-sequence
// starts and how many bytes it holds -- exactly what the write pass would otherwise find again with a warm-up and a
// count of its own (1.5 of its 2.5 walks over the bits).  So the count pass leaves one word per lane and segment in
// device scratch (128 B per KiB of compressed data) and the write pass of a span that has records goes straight from
// staging to the prefix sum.  Word = start bit in the lane's row (9 bits) | REC_INVALID | REC_EOB | bytes << 16
// (a lane holds at most 27 run tokens of 258 bytes).  Spans beyond the pool (fdb_set_split_scratch) count again.
enum : uint32_t { REC_INVALID = 0x4000u, REC_EOB = 0x8000u };
static const uint32_t K4_REC_WORDS = K4_SPAN_SEGS * 32;

// One span of one split stream (device scratch, filled by the three passes in turn).
struct K4Item {
    uint32_t stream, j;
    uint32_t flags;    // SP_* (count pass; SP_SKIP by the scan)
    uint32_t prev_nz;  // scan: state of the byte before the span
    uint64_t p_start, p_end, bytes, trailer_byte;  // count pass
    uint64_t o0;       // scan: output bytes before the span
    uint64_t s1, s2;   // write pass: adler partial sums (s2 reduced mod 65521)
};
// Per-batch bookkeeping of the split path (device scratch).
struct K4Split {
    K4Item* items;
    uint32_t item_cap;
    uint32_t* n_items;     // counter
    uint32_t* item0;       // [n] first item of stream i, or K4_NO_ITEM
    uint32_t* nspans;      // [n] spans reserved for stream i
    uint32_t* need;        // [n] spans up to and including the one with the end-of-block code (scan)
    uint32_t* done;        // [n] spans written so far
    uint32_t* failed;      // [n] some span's write pass gave up
    uint32_t* next_count;  // work counters of the three persistent passes
    uint32_t* next_scan;
    uint32_t* next_write;
    uint64_t min_bytes;    // streams at least this long are split
    uint32_t* rec;         // lane records (K4_REC_WORDS per item) of the first rec_items items, or null
    uint32_t rec_items;
};

struct K4Span {
    uint64_t seg_word;   // first word of lane 0's sub-sequence in the span's first segment
    uint64_t stop_word;  // no segment starts at or beyond this word
    uint64_t p0;         // bit where decoding starts: exact, or a guess when `discard`
    uint64_t o0;         // K4_WRITE: output bytes before the span
    uint32_t prev_nz;    // K4_WRITE: "the byte before the span is non-zero, or there is none"
    uint32_t discard;    // K4_COUNT: the first segment only synchronises
};
struct K4SpanOut {
    uint64_t p_start;       // K4_COUNT: exact bit the counting started from
    uint64_t p_end;         // first token of the next span, or the position of the end-of-block code
    uint64_t bytes;         // K4_COUNT: bytes produced by [p_start, p_end)
    uint64_t trailer_byte;  // SP_EOB: where the adler32 trailer starts (relative to the 16-byte aligned base)
    uint32_t flags;         // SP_*
    AdlerAcc ad;            // K4_WRITE: this lane's partial sums over the span's bytes
};

// K4_WHOLE: returns ST_OK / ST_WRONG_CHECKSUM, or ST_PENDING_GENERAL when the stream must go to K3.
// K4_COUNT / K4_WRITE: ST_OK with *so filled, or ST_PENDING_GENERAL.
template <int MODE>
FDB_DEVICE int32_t inflate_uf_run(const UfTabs& t, const uint32_t* hdr, K4Warp& ws, const K4Stream& s, uint32_t flags,
                                  uint64_t* out_len, uint64_t* consumed, const K4Span* sp, K4SpanOut* so,
                                  uint32_t* rec = nullptr) {
    const unsigned lane = simt::lane_id();
    uint32_t* stg = ws.stg;
    const simt::saddr row = simt::smem_addr(ws.stg + lane);
    uint8_t* win = ws.win;
    const simt::saddr win_s = simt::smem_addr(ws.win);
    if (MODE == K4_WHOLE) {
        *out_len = 0;
        *consumed = 0;
    }

    // ---- header must be the ultra-fast constant (ultrafast.rs:82-91) ----
    if (s.n < 54 + 2 + 4) return ST_PENDING_GENERAL;
    if (MODE == K4_WHOLE) {
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
    // Every match of the format replicates the previous byte, and the encoder only ever emits one
    // after a zero.  prev_nz = "the previous byte is non-zero, or there is none": a run token in
    // that state is not decoded here (K3 replicates the byte, or reports DistanceTooFarBack).
    uint32_t prev_nz = 1;
    uint64_t win_vo = 0;  // virtual output position of win[0] (multiple of 16)
    AdlerAcc ad = {0, 0};
    uint64_t lo_vo = oalign;  // virtual output position of the first byte this call may store
    uint32_t sync_seg = 0;    // K4_COUNT: the current segment only synchronises
    uint32_t first_seen = 0;  // K4_COUNT: some segment of the span has produced output
    if (MODE != K4_WHOLE) {
        seg_word = sp->seg_word;
        p0 = sp->p0;
        so->p_start = sp->p0;
        so->p_end = 0;
        so->bytes = 0;
        so->trailer_byte = 0;
        so->flags = 0;
    }
    if (MODE == K4_COUNT) {
        prev_nz = 0;  // what the first token needs of the byte before it is recorded (SP_FIRSTRUN), not tested
        sync_seg = sp->discard;
    }
    if (MODE == K4_WRITE) {
        o0 = sp->o0;
        prev_nz = sp->prev_nz;
        lo_vo = oalign + o0;
        win_vo = lo_vo & ~(uint64_t)15;
    }

    // zero the output window
    for (uint32_t v = lane; v < (K4_WIN + 16) / 16; v += 32) ((uint4*)win)[v] = make_uint4(0, 0, 0, 0);
    simt::syncwarp();

    // Store the finished vectors win[0 .. 16*nvec) at virtual position win_vo, feed adler32, and zero
    // them again.  Bytes outside [oalign, stream_end_vo) (first / last vector of the stream) are masked.
    auto flush_vectors = [&](uint32_t nvec, uint64_t stream_end_vo) {
        if (win_vo >= lo_vo && win_vo + 16ull * nvec <= stream_end_vo) {
            // every vector lies inside the stream (all segments but the first / last of a stream)
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

    for (;;) {
        if (MODE != K4_WHOLE && seg_word >= sp->stop_word) {
            // the span ends here, on the first token boundary at or after its last segment
            so->p_end = p0;
            if (MODE == K4_COUNT) {
                so->bytes = o0;
                so->flags |= prev_nz ? SP_LASTNZ : 0u;
            }
            if (MODE == K4_WRITE) {
                const uint32_t left = (uint32_t)(oalign + o0 - win_vo);
                flush_vectors((left + 15) / 16, oalign + o0);
                simt::syncwarp();
                so->ad = ad;
            }
            return ST_OK;
        }
        if ((seg_word << 5) >= vend) return ST_PENDING_GENERAL;  // ran off the input without an EOB
        const uint64_t s0 = seg_word - K4_WARM;                  // first staged word (virtual word index)

        // ---- 1. stage: coalesced 16-byte loads, every word scattered to the row(s) that can see it ----
        simt::syncwarp();
        // (lane l takes vectors 2l, 2l+1, then 64+l: within one store instruction every lane writes a
        // different row, i.e. a different bank)
        const bool seg_inside = (s0 << 2) >= first_byte && (s0 << 2) + 4ull * K4_SEG_WORDS <= end_byte;
        if (seg_inside) {
            // the next segment's lines are needed in a few thousand cycles: start them towards L2 now
            const uint64_t pf = (s0 << 2) + 4u * 32u * K4_SUBW + 128u * lane;
            if (lane < 9 && pf < end_byte) simt::prefetch_l2(abase + pf);  // (only lines that hold stream bytes)
        }
#if K4_STAGE_HOIST
        // all loads of the segment first (a load issued behind the stores of the vector before it waits out its own trip
        // to L2); needs the registers of a 28-warp CTA
        uint4 q3[3];
#pragma unroll
        for (uint32_t it = 0; it < 3; it++) {
            const uint32_t v = it < 2 ? 2u * lane + it : 64u + lane;
            q3[it] = make_uint4(0, 0, 0, 0);
            if (seg_inside && v < K4_SEG_WORDS / 4) q3[it] = simt::ldg128((const uint4*)(abase + (s0 << 2) + 16ull * v));
        }
#pragma unroll
#endif
        for (uint32_t it = 0; it < 3; it++) {
            const uint32_t v = it < 2 ? 2u * lane + it : 64u + lane;
            if (v >= K4_SEG_WORDS / 4) continue;
            uint64_t byte0 = (s0 << 2) + 16ull * v;  // relative to abase
            uint4 q = make_uint4(0, 0, 0, 0);
            if (seg_inside) {
#if K4_STAGE_HOIST
                q = q3[it];
#else
                q = simt::ldg128((const uint4*)(abase + byte0));
#endif
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
            // staged word t = 4v + j is word (t & 7) of row (t >> 3) and word (t & 7) + 8 of the row before
            const uint32_t r1 = v >> 1, c1 = (v & 1) * 4;
            const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (uint32_t j = 0; j < 4; j++) {
                const uint32_t wr = simt::brev(w4[j]);  // MSB-first from here on (LaneBits)
                if (r1 < 32) stg[(c1 + j) * 32 + r1] = wr;
                if (r1 >= 1 && c1 + j + 8 < K4_ROWW) stg[(c1 + j + 8) * 32 + (r1 - 1)] = wr;
            }
        }
        simt::syncwarp();

        // ---- 2. count ----  (or, for a span whose count pass left lane records, just read them)
        const bool replay = MODE == K4_WRITE && rec != nullptr;
        const uint32_t rec_seg = MODE == K4_WHOLE ? 0u : (uint32_t)((seg_word - (sp->stop_word - K4_SPAN_WORDS)) / (32 * K4_SUBW));
        uint32_t start;
        LaneCount c;
        uint32_t eob_lane = 32;
        if (replay) {
            const uint32_t r = rec[rec_seg * 32 + lane];
            start = (r & REC_INVALID) ? K4_INVALID : (r & 0x1ffu);
            c.end = K4_INVALID;  // (known after the write loop: where the lane's reader stops)
            c.cnt = r >> 16;
            c.flags = 0;
            const uint32_t em = simt::ballot((r & REC_EOB) != 0);
            if (em) eob_lane = simt::ffs(em) - 1;
        } else {
            start = warm_up(t, row, lane != 0);
            if (lane == 0) start = (uint32_t)(p0 - (s0 << 5));
            c = count_tokens(t, row, start, start != K4_INVALID);

            // ---- 3. verify the chain: my start must be my predecessor's end (rows are 32*SUBW bits apart) ----
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
                LaneCount c2 = count_tokens(t, row, start, redo);
                if (mismatch) c = c2;  // (unresolved lanes get end = INVALID, cnt = 0, flags = 0)
            }
            if (lane > eob_lane) {
                c.cnt = 0;
                c.flags = 0;
            }
            if (MODE == K4_COUNT && sync_seg) {
                // only where this segment ends matters; a walk that began on a guess may have read anything
                if (eob_lane < 32) {  // the guessed walk read an end-of-block code before it synchronised: guess again
                    so->flags |= SP_SYNCFAIL;
                    return ST_PENDING_GENERAL;
                }
                sync_seg = 0;
                p0 = ((s0 + (uint64_t)K4_SUBW * 31) << 5) + simt::shfl(c.end, 31);
                so->p_start = p0;
                seg_word += 32 * K4_SUBW;
                continue;
            }
            if (simt::any((c.flags & CF_BAD) != 0)) return ST_PENDING_GENERAL;
        }

        // ---- 4. scan ----
        const uint32_t incl = simt::scan_incl_add(c.cnt);
        const uint64_t seg_bytes = simt::shfl(incl, 31);
        if (o0 + seg_bytes > s.cap) return ST_PENDING_GENERAL;  // K3 reports OutputTooLarge
        uint64_t op = oalign + o0 + (incl - c.cnt);             // my virtual output position
        const uint64_t my_end_vo = op + c.cnt;
        if (!replay) {  // (the count pass of a span with records has checked all of this)
            // a lane that opens with a run needs a zero byte before it: the last byte of the nearest
            // lane below that produced anything, else the last byte of the previous segment
            const uint32_t has_mask = simt::ballot(c.cnt != 0);
            const uint32_t nz_mask = simt::ballot(c.cnt != 0 && (c.flags & CF_LASTNZ) != 0);
            const uint32_t below = has_mask & simt::lanemask_lt();
            const uint32_t pred_nz = below ? ((nz_mask >> (31u - simt::clz(below))) & 1u) : prev_nz;
            if (simt::any(c.cnt != 0 && (c.flags & CF_FIRSTRUN) != 0 && pred_nz != 0)) return ST_PENDING_GENERAL;
            if (MODE == K4_COUNT && !first_seen && has_mask) {
                // does the span open with a run?  (its first token is the first token of the first lane that
                // produced anything; whether the byte before it is a zero is known to the span before)
                first_seen = 1;
                const uint32_t fr_mask = simt::ballot(c.cnt != 0 && (c.flags & CF_FIRSTRUN) != 0);
                if ((fr_mask >> (simt::ffs(has_mask) - 1u)) & 1u) so->flags |= SP_FIRSTRUN;
            }
            if (has_mask) prev_nz = (nz_mask >> (31u - simt::clz(has_mask))) & 1u;
        }
        if (MODE == K4_COUNT && rec) {  // what the write pass needs of this segment
            const bool valid = start != K4_INVALID && lane <= eob_lane;
            rec[rec_seg * 32 + lane] = valid ? (start | (lane == eob_lane ? REC_EOB : 0u) | (c.cnt << 16)) : REC_INVALID;
        }

        // ---- 5. write ----
        // The window base slides with the output: it is the 16-byte vector holding the first byte of
        // this segment, so a segment whose output fits in K4_WIN is written by all lanes at once.
        // Only literals are stored; runs are zeros and the window is zero-initialised.
        const uint64_t seg_end_vo = oalign + o0 + seg_bytes;
        LaneBits b;
        lb_start(b, row, start != K4_INVALID ? start : 0u);
        uint32_t fin = (start == K4_INVALID || lane > eob_lane) ? 1u : 0u;  // no more tokens to decode
        // A run replicates the byte before it, and this decoder only knows runs of zeros: a run behind a non-zero byte
        // sends the stream to K3.  Across lanes that is settled before anything is written (CF_FIRSTRUN / CF_LASTNZ);
        // inside a lane the write loop looks at the byte it stored last, which is still in the window -- or, where the
        // window has moved on in between, at what the careful path remembers of it.
        uint32_t run_bad = 0;   // some run of this lane follows a non-zero byte
        uint32_t gone_nz = 0;   // my last byte has left the window and was non-zero
        for (; MODE != K4_COUNT;) {
            const uint64_t wend = win_vo + K4_WIN;
            const bool mine = !fin && op < wend;
            uint32_t wp = mine ? (uint32_t)(op - win_vo) : 0u;
#if K4_WORD_STORES
            simt::saddr tail_at = 0;
            uint32_t tail_word = 0;
            // (read before any lane stores into the window in this round: a lane on the careful path may store bytes
            // into the same word, above the ones kept here)
            const uint32_t seed = mine ? ww_seed(win_s + wp) : 0u;
            simt::syncwarp();
#endif
            if (mine && my_end_vo <= wend) {
                // fast path: everything this lane still has to write fits in the window
#if K4_WORD_STORES
                WinWriter ww;
                ww_start(ww, win_s + wp, seed);
                const simt::saddr wfirst = ww.wptr;
                // (a run: the byte before it is mine and in the window unless nothing of mine is there yet)
                auto run_token = [&](uint32_t len) {
                    ww_run(ww);
                    run_bad |= ww.wptr != wfirst ? simt::lds8(ww.wptr - 1u) : gone_nz;
                    ww.wptr += len;
                };
                // The lane knows how many bytes it has to produce, so the walk is bounded by bytes, not by bits: blind
                // pairs of entries while six more bytes are mine, then entries cut to what is left -- no token-at-a-time
                // tail (it ran at half the lanes).  Where the walk ends in bits only matters to a span's write pass.
                const simt::saddr wlast = win_s + (uint32_t)(my_end_vo - win_vo);
                while (!fin && ww.wptr + 6u <= wlast) {
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
                            run_token(len);
                        }
                    } else {
                        ww_put(ww, e1);
                        const uint32_t e2 = wt_at(t, lb_skip(bits, e1));  // special: 0 bytes, 0 bits -> next trip
                        ww_put(ww, e2 >= UW_LITERAL_MIN ? e2 : 0u);
                        n = (e1 & 31u) + (e2 & 31u);
                    }
                    lb_advance(b, n);
                }
                while (!fin && ww.wptr < wlast) {
                    const uint32_t bits = lb_peek(b);
                    const uint32_t e1 = wt_at(t, bits);
                    uint32_t n;
                    if (e1 < UW_LITERAL_MIN) {
                        const uint32_t w = e1 >> UW_SPECIAL_SHIFT;
                        if (w & UW_EOB) {  // (the count pass saw more bytes before it: cannot happen)
                            fin = 1;
                            n = 0;
                        } else {
                            uint32_t len, bd;
                            uf_long_run(w, bits, n, len, bd);
                            run_token(len);
                        }
                    } else {
                        n = ww_put_upto<MODE == K4_WRITE>(ww, t, e1, bits, (uint32_t)(wlast - ww.wptr));
                        const uint32_t bits2 = lb_skip(bits, e1);
                        const uint32_t e2 = wt_at(t, bits2);
                        const uint32_t rem = (uint32_t)(wlast - ww.wptr);
                        if (rem != 0 && e2 >= UW_LITERAL_MIN) n += ww_put_upto<MODE == K4_WRITE>(ww, t, e2, bits2, rem);
                    }
                    lb_advance(b, n);
                }
                fin = 1;
                wp = (uint32_t)(ww.wptr - win_s);
                if (wp & 3u) {  // my unfinished last word, OR-ed in below
                    tail_at = ww.wptr & ~(simt::saddr)3;
                    tail_word = ww.acc;
                }
#else
                simt::saddr wptr = win_s + wp;
                const simt::saddr wfirst = wptr;
                // (a run: the byte before it is mine and in the window unless nothing of mine is there yet)
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
                        // exact stores: the byte after this lane's last one belongs to the next lane
                        simt::sts8(wptr, e1 >> 5);
                        simt::sts8_if(wptr + 1, e1 >> 13, (int32_t)e1 < 0);
                        simt::sts8_if(wptr + 2, e1 >> 21, e1 >= (3u << 30));
                        wptr += e1 >> 30;
                        const uint32_t e2 = wt_at(t, lb_skip(bits, e1));  // bits >> n1; special: 0 bytes, 0 bits -> next trip
                        simt::sts8_if(wptr, e2 >> 5, e2 >= (1u << 30));
                        simt::sts8_if(wptr + 1, e2 >> 13, (int32_t)e2 < 0);
                        simt::sts8_if(wptr + 2, e2 >> 21, e2 >= (3u << 30));
                        wptr += e2 >> 30;
                        n = (e1 & 31u) + (e2 & 31u);
                    }
                    lb_advance(b, n);
                }
                while (!fin && b.rp <= K4_LIM_HI - 12u) {  // whole entries that still end at or before LIM_HI
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
                while (!fin && b.rp < K4_LIM_HI) {  // single tokens up to the first boundary >= LIM_HI
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
#endif
            } else if (mine) {
                // careful path: this lane's output crosses the window end (long runs); single tokens,
                // stop at the window end and resume after the flush
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
#if K4_WORD_STORES
            if (tail_word) simt::atoms_or(tail_at, tail_word);
            simt::syncwarp();
#endif
            if (seg_end_vo < wend) break;  // the rest of this segment fits: leave it in the window
            // the window is complete: flush all of it and slide by K4_WIN
            flush_vectors(K4_WIN / 16, ~0ull);
            simt::syncwarp();
            win_vo = wend;
        }
        if (MODE != K4_COUNT && simt::any(run_bad != 0)) return ST_PENDING_GENERAL;  // K3 replicates the byte

        // ---- next segment or finish ----
        if (replay) c.end = b.rp;  // first token boundary at or after the lane's sub-sequence, or the end-of-block code
        adler_fold(ad);
        o0 += seg_bytes;
        if (eob_lane < 32) {
            const uint32_t eob_rel = simt::shfl(c.end, eob_lane);
            const uint64_t eob_end = ((s0 + (uint64_t)K4_SUBW * eob_lane) << 5) + eob_rel + 12;  // EOB = 12 bits
            const uint64_t trailer_byte = (eob_end + 7) >> 3;   // relative to abase
            if (trailer_byte + 4 > end_byte) return ST_PENDING_GENERAL;  // truncated: K3 reports it
            if (MODE != K4_WHOLE) {
                so->p_end = eob_end - 12;
                so->trailer_byte = trailer_byte;
                so->flags |= SP_EOB;
            }
            if (MODE == K4_COUNT) {
                so->bytes = o0;
                so->flags |= prev_nz ? SP_LASTNZ : 0u;
                return ST_OK;
            }
            uint32_t left = (uint32_t)(oalign + o0 - win_vo);
            flush_vectors((left + 15) / 16, oalign + o0);
            simt::syncwarp();
            if (MODE == K4_WRITE) {
                so->ad = ad;
                return ST_OK;
            }
            const uint8_t* tr = abase + trailer_byte;
            uint32_t stored = ((uint32_t)simt::ldg8(tr) << 24) | ((uint32_t)simt::ldg8(tr + 1) << 16) |
                              ((uint32_t)simt::ldg8(tr + 2) << 8) | (uint32_t)simt::ldg8(tr + 3);
            uint32_t got = adler_finish_warp(ad, o0);
            *out_len = o0;
            *consumed = trailer_byte + 4 - first_byte;
            if (!(flags & FLAG_IGNORE_ADLER32) && got != stored) return ST_WRONG_CHECKSUM;
            return ST_OK;
        }
        // flush the finished vectors of this segment and slide the window base to the vector that
        // holds the next output byte (its already-written bytes move to win[0..16))
        if (MODE != K4_COUNT) {
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

#if K4_PAIR_UNITS
FDB_DEVICE int32_t inflate_uf_whole2(const UfTabs& t, const uint32_t* hdr, K4Warp& ws, const K4Stream& s, uint32_t flags,
                                     uint64_t* out_len, uint64_t* consumed);  // inflate_uf2.cuh
#endif
FDB_DEVICE int32_t inflate_uf_stream(const UfTabs& t, const uint32_t* hdr, K4Warp& ws, const K4Stream& s,
                                     uint32_t flags, uint64_t* out_len, uint64_t* consumed) {
#if K4_PAIR_UNITS
    return inflate_uf_whole2(t, hdr, ws, s, flags, out_len, consumed);
#else
    return inflate_uf_run<K4_WHOLE>(t, hdr, ws, s, flags, out_len, consumed, nullptr, nullptr);
#endif
}

// Persistent kernel, one CTA per SM.  Streams the fast path declines are appended to worklist[]
// (count in *work_count) with status ST_PENDING_GENERAL; the host launches K3 over that list next,
// on the same stream.
FDB_GLOBAL void FDB_LAUNCH_BOUNDS(K4_WARPS * 32, 1)
    inflate_uf_kernel(InflateBatch b, const UfDecTables* tables, uint32_t* next, uint32_t* worklist,
                      uint32_t* work_count, const uint32_t* split_item0) {
    FDB_DYN_SMEM(smem_raw);
    K4Smem& sm = *reinterpret_cast<K4Smem*>(smem_raw);
    FDB_SHARED uint32_t hdr[14];
    for (uint32_t i = threadIdx.x; i < 4096; i += blockDim.x) {
        sm.wt[i] = tables->wt[i];
        sm.ct[i] = tables->ct[i];
        sm.bt[i] = tables->bt[i];
    }
    if (threadIdx.x < 14) hdr[threadIdx.x] = tables->header[threadIdx.x];
    simt::syncthreads();
    const unsigned lane = simt::lane_id();
    K4Warp& ws = sm.warp[simt::warp_in_block()];
    const UfTabs t = {simt::smem_addr(sm.wt), simt::smem_addr(sm.ct), simt::smem_addr(sm.bt)};
    // The first stream of every warp is fixed: stream c + grid * w for warp w of CTA c, so that a batch of about one
    // stream per resident warp (the bench: 4096 streams on 4144 warps) gives every SM the same number of streams.
    // Handed out through the counter, the streams go to whichever warps ask first, an SM ends up with anything from
    // ~22 to 32 of them, and the kernel lasts as long as the fullest SM.  Further streams come from the counter.
    const uint32_t slots = gridDim.x * K4_WARPS;
    bool first = true;
    for (;;) {
        uint32_t i = blockIdx.x + gridDim.x * simt::warp_in_block();
        if (!first) {
            if (lane == 0) i = slots + simt::atomic_add(next, 1u);
            i = simt::shfl(i, 0);
        }
        first = false;
        if (i >= b.n) break;
        if (split_item0 && split_item0[i] != K4_NO_ITEM) continue;  // decoded span by span (below)
        K4Stream s = {b.in_base + b.in_off[i], b.in_len[i], b.out_base + b.out_off[i], b.out_cap[i]};
        uint64_t out_len = 0, consumed = 0;
        int32_t st = inflate_uf_stream(t, hdr, ws, s, b.flags, &out_len, &consumed);
        if (lane == 0) {
            b.status[i] = st;
            b.out_len[i] = out_len;
            if (b.consumed) b.consumed[i] = consumed;
            if (st == ST_PENDING_GENERAL) worklist[simt::atomic_add(work_count, 1u)] = i;
        }
        simt::syncwarp();
    }
}

// ---- split path, pass 0: which streams are decoded span by span --------------------------------
FDB_DEVICE void k4_geometry(const uint8_t* in, uint64_t n, uint64_t* first_byte, uint64_t* seg0, uint64_t* end_word) {
    *first_byte = (uint64_t)((uintptr_t)in & 15u);
    const uint64_t vstart = *first_byte * 8 + 53 * 8 + 5;
    *seg0 = ((vstart >> 5) >> 2) << 2;
    *end_word = ((*first_byte + n) * 8 + 31) >> 5;
}

FDB_GLOBAL void inflate_uf_plan_kernel(InflateBatch b, const UfDecTables* tables, K4Split sp) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    sp.item0[i] = K4_NO_ITEM;
    sp.nspans[i] = 0;
    sp.need[i] = 0;
    sp.done[i] = 0;
    sp.failed[i] = 0;
    const uint64_t n = b.in_len[i];
    if (n < sp.min_bytes) return;
    const uint8_t* in = b.in_base + b.in_off[i];
    for (uint32_t j = 0; j < 54; j++) {  // the constant header (ultrafast.rs:82-91)
        uint32_t want = (tables->header[j >> 2] >> (8u * (j & 3u))) & 0xffu;
        uint32_t got = simt::ldg8(in + j);
        if (j == 53) got &= 0x1fu;
        if (got != want) return;
    }
    uint64_t first_byte, seg0, end_word;
    k4_geometry(in, n, &first_byte, &seg0, &end_word);
    const uint64_t S = (end_word - seg0 + K4_SPAN_WORDS - 1) / K4_SPAN_WORDS;
    if (S < 2 || S > 0x7fffffffull) return;
    const uint32_t base = simt::atomic_add(sp.n_items, (uint32_t)S);
    if ((uint64_t)base + S > sp.item_cap) return;  // scratch exhausted: this stream stays with one warp
    sp.item0[i] = base;
    sp.nspans[i] = (uint32_t)S;
    for (uint32_t j = 0; j < (uint32_t)S; j++) {
        K4Item& it = sp.items[base + j];
        it.stream = i;
        it.j = j;
        it.flags = SP_DEAD;
    }
}

FDB_DEVICE uint32_t k4_item_count(const K4Split& sp) {
    const uint32_t c = *sp.n_items;
    return c < sp.item_cap ? c : sp.item_cap;
}
// items past the capacity were never reserved as a whole stream: an item is live iff its stream points at it
FDB_DEVICE bool k4_item_live(const K4Split& sp, uint32_t idx, const K4Item& it, uint32_t n) {
    return it.stream < n && sp.item0[it.stream] != K4_NO_ITEM && sp.item0[it.stream] + it.j == idx;
}

struct K4Tables {
    UfTabs t;
    K4Warp* ws;
};
FDB_DEVICE K4Tables k4_load_tables(unsigned char* smem_raw, const UfDecTables* tables) {
    K4Smem& sm = *reinterpret_cast<K4Smem*>(smem_raw);
    for (uint32_t i = threadIdx.x; i < 4096; i += blockDim.x) {
        sm.wt[i] = tables->wt[i];
        sm.ct[i] = tables->ct[i];
        sm.bt[i] = tables->bt[i];
    }
    simt::syncthreads();
    K4Tables r = {{simt::smem_addr(sm.wt), simt::smem_addr(sm.ct), simt::smem_addr(sm.bt)}, &sm.warp[simt::warp_in_block()]};
    return r;
}

// ---- pass 1: synchronise and count every span ------------------------------------------------
FDB_GLOBAL void FDB_LAUNCH_BOUNDS(K4_WARPS * 32, 1)
    inflate_uf_split_count_kernel(InflateBatch b, const UfDecTables* tables, K4Split sp) {
    FDB_DYN_SMEM(smem_raw);
    const uint32_t count = k4_item_count(sp);
    if (b.split_out && blockIdx.x == 0 && threadIdx.x == 0) *b.split_out = count;
    if (count == 0) return;
    const K4Tables kt = k4_load_tables(smem_raw, tables);
    const unsigned lane = simt::lane_id();
    for (;;) {
        uint32_t idx = 0;
        if (lane == 0) idx = simt::atomic_add(sp.next_count, 1u);
        idx = simt::shfl(idx, 0);
        if (idx >= count) break;
        K4Item& it = sp.items[idx];
        if (!k4_item_live(sp, idx, it, b.n)) continue;
        const uint32_t i = it.stream, j = it.j;
        K4Stream s = {b.in_base + b.in_off[i], b.in_len[i], b.out_base + b.out_off[i], b.out_cap[i]};
        uint64_t first_byte, seg0, end_word;
        k4_geometry(s.in, s.n, &first_byte, &seg0, &end_word);
        K4Span span;
        span.stop_word = seg0 + (uint64_t)(j + 1) * K4_SPAN_WORDS;
        span.o0 = 0;
        span.prev_nz = 0;
        if (j == 0) {
            span.seg_word = seg0;
            span.p0 = first_byte * 8 + 53 * 8 + 5;
            span.discard = 0;
        } else {
            span.seg_word = seg0 + (uint64_t)j * K4_SPAN_WORDS - 32 * K4_SUBW;  // one segment early
            span.p0 = span.seg_word << 5;
            span.discard = 1;
        }
        K4SpanOut so;
        uint32_t* const rec = (sp.rec && idx < sp.rec_items) ? sp.rec + (size_t)idx * K4_REC_WORDS : nullptr;
        int32_t st = inflate_uf_run<K4_COUNT>(kt.t, nullptr, *kt.ws, s, b.flags, nullptr, nullptr, &span, &so, rec);
        // A walk that starts on a guessed bit can run into an end-of-block pattern before it reaches a true
        // token boundary (likely only on data made of 12-bit codes); another starting bit reads other tokens.
        for (uint32_t retry = 1; retry < 10 && st != ST_OK && (so.flags & SP_SYNCFAIL); retry++) {
            span.p0 = (span.seg_word << 5) + 3u * retry;
            st = inflate_uf_run<K4_COUNT>(kt.t, nullptr, *kt.ws, s, b.flags, nullptr, nullptr, &span, &so, rec);
        }
        if (lane == 0) {
            it.p_start = so.p_start;
            it.p_end = so.p_end;
            it.bytes = so.bytes;
            it.trailer_byte = so.trailer_byte;
            it.flags = st == ST_OK ? so.flags : SP_DEAD;
        }
        simt::syncwarp();
    }
}

// ---- pass 2: chain check and output offsets, one warp per split stream -------------------------
// Span 0 starts at the exact first data bit; span j is proven by p_start[j] == p_end[j-1].  The stream is
// handed to the general kernel if the chain breaks anywhere before the span that holds the end-of-block
// code, if a span opens with a run behind a non-zero byte, or if the output does not fit the slot.
FDB_GLOBAL void inflate_uf_split_scan_kernel(InflateBatch b, K4Split sp, uint32_t* worklist, uint32_t* work_count) {
    const unsigned lane = simt::lane_id();
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = simt::atomic_add(sp.next_scan, 1u);
        i = simt::shfl(i, 0);
        if (i >= b.n) break;
        const uint32_t base = sp.item0[i];
        if (base == K4_NO_ITEM) continue;
        const uint32_t S = sp.nspans[i];
        bool ok = true;
        uint32_t need = 0;           // spans up to and including the end-of-block span
        uint64_t carry_end = 0;      // p_end of the span before this group of 32
        uint32_t carry_nz = 1;       // "the byte before this group is non-zero, or there is none"
        uint64_t carry_bytes = 0;
        for (uint32_t g = 0; g < S && ok && need == 0; g += 32) {
            const uint32_t j = g + lane;
            const bool have = j < S;
            K4Item it;
            it.flags = SP_DEAD;
            it.p_start = it.p_end = it.bytes = 0;
            if (have) it = sp.items[base + j];
            const uint32_t eob_mask = simt::ballot(have && !(it.flags & SP_DEAD) && (it.flags & SP_EOB));
            const uint32_t upto = eob_mask ? simt::ffs(eob_mask) - 1u : 31u;  // last lane that matters in this group
            uint64_t prev_end = simt::shfl_up(it.p_end, 1);
            uint32_t prev_flags = simt::shfl_up(it.flags, 1);
            if (lane == 0) {
                prev_end = carry_end;
                prev_flags = carry_nz ? SP_LASTNZ : 0u;
            }
            const bool relevant = have && lane <= upto;
            bool bad = relevant && (it.flags & SP_DEAD);
            bad = bad || (relevant && j > 0 && it.p_start != prev_end);
            bad = bad || (relevant && (it.flags & SP_FIRSTRUN) && (prev_flags & SP_LASTNZ));
            if (simt::any(bad)) ok = false;
            const uint64_t incl = simt::scan_incl_add((uint64_t)(relevant ? it.bytes : 0ull));
            const uint64_t o0 = carry_bytes + incl - (relevant ? it.bytes : 0ull);
            if (simt::any(relevant && o0 + it.bytes > b.out_cap[i])) ok = false;
            if (ok && relevant) {
                sp.items[base + j].o0 = o0;
                sp.items[base + j].prev_nz = (prev_flags & SP_LASTNZ) ? 1u : 0u;
            }
            if (eob_mask) need = g + upto + 1u;
            // (only full groups carry over)
            carry_end = simt::shfl(it.p_end, 31);
            carry_nz = (simt::shfl(it.flags, 31) & SP_LASTNZ) ? 1u : 0u;
            carry_bytes += simt::shfl(incl, 31);
        }
        if (need == 0) ok = false;  // no end of block anywhere: truncated, the general kernel reports it
        if (ok) {
            for (uint32_t j = need + lane; j < S; j += 32) sp.items[base + j].flags |= SP_SKIP;  // behind the end of block
            if (lane == 0) sp.need[i] = need;
        } else {
            for (uint32_t j = lane; j < S; j += 32) sp.items[base + j].flags |= SP_SKIP;
            if (lane == 0) {
                b.status[i] = ST_PENDING_GENERAL;
                b.out_len[i] = 0;
                if (b.consumed) b.consumed[i] = 0;
                worklist[simt::atomic_add(work_count, 1u)] = i;
            }
        }
        simt::syncwarp();
    }
}

// ---- pass 3: decode every span into its part of the slot; the last one to finish closes the stream ----
FDB_GLOBAL void FDB_LAUNCH_BOUNDS(K4_WARPS * 32, 1)
    inflate_uf_split_write_kernel(InflateBatch b, const UfDecTables* tables, K4Split sp, uint32_t* worklist,
                                  uint32_t* work_count) {
    FDB_DYN_SMEM(smem_raw);
    const uint32_t count = k4_item_count(sp);
    if (count == 0) return;
    const K4Tables kt = k4_load_tables(smem_raw, tables);
    const unsigned lane = simt::lane_id();
    for (;;) {
        uint32_t idx = 0;
        if (lane == 0) idx = simt::atomic_add(sp.next_write, 1u);
        idx = simt::shfl(idx, 0);
        if (idx >= count) break;
        K4Item& it = sp.items[idx];
        if (!k4_item_live(sp, idx, it, b.n) || (it.flags & SP_SKIP)) continue;
        const uint32_t i = it.stream, j = it.j;
        K4Stream s = {b.in_base + b.in_off[i], b.in_len[i], b.out_base + b.out_off[i], b.out_cap[i]};
        uint64_t first_byte, seg0, end_word;
        k4_geometry(s.in, s.n, &first_byte, &seg0, &end_word);
        K4Span span;
        span.seg_word = seg0 + (uint64_t)j * K4_SPAN_WORDS;
        span.stop_word = span.seg_word + K4_SPAN_WORDS;
        span.p0 = it.p_start;
        span.o0 = it.o0;
        span.prev_nz = it.prev_nz;
        span.discard = 0;
        K4SpanOut so;
        uint32_t* const rec = (sp.rec && idx < sp.rec_items) ? sp.rec + (size_t)idx * K4_REC_WORDS : nullptr;
        const int32_t st = inflate_uf_run<K4_WRITE>(kt.t, nullptr, *kt.ws, s, b.flags, nullptr, nullptr, &span, &so, rec);
        uint64_t s1 = 0, s2 = 0;
        if (st == ST_OK) {
            s1 = simt::reduce_add(so.ad.s1);
            s2 = simt::reduce_add(so.ad.s2 % ADLER_MOD) % ADLER_MOD;
        }
        uint32_t prior = 0;
        if (lane == 0) {
            it.s1 = s1;
            it.s2 = s2;
            // the walk must end where the count pass said it would
            if (st != ST_OK || so.p_end != it.p_end) sp.failed[i] = 1;
            simt::threadfence();
            prior = simt::atomic_add(&sp.done[i], 1u);
        }
        prior = simt::shfl(prior, 0);
        const uint32_t need = sp.need[i];
        if (prior + 1 != need) continue;
        // every span of the stream is written: checksum, lengths, status
        simt::threadfence();
        const K4Item* its = sp.items + sp.item0[i];
        uint64_t t1 = 0, t2 = 0;
        for (uint32_t k = lane; k < need; k += 32) {
            t1 += ((volatile const K4Item*)its)[k].s1;
            t2 += ((volatile const K4Item*)its)[k].s2;
        }
        t1 = simt::reduce_add(t1);
        t2 = simt::reduce_add(t2) % ADLER_MOD;
        const K4Item& last = its[need - 1];
        const uint64_t total = ((volatile const K4Item&)last).o0 + ((volatile const K4Item&)last).bytes;
        const uint64_t trailer_byte = ((volatile const K4Item&)last).trailer_byte;
        if (lane == 0) {
            const uint8_t* abase = (const uint8_t*)((uintptr_t)s.in & ~(uintptr_t)15);
            const uint8_t* tr = abase + trailer_byte;
            const uint32_t stored = ((uint32_t)simt::ldg8(tr) << 24) | ((uint32_t)simt::ldg8(tr + 1) << 16) |
                                    ((uint32_t)simt::ldg8(tr + 2) << 8) | (uint32_t)simt::ldg8(tr + 3);
            const uint32_t s1m = (uint32_t)(t1 % ADLER_MOD), s2m = (uint32_t)t2, nm = (uint32_t)(total % ADLER_MOD);
            const uint32_t A = (1u + s1m) % ADLER_MOD;
            const uint32_t B = (uint32_t)(((uint64_t)nm + (uint64_t)nm * s1m % ADLER_MOD + ADLER_MOD - s2m) % ADLER_MOD);
            const uint32_t got = (B << 16) | A;
            if (((volatile uint32_t*)sp.failed)[i]) {
                b.status[i] = ST_PENDING_GENERAL;
                b.out_len[i] = 0;
                if (b.consumed) b.consumed[i] = 0;
                worklist[simt::atomic_add(work_count, 1u)] = i;
            } else {
                b.out_len[i] = total;
                if (b.consumed) b.consumed[i] = trailer_byte + 4 - first_byte;
                b.status[i] = (!(b.flags & FLAG_IGNORE_ADLER32) && got != stored) ? ST_WRONG_CHECKSUM : ST_OK;
            }
        }
        simt::syncwarp();
    }
}

}  // namespace fdb
