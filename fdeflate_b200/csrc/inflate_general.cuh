// inflate_general.cuh -- K3: general zlib inflate, ONE WARP PER STREAM.
//
// Replaces (whole-buffer semantics of) the reference's Decompressor::read state machine
//   src/decompress.rs:179-337  (states), :344-438 (block header), :440-555 (code lengths),
//   :561-606 (build_tables), :611-1018 (read_compressed) and src/huffman.rs:18-184 (build_table)
// as driven by decompress_to_vec_bounded (src/decompress.rs:1111-1144).
//
// Design (B200): the warp executes the sequential decode loop converged (all lanes carry the same
// scalar state, table lookups are shared-memory broadcasts) and uses its 32 lanes where the data
// allows it:
//   * input  : 128-byte coalesced chunk loads, one 32-bit word per lane, double-buffered in
//              registers; the 64-bit bit reservoir is refilled with a warp shuffle
//   * tables : 4096-entry litlen table with two-literal entries + 512-entry distance table, rebuilt
//              in shared memory per dynamic block by an index-major warp-cooperative canonical
//              decode (every lane fills 128 entries; no serial codeword walk)
//   * matches: out[p+i] = out[p-dist + i mod dist] for all i in parallel (overlap-safe because
//              the source never leaves the already-final region)
//   * adler32: position-weighted sums over the produced bytes (adler.cuh)
//
// Error semantics follow the reference bit for bit in terms of `avail` = bits left in the stream
// (the reference's nbits gates only bind when its reservoir holds every remaining bit):
// see DESIGN.md "Inflate semantics".
#pragma once
#include <stddef.h>

#include "simt.h"
#include "fdb_common.h"
#include "adler.cuh"

namespace fdb {

// Primary litlen table: 2^K3_TB entries indexed by the next K3_TB stream bits (the reference's has 4096, tables of
// 12 bits; a smaller one keeps more warps on an SM: the table is most of a warp's shared memory).  Codes longer than
// K3_TB bits are decoded canonically.  Which literals share a table entry only shows at a truncated end of input
// (the reference wants the bits of both literals of a pair, decompress.rs:852); the careful loop restates that rule
// for the reference's 12-bit pairing whatever K3_TB is.
#ifndef K3_TABLE_BITS
#define K3_TABLE_BITS 10  // measured on zlib-6 tiles: 12 bits (8 warps / SM) 55 GB/s, 10 bits (14 warps / SM) 90 GB/s
#endif
static const uint32_t K3_TB = K3_TABLE_BITS;
static const uint32_t K3_TMASK = (1u << K3_TB) - 1u;
static_assert(K3_TB >= 9 && K3_TB <= 12, "fixed-code literals need 9 bits; the entry format has 4 bits for a length");

#ifndef K3_PARKED_MATCHES
#define K3_PARKED_MATCHES 640  // matches one segment of the parallel block decode may park (more: the sequential reader takes over)
#endif

struct K3Smem {
    uint32_t litlen[1u << K3_TB];
    uint32_t dist[512];
    uint32_t first[3][16];  // canonical first code per length  [0]=litlen [1]=dist [2]=code-length code
    uint32_t lim[3][16];    // (first+cnt) left-aligned to the table width
    uint16_t off[3][16];    // offset of each length in sorted[]
    uint16_t cnt[3][16];
    uint32_t hist[16];
    uint16_t sorted_lit[288];
    uint16_t sorted_dist[32];
    uint16_t sorted_cl[32];
    uint8_t lens[320];  // litlen lengths at 0..287, distance lengths at 288..319 (reference layout)
    uint8_t cl_lens[32];
    uint8_t cl_table[128];  // (sym << 3) | len
    uint32_t info[4];       // [0] max_len [1] complete flag
    // parallel block decode (decode_block_parallel): transposed per-lane staging rows and the
    // parked matches of one segment in stream order
    uint32_t pstg[17 * 32];
    uint2 pmatch[K3_PARKED_MATCHES];      // {x = destination relative to the segment's first byte, y = len | dist << 16}
};

// ---- bit reader: warp-shuffle reservoir over coalesced chunk loads ---------------------------
struct BitReader {
    const uint8_t* abase;  // 4-byte aligned base address
    uint64_t first_byte;   // first valid byte, relative to abase
    uint64_t end_byte;     // one past the last valid byte, relative to abase
    uint64_t bb;           // reservoir (LSB first); bits past the end of input read as 0
    uint32_t nb;           // bits in reservoir (counting the zero padding)
    uint64_t widx;         // next word to pull into the reservoir
    uint32_t cur, nxt;     // this lane's word of the current / next 32-word chunk
    uint64_t pos;          // bits consumed since the start of the stream
    uint64_t tot;          // 8 * in_len
};

FDB_DEVICE uint32_t br_load_word(const BitReader& r, uint64_t w) {
    uint64_t o = w << 2;
    if (o + 4 <= r.first_byte || o >= r.end_byte) return 0;
    uint32_t v = simt::ldg32((const uint32_t*)(r.abase + o));
    if (o < r.first_byte) v &= 0xffffffffu << (8u * (uint32_t)(r.first_byte - o));
    if (o + 4 > r.end_byte) v &= 0xffffffffu >> (8u * (uint32_t)(o + 4 - r.end_byte));
    return v;
}

FDB_DEVICE void br_pull(BitReader& r) {  // append one 32-bit word to the reservoir
    uint32_t w = simt::shfl(r.cur, (unsigned)(r.widx & 31));
    r.bb |= (uint64_t)w << r.nb;
    r.nb += 32;
    r.widx++;
    if ((r.widx & 31) == 0) {
        r.cur = r.nxt;
        r.nxt = br_load_word(r, ((r.widx >> 5) + 1) * 32 + simt::lane_id());
    }
}

FDB_DEVICE void br_seek(BitReader& r, uint64_t bitpos) {
    uint64_t abit = r.first_byte * 8 + bitpos;
    r.widx = abit >> 5;
    uint64_t c = r.widx >> 5;
    r.cur = br_load_word(r, c * 32 + simt::lane_id());
    r.nxt = br_load_word(r, (c + 1) * 32 + simt::lane_id());
    r.bb = 0;
    r.nb = 0;
    br_pull(r);
    uint32_t sh = (uint32_t)(abit & 31);
    r.bb >>= sh;
    r.nb -= sh;
    r.pos = bitpos;
    if (r.nb <= 32) br_pull(r);
}

FDB_DEVICE void br_init(BitReader& r, const uint8_t* in, uint64_t n) {
    uintptr_t a = (uintptr_t)in;
    r.abase = (const uint8_t*)(a & ~(uintptr_t)3);
    r.first_byte = (uint64_t)(a & 3);
    r.end_byte = r.first_byte + n;
    r.tot = n * 8;
    br_seek(r, 0);
}

FDB_DEVICE void br_refill(BitReader& r) {  // afterwards nb >= 33
    while (r.nb <= 32) br_pull(r);
}
FDB_DEVICE uint32_t br_peek(const BitReader& r, uint32_t k) {  // k <= 32
    return (uint32_t)(r.bb & ((1ull << k) - 1ull));
}
FDB_DEVICE void br_consume(BitReader& r, uint32_t k) {
    r.bb >>= k;
    r.nb -= k;
    r.pos += k;
}
FDB_DEVICE uint64_t br_avail(const BitReader& r) { return r.tot - r.pos; }

// ---- warp-cooperative canonical Huffman set-up (replaces huffman.rs:28-84) --------------------
// lens[0..nsym) in shared memory.  Fills first/lim/off/cnt[which] and sorted[]; returns through
// info[0] = max_len (>=1), info[1] = 1 if the code is complete (Kraft sum == 2^max_len).
FDB_DEVICE void canon_setup(K3Smem& s, int which, const uint8_t* lens, uint32_t nsym, uint16_t* sorted,
                            uint32_t table_bits) {
    const unsigned lane = simt::lane_id();
    if (lane < 16) s.hist[lane] = 0;
    simt::syncwarp();
    for (uint32_t i = lane; i < nsym; i += 32) simt::atomic_add(&s.hist[lens[i]], 1u);
    simt::syncwarp();
    if (lane == 0) {
        uint32_t max_len = 15;
        while (max_len > 1 && s.hist[max_len] == 0) max_len--;
        uint32_t used = 0, code = 0, o = 0;
        for (uint32_t L = 1; L <= 15; L++) {
            uint32_t c = s.hist[L];
            if (L <= max_len) used = (used << 1) + c;
            code <<= 1;
            s.first[which][L] = code;
            s.cnt[which][L] = (uint16_t)c;
            s.off[which][L] = (uint16_t)o;
            code += c;
            o += c;
            // exclusive upper bound of length-L codes, left-aligned to table_bits (only L<=table_bits used)
            s.lim[which][L] = (L <= table_bits) ? (code << (table_bits - L)) : 0u;
        }
        s.info[0] = max_len;
        s.info[1] = (used == (1u << max_len)) ? 1u : 0u;
    }
    simt::syncwarp();
    // sorted[]: symbols ordered by (length, symbol) -- ballot compaction per length
    uint32_t max_len = s.info[0];
    for (uint32_t L = 1; L <= max_len; L++) {
        if (s.cnt[which][L] == 0) continue;
        uint32_t base = s.off[which][L];
        for (uint32_t i0 = 0; i0 < nsym; i0 += 32) {
            uint32_t i = i0 + lane;
            bool p = (i < nsym) && (lens[i] == L);
            uint32_t m = simt::ballot(p);
            if (p) sorted[base + simt::popc(m & simt::lanemask_lt())] = (uint16_t)i;
            base += simt::popc(m);
        }
    }
    simt::syncwarp();
}

// canonical decode of a `maxbits`-bit window for codes longer than the primary table
// (replaces the secondary tables of huffman.rs:139-181).  v = next bits of the stream, LSB first.
FDB_DEVICE bool canon_long_decode(const K3Smem& s, int which, const uint16_t* sorted, uint32_t v,
                                  uint32_t from_len, uint32_t* sym, uint32_t* nbits) {
    uint32_t msb = simt::brev(v) >> 17;  // first 15 stream bits, MSB first
    for (uint32_t L = from_len; L <= 15; L++) {
        uint32_t c = msb >> (15 - L);
        uint32_t d = c - s.first[which][L];
        if (d < s.cnt[which][L]) {
            *sym = sorted[s.off[which][L] + d];
            *nbits = L;
            return true;
        }
    }
    return false;
}

// primary litlen table, index-major: every lane decodes 128 of the 4096 indices canonically, then a
// second pass upgrades single literals to two-literal entries (replaces huffman.rs:92-136).
FDB_DEVICE void build_litlen_table(K3Smem& s) {
    const unsigned lane = simt::lane_id();
    for (uint32_t idx = lane; idx < (1u << K3_TB); idx += 32) {
        uint32_t v = simt::brev(idx) >> (32 - K3_TB);
        uint32_t e = 0;  // none of LIT/LEN/EOB => long code
        for (uint32_t L = 1; L <= K3_TB; L++) {
            if (v < s.lim[0][L]) {
                uint32_t c = v >> (K3_TB - L);
                e = make_litlen_entry(s.sorted_lit[s.off[0][L] + c - s.first[0][L]], L);
                break;
            }
        }
        s.litlen[idx] = e;
    }
    simt::syncwarp();
    // second pass, 128 indices at a time: read (the entry and the one its second literal would come
    // from), barrier, write -- so no lane reads an entry another lane is upgrading at that moment.
    // An entry read here may already be a pair from an earlier round: its first literal and that
    // literal's bit count are the same in both forms.
    for (uint32_t base = 0; base < (1u << K3_TB); base += 128) {
        uint32_t ne[4];
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t idx = base + 32 * k + lane;
            const uint32_t e = s.litlen[idx];
            ne[k] = e;
            if (!(e & LL_LIT) || (e & LL_LIT2)) continue;
            const uint32_t l1 = e & 15u;
            if (l1 >= K3_TB) continue;
            const uint32_t e2 = s.litlen[idx >> l1];
            if (!(e2 & LL_LIT)) continue;
            const uint32_t l2 = (e2 >> 24) & 15u;
            if (l1 + l2 > K3_TB) continue;
            ne[k] = make_litlen_pair(e, (e2 >> 8) & 0xffu, l1, l2);
        }
        simt::syncwarp();
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) s.litlen[base + 32 * k + lane] = ne[k];
        simt::syncwarp();
    }
}

FDB_DEVICE void build_dist_table(K3Smem& s) {
    const unsigned lane = simt::lane_id();
    for (uint32_t idx = lane; idx < 512; idx += 32) {
        uint32_t v = simt::brev(idx) >> 23;
        uint32_t e = DS_LONG;
        for (uint32_t L = 1; L <= 9; L++) {
            if (v < s.lim[1][L]) {
                uint32_t c = v >> (9 - L);
                e = make_dist_entry(s.sorted_dist[s.off[1][L] + c - s.first[1][L]], L);
                break;
            }
        }
        s.dist[idx] = e;
    }
    simt::syncwarp();
}

// litlen + distance tables from s.lens (reference decompress.rs:561-606). Returns a Status.
FDB_DEVICE int32_t build_block_tables(K3Smem& s, uint32_t hlit, uint32_t* eof_code, uint32_t* eof_bits) {
    const unsigned lane = simt::lane_id();
    if (s.lens[256] == 0) return ST_BAD_LITERAL_LENGTH_HUFFMAN_TREE;  // :563-566
    canon_setup(s, 0, s.lens, hlit, s.sorted_lit, K3_TB);
    if (!s.info[1]) return ST_BAD_CODE_LENGTH_HUFFMAN_TREE;  // :570-580 (sic: the reference's variant)
    build_litlen_table(s);
    {  // eof code = bit-reversed canonical code of symbol 256 (:582-584)
        uint32_t L = s.lens[256];
        uint32_t o = s.off[0][L], c = s.cnt[0][L];
        uint32_t found = 0;
        for (uint32_t i0 = 0; i0 < c; i0 += 32) {
            uint32_t i = i0 + lane;
            uint32_t m = simt::ballot(i < c && s.sorted_lit[o + i] == 256);
            if (m) found = i0 + simt::ffs(m) - 1;
        }
        uint32_t code = s.first[0][L] + found;
        *eof_code = simt::brev(code) >> (32 - L);
        *eof_bits = L;
    }
    // distance code (:587-603, huffman.rs:40-59)
    uint32_t nz = 0;
    {
        uint32_t m = simt::ballot(s.lens[288 + lane] != 0);
        nz = m;
    }
    if (nz == 0) {
        for (uint32_t idx = lane; idx < 512; idx += 32) s.dist[idx] = 0;  // any match => InvalidDistanceCode
        simt::syncwarp();
        return ST_OK;
    }
    canon_setup(s, 1, s.lens + 288, 32, s.sorted_dist, 9);
    if (s.info[0] == 1 && s.cnt[1][1] == 1) {  // exactly one 1-bit code
        uint32_t sym = s.sorted_dist[s.off[1][1]];
        uint32_t e = make_dist_entry(sym, 1);
        for (uint32_t idx = lane; idx < 512; idx += 32) s.dist[idx] = (idx & 1) ? 0u : e;
        simt::syncwarp();
        return ST_OK;
    }
    if (!s.info[1]) return ST_BAD_DISTANCE_HUFFMAN_TREE;
    build_dist_table(s);
    return ST_OK;
}

FDB_DEVICE void load_fixed_lengths(K3Smem& s) {  // RFC 1951 3.2.6 (reference tables.rs:207-232)
    const unsigned lane = simt::lane_id();
    for (uint32_t i = lane; i < 320; i += 32) {
        uint8_t L = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : i < 288 ? 8 : 5;
        s.lens[i] = L;
    }
    simt::syncwarp();
}

// Warp-wide copy of n bytes between two byte addresses of any alignment (the payload of a stored block): head bytes up
// to the first 16-byte boundary of the destination, then 16-byte vectors -- two aligned source vectors funnel-shifted
// into one aligned store, four vectors per lane in flight --, then the tail bytes.  src must be readable up to the next
// 16-byte boundary past src + n (the contract of every input buffer, include/fdeflate_b200.h).
FDB_DEVICE void warp_copy_bytes(uint8_t* dst, const uint8_t* src, uint64_t n) {
    const unsigned lane = simt::lane_id();
    uint64_t head = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;
    if (head > n) head = n;
    if (lane < head) dst[lane] = simt::ldg8(src + lane);
    const uint64_t nvec = (n - head) >> 4;
    const uint8_t* s0 = src + head;
    const uint32_t sh = (uint32_t)((uintptr_t)s0 & 15u);
    const uint4* a0 = (const uint4*)(s0 - sh);
    uint4* d0 = (uint4*)(dst + head);
    const uint32_t s4 = sh >> 2, sb = 8u * (sh & 3u);
    auto shifted = [&](const uint4& lo, const uint4& hi) -> uint4 {
        const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        uint4 r;
        r.x = simt::funnel_r(w[s4], w[s4 + 1], sb);
        r.y = simt::funnel_r(w[s4 + 1], w[s4 + 2], sb);
        r.z = simt::funnel_r(w[s4 + 2], w[s4 + 3], sb);
        r.w = simt::funnel_r(w[s4 + 3], w[s4 + 4], sb);
        return r;
    };
    uint64_t v = lane;
    if (sh == 0) {
        for (; v + 96 < nvec; v += 128) {
            const uint4 q0 = simt::ldg128(a0 + v), q1 = simt::ldg128(a0 + v + 32), q2 = simt::ldg128(a0 + v + 64), q3 = simt::ldg128(a0 + v + 96);
            d0[v] = q0;
            d0[v + 32] = q1;
            d0[v + 64] = q2;
            d0[v + 96] = q3;
        }
        for (; v < nvec; v += 32) d0[v] = simt::ldg128(a0 + v);
    } else {
        for (; v + 96 < nvec; v += 128) {
            const uint4 l0 = simt::ldg128(a0 + v), h0 = simt::ldg128(a0 + v + 1);
            const uint4 l1 = simt::ldg128(a0 + v + 32), h1 = simt::ldg128(a0 + v + 33);
            const uint4 l2 = simt::ldg128(a0 + v + 64), h2 = simt::ldg128(a0 + v + 65);
            const uint4 l3 = simt::ldg128(a0 + v + 96), h3 = simt::ldg128(a0 + v + 97);
            d0[v] = shifted(l0, h0);
            d0[v + 32] = shifted(l1, h1);
            d0[v + 64] = shifted(l2, h2);
            d0[v + 96] = shifted(l3, h3);
        }
        for (; v < nvec; v += 32) {
            const uint4 l0 = simt::ldg128(a0 + v), h0 = simt::ldg128(a0 + v + 1);
            d0[v] = shifted(l0, h0);
        }
    }
    const uint64_t tail0 = head + (nvec << 4);
    if (tail0 + lane < n) dst[tail0 + lane] = simt::ldg8(src + tail0 + lane);
    simt::syncwarp();
}

// ---- one compressed block (reference decompress.rs:611-1018, careful-loop semantics) ----------
struct OutCursor {
    uint8_t* out;
    uint64_t pos;
    uint64_t cap;
    uint64_t lo;  // first position that holds output (0 except for a streaming decoder that has dropped old history):
                  // a distance reaches back to `lo` at most
};

// What a streaming decoder carries from one fdb_stream_read_batch call to the next besides its tables
// (decode_block only; the block-level state is K3StreamState below).
struct K3Resume {
    uint32_t pend_len, pend_dist;  // the part of a match that did not fit the caller's room (reference QueuedOutput, :1066-1070)
};

// Matches are not executed one by one: the warp keeps decoding and parks up to 32 of them, one per
// lane, then executes the whole batch at once.  A match whose source ends before the first parked
// destination cannot depend on any match of the batch ("early": every lane copies its own, all loads
// in flight together -- the global round trip is paid once per batch instead of once per match);
// the others ("late": overlapping or near sources, e.g. distance-1 runs) follow in stream order with
// the warp-wide copy.  Literals are stored directly; they never alias a parked destination.
struct MatchQueue {
    uint64_t dst;   // this lane's parked match: output position,
    uint32_t len;   //   bytes to produce (already clipped to the slot),
    uint32_t dist;  //   distance
    uint32_t qn;    // matches parked (warp-uniform)
};

FDB_DEVICE void mq_flush(uint8_t* out, MatchQueue& q) {
    const unsigned lane = simt::lane_id();
    if (q.qn == 0) return;
    simt::syncwarp();  // literal stores by lane 0/1 must be visible to the loads below
    const bool active = lane < q.qn;
    const uint64_t first_dst = simt::shfl(q.dst, 0);
    const bool early = active && (q.dst - q.dist + q.len <= first_dst);
    if (early) {
        uint8_t* d = out + q.dst;
        const uint8_t* s = d - q.dist;
        uint32_t i = 0;
        for (; i + 8 <= q.len; i += 8) {  // eight independent loads, then eight stores
            uint8_t b0 = s[i], b1 = s[i + 1], b2 = s[i + 2], b3 = s[i + 3];
            uint8_t b4 = s[i + 4], b5 = s[i + 5], b6 = s[i + 6], b7 = s[i + 7];
            d[i] = b0; d[i + 1] = b1; d[i + 2] = b2; d[i + 3] = b3;
            d[i + 4] = b4; d[i + 5] = b5; d[i + 6] = b6; d[i + 7] = b7;
        }
        uint8_t t[8];
#pragma unroll
        for (uint32_t j = 0; j < 8; j++) t[j] = (i + j < q.len) ? s[i + j] : (uint8_t)0;
#pragma unroll
        for (uint32_t j = 0; j < 8; j++)
            if (i + j < q.len) d[i + j] = t[j];
    }
    uint32_t late = simt::ballot(active && !early);
    while (late) {
        const uint32_t k = simt::ffs(late) - 1;
        late &= late - 1;
        const uint64_t kd = simt::shfl(q.dst, k);
        const uint32_t n = simt::shfl(q.len, k), d32 = simt::shfl(q.dist, k);
        simt::syncwarp();  // everything before this match in stream order is complete
        uint8_t* d = out + kd;
        const uint8_t* s = d - d32;
        if (d32 >= n) {
            for (uint32_t i = lane; i < n; i += 32) d[i] = s[i];
        } else {
            for (uint32_t i = lane; i < n; i += 32) d[i] = s[i % d32];
        }
    }
    simt::syncwarp();
    q.qn = 0;
}

// ---- parallel block decode ---------------------------------------------------------------------
// The scheme of K4 (inflate_uf.cuh) with the block's OWN tables: per segment of 32 x 8 words every lane
// decodes a different sub-sequence of the block.  Lanes warm up on the 128 bits before their
// sub-sequence (Huffman codes self-synchronise), count the bytes and matches of their tokens, the chain
// "my start == my predecessor's end" is verified from lane 0 (exact) upwards, scans give every token
// its output position, then lanes decode again: literals go straight to their final place, matches
// are parked in stream order and executed 32 at a time (mq-style early / late split).
// It only ever COMMITS regular segments.  On anything else -- a code or distance it cannot decode,
// a distance too far back, the slot or the input running out, too many matches -- it stops at the last
// committed token boundary and the sequential reader (exact reference semantics) carries on from there.
static const uint32_t P_SUBW = 8, P_WARM = 4, P_TAILW = 4;
static const uint32_t P_ROWW = P_WARM + P_SUBW + P_TAILW;        // 16 words seen by one lane (+1 look-ahead)
static const uint32_t P_SEG_WORDS = P_WARM + 32 * P_SUBW + 4;    // staged per segment (vectors of 4)
static const uint32_t P_LIM_LO = 32u * P_WARM, P_LIM_HI = 32u * (P_WARM + P_SUBW);
static const uint32_t P_MAXM = K3_PARKED_MATCHES;
static const uint32_t P_INVALID = 0xffffffffu;

struct GLane {  // lane-private LSB-first reader over a transposed staging row (cf. LaneBits in inflate_uf.cuh)
    uint32_t w0, w1, w2, rp;
    simt::saddr nx;
};
FDB_DEVICE void gl_start(GLane& b, simt::saddr row, uint32_t rp) {
    b.rp = rp;
    const simt::saddr p = row + (rp >> 5) * 128u;
    b.w0 = simt::lds32(p);
    b.w1 = simt::lds32(p + 128u);
    b.w2 = simt::lds32(p + 256u);
    b.nx = p + 384u;
}
FDB_DEVICE void gl_advance(GLane& b, uint32_t n) {  // n <= 48
#if !defined(FDB_EMUL)
    // the two conditional word shifts as predicated instructions (cf. lb_advance in inflate_uf.cuh: the compiler's
    // version moves the three words through temporaries)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 x, y;\n\t"
        "setp.ge.u32 p, %5, 32;\n\t"
        "@p mov.u32 %0, %1;\n\t@p mov.u32 %1, %2;\n\t@p ld.shared.u32 %2, [%3];\n\t@p add.u32 %3, %3, 128;\n\t"
        "@p add.u32 %4, %4, 32;\n\t@p sub.u32 %5, %5, 32;\n\t"
        "add.u32 y, %4, %5;\n\txor.b32 x, y, %4;\n\tand.b32 x, x, 32;\n\tsetp.ne.u32 p, x, 0;\n\t"
        "@p mov.u32 %0, %1;\n\t@p mov.u32 %1, %2;\n\t@p ld.shared.u32 %2, [%3];\n\t@p add.u32 %3, %3, 128;\n\t"
        "mov.u32 %4, y;\n\t}"
        : "+r"(b.w0), "+r"(b.w1), "+r"(b.w2), "+r"(b.nx), "+r"(b.rp), "+r"(n)
        :
        : "memory");
    return;
#endif
    if (n >= 32) {
        b.w0 = b.w1;
        b.w1 = b.w2;
        b.w2 = simt::lds32(b.nx);
        b.nx += 128u;
        b.rp += 32;
        n -= 32;
    }
    const uint32_t nrp = b.rp + n;
    if ((nrp ^ b.rp) & 32u) {
        b.w0 = b.w1;
        b.w1 = b.w2;
        b.w2 = simt::lds32(b.nx);
        b.nx += 128u;
    }
    b.rp = nrp;
}

// shared-window addresses of the block's tables
struct GTabs {
    simt::saddr litlen, dist;
};

enum : uint32_t { GT_LIT = 0, GT_MATCH = 1, GT_EOB = 2, GT_BAD = 3 };
struct GTok {
    uint32_t kind;
    uint32_t nbits;  // bits of the token (both literals of a pair)
    uint32_t bytes;  // bytes it produces
    uint32_t lit;    // literal(s), first in the low byte
    uint32_t dist;
};
// One table entry at the reader's position -- a token, or a pair of literals.  Pairs are NOT split at
// sub-sequence boundaries: lanes hand over on ENTRY boundaries, so the verified chain is exactly the
// sequence of entries the sequential decoder would take from the same start.  (Which literal opens a
// pair matters at a truncated end of input: the reference wants the bits of both literals of an entry,
// decompress.rs:852, so status and output there depend on the pairing.)
FDB_DEVICE void g_token(const K3Smem& s, const GTabs& tb, const GLane& b, GTok& t) {
    const uint32_t bits = simt::funnel_r(b.w0, b.w1, b.rp);
    const uint32_t e = simt::lds32_ro(tb.litlen + ((bits & K3_TMASK) << 2));
    uint32_t nbits = e & 15u;
    t.dist = 0;
    if (e & LL_LIT) {
        t.kind = GT_LIT;
        t.nbits = nbits;
        t.bytes = e >> 28;  // 1 or 2
        t.lit = (e >> 8) & 0xffffu;  // (the second byte of a single is 0)
        return;
    }
    uint32_t len_base, len_extra;
    if (e & LL_EOB) {  // (includes the fixed-code 286/287 quirk)
        t.kind = GT_EOB;
        t.nbits = nbits;
        t.bytes = 0;
        t.lit = 0;
        return;
    } else if (e & LL_LEN) {
        len_base = (e >> 16) & 0x1ffu;
        len_extra = (e >> 8) & 7u;
    } else {  // code longer than the table
        uint32_t sym = 0;
        if (!canon_long_decode(s, 0, s.sorted_lit, bits & 0x7fffu, K3_TB + 1, &sym, &nbits)) {
            t.kind = GT_BAD;
            t.nbits = 0;
            t.bytes = 0;
            t.lit = 0;
            return;
        }
        if (sym <= 256) {
            t.kind = sym < 256 ? GT_LIT : GT_EOB;
            t.nbits = nbits;
            t.bytes = sym < 256 ? 1u : 0u;
            t.lit = sym & 0xffu;
            return;
        }
        len_base = len_sym_base(sym);
        len_extra = len_sym_extra(sym);
    }
    const uint32_t n1 = nbits + len_extra;  // <= 20
    const uint32_t length = len_base + ((bits >> nbits) & ((1u << len_extra) - 1u));
    const uint32_t hi = simt::funnel_r(b.w1, b.w2, b.rp);
    const uint32_t dw = simt::funnel_r(bits, hi, n1);  // the 32 bits behind the length
    const uint32_t de = simt::lds32_ro(tb.dist + ((dw & 0x1ffu) << 2));
    uint32_t dbits, dextra, dbase;
    if (de & DS_VALID) {
        dbits = de & 15u;
        dextra = (de >> 4) & 15u;
        dbase = de >> 16;
    } else {
        uint32_t dsym = 0;
        if (!(de & DS_LONG) || !canon_long_decode(s, 1, s.sorted_dist, dw & 0x7fffu, 10, &dsym, &dbits) ||
            dsym >= 30) {
            t.kind = GT_BAD;
            t.nbits = 0;
            t.bytes = 0;
            t.lit = 0;
            return;
        }
        dextra = dist_sym_extra(dsym);
        dbase = dist_sym_base(dsym);
    }
    t.kind = GT_MATCH;
    t.nbits = n1 + dbits + dextra;  // <= 48
    t.bytes = length;
    t.lit = 0;
    t.dist = dbase + ((dw >> dbits) & ((1u << dextra) - 1u));
}

struct GCount {
    uint32_t end;    // first token boundary >= LIM_HI (row-relative), or the position of the EOB code
    uint32_t cnt;    // bytes produced by the tokens in [start, end)
    uint32_t nm;     // matches among them
    uint32_t flags;  // GF_*
    uint32_t eobn;   // bits of the EOB code (with GF_EOB)
};
enum : uint32_t { GF_EOB = 1, GF_BAD = 2 };

FDB_DEVICE GCount g_count(const K3Smem& s, const GTabs& tb, simt::saddr row, uint32_t start, bool active) {
    GCount c = {P_INVALID, 0, 0, 0, 0};
    GLane b;
    gl_start(b, row, active ? start : 0u);
    uint32_t cnt = 0, nm = 0, flags = 0, eobn = 0;
    bool stop = !active;
    while (!stop && b.rp < P_LIM_HI) {
        GTok t;
        g_token(s, tb, b, t);
        if (t.kind >= GT_EOB) {
            flags |= t.kind == GT_EOB ? GF_EOB : GF_BAD;
            eobn = t.nbits;
            stop = true;
        } else {
            cnt += t.bytes;
            nm += t.kind;  // GT_MATCH == 1
            gl_advance(b, t.nbits);
        }
    }
    if (active) {
        c.end = b.rp;
        c.cnt = cnt;
        c.nm = nm;
        c.flags = flags;
        c.eobn = eobn;
    }
    return c;
}

FDB_DEVICE uint32_t g_warm_up(const K3Smem& s, const GTabs& tb, simt::saddr row, bool active) {
    GLane b;
    gl_start(b, row, 0u);
    bool stop = !active, dead = false;
    while (!stop && b.rp < P_LIM_LO) {
        GTok t;
        g_token(s, tb, b, t);
        if (t.kind >= GT_EOB) {  // speculative end of block / undecodable: this lane has no valid guess
            dead = true;
            stop = true;
        } else {
            gl_advance(b, t.nbits);
        }
    }
    return (active && !dead) ? b.rp : P_INVALID;
}

// Returns true when the block ended here (EOB consumed, r positioned behind it); false when the
// sequential decoder has to continue from (r, o), which then sit on the last committed token boundary.
FDB_DEVICE bool decode_block_parallel(K3Smem& s, BitReader& r, OutCursor& o) {
    const unsigned lane = simt::lane_id();
    uint32_t* stg = s.pstg;
    const simt::saddr row = simt::smem_addr(s.pstg + lane);
    const GTabs tb = {simt::smem_addr(s.litlen), simt::smem_addr(s.dist)};
    // virtual bit / byte coordinates relative to a 16-byte aligned base at or below the stream
    // (the base lies 64 bytes further down so that the warm-up words of the first segment have
    // non-negative indices; nothing below the stream's first byte is ever loaded)
    const uint8_t* in = r.abase + r.first_byte;
    const uint8_t* abase = (const uint8_t*)((uintptr_t)in & ~(uintptr_t)15) - 64;
    const uint64_t first_byte = (uint64_t)((uintptr_t)in & 15u) + 64;
    const uint64_t end_byte = first_byte + (r.tot >> 3);
    uint64_t p0 = first_byte * 8 + r.pos;  // true bit position of the next token
    bool ended = false;

    for (;;) {
        // enough input for a whole segment plus look-ahead, and room for the typical output?
        const uint64_t seg_word = ((p0 >> 5) >> 2) << 2;
        const uint64_t s0 = seg_word - P_WARM;
        if (((s0 + P_SEG_WORDS) << 2) + 8 > end_byte) break;  // near the end of the input: sequential
        if (o.cap - o.pos < 4096) break;

        // ---- stage (as in K4: lane l takes vectors 2l, 2l+1, then 64+l; every store hits its own bank) ----
        simt::syncwarp();
        for (uint32_t it = 0; it < 3; it++) {
            const uint32_t v = it < 2 ? 2u * lane + it : 64u + lane;
            if (v >= P_SEG_WORDS / 4) continue;
            const uint64_t byte0 = (s0 << 2) + 16ull * v;
            uint4 q = make_uint4(0, 0, 0, 0);
            if (byte0 + 16 > first_byte) {
                q = simt::ldg128((const uint4*)(abase + byte0));
                if (byte0 < first_byte) {  // bytes in front of the stream read as 0
                    uint32_t w[4] = {q.x, q.y, q.z, q.w};
                    for (uint32_t j = 0; j < 16; j++)
                        if (byte0 + j < first_byte) w[j >> 2] &= ~(0xffu << (8u * (j & 3u)));
                    q = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
            const uint32_t r1 = v >> 1, c1 = (v & 1) * 4;
            const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (uint32_t j = 0; j < 4; j++) {
                if (r1 < 32) stg[(c1 + j) * 32 + r1] = w4[j];
                if (r1 >= 1 && r1 <= 32) stg[(c1 + j + 8) * 32 + (r1 - 1)] = w4[j];
            }
        }
        simt::syncwarp();

        // ---- warm up, count ----
        uint32_t start = g_warm_up(s, tb, row, lane != 0);
        if (lane == 0) start = (uint32_t)(p0 - (s0 << 5));
        GCount c = g_count(s, tb, row, start, start != P_INVALID);

        // ---- verify the chain (rows are 32 * SUBW bits apart) ----
        uint32_t eob_lane = 32;
        for (;;) {
            const uint32_t prev_end = simt::shfl_up(c.end, 1);
            const uint32_t prev_flags = simt::shfl_up(c.flags, 1);
            const uint32_t want = prev_end - 32u * P_SUBW;
            const bool mismatch = (lane > 0) && (prev_end == P_INVALID || start != want || (prev_flags & (GF_EOB | GF_BAD)));
            const uint32_t mm = simt::ballot(mismatch);
            const uint32_t em = simt::ballot((c.flags & (GF_EOB | GF_BAD)) != 0 && start != P_INVALID);
            const uint32_t first_mis = mm ? simt::ffs(mm) - 1 : 32;
            const uint32_t first_eob = em ? simt::ffs(em) - 1 : 32;
            if (first_eob < first_mis) {  // the block ends (or turns undecodable) inside a verified lane
                eob_lane = first_eob;
                break;
            }
            if (first_mis == 32) break;
            const bool redo = mismatch && !(prev_flags & (GF_EOB | GF_BAD)) && prev_end != P_INVALID;
            if (mismatch) start = redo ? want : P_INVALID;
            GCount c2 = g_count(s, tb, row, start, redo);
            if (mismatch) c = c2;
        }
        if (lane > eob_lane) {
            c.cnt = 0;
            c.nm = 0;
            c.flags = 0;
        }
        // an undecodable token in a verified lane: commit nothing of this segment
        if (simt::any((c.flags & GF_BAD) != 0)) break;

        // ---- scan ----
        const uint32_t incl = simt::scan_incl_add(c.cnt);
        const uint32_t seg_bytes = simt::shfl(incl, 31);
        const uint32_t mincl = simt::scan_incl_add(c.nm);
        const uint32_t seg_matches = simt::shfl(mincl, 31);
        if ((uint64_t)seg_bytes + 2 > o.cap - o.pos || seg_matches > P_MAXM) break;
        const uint64_t o0 = o.pos;

        // ---- write: literals in place, matches parked in stream order ----
        {
            GLane b;
            gl_start(b, row, start != P_INVALID ? start : 0u);
            uint32_t rel = incl - c.cnt;     // my first byte, relative to o0
            uint32_t mi = mincl - c.nm;      // my first match
            bool fin = start == P_INVALID || lane > eob_lane;
            bool bad = false;
            while (!fin && b.rp < P_LIM_HI) {
                GTok t;
                g_token(s, tb, b, t);
                if (t.kind == GT_LIT) {
                    o.out[o0 + rel] = (uint8_t)t.lit;
                    if (t.bytes == 2) o.out[o0 + rel + 1] = (uint8_t)(t.lit >> 8);
                } else if (t.kind == GT_MATCH) {
                    if ((uint64_t)t.dist > o0 + rel - o.lo) bad = true;  // DistanceTooFarBack: the sequential decoder reports it
                    s.pmatch[mi] = make_uint2(rel, t.bytes | (t.dist << 16));
                    mi++;
                } else {
                    fin = true;
                }
                rel += t.bytes;
                if (!fin) gl_advance(b, t.nbits);
            }
            if (simt::any(bad)) break;
        }
        simt::syncwarp();

        // ---- execute the parked matches ----
        // One lane copies one match: `len` independent byte loads, then the stores.
        auto copy_own = [&](uint64_t dst, uint32_t mlen, uint32_t mdist) {
            uint8_t* d = o.out + dst;
            const uint8_t* sp = d - mdist;
            uint32_t i = 0;
            for (; i + 8 <= mlen; i += 8) {
                uint8_t b0 = sp[i], b1 = sp[i + 1], b2 = sp[i + 2], b3 = sp[i + 3];
                uint8_t b4 = sp[i + 4], b5 = sp[i + 5], b6 = sp[i + 6], b7 = sp[i + 7];
                d[i] = b0; d[i + 1] = b1; d[i + 2] = b2; d[i + 3] = b3;
                d[i + 4] = b4; d[i + 5] = b5; d[i + 6] = b6; d[i + 7] = b7;
            }
            uint8_t tb8[8];
#pragma unroll
            for (uint32_t j = 0; j < 8; j++) tb8[j] = (i + j < mlen) ? sp[i + j] : (uint8_t)0;
#pragma unroll
            for (uint32_t j = 0; j < 8; j++)
                if (i + j < mlen) d[i + j] = tb8[j];
        };
        // (A segment-wide pass for matches whose source ends before the segment's first destination was
        // tried and measured slower: 47 vs 54 GB/s on level-6 tiles.)
        // In stream order, 32 at a time: "early" = source ends before the first destination of the batch,
        // executed lane-parallel; "late" = overlapping / near sources, one after another with the
        // warp-wide copy.
        for (uint32_t m0 = 0; m0 < seg_matches; m0 += 32) {
            const uint32_t k = m0 + lane;
            const uint2 rec = k < seg_matches ? s.pmatch[k] : make_uint2(0, 0);
            const uint32_t mlen = rec.y & 0xffffu, mdist = rec.y >> 16;
            const bool active = mlen != 0;
            const uint32_t am = simt::ballot(active);
            if (am == 0) continue;
            const uint64_t dst = o0 + rec.x;
            const uint64_t first_dst = o0 + simt::shfl(rec.x, simt::ffs(am) - 1);
            simt::syncwarp();  // literals and earlier batches are complete
            const bool early = active && (dst - mdist + mlen <= first_dst);
            if (early) copy_own(dst, mlen, mdist);
            uint32_t late = simt::ballot(active && !early);
            while (late) {
                const uint32_t kk = simt::ffs(late) - 1;
                late &= late - 1;
                const uint64_t kd = simt::shfl(dst, kk);
                const uint32_t n = simt::shfl(mlen, kk), d32 = simt::shfl(mdist, kk);
                simt::syncwarp();
                uint8_t* d = o.out + kd;
                const uint8_t* sp = d - d32;
                if (d32 >= n) {
                    for (uint32_t i = lane; i < n; i += 32) d[i] = sp[i];
                } else {
                    for (uint32_t i = lane; i < n; i += 32) d[i] = sp[i % d32];
                }
            }
        }
        simt::syncwarp();

        // ---- commit ----
        o.pos = o0 + seg_bytes;
        if (eob_lane < 32) {
            const uint32_t eob_rel = simt::shfl(c.end, eob_lane), eob_n = simt::shfl(c.eobn, eob_lane);
            p0 = ((s0 + (uint64_t)P_SUBW * eob_lane) << 5) + eob_rel + eob_n;  // behind the EOB code
            ended = true;
            break;
        }
        p0 = ((s0 + (uint64_t)P_SUBW * 31) << 5) + simt::shfl(c.end, 31);
    }
    br_seek(r, p0 - first_byte * 8);
    return ended;
}

// Fast path of decode_block: the same tokens through a lean, warp-uniform reader (three words in
// registers, the next one fetched a word ahead with a shuffle), taken only while the input has at
// least 64 more bits and the slot room for a longest match plus a literal pair.  It never reports
// anything: at the end of the block's reach, an end-of-block code, a code longer than the tables, an
// invalid code or a distance too far back it stops AT the token boundary and the careful loop below
// (the reference's careful loop, decompress.rs:836-1007) takes over from there.
FDB_DEVICE void decode_block_fast(K3Smem& s, BitReader& r, OutCursor& o, MatchQueue& mq) {
    const unsigned lane = simt::lane_id();
    const uint64_t avail = br_avail(r);
    if (avail < 128 || o.cap - o.pos < 264) return;
    const uint64_t budget = avail - 64;
    const uint32_t max_used = budget > 0x7fffff00ull ? 0x7fffff00u : (uint32_t)budget;
    const uint64_t out_limit = o.cap - 262;  // a token may add up to 258 bytes

    // ---- reader over the 128-byte input chunks (one word per lane) ----
    const uint64_t abit = r.first_byte * 8 + r.pos;
    uint64_t fetch = abit >> 5;  // next word to pull
    uint64_t c = fetch >> 5;     // chunk held in `cur`
    uint32_t cur = br_load_word(r, c * 32 + lane), nxt = br_load_word(r, (c + 1) * 32 + lane);
    auto pull = [&]() -> uint32_t {
        const uint32_t w = simt::shfl(cur, (unsigned)(fetch & 31));
        fetch++;
        if ((fetch & 31) == 0) {
            cur = nxt;
            c++;
            nxt = br_load_word(r, (c + 1) * 32 + lane);
        }
        return w;
    };
    uint32_t w0 = pull(), w1 = pull(), w2 = pull();
    uint32_t rp = (uint32_t)(abit & 31);  // bit offset of the next token inside w0 (kept < 32 by advance)
    auto advance = [&](uint32_t n) {     // n < 32
        const uint32_t nrp = rp + n;
        if (nrp & 32u) {
            w0 = w1;
            w1 = w2;
            w2 = pull();
        }
        rp = nrp & 31u;
    };

    uint32_t used = 0;
    uint64_t pos = o.pos;
    while (used <= max_used && pos <= out_limit) {
        const uint32_t bits = simt::funnel_r(w0, w1, rp);
        const uint32_t e = s.litlen[bits & K3_TMASK];
        const uint32_t n = e & 15u;
        if (e & LL_LIT) {
            const uint32_t k = e >> 28;  // 1 or 2 literals
            if (lane < k) o.out[pos + lane] = (uint8_t)(e >> (8u + 8u * lane));
            pos += k;
            used += n;
            advance(n);
            continue;
        }
        if (!(e & LL_LEN)) break;  // end of block, 286/287, or a code longer than the table
        const uint32_t xb = (e >> 8) & 7u;
        const uint32_t length = ((e >> 16) & 0x1ffu) + ((bits >> n) & ((1u << xb) - 1u));
        const uint32_t n1 = n + xb;  // <= 17
        // the 32 bits behind the length: a second window over (w1, w2), shifted into place
        const uint32_t hi = simt::funnel_r(w1, w2, rp);
        const uint32_t dbitsw = simt::funnel_r(bits, hi, n1);
        const uint32_t de = s.dist[dbitsw & 0x1ffu];
        if (!(de & DS_VALID)) break;  // long or invalid distance code
        const uint32_t dbits = de & 15u, dextra = (de >> 4) & 15u;
        const uint32_t dist = (de >> 16) + ((dbitsw >> dbits) & ((1u << dextra) - 1u));
        if (dist > pos - o.lo) break;  // DistanceTooFarBack, reported by the careful loop
        if (lane == mq.qn) {
            mq.dst = pos;
            mq.len = length;
            mq.dist = dist;
        }
        mq.qn++;
        pos += length;
        used += n1 + dbits + dextra;
        advance(n1);
        advance(dbits + dextra);  // <= 22
        if (mq.qn == 32) mq_flush(o.out, mq);
    }
    o.pos = pos;
    br_seek(r, r.pos + used);
}

// rs == nullptr: whole-buffer semantics (decompress_to_vec_bounded).  rs != nullptr: a streaming decoder -- where the
// whole-buffer call would report InsufficientInput or OutputTooLarge, the reader goes back to the start of the token it
// could not finish and returns ST_STREAM_NEED_INPUT / ST_STREAM_OUTPUT_FULL; the next call carries on from there with
// more input / more room (what the reference's read() does with its bit reservoir and its QueuedOutput, :194-219).
// exact: the reference's table entries, one at a time (no fast or parallel path).  With a primary table smaller than
// the reference's 4096 entries (K3_TB < 12) different literals share an entry than there, and the reference wants the
// bits of BOTH literals of an entry before it emits either (:852): at a truncated end of input, status and output
// depend on where its entries begin, which follows from the greedy pairing since the last non-literal token.  The
// kernel therefore decodes a stream that did NOT end with Ok near the end of its input a second time in this mode,
// which restates the 12-bit pairing token by token -- error streams only; valid streams never get here.
FDB_DEVICE int32_t decode_block(K3Smem& s, BitReader& r, OutCursor& o, uint32_t eof_code, uint32_t eof_bits,
                                bool* too_large, K3Resume* rs = nullptr, bool exact = false) {
    const unsigned lane = simt::lane_id();
    const uint32_t eof_mask = (1u << eof_bits) - 1u;
    MatchQueue mq = {0, 0, 0, 0};
    // is the token at bit offset `off` of the (zero-padded) window a literal?  (codes up to 15 bits)
    auto lit_at = [&](uint32_t off, uint32_t* sym, uint32_t* len) -> bool {
        const uint32_t v = br_peek(r, 32) >> off;
        const uint32_t e = s.litlen[v & K3_TMASK];
        if (e & LL_LIT) {
            *sym = (e >> 8) & 0xffu;
            *len = (e >> 24) & 15u;
            return true;
        }
        if (e & (LL_LEN | LL_EOB)) return false;
        uint32_t sy = 0, nb = 0;
        if (!canon_long_decode(s, 0, s.sorted_lit, v & 0x7fffu, K3_TB + 1, &sy, &nb) || sy >= 256) return false;
        *sym = sy;
        *len = nb;
        return true;
    };
// every way out of the block first completes the parked matches
#define K3_RETURN(x)          \
    do {                      \
        mq_flush(o.out, mq);  \
        return (x);           \
    } while (0)
#ifndef K3_NO_PARALLEL
    if (!exact && decode_block_parallel(s, r, o)) return ST_OK;
#endif
    for (;;) {
        // as far as the fast path gets, then ONE token (or the end of the block) the careful way
#ifndef K3_NO_FAST
        if (!exact) decode_block_fast(s, r, o, mq);
#endif
        br_refill(r);
        uint64_t avail = br_avail(r);
        const uint64_t tok_pos = r.pos;  // where this token starts
// the input ends inside this token
#define K3_STARVED()                              \
    do {                                          \
        if (rs) {                                 \
            br_seek(r, tok_pos);                  \
            K3_RETURN(ST_STREAM_NEED_INPUT);      \
        }                                         \
        K3_RETURN(ST_INSUFFICIENT_INPUT);         \
    } while (0)
        if (o.pos == o.cap) {
            // output exactly full: only an end-of-block code may follow (EOB peek, :1009-1015)
            if (avail >= 15 && (br_peek(r, 15) & eof_mask) == eof_code) {
                br_consume(r, eof_bits);
                K3_RETURN(ST_OK);
            }
            if (rs) K3_RETURN(ST_STREAM_OUTPUT_FULL);
            *too_large = true;
            K3_RETURN(ST_OK);
        }
        if (exact) {
            // one entry of the reference's 12-bit table: a literal of <= 12 bits, paired with the literal behind it
            // when both codes fit in 12 bits together (huffman.rs:110-130)
            uint32_t a = 0, l1 = 0;
            if (lit_at(0, &a, &l1) && l1 <= 12) {
                uint32_t b2 = 0, l2 = 0;
                const bool pair = l1 < 12 && lit_at(l1, &b2, &l2) && l1 + l2 <= 12;
                const uint32_t nb = pair ? l1 + l2 : l1;
                if (avail < nb) K3_STARVED();
                if (lane == 0) o.out[o.pos] = (uint8_t)a;
                if (pair && o.pos + 1 == o.cap) {  // the second literal is queued in the reference (:866-876)
                    o.pos += 1;
                    if (rs) {
                        br_consume(r, l1);
                        K3_RETURN(ST_STREAM_OUTPUT_FULL);
                    }
                    br_consume(r, nb);
                    *too_large = true;
                    K3_RETURN(ST_OK);
                }
                if (pair && lane == 1) o.out[o.pos + 1] = (uint8_t)b2;
                o.pos += pair ? 2 : 1;
                br_consume(r, nb);
                continue;
            }
        }
        uint32_t e = s.litlen[br_peek(r, K3_TB)];
        uint32_t nbits = e & 15u;
        if (e & LL_LIT) {  // :846-877
            if (avail < nbits) K3_STARVED();
            bool two = (e & LL_LIT2) != 0;
            if (lane == 0) o.out[o.pos] = (uint8_t)(e >> 8);
            if (two && o.pos + 1 == o.cap) {  // second literal does not fit (queued in the reference)
                o.pos += 1;
                if (rs) {  // take the first literal only; the second one is decoded again by the next call
                    br_consume(r, (e >> 24) & 15u);
                    K3_RETURN(ST_STREAM_OUTPUT_FULL);
                }
                br_consume(r, nbits);
                *too_large = true;
                K3_RETURN(ST_OK);
            }
            if (two && lane == 1) o.out[o.pos + 1] = (uint8_t)(e >> 16);
            o.pos += two ? 2 : 1;
            br_consume(r, nbits);
            continue;
        }
        uint32_t len_base, len_extra;
        if (e & LL_EOB) {  // :910-918 (includes the 286/287 quirk)
            if (avail < nbits) K3_STARVED();
            br_consume(r, nbits);
            K3_RETURN(ST_OK);
        } else if (e & LL_LEN) {
            len_base = (e >> 16) & 0x1ffu;
            len_extra = (e >> 8) & 7u;
        } else {  // code longer than 12 bits (:886-909)
            uint32_t sym = 0;
            if (!canon_long_decode(s, 0, s.sorted_lit, br_peek(r, 15), K3_TB + 1, &sym, &nbits))
                K3_RETURN(ST_INVALID_LITERAL_LENGTH_CODE);
            if (avail < nbits) K3_STARVED();
            if (sym < 256) {
                if (lane == 0) o.out[o.pos] = (uint8_t)sym;
                o.pos += 1;
                br_consume(r, nbits);
                continue;
            } else if (sym == 256) {
                br_consume(r, nbits);
                K3_RETURN(ST_OK);
            }
            len_base = len_sym_base(sym);
            len_extra = len_sym_extra(sym);
        }
        // length + distance (:919-967); the reference tests the total against nbits before consuming
        uint32_t le_bits = nbits + len_extra;  // <= 20
        uint32_t length = len_base + ((br_peek(r, le_bits) >> nbits) & ((1u << len_extra) - 1u));
        br_consume(r, le_bits);
        br_refill(r);
        uint32_t de = s.dist[br_peek(r, 9)];
        uint32_t dbits, dextra, dbase;
        if (de & DS_VALID) {
            dbits = de & 15u;
            dextra = (de >> 4) & 15u;
            dbase = de >> 16;
        } else if (avail > (uint64_t)le_bits + 9) {  // :932-951
            if (!(de & DS_LONG)) K3_RETURN(ST_INVALID_DISTANCE_CODE);
            uint32_t dsym = 0;
            if (!canon_long_decode(s, 1, s.sorted_dist, br_peek(r, 15), 10, &dsym, &dbits))
                K3_RETURN(ST_INVALID_DISTANCE_CODE);
            if (dsym >= 30) K3_RETURN(ST_INVALID_DISTANCE_CODE);
            dextra = dist_sym_extra(dsym);
            dbase = dist_sym_base(dsym);
        } else {
            K3_STARVED();  // `break` with the input exhausted
        }
        uint32_t dd_bits = dbits + dextra;  // <= 28
        uint64_t dist = dbase + ((br_peek(r, dd_bits) >> dbits) & ((1u << dextra) - 1u));
        if (avail < (uint64_t)le_bits + dd_bits) K3_STARVED();                          // :961
        if (dist > o.pos - o.lo) K3_RETURN(ST_DISTANCE_TOO_FAR_BACK);                  // :963
        br_consume(r, dd_bits);

        uint64_t room = o.cap - o.pos;
        uint32_t n = length < room ? length : (uint32_t)room;
        if (lane == mq.qn) {
            mq.dst = o.pos;
            mq.len = n;
            mq.dist = (uint32_t)dist;
        }
        mq.qn++;
        o.pos += n;
        if (n < length) {  // remainder would be queued (:797-801, :823-827) => output too large
            if (rs) {
                rs->pend_len = length - n;
                rs->pend_dist = (uint32_t)dist;
                K3_RETURN(ST_STREAM_OUTPUT_FULL);
            }
            *too_large = true;
            K3_RETURN(ST_OK);
        }
        if (mq.qn == 32) mq_flush(o.out, mq);
    }
#undef K3_STARVED
#undef K3_RETURN
}

// The header of a dynamic block, its 3 type bits included (:415-434, :440-555): HLIT / HDIST / HCLEN, the code-length
// code, the HLIT + HDIST code lengths into s.lens (litlen at 0..287, distance at 288..319).  ST_INSUFFICIENT_INPUT
// when the input ends inside the header (a streaming decoder then parses it again from its first bit).
FDB_DEVICE int32_t read_dynamic_header(K3Smem& s, BitReader& r, uint32_t* hlit_out) {
    const unsigned lane = simt::lane_id();
    if (br_avail(r) < 17) return ST_INSUFFICIENT_INPUT;
    uint32_t hlit = (br_peek(r, 8) >> 3) + 257;
    uint32_t hdist = (br_peek(r, 13) >> 8) + 1;
    uint32_t hclen = (br_peek(r, 17) >> 13) + 4;
    if (hlit > 286) return ST_INVALID_HLIT;
    if (hdist > 30) return ST_INVALID_HDIST;
    br_consume(r, 17);
    if (br_avail(r) < 3ull * hclen) return ST_INSUFFICIENT_INPUT;
    if (lane < 32) s.cl_lens[lane] = 0;
    simt::syncwarp();
    for (uint32_t i = 0; i < hclen; i++) {
        br_refill(r);
        // reference tables.rs:63-65 CLCL_ORDER, packed 5 bits per entry
        const uint64_t order_lo = 0x22caa324e804a30ull;  // entries 0..11
        const uint64_t order_hi = 0x3c2e1346cull;  // entries 12..18
        uint32_t sym = (i < 12) ? (uint32_t)((order_lo >> (5 * i)) & 31u)
                                : (uint32_t)((order_hi >> (5 * (i - 12))) & 31u);
        if (lane == 0) s.cl_lens[sym] = (uint8_t)br_peek(r, 3);
        br_consume(r, 3);
    }
    simt::syncwarp();
    canon_setup(s, 2, s.cl_lens, 19, s.sorted_cl, 7);
    if (!s.info[1]) return ST_BAD_CODE_LENGTH_HUFFMAN_TREE;  // :462-472
    for (uint32_t idx = lane; idx < 128; idx += 32) {
        uint32_t v = simt::brev(idx) >> 25;
        uint32_t e = 0;
        for (uint32_t L = 1; L <= 7; L++) {
            if (v < s.lim[2][L]) {
                uint32_t c = v >> (7 - L);
                e = ((uint32_t)s.sorted_cl[s.off[2][L] + c - s.first[2][L]] << 3) | L;
                break;
            }
        }
        s.cl_table[idx] = (uint8_t)e;
    }
    simt::syncwarp();
    // code lengths (:479-539)
    uint32_t total = hlit + hdist, nread = 0;
    while (nread < total) {
        br_refill(r);
        uint64_t avail = br_avail(r);
        if (avail < 7) return ST_INSUFFICIENT_INPUT;
        uint32_t ce = s.cl_table[br_peek(r, 7)];
        uint32_t clen = ce & 7u, sym = ce >> 3;
        if (sym <= 15) {
            if (lane == 0) s.lens[nread] = (uint8_t)sym;
            nread += 1;
            br_consume(r, clen);
        } else {
            uint32_t base_repeat = sym == 18 ? 11u : 3u;
            uint32_t extra = sym == 16 ? 2u : sym == 17 ? 3u : 7u;
            if (avail < clen + extra) return ST_INSUFFICIENT_INPUT;
            uint32_t value = 0;
            if (sym == 16) {
                if (nread == 0) return ST_INVALID_CODE_LENGTH_REPEAT;
                simt::syncwarp();
                value = s.lens[nread - 1];
            }
            uint32_t repeat = (br_peek(r, clen + extra) >> clen) + base_repeat;
            if (nread + repeat > total) return ST_INVALID_CODE_LENGTH_REPEAT;
            simt::syncwarp();
            if (lane < repeat) s.lens[nread + lane] = (uint8_t)value;
            for (uint32_t i = 32 + lane; i < repeat; i += 32) s.lens[nread + i] = (uint8_t)value;
            nread += repeat;
            br_consume(r, clen + extra);
        }
    }
    simt::syncwarp();
    // split into litlen[0..288) and dist[288..320) (:541-549); hdist <= 30 so no overlap issues
    uint8_t dl = (lane < hdist) ? s.lens[hlit + lane] : (uint8_t)0;
    simt::syncwarp();
    for (uint32_t i = hlit + lane; i < 288; i += 32) s.lens[i] = 0;
    s.lens[288 + lane] = dl;
    simt::syncwarp();
    *hlit_out = hlit;
    return ST_OK;
}

// ---- whole stream --------------------------------------------------------------------------
FDB_DEVICE int32_t inflate_stream_general_impl(K3Smem& s, BitReader& r, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap,
                                               uint32_t flags, uint64_t* out_len, uint64_t* consumed, bool exact) {
    br_init(r, in, n);
    OutCursor o = {out, 0, cap, 0};
    *out_len = 0;
    *consumed = 0;

    // zlib header (:226-244)
    if (br_avail(r) < 16) return ST_INSUFFICIENT_INPUT;
    {
        uint32_t cmf = br_peek(r, 8), flg = br_peek(r, 16) >> 8;
        if ((cmf & 0x0f) != 0x08 || (cmf & 0xf0) > 0x70 || (flg & 0x20) != 0 || ((cmf << 8) | flg) % 31 != 0)
            return ST_BAD_ZLIB_HEADER;
        br_consume(r, 16);
    }

    bool fixed_ready = false;
    uint32_t fixed_eof_code = 0, fixed_eof_bits = 7;
    bool too_large = false;
    for (;;) {
        br_refill(r);
        if (br_avail(r) < 10) {  // :346
            *out_len = o.pos;
            return ST_INSUFFICIENT_INPUT;
        }
        uint32_t hdr = br_peek(r, 3);
        bool last = (hdr & 1) != 0;
        uint32_t btype = hdr >> 1;
        int32_t st = ST_OK;
        if (btype == 0) {  // stored (:353-370, :271-305)
            uint32_t align = (8u - (uint32_t)((r.pos + 3) & 7)) & 7u;
            if (br_avail(r) < 3 + 32 + align) {
                *out_len = o.pos;
                return ST_INSUFFICIENT_INPUT;
            }
            br_consume(r, 3 + align);
            br_refill(r);
            uint32_t v = br_peek(r, 32);
            uint32_t len = v & 0xffffu, nlen = v >> 16;
            if (nlen != (len ^ 0xffffu)) return ST_INVALID_UNCOMPRESSED_BLOCK_LENGTH;
            br_consume(r, 32);
            uint64_t src_byte = r.pos >> 3;
            uint64_t in_left = n - src_byte;
            uint64_t room = o.cap - o.pos;
            uint64_t ncopy = len;
            if (ncopy > in_left) ncopy = in_left;
            if (ncopy > room) ncopy = room;
            warp_copy_bytes(o.out + o.pos, in + src_byte, ncopy);
            o.pos += ncopy;
            if (ncopy < len) {
                *out_len = o.pos;
                // decompress.rs:1128-1136: a full output is reported before an exhausted input
                return (o.pos == o.cap) ? ST_OUTPUT_TOO_LARGE : ST_INSUFFICIENT_INPUT;
            }
            br_seek(r, (src_byte + len) * 8);
        } else if (btype == 1) {  // fixed (:371-414)
            br_consume(r, 3);
            if (!fixed_ready) {
                load_fixed_lengths(s);
                st = build_block_tables(s, 288, &fixed_eof_code, &fixed_eof_bits);
                fixed_ready = true;
            }
            if (st == ST_OK) st = decode_block(s, r, o, fixed_eof_code, fixed_eof_bits, &too_large, nullptr, exact);
        } else if (btype == 2) {  // dynamic (:415-434, :440-555)
            uint32_t hlit = 0;
            st = read_dynamic_header(s, r, &hlit);
            fixed_ready = false;
            if (st == ST_INSUFFICIENT_INPUT) {
                *out_len = o.pos;
                return st;
            }
            if (st != ST_OK) return st;
            uint32_t eof_code = 0, eof_bits = 0;
            st = build_block_tables(s, hlit, &eof_code, &eof_bits);
            if (st == ST_OK) st = decode_block(s, r, o, eof_code, eof_bits, &too_large, nullptr, exact);
        } else {
            return ST_INVALID_BLOCK_TYPE;  // :435
        }
        if (st != ST_OK) {
            *out_len = o.pos;
            return st;
        }
        if (too_large) {
            *out_len = o.pos;
            return ST_OUTPUT_TOO_LARGE;
        }
        if (last) break;
    }

    // checksum (:306-326)
    *out_len = o.pos;
    uint32_t align = (8u - (uint32_t)(r.pos & 7)) & 7u;
    if (br_avail(r) < 32 + align) return ST_INSUFFICIENT_INPUT;
    br_consume(r, align);
    br_refill(r);
    uint32_t stored = br_peek(r, 32);
    stored = simt::byte_perm(stored, 0, 0x0123);  // big-endian on the wire
    br_consume(r, 32);
    *consumed = r.pos >> 3;
    if (!(flags & FLAG_IGNORE_ADLER32)) {
        simt::syncwarp();
        uint32_t got = warp_adler32(out, o.pos);
        if (got != stored) return ST_WRONG_CHECKSUM;
    }
    return ST_OK;
}

FDB_DEVICE int32_t inflate_stream_general(K3Smem& s, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap,
                                          uint32_t flags, uint64_t* out_len, uint64_t* consumed) {
    BitReader r;
    int32_t st = inflate_stream_general_impl(s, r, in, n, out, cap, flags, out_len, consumed, false);
    // (see decode_block: a stream that failed within 64 bits of the end of its input is decoded again entry by entry)
    if (K3_TB < 12 && st != ST_OK && br_avail(r) < 64) {
        simt::syncwarp();
        st = inflate_stream_general_impl(s, r, in, n, out, cap, flags, out_len, consumed, true);
    }
    return st;
}

// Persistent kernel: one warp per CTA, warps pull stream indices from a device counter.
// worklist == nullptr: all streams 0..n-1; else the first *work_count entries of worklist.
#ifndef K3_MIN_CTAS
#define K3_MIN_CTAS 14  // one warp per CTA: what 15.5 KB of shared memory per warp allow (128 registers, no spills)
#endif
FDB_GLOBAL void FDB_LAUNCH_BOUNDS(32, K3_MIN_CTAS)
    inflate_general_kernel(InflateBatch b, const uint32_t* worklist, const uint32_t* work_count, uint32_t* next) {
    FDB_DYN_SMEM(smem_raw);
    K3Smem& s = *reinterpret_cast<K3Smem*>(smem_raw);
    const unsigned lane = simt::lane_id();
    const uint32_t count = worklist ? *work_count : b.n;
    if (b.general_out && blockIdx.x == 0 && lane == 0) *b.general_out = worklist ? count : 0u;
    for (;;) {
        uint32_t idx = 0;
        if (lane == 0) idx = simt::atomic_add(next, 1u);
        idx = simt::shfl(idx, 0);
        if (idx >= count) break;
        uint32_t i = worklist ? worklist[idx] : idx;
        uint64_t out_len = 0, consumed = 0;
        const uint64_t cap = b.out_cap[i];
        int32_t st = inflate_stream_general(s, b.in_base + b.in_off[i], b.in_len[i], b.out_base + b.out_off[i], cap,
                                            b.flags, &out_len, &consumed);
        // decompress_to_vec_bounded tests "output_index == maxlen" before "input exhausted"
        // (reference decompress.rs:1128 vs :1135)
        if (st == ST_INSUFFICIENT_INPUT && out_len == cap) st = ST_OUTPUT_TOO_LARGE;
        if (lane == 0) {
            b.status[i] = st;
            b.out_len[i] = out_len;
            if (b.consumed) b.consumed[i] = consumed;
        }
        simt::syncwarp();
    }
}

// ---- streaming decoders: Decompressor::read kept on the device between calls -------------------------------------
// (reference src/decompress.rs:96-113 struct, :158-219 read contract, :1066-1070 QueuedOutput)
//
// One K3StreamState per in-flight decoder lives in device memory owned by the context: where in the stream the decoder
// stands (zlib header / block header / stored body / Huffman body / checksum), the tables of the block it is inside, a
// match cut short by a full output, the running adler32.  Next to it a device buffer holds the decoder's recent output
// (at least the last 32 KiB: the window matches reach into) with the new bytes appended behind it.
// fdb_stream_read_batch hands every decoder the bytes it has not parsed yet plus the new input; the decoder resumes at
// its checkpoint -- always a token boundary (or the first bit of a block header, which is parsed again once more of
// it has arrived) -- decodes until the input ends inside the next item or the caller's room is full, and leaves a new
// checkpoint.  Work per call is proportional to the bytes of that call (plus one 20 KB table restore), so feeding a
// stream byte by byte is linear in its length, not quadratic.
enum : uint32_t { PH_ZLIB = 0, PH_BLOCK = 1, PH_STORED = 2, PH_HUFF = 3, PH_CHECKSUM = 4, PH_DONE = 5, PH_ERROR = 6 };
static const uint32_t K3_TABLE_WORDS = (uint32_t)(offsetof(K3Smem, pstg) / 4);  // everything decode_block reads

struct K3StreamState {
    uint32_t phase;
    uint32_t last;         // the block the decoder is inside is the final one
    uint32_t stored_left;  // PH_STORED: bytes of the block still to copy
    uint32_t pend_len, pend_dist;
    uint32_t eof_code, eof_bits;
    uint32_t adler_a, adler_b;  // adler32 of everything produced so far
    int32_t error;              // PH_ERROR: the status every later call reports
    uint64_t total_out;
    uint32_t tables[K3_TABLE_WORDS];  // PH_HUFF: the shared-memory tables of the block
};

// one call of one decoder
struct K3StreamJob {
    K3StreamState* state;
    const uint8_t* in;    // the bytes not parsed yet, then the new ones
    uint64_t in_len;
    uint32_t start_bit;   // bits of in[0] already consumed (0..7)
    uint32_t flags;
    uint8_t* buf;         // buf[lo .. pos) = the most recent output, new output goes to buf[pos .. pos + room)
    uint64_t pos, lo, room;
    // results
    uint64_t produced, consumed_bits;
    int32_t status;       // ST_OK = stream complete; ST_STREAM_NEED_INPUT / ST_STREAM_OUTPUT_FULL; or an error
};

FDB_DEVICE void inflate_stream_resume(K3Smem& s, K3StreamJob& job) {
    const unsigned lane = simt::lane_id();
    K3StreamState& ss = *job.state;
    uint32_t phase = ss.phase, last = ss.last, stored_left = ss.stored_left;
    uint32_t eof_code = ss.eof_code, eof_bits = ss.eof_bits;
    K3Resume rs = {ss.pend_len, ss.pend_dist};
    int32_t st = ST_OK;
    BitReader r;
    br_init(r, job.in, job.in_len);
    if (job.start_bit) br_seek(r, job.start_bit);
    OutCursor o = {job.buf, job.pos, job.pos + job.room, job.lo};
    uint32_t stored_sum = 0;
    bool have_sum = false;
    simt::syncwarp();
    if (phase == PH_DONE) {
        st = ST_OK;
    } else if (phase == PH_ERROR) {
        st = ss.error;
    } else {
        if (phase == PH_HUFF) {  // back inside a block: its tables
            uint32_t* dst = reinterpret_cast<uint32_t*>(&s);
            for (uint32_t i = lane; i < K3_TABLE_WORDS; i += 32) dst[i] = ss.tables[i];
            simt::syncwarp();
            if (rs.pend_len) {  // the rest of the match the last call could not finish
                const uint64_t room = o.cap - o.pos;
                const uint32_t n = rs.pend_len < room ? rs.pend_len : (uint32_t)room;
                uint8_t* d = o.out + o.pos;
                const uint8_t* src = d - rs.pend_dist;
                if (rs.pend_dist >= n) {
                    for (uint32_t i = lane; i < n; i += 32) d[i] = src[i];
                } else {
                    for (uint32_t i = lane; i < n; i += 32) d[i] = src[i % rs.pend_dist];
                }
                simt::syncwarp();
                o.pos += n;
                rs.pend_len -= n;
            }
        }
        for (;;) {
            if (phase == PH_ZLIB) {  // :226-244
                if (br_avail(r) < 16) {
                    st = ST_STREAM_NEED_INPUT;
                    break;
                }
                uint32_t cmf = br_peek(r, 8), flg = br_peek(r, 16) >> 8;
                if ((cmf & 0x0f) != 0x08 || (cmf & 0xf0) > 0x70 || (flg & 0x20) != 0 || ((cmf << 8) | flg) % 31 != 0) {
                    st = ST_BAD_ZLIB_HEADER;
                    break;
                }
                br_consume(r, 16);
                phase = PH_BLOCK;
            } else if (phase == PH_BLOCK) {
                const uint64_t hdr_pos = r.pos;  // a header the input ends inside is parsed again from here
                br_refill(r);
                if (br_avail(r) < 10) {  // :346
                    st = ST_STREAM_NEED_INPUT;
                    break;
                }
                const uint32_t hdr = br_peek(r, 3);
                last = hdr & 1u;
                const uint32_t btype = hdr >> 1;
                if (btype == 0) {  // stored (:353-370)
                    const uint32_t align = (8u - (uint32_t)((r.pos + 3) & 7)) & 7u;
                    if (br_avail(r) < 3 + 32 + align) {
                        st = ST_STREAM_NEED_INPUT;
                        break;
                    }
                    br_consume(r, 3 + align);
                    br_refill(r);
                    const uint32_t v = br_peek(r, 32);
                    const uint32_t len = v & 0xffffu, nlen = v >> 16;
                    if (nlen != (len ^ 0xffffu)) {
                        st = ST_INVALID_UNCOMPRESSED_BLOCK_LENGTH;
                        break;
                    }
                    br_consume(r, 32);
                    stored_left = len;
                    phase = PH_STORED;
                } else if (btype == 1) {  // fixed (:371-414)
                    br_consume(r, 3);
                    load_fixed_lengths(s);
                    st = build_block_tables(s, 288, &eof_code, &eof_bits);
                    if (st != ST_OK) break;
                    phase = PH_HUFF;
                } else if (btype == 2) {  // dynamic (:415-434, :440-555)
                    uint32_t hlit = 0;
                    st = read_dynamic_header(s, r, &hlit);
                    if (st == ST_INSUFFICIENT_INPUT) {
                        br_seek(r, hdr_pos);
                        st = ST_STREAM_NEED_INPUT;
                        break;
                    }
                    if (st != ST_OK) break;
                    st = build_block_tables(s, hlit, &eof_code, &eof_bits);
                    if (st != ST_OK) break;
                    phase = PH_HUFF;
                } else {
                    st = ST_INVALID_BLOCK_TYPE;  // :435
                    break;
                }
            } else if (phase == PH_STORED) {  // :271-305; the reader stands on a byte boundary
                const uint64_t src_byte = r.pos >> 3;
                const uint64_t in_left = job.in_len - src_byte;
                const uint64_t room = o.cap - o.pos;
                uint64_t ncopy = stored_left;
                if (ncopy > in_left) ncopy = in_left;
                if (ncopy > room) ncopy = room;
                warp_copy_bytes(o.out + o.pos, job.in + src_byte, ncopy);
                o.pos += ncopy;
                stored_left -= (uint32_t)ncopy;
                br_seek(r, (src_byte + ncopy) * 8);
                if (stored_left) {
                    st = (o.pos == o.cap) ? ST_STREAM_OUTPUT_FULL : ST_STREAM_NEED_INPUT;
                    break;
                }
                phase = last ? PH_CHECKSUM : PH_BLOCK;
            } else if (phase == PH_HUFF) {
                if (rs.pend_len) {  // (the room was used up by the pending match)
                    st = ST_STREAM_OUTPUT_FULL;
                    break;
                }
                bool too_large = false;
                st = decode_block(s, r, o, eof_code, eof_bits, &too_large, &rs);
                if (st != ST_OK) break;  // need input / output full (the checkpoint is r.pos), or an error
                phase = last ? PH_CHECKSUM : PH_BLOCK;
            } else {  // PH_CHECKSUM (:306-326)
                const uint32_t align = (8u - (uint32_t)(r.pos & 7)) & 7u;
                if (br_avail(r) < 32 + align) {
                    st = ST_STREAM_NEED_INPUT;
                    break;
                }
                br_consume(r, align);
                br_refill(r);
                stored_sum = simt::byte_perm(br_peek(r, 32), 0, 0x0123);  // big-endian on the wire
                br_consume(r, 32);
                have_sum = true;
                phase = PH_DONE;
                st = ST_OK;
                break;
            }
        }
    }
    simt::syncwarp();
    // adler32 of the new bytes, appended to the running value: a' = a + sum d, b' = b + n a + sum (n - i) d_i
    const uint64_t produced = o.pos - job.pos;
    uint32_t a = ss.adler_a, b = ss.adler_b;
    if (produced && !(job.flags & FLAG_IGNORE_ADLER32)) {
        const uint8_t* d = job.buf + job.pos;
        for (uint64_t base = 0; base < produced; base += 4096) {  // (4096 * 255 * 4096 < 2^32)
            const uint32_t n = produced - base < 4096 ? (uint32_t)(produced - base) : 4096u;
            uint32_t s1 = 0, s2 = 0;
            for (uint32_t i = lane; i < n; i += 32) {
                const uint32_t v = simt::ldcg8(d + base + i);
                s1 += v;
                s2 += (n - i) * v;
            }
            s1 = simt::reduce_add(s1);
            s2 = simt::reduce_add(s2 % ADLER_MOD) % ADLER_MOD;
            b = (uint32_t)((b + (uint64_t)n * a + s2) % ADLER_MOD);
            a = (a + s1) % ADLER_MOD;
        }
    }
    if (have_sum && !(job.flags & FLAG_IGNORE_ADLER32) && ((b << 16) | a) != stored_sum) st = ST_WRONG_CHECKSUM;
    if (st > 0) phase = PH_ERROR;
    // the checkpoint
    if (phase == PH_HUFF && ss.phase != PH_ERROR && ss.phase != PH_DONE) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&s);
        for (uint32_t i = lane; i < K3_TABLE_WORDS; i += 32) ss.tables[i] = src[i];
    }
    simt::syncwarp();
    if (lane == 0 && ss.phase != PH_ERROR && ss.phase != PH_DONE) {
        ss.phase = phase;
        ss.last = last;
        ss.stored_left = stored_left;
        ss.pend_len = rs.pend_len;
        ss.pend_dist = rs.pend_dist;
        ss.eof_code = eof_code;
        ss.eof_bits = eof_bits;
        ss.adler_a = a;
        ss.adler_b = b;
        ss.total_out += produced;
        if (st > 0) ss.error = st;
    }
    if (lane == 0) {
        job.produced = produced;
        job.consumed_bits = r.pos;
        job.status = st;
    }
    simt::syncwarp();
}

// one warp per decoder
FDB_GLOBAL void FDB_LAUNCH_BOUNDS(32, 1) inflate_stream_kernel(K3StreamJob* jobs, uint32_t n, uint32_t* next) {
    FDB_DYN_SMEM(smem_raw);
    K3Smem& s = *reinterpret_cast<K3Smem*>(smem_raw);
    const unsigned lane = simt::lane_id();
    for (;;) {
        uint32_t idx = 0;
        if (lane == 0) idx = simt::atomic_add(next, 1u);
        idx = simt::shfl(idx, 0);
        if (idx >= n) break;
        inflate_stream_resume(s, jobs[idx]);
    }
}

// keep the last `keep` bytes of buf[0 .. pos) at the front of the buffer (the window of a decoder whose buffer is full)
FDB_GLOBAL void stream_compact_kernel(uint8_t* buf, uint64_t pos, uint64_t keep) {
    const uint64_t shift = pos - keep;  // > 0
    // dst < src: chunks in ascending order, every chunk read completely before it is written
    for (uint64_t base = 0; base < keep; base += blockDim.x) {
        const uint64_t i = base + threadIdx.x;
        uint8_t v = 0;
        if (i < keep) v = buf[shift + i];
        simt::syncthreads();
        if (i < keep) buf[i] = v;
        simt::syncthreads();
    }
}

}  // namespace fdb
