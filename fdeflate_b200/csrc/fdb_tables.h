// fdb_tables.h -- host-side construction of the constant tables of the ultra-fast format.
//
// The format is defined by two constants of the reference: the code lengths HUFFMAN_LENGTHS
// (src/tables.rs:7-20, a length-limited Huffman code trained on filtered PNG data) and the 54-byte
// stream header that announces exactly that code (src/compress/ultrafast.rs:82-86).  Codes are the
// canonical codes of those lengths, bit-reversed for LSB-first packing (src/lib.rs:103-127).
#pragma once
#include <stdint.h>
#include <string.h>

#include "fdb_common.h"

namespace fdb {

// HUFFMAN_LENGTHS as (length, repeat) runs over symbols 0..285
static const uint8_t UF_LEN_RUNS[][2] = {
    {2, 1},  {3, 1},  {4, 1},  {5, 2},  {6, 2},  {7, 3}, {8, 5},  {9, 7},  {10, 9}, {11, 12}, {12, 171}, {11, 10},
    {10, 1}, {11, 1}, {10, 9}, {9, 5},  {8, 1},  {9, 1}, {8, 5},  {7, 3},  {6, 3},  {5, 1},   {4, 1},    {3, 1},
    {12, 3}, {9, 2},  {11, 1}, {10, 1}, {11, 2}, {10, 1}, {11, 6}, {12, 1}, {11, 1}, {12, 11}, {9, 1},
};

static const char UF_HEADER_HEX[] =
    "7801edc003a0245996c6f1ff77ee8dc8cca7724b63ae6ddbb66ddbb66ddbb66d698c9e964aaf9e323322eef976b76a7aa6873b6b"
    "d50f";

struct UfHostTables {
    uint8_t len[286];
    uint16_t code[286];
    uint32_t lit_tok[256];
    uint32_t tail_tok[258];
    uint32_t header[14];
    uint32_t wt[4096];  // UW write table (fdb_common.h)
    uint32_t ct[4096];  // UC count table
    uint16_t bt[4096];  // UB boundary table
};

static inline uint32_t rev_bits(uint32_t v, uint32_t n) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < n; i++)
        if (v & (1u << i)) r |= 1u << (n - 1 - i);
    return r;
}

static inline bool build_uf_host_tables(UfHostTables& t) {
    memset(&t, 0, sizeof t);
    uint32_t sym = 0;
    for (size_t r = 0; r < sizeof(UF_LEN_RUNS) / sizeof(UF_LEN_RUNS[0]); r++)
        for (uint32_t k = 0; k < UF_LEN_RUNS[r][1]; k++) {
            if (sym >= 286) return false;
            t.len[sym++] = UF_LEN_RUNS[r][0];
        }
    if (sym != 286) return false;

    // canonical codes (RFC 1951 3.2.2), bit-reversed
    uint32_t count[16] = {0}, next[16] = {0};
    for (uint32_t s = 0; s < 286; s++) count[t.len[s]]++;
    {
        uint32_t c = 0, kraft = 0;
        for (uint32_t L = 1; L <= 15; L++) {
            c <<= 1;
            next[L] = c;  // first code of length L
            c += count[L];
            kraft += count[L] << (15 - L);
        }
        if (kraft != (1u << 15)) return false;  // the code must be complete
    }
    for (uint32_t s = 0; s < 286; s++) {
        uint32_t L = t.len[s];
        t.code[s] = (uint16_t)rev_bits(next[L]++, L);
    }

    // encoder tokens
    for (uint32_t b = 0; b < 256; b++) t.lit_tok[b] = (uint32_t)t.code[b] | ((uint32_t)t.len[b] << 16);
    // tail of a zero run, r = (R - 1) mod 258 bytes still owed after the sym-285 chain
    // (reference src/compress/ultrafast.rs:54-64)
    t.tail_tok[0] = 0;
    for (uint32_t r = 1; r < 258; r++) {
        if (r <= 4) {
            t.tail_tok[r] = (2u * r) << 24;  // r literal zeros, code 00 each
        } else {
            // length r as (symbol, extra bits) per RFC 1951 3.2.5, then the 1-bit distance code 0
            uint32_t s = 257;
            while (s < 285 && !(r >= len_sym_base(s) && r < len_sym_base(s) + (1u << len_sym_extra(s)))) s++;
            uint32_t xb = len_sym_extra(s);
            uint32_t extra = (r - len_sym_base(s)) & ((1u << xb) - 1u);
            uint32_t bits = (uint32_t)t.code[s] | (extra << t.len[s]);
            uint32_t nbits = t.len[s] + xb + 1u;
            t.tail_tok[r] = bits | (nbits << 24);
        }
    }

    // header bytes
    uint8_t hb[56] = {0};
    for (int i = 0; i < 54; i++) {
        auto nib = [](char c) -> uint32_t { return c <= '9' ? (uint32_t)(c - '0') : (uint32_t)(c - 'a' + 10); };
        hb[i] = (uint8_t)((nib(UF_HEADER_HEX[2 * i]) << 4) | nib(UF_HEADER_HEX[2 * i + 1]));
    }
    memcpy(t.header, hb, 56);

    // decode tables (formats in fdb_common.h): walk the tokens that lie completely inside the 12 index bits
    auto first_sym = [&](uint32_t v, uint32_t avail, uint32_t* L) -> int {
        for (uint32_t s = 0; s < 286; s++) {
            uint32_t l = t.len[s];
            if (l <= avail && (v & ((1u << l) - 1u)) == t.code[s]) {
                *L = l;
                return (int)s;
            }
        }
        return -1;  // the next code does not fit in the bits that are left
    };
    for (uint32_t idx = 0; idx < 4096; idx++) {
        uint32_t L = 0;
        const int s = first_sym(idx, 12, &L);
        if (s < 0) return false;  // every code is <= 12 bits
        if (s < 256) {
            uint32_t pos = 0, k = 0, bytes = 0, wbits = 0, wk = 0, last = 0, ends = 0;
            for (;;) {
                uint32_t l2 = 0;
                const int s2 = first_sym(idx >> pos, 12 - pos, &l2);
                if (s2 < 0 || s2 >= 256 || k == 6) break;
                if (k < 3) {
                    bytes |= (uint32_t)s2 << (8 * k);
                    wbits = pos + l2;
                    wk = k + 1;
                }
                pos += l2;
                ends |= 1u << (pos - 1u);
                k++;
                last = (uint32_t)s2;
            }
            t.wt[uf_slot(idx)] = wbits | (bytes << 5) | (wk << 30);
            t.ct[uf_slot(idx)] = pos | (k << UC_CNT_SHIFT) | (L << UC_FIRST_SHIFT) | (last ? UC_ENDNZ : 0u) | (s ? UC_FIRSTNZ : 0u);
            t.bt[uf_slot(idx)] = (uint16_t)ends;
        } else if (s == 256) {
            t.wt[uf_slot(idx)] = (L | UW_EOB) << UW_SPECIAL_SHIFT;
            t.ct[uf_slot(idx)] = 0;
        } else {
            const uint32_t xb = len_sym_extra((uint32_t)s), base = len_sym_base((uint32_t)s), tot = L + xb + 1u;
            const uint32_t len = base + ((idx >> L) & ((1u << xb) - 1u));  // (only meaningful when tot <= 12)
            // the write loop takes every run through its special path (where it looks at the byte before the run);
            // the count loop steps over a run that fits in the 12 bits like over literals
            t.wt[uf_slot(idx)] = (L | (xb << 4) | (base << 8)) << UW_SPECIAL_SHIFT;
            if (tot <= 12 && len <= 12 && ((idx >> (L + xb)) & 1u) == 0)  // (symbol 285 = 258 bytes stays special)
                t.ct[uf_slot(idx)] = tot | (len << UC_CNT_SHIFT) | (tot << UC_FIRST_SHIFT) | UC_RUN;
            else
                t.ct[uf_slot(idx)] = 0;
        }
    }
    return true;
}

}  // namespace fdb
