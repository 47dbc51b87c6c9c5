// crc32.cuh -- K10: CRC-32 (ISO-HDLC / zlib / PNG chunk CRC, reflected polynomial 0xEDB88320) of a batch of byte
// ranges, ONE WARP PER RANGE (PNG chunk framing).  Not part of image-rs/fdeflate (the `png`
// crate checks chunk CRCs with the crc32fast crate); the algorithm is the standard one and the known answers are
// zlib's crc32().
//
// A CRC is linear over GF(2): crc(A || B) = x^(8|B|) * crc(A) + crc(B)  (mod p), which holds for the finished values
// with their 0xffffffff pre/post conditioning.  So the 32 lanes of a warp take one contiguous block each (table
// driven, four bytes per step: "slicing by 4"), and lane 0 folds the 32 block CRCs left to right with carry-less
// multiplications by x^(8 * block length) mod p.
#pragma once
#include "simt.h"
#include "fdb_common.h"

namespace fdb {

static const uint32_t CRC_POLY = 0xedb88320u;

struct CrcTables {
    uint32_t t[4][256];   // slicing-by-4 tables
    uint32_t x2n[32];     // x^(2^k) mod p, reflected
};

// a(x) * b(x) mod p(x) on reflected 32-bit polynomials (bit 31 = x^0)
FDB_HD uint32_t crc_mulmod(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1u)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ 0xedb88320u : b >> 1;
    }
    return p;
}
// x^(n * 2^k) mod p
FDB_HD uint32_t crc_xpow(const uint32_t* x2n, uint64_t n, uint32_t k) {
    uint32_t p = 1u << 31;
    while (n) {
        if (n & 1u) p = crc_mulmod(x2n[k & 31u], p);
        n >>= 1;
        k++;
    }
    return p;
}
static inline void build_crc_tables(CrcTables& c) {
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t v = i;
        for (int k = 0; k < 8; k++) v = (v & 1u) ? (v >> 1) ^ 0xedb88320u : v >> 1;
        c.t[0][i] = v;
    }
    for (uint32_t i = 0; i < 256; i++)
        for (int j = 1; j < 4; j++) c.t[j][i] = (c.t[j - 1][i] >> 8) ^ c.t[0][c.t[j - 1][i] & 0xffu];
    uint32_t p = 1u << 30;  // x^1
    c.x2n[0] = p;
    for (int k = 1; k < 32; k++) c.x2n[k] = p = crc_mulmod(p, p);
}

struct CrcBatch {
    const uint8_t* base;
    const uint64_t* off;  // [n]
    const uint64_t* len;  // [n]
    uint32_t* crc;        // [n] out
    uint32_t n;
    uint32_t seed;        // CRC of what precedes every range (0 = nothing), e.g. crc32("IDAT")
};

// finished CRC of p[0..n) continuing from the finished value `crc`
FDB_DEVICE uint32_t crc_block(const uint32_t (*t)[256], const uint8_t* p, uint64_t n, uint32_t crc) {
    uint32_t r = ~crc;
    while (n && ((uintptr_t)p & 3u)) {
        r = (r >> 8) ^ t[0][(r ^ simt::ldg8(p)) & 0xffu];
        p++;
        n--;
    }
    const uint32_t* w = (const uint32_t*)p;
    for (; n >= 4; n -= 4) {
        r ^= simt::ldg32(w++);
        r = t[3][r & 0xffu] ^ t[2][(r >> 8) & 0xffu] ^ t[1][(r >> 16) & 0xffu] ^ t[0][r >> 24];
    }
    p = (const uint8_t*)w;
    for (; n; n--) r = (r >> 8) ^ t[0][(r ^ simt::ldg8(p++)) & 0xffu];
    return ~r;
}

static const int CRC_WARPS = 8;

FDB_GLOBAL void FDB_LAUNCH_BOUNDS(CRC_WARPS * 32, 4) crc32_kernel(CrcBatch b, const CrcTables* tables, uint32_t* next) {
    FDB_SHARED CrcTables s;
    for (uint32_t i = threadIdx.x; i < sizeof(CrcTables) / 4; i += blockDim.x) ((uint32_t*)&s)[i] = ((const uint32_t*)tables)[i];
    simt::syncthreads();
    const unsigned lane = simt::lane_id();
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = simt::atomic_add(next, 1u);
        i = simt::shfl(i, 0);
        if (i >= b.n) break;
        const uint8_t* p = b.base + b.off[i];
        const uint64_t n = b.len[i];
        uint32_t result;
        if (n < 2048) {
            result = lane == 0 ? crc_block(s.t, p, n, b.seed) : 0u;
        } else {
            const uint64_t S = (n / 32) & ~(uint64_t)3;                  // block length of lanes 0..30
            const uint64_t my_len = lane < 31 ? S : n - 31 * S;
            const uint32_t mine = crc_block(s.t, p + lane * S, my_len, lane == 0 ? b.seed : 0u);
            const uint32_t xs = crc_xpow(s.x2n, S, 3), xl = crc_xpow(s.x2n, n - 31 * S, 3);  // x^(8 S), x^(8 * last)
            uint32_t acc = simt::shfl(mine, 0);
            for (unsigned l = 1; l < 32; l++) {
                const uint32_t c = simt::shfl(mine, l);
                acc = crc_mulmod(l < 31 ? xs : xl, acc) ^ c;
            }
            result = acc;
        }
        if (lane == 0) b.crc[i] = result;
    }
}

// ---- sum of a batch's stream lengths (input of the span / segment planners) ----
FDB_GLOBAL void batch_total_kernel(const uint64_t* len, uint32_t n, uint64_t* total) {
    uint64_t s = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += len[i];
    s = simt::reduce_add(s);
    if (simt::lane_id() == 0 && s) simt::atomic_add(total, s);
}

// ---- gather: byte ranges copied to new places (IDAT payloads of one file -> one contiguous zlib stream) ----
struct GatherItem {
    uint64_t src, dst, len;  // offsets into src_base / dst_base
};
FDB_GLOBAL void gather_kernel(const uint8_t* src_base, uint8_t* dst_base, const GatherItem* items, uint32_t n, uint32_t* next) {
    const unsigned lane = simt::lane_id();
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = simt::atomic_add(next, 1u);
        i = simt::shfl(i, 0);
        if (i >= n) break;
        const uint8_t* s = src_base + items[i].src;
        uint8_t* d = dst_base + items[i].dst;
        uint64_t len = items[i].len;
        if ((((uintptr_t)s ^ (uintptr_t)d) & 3u) == 0) {  // same alignment: whole words in the middle
            uint64_t head = (4u - (uint32_t)((uintptr_t)d & 3u)) & 3u;
            if (head > len) head = len;
            if (lane < head) d[lane] = simt::ldg8(s + lane);
            const uint64_t words = (len - head) >> 2;
            const uint32_t* sw = (const uint32_t*)(s + head);
            uint32_t* dw = (uint32_t*)(d + head);
            for (uint64_t k = lane; k < words; k += 32) dw[k] = simt::ldg32(sw + k);
            const uint64_t done = head + (words << 2);
            if (done + lane < len) d[done + lane] = simt::ldg8(s + done + lane);
        } else {
            for (uint64_t k = lane; k < len; k += 32) d[k] = simt::ldg8(s + k);
        }
    }
}

}  // namespace fdb
