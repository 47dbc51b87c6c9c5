// deflate_stored.cuh -- K5: "stored" (level 0) zlib streams, one CTA per stream.
//
// Replaces Compressor::new(w, 0, true) + one write_data(whole input) + finish() of the reference
// (src/compress/mod.rs:69-101, :126-156, :194-214, :234-268): header 78 01; k non-final blocks
// `00 ff ff 00 00` + 65535 bytes; then a final stored block `01 LEN ~LEN` + data, or -- when nothing
// is left (empty input or a length that is a multiple of 65535) -- the empty final fixed block
// `03 00`; adler32 big-endian.
#pragma once
#include "simt.h"
#include "fdb_common.h"
#include "adler.cuh"

namespace fdb {

static const int STORED_THREADS = 256;

FDB_DEVICE uint64_t stored_len(uint64_t n) {
    uint64_t full = n / 65535, rem = n % 65535;
    return 2 + full * (5 + 65535) + (rem ? 5 + rem : 2) + 4;
}

FDB_GLOBAL void FDB_LAUNCH_BOUNDS(STORED_THREADS, 1) deflate_stored_kernel(DeflateBatch b, uint32_t* next) {
    FDB_SHARED uint64_t red1[STORED_THREADS / 32], red2[STORED_THREADS / 32];
    FDB_SHARED uint32_t cur;
    const unsigned tid = threadIdx.x, lane = simt::lane_id(), warp = simt::warp_in_block();
    for (;;) {
        if (tid == 0) cur = simt::atomic_add(next, 1u);
        simt::syncthreads();
        const uint32_t i = cur;
        simt::syncthreads();
        if (i >= b.n) break;
        const uint8_t* in = b.in_base + b.in_off[i];
        uint8_t* out = b.out_base + b.out_off[i];
        const uint64_t n = b.in_len[i], cap = b.out_cap[i];
        const uint64_t total = stored_len(n);
        if (total > cap) {
            if (tid == 0) {
                b.out_len[i] = 0;
                b.status[i] = ST_OUTPUT_BUFFER_TOO_SMALL;
            }
            continue;
        }
        const uint64_t full = n / 65535, rem = n % 65535;
        // block headers
        for (uint64_t k = tid; k < full; k += STORED_THREADS) {
            uint8_t* h = out + 2 + k * 65540;
            h[0] = 0x00; h[1] = 0xff; h[2] = 0xff; h[3] = 0x00; h[4] = 0x00;
        }
        if (tid == 0) {
            out[0] = 0x78;
            out[1] = 0x01;
            uint8_t* h = out + 2 + full * 65540;
            if (rem) {
                h[0] = 0x01;
                h[1] = (uint8_t)rem; h[2] = (uint8_t)(rem >> 8);
                h[3] = (uint8_t)~rem; h[4] = (uint8_t)(~rem >> 8);
            } else {
                h[0] = 0x03; h[1] = 0x00;
            }
        }
        // payload: input byte p lands at 2 + 5 * (p / 65535 + 1) + p
        AdlerAcc ad = {0, 0};
        const bool aligned = ((uintptr_t)in & 15u) == 0;
        const uint64_t nvec = aligned ? (n >> 4) : 0;
        for (uint64_t v = tid; v < nvec; v += STORED_THREADS) {
            uint4 q = simt::ldg128((const uint4*)in + v);
            uint64_t p = v << 4;
            adler_add16(ad, q, p);
            adler_fold(ad);
            uint32_t w[4] = {q.x, q.y, q.z, q.w};
            uint64_t blk = p / 65535;
            uint64_t next_edge = (blk + 1) * 65535;
            uint64_t d = 2 + 5 * (blk + 1) + p;
#pragma unroll
            for (uint32_t j = 0; j < 16; j++) {
                if (p + j == next_edge) d += 5;
                out[d + j] = (uint8_t)(w[j >> 2] >> (8u * (j & 3u)));
            }
        }
        for (uint64_t p = (nvec << 4) + tid; p < n; p += STORED_THREADS) {
            uint8_t v = simt::ldg8(in + p);
            adler_add1(ad, v, p);
            out[2 + 5 * (p / 65535 + 1) + p] = v;
        }
        // block-wide adler reduction
        uint64_t s1 = simt::reduce_add(ad.s1);
        uint64_t s2 = simt::reduce_add(ad.s2 % ADLER_MOD);
        if (lane == 0) { red1[warp] = s1; red2[warp] = s2; }
        simt::syncthreads();
        if (tid == 0) {
            uint64_t t1 = 0, t2 = 0;
            for (int w = 0; w < STORED_THREADS / 32; w++) { t1 += red1[w]; t2 += red2[w]; }
            uint32_t s1m = (uint32_t)(t1 % ADLER_MOD), s2m = (uint32_t)(t2 % ADLER_MOD), nm = (uint32_t)(n % ADLER_MOD);
            uint32_t A = (1u + s1m) % ADLER_MOD;
            uint32_t B = (uint32_t)(((uint64_t)nm + (uint64_t)nm * s1m % ADLER_MOD + ADLER_MOD - s2m) % ADLER_MOD);
            uint8_t* t = out + total - 4;
            t[0] = (uint8_t)(B >> 8); t[1] = (uint8_t)B; t[2] = (uint8_t)(A >> 8); t[3] = (uint8_t)A;
            b.out_len[i] = total;
            b.status[i] = ST_OK;
        }
        simt::syncthreads();
    }
}

}  // namespace fdb
