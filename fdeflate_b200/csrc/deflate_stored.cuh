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
        // payload: block k's bytes move from in + 65535 k to out + 2 + 65540 k + 5.  Source and destination are
        // misaligned against each other by a different amount in every block, so each block is copied as: a few
        // head bytes up to the first 16-byte boundary of the DESTINATION, then whole vectors (two aligned
        // 16-byte loads funnel-shifted into one aligned 16-byte streaming store), then the tail bytes.
        // adler32 rides along on the bytes in registers.
        AdlerAcc ad = {0, 0};
        const uint64_t nblocks = full + (rem ? 1 : 0);
        for (uint64_t k = 0; k < nblocks; k++) {
            const uint64_t p0 = k * 65535;
            const uint32_t len = k < full ? 65535u : (uint32_t)rem;
            const uint8_t* src = in + p0;
            uint8_t* dst = out + 2 + k * 65540 + 5;
            uint32_t head = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;
            if (head > len) head = len;
            const uint32_t nvec = (len - head) >> 4;
            const uint32_t tail0 = head + (nvec << 4);
            if (tid < head) {
                const uint8_t v = simt::ldg8(src + tid);
                adler_add1(ad, v, p0 + tid);
                dst[tid] = v;
            }
            if (tid >= 32 && tid - 32 < len - tail0) {  // (another warp than the head's)
                const uint32_t q = tail0 + (tid - 32);
                const uint8_t v = simt::ldg8(src + q);
                adler_add1(ad, v, p0 + q);
                dst[q] = v;
            }
            const uint8_t* sh_src = src + head;
            const uint32_t sh = (uint32_t)((uintptr_t)sh_src & 15u);
            const uint4* a0 = (const uint4*)(sh_src - sh);
            uint4* d0 = (uint4*)(dst + head);
            const uint32_t s4 = sh >> 2, sb = 8u * (sh & 3u);
            for (uint32_t v = tid; v < nvec; v += STORED_THREADS) {
                const uint4 lo = simt::ldg128(a0 + v);
                uint4 r = lo;
                if (sh) {
                    const uint4 hi = simt::ldg128(a0 + v + 1);
                    switch (s4) {  // (uniform over the block)
                        case 0:
                            r = make_uint4(simt::funnel_r(lo.x, lo.y, sb), simt::funnel_r(lo.y, lo.z, sb),
                                           simt::funnel_r(lo.z, lo.w, sb), simt::funnel_r(lo.w, hi.x, sb));
                            break;
                        case 1:
                            r = make_uint4(simt::funnel_r(lo.y, lo.z, sb), simt::funnel_r(lo.z, lo.w, sb),
                                           simt::funnel_r(lo.w, hi.x, sb), simt::funnel_r(hi.x, hi.y, sb));
                            break;
                        case 2:
                            r = make_uint4(simt::funnel_r(lo.z, lo.w, sb), simt::funnel_r(lo.w, hi.x, sb),
                                           simt::funnel_r(hi.x, hi.y, sb), simt::funnel_r(hi.y, hi.z, sb));
                            break;
                        default:
                            r = make_uint4(simt::funnel_r(lo.w, hi.x, sb), simt::funnel_r(hi.x, hi.y, sb),
                                           simt::funnel_r(hi.y, hi.z, sb), simt::funnel_r(hi.z, hi.w, sb));
                            break;
                    }
                }
                adler_add16(ad, r, p0 + head + ((uint64_t)v << 4));
                simt::stcs128(d0 + v, r);
            }
            adler_fold(ad);
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
