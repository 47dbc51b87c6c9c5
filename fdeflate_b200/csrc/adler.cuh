// adler.cuh -- Adler-32 (RFC 1950) as position-weighted sums, so that any partition of the bytes
// over lanes / tiles reduces with plain additions:
//     S1 = sum b_k            S2 = sum k * b_k           (k = 0-based position in the stream)
//     A  = (1 + S1) mod 65521
//     B  = (n + n*S1 - S2) mod 65521
// Replaces the reference's `simd_adler32::Adler32::{write,finish}` calls
// (reference src/decompress.rs:311,318,332; src/compress/ultrafast.rs:95,176).
#pragma once
#include "simt.h"
#include "fdb_common.h"

namespace fdb {

struct AdlerAcc {
    uint64_t s1;
    uint64_t s2;
};

// byte sums of one 16-byte vector that sits at stream position `pos`
FDB_DEVICE void adler_add16(AdlerAcc& a, uint4 q, uint64_t pos) {
    uint32_t t1 = simt::dp4a_u(q.x, 0x01010101u, 0);
    t1 = simt::dp4a_u(q.y, 0x01010101u, t1);
    t1 = simt::dp4a_u(q.z, 0x01010101u, t1);
    t1 = simt::dp4a_u(q.w, 0x01010101u, t1);
    uint32_t t2 = simt::dp4a_u(q.x, 0x03020100u, 0);
    t2 = simt::dp4a_u(q.y, 0x07060504u, t2);
    t2 = simt::dp4a_u(q.z, 0x0b0a0908u, t2);
    t2 = simt::dp4a_u(q.w, 0x0f0e0d0cu, t2);
    a.s1 += t1;
    a.s2 += pos * (uint64_t)t1 + t2;
}
// s2 grows by < 2^12 * pos per vector; callers that stream more than 2^40 weighted bytes through one
// accumulator call this now and then (once per segment / warp step is plenty)
FDB_DEVICE void adler_fold(AdlerAcc& a) {
    if (a.s2 >> 60) a.s2 %= ADLER_MOD;
}

FDB_DEVICE void adler_add1(AdlerAcc& a, uint32_t byte, uint64_t pos) {
    a.s1 += byte;
    a.s2 += pos * (uint64_t)byte;
    if (a.s2 >> 62) a.s2 %= ADLER_MOD;
}

// warp-wide finish: every lane passes its partial sums, every lane gets the checksum of n bytes
FDB_DEVICE uint32_t adler_finish_warp(AdlerAcc a, uint64_t n) {
    uint64_t s1 = simt::reduce_add(a.s1);
    uint64_t s2 = simt::reduce_add(a.s2 % ADLER_MOD);
    uint32_t s1m = (uint32_t)(s1 % ADLER_MOD);
    uint32_t s2m = (uint32_t)(s2 % ADLER_MOD);
    uint32_t nm = (uint32_t)(n % ADLER_MOD);
    uint32_t A = (1u + s1m) % ADLER_MOD;
    uint32_t B = (uint32_t)(((uint64_t)nm + (uint64_t)nm * s1m % ADLER_MOD + ADLER_MOD - s2m) % ADLER_MOD);
    return (B << 16) | A;
}

// Adler-32 of buf[0..n) read back by the whole warp with 16-byte loads.  buf was written by this
// kernel, so plain (coherent) loads are used, never the read-only path.
FDB_DEVICE uint32_t warp_adler32(const uint8_t* buf, uint64_t n) {
    const unsigned lane = simt::lane_id();
    AdlerAcc a = {0, 0};
    uint64_t head = (16u - (uint32_t)((uintptr_t)buf & 15u)) & 15u;
    if (head > n) head = n;
    if (lane < head) adler_add1(a, buf[lane], lane);
    uint64_t nvec = (n - head) >> 4;
    const uint4* v = (const uint4*)(buf + head);
    for (uint64_t i = lane; i < nvec; i += 32) {
        adler_add16(a, v[i], head + (i << 4));
        adler_fold(a);
    }
    uint64_t done = head + (nvec << 4);
    if (done + lane < n) adler_add1(a, buf[done + lane], done + lane);
    return adler_finish_warp(a, n);
}

}  // namespace fdb
