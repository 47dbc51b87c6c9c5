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
// The payload travels through shared memory: one thread queues bulk asynchronous copies (cp.async.bulk, completion
// on an mbarrier) of the next chunks of the input while all threads realign the chunk that has landed and stream it
// out.  Round 1 loaded every vector twice through the load/store unit (lo / hi of the funnel shift) and had one
// 16-vector iteration of loads in flight per thread: 0.70 of the HBM roofline, long-scoreboard stalls 28 per issue.
static const uint32_t STORED_STAGES = 4;
static const uint32_t STORED_CHUNK_VECS = 1024;                        // 16 KiB of payload per stage
static const uint32_t STORED_STAGE_BYTES = 16 * STORED_CHUNK_VECS + 32;  // + the vector the last funnel shift reaches into
struct StoredSmem {
    unsigned char stage[STORED_STAGES][STORED_STAGE_BYTES];
    uint64_t bar[STORED_STAGES];
};

FDB_DEVICE uint64_t stored_len(uint64_t n) {
    uint64_t full = n / 65535, rem = n % 65535;
    return 2 + full * (5 + 65535) + (rem ? 5 + rem : 2) + 4;
}

// Block k of a stream: its 16-byte vectors in DESTINATION alignment (the bytes before the first and after the last
// one are copied byte-wise), and the aligned source vectors they are cut from.
struct StoredBlock {
    const uint8_t* src;   // first payload byte of the block
    uint8_t* dst;         // where it goes
    uint32_t len, head, nvec, sh;
    const uint8_t* a0;    // 16-byte aligned source address of vector 0's `lo`
};
FDB_DEVICE StoredBlock stored_block(const uint8_t* in, uint8_t* out, uint64_t k, uint64_t full, uint64_t rem) {
    StoredBlock b;
    b.len = k < full ? 65535u : (uint32_t)rem;
    b.src = in + k * 65535;
    b.dst = out + 2 + k * 65540 + 5;
    b.head = (16u - (uint32_t)((uintptr_t)b.dst & 15u)) & 15u;
    if (b.head > b.len) b.head = b.len;
    b.nvec = (b.len - b.head) >> 4;
    const uint8_t* sh_src = b.src + b.head;
    b.sh = (uint32_t)((uintptr_t)sh_src & 15u);
    b.a0 = sh_src - b.sh;
    return b;
}

FDB_GLOBAL void FDB_LAUNCH_BOUNDS(STORED_THREADS, 3) deflate_stored_kernel(DeflateBatch b, uint32_t* next) {
    FDB_DYN_SMEM(smem_raw);
    StoredSmem& sm = *reinterpret_cast<StoredSmem*>(smem_raw);
    FDB_SHARED uint64_t red1[STORED_THREADS / 32], red2[STORED_THREADS / 32];
    FDB_SHARED uint32_t cur;
    const unsigned tid = threadIdx.x, lane = simt::lane_id(), warp = simt::warp_in_block();
    if (tid == 0) {
        for (uint32_t s = 0; s < STORED_STAGES; s++) simt::mbar_init(simt::smem_addr(&sm.bar[s]), 1);
        simt::mbar_fence_init();
    }
    simt::syncthreads();
    uint32_t uses = 0;  // chunks this CTA has pushed through its stages so far (stage = uses % STAGES, parity = uses / STAGES)
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
        // head bytes up to the first 16-byte boundary of the DESTINATION, then whole vectors (two aligned source
        // vectors funnel-shifted into one aligned 16-byte streaming store), then the tail bytes.  The vectors of all
        // blocks of the stream form one sequence of chunks that the copy engine keeps STAGES - 1 ahead of the threads.
        // adler32 rides along on the bytes in registers.
        AdlerAcc ad = {0, 0};
        const uint64_t nblocks = full + (rem ? 1 : 0);
        // producer cursor (used by thread 0 only): the next chunk to queue
        uint64_t pk = 0;
        uint32_t pc = 0, queued = 0;
        auto queue_next = [&]() {  // thread 0: start the copy of the next chunk, if the stream has one
            while (pk < nblocks) {
                const StoredBlock bl = stored_block(in, out, pk, full, rem);
                if (pc * STORED_CHUNK_VECS >= bl.nvec) {
                    pk++;
                    pc = 0;
                    continue;
                }
                const uint32_t cnt = bl.nvec - pc * STORED_CHUNK_VECS < STORED_CHUNK_VECS ? bl.nvec - pc * STORED_CHUNK_VECS : STORED_CHUNK_VECS;
                const uint32_t bytes = 16u * (cnt + (bl.sh ? 1u : 0u));
                const uint32_t st = (uses + queued) % STORED_STAGES;
                const simt::saddr bar = simt::smem_addr(&sm.bar[st]);
                simt::mbar_expect_tx(bar, bytes);
                simt::bulk_g2s(simt::smem_addr(sm.stage[st]), bl.a0 + 16ull * pc * STORED_CHUNK_VECS, bytes, bar);
                queued++;
                pc++;
                return;
            }
        };
        if (tid == 0)
            for (uint32_t s = 0; s + 1 < STORED_STAGES; s++) queue_next();
        uint32_t done = 0;  // chunks of this stream consumed so far (same in every thread)
        for (uint64_t k = 0; k < nblocks; k++) {
            const uint64_t p0 = k * 65535;
            const StoredBlock bl = stored_block(in, out, k, full, rem);
            const uint32_t tail0 = bl.head + (bl.nvec << 4);
            if (tid < bl.head) {
                const uint8_t v = simt::ldg8(bl.src + tid);
                adler_add1(ad, v, p0 + tid);
                bl.dst[tid] = v;
            }
            if (tid >= 32 && tid - 32 < bl.len - tail0) {  // (another warp than the head's)
                const uint32_t q = tail0 + (tid - 32);
                const uint8_t v = simt::ldg8(bl.src + q);
                adler_add1(ad, v, p0 + q);
                bl.dst[q] = v;
            }
            uint4* d0 = (uint4*)(bl.dst + bl.head);
            const uint32_t s4 = bl.sh >> 2, sb = 8u * (bl.sh & 3u);
            for (uint32_t c0 = 0; c0 < bl.nvec; c0 += STORED_CHUNK_VECS) {
                if (tid == 0) queue_next();  // into the stage the chunk before this one has just left
                const uint32_t st = (uses + done) % STORED_STAGES;
                simt::mbar_wait(simt::smem_addr(&sm.bar[st]), ((uses + done) / STORED_STAGES) & 1u);
                const simt::saddr sbase = simt::smem_addr(sm.stage[st]);
                const uint32_t cnt = bl.nvec - c0 < STORED_CHUNK_VECS ? bl.nvec - c0 : STORED_CHUNK_VECS;
#pragma unroll 2
                for (uint32_t v = tid; v < cnt; v += STORED_THREADS) {
                    const uint4 lo = simt::lds128(sbase + 16u * v);
                    uint4 r = lo;
                    if (bl.sh) {
                        const uint4 hi = simt::lds128(sbase + 16u * v + 16u);
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
                    adler_add16(ad, r, p0 + bl.head + ((uint64_t)(c0 + v) << 4));
                    simt::stcs128(d0 + c0 + v, r);
                }
                done++;
                simt::syncthreads();  // every thread is done with this stage: it may be refilled
            }
            adler_fold(ad);
        }
        uses += done;
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
