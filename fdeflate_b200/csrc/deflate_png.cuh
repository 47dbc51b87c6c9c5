// deflate_png.cuh -- the PNG row filter FUSED into the ultra-fast encoder's prologue.
//
// png_filter_kernel (K9) writes the filtered image -- h rows of (type byte + stride bytes) -- to device memory and the
// encoder reads it back: one extra write and one extra read of the whole image through HBM.  The forward filter has no
// recurrence (every predictor reads RAW pixels: left, up, up-left), so the encoder can compute the filtered bytes where
// it stages them: deflate_ufb_stream takes its input through a source object, and UbPngSrc below returns bytes
// [g, g + 16) of the VIRTUAL filtered stream of an image, computed from the raw rows (read through L1 / L2; the row
// above was read a row earlier by the same warp).  Four byte channels at a time where a 32-bit word lies inside one
// row's pixels and right of its first pixel (png_*4 of png_filter.cuh); byte by byte for the few words per row that
// hold the type byte, wrap to the next row, or touch the first pixel (whose "left" is outside the image).
// Same bytes as filter-then-deflate for filter modes 0..4 (one type on every row); the adaptive mode 5 needs a pass
// over each row before its type is known and keeps the two-kernel path.
#pragma once
#include "deflate_ufb.cuh"
#include "png_filter.cuh"

namespace fdb {

struct UbPngSrc {
    const uint8_t* raw;  // h rows of `stride` bytes
    uint32_t stride, L /* = stride + 1 */, bpp, type, height;
    uint32_t n32;        // height * L (< 2^32: the host takes this path only for such images)

    // one byte: position p (0 = the type byte) of row y
    FDB_MEMBER uint32_t at(uint32_t y, uint32_t p) const {
        if (p == 0) return type;
        const uint32_t x = p - 1u;
        const uint8_t* r = raw + (size_t)y * stride + x;
        const uint32_t v = simt::ldg8(r);
        const uint32_t a = x >= bpp ? (uint32_t)simt::ldg8(r - bpp) : 0u;
        const uint32_t b = y ? (uint32_t)simt::ldg8(r - stride) : 0u;
        const uint32_t c = (y && x >= bpp) ? (uint32_t)simt::ldg8(r - stride - bpp) : 0u;
        return (v - png_predict(type, a, b, c)) & 0xffu;
    }
    FDB_MEMBER uint32_t byte(uint64_t g) const {
        if (g >= n32) return 0u;
        const uint32_t y = (uint32_t)g / L;
        return at(y, (uint32_t)g - y * L);
    }
    FDB_MEMBER void prefetch(uint64_t g) const {
        const uint64_t r = g - g / L;  // about where the raw bytes of stream position g are
        if (r < (uint64_t)height * stride) simt::prefetch_l2(raw + r);
    }
    // four bytes at any address (two aligned words and a funnel shift; one word when the address is aligned)
    FDB_MEMBER uint32_t ldu32(const uint8_t* p) const {
        const uint32_t sh = 8u * (uint32_t)((uintptr_t)p & 3u);
        const uint32_t* w = (const uint32_t*)((uintptr_t)p & ~(uintptr_t)3);
        const uint32_t lo = simt::ldg32(w);
        const uint32_t hi = sh ? simt::ldg32(w + 1) : 0u;
        return simt::funnel_r(lo, hi, sh);
    }
    FDB_MEMBER uint4 load16(uint64_t g) const {
        uint32_t w[4];
        uint32_t y = (uint32_t)g / L, p = (uint32_t)g - y * L;
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            if (y < height && p >= 1u + bpp && p + 4u <= L) {
                // the word lies in row y's pixels, right of the first pixel
                const uint8_t* r = raw + (size_t)y * stride + (p - 1u);
                const uint32_t cur = ldu32(r);
                uint32_t pred = 0;
                if (type == 1) {
                    pred = ldu32(r - bpp);
                } else if (type == 2) {
                    pred = y ? ldu32(r - stride) : 0u;
                } else if (type == 3) {
                    pred = png_avg4(ldu32(r - bpp), y ? ldu32(r - stride) : 0u);
                } else if (type == 4) {
                    const uint32_t a = ldu32(r - bpp);
                    pred = y ? png_paeth4(a, ldu32(r - stride), ldu32(r - stride - bpp)) : a;  // paeth(a, 0, 0) = a
                }
                w[k] = png_sub4(cur, pred);
            } else {
                uint32_t v = 0, yy = y, pp = p;
                for (uint32_t j = 0; j < 4; j++) {
                    while (pp >= L) {
                        pp -= L;
                        yy++;
                    }
                    if (yy < height) v |= at(yy, pp) << (8u * j);
                    pp++;
                }
                w[k] = v;
            }
            p += 4u;
            while (p >= L) {
                p -= L;
                y++;
            }
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
};

// deflate_ufb_kernel over raw images: b.in_* describe the RAW images (in_len is not used), one filter type for all rows.
// fstatus[i] = ST_OK / ST_PNG_BAD_GEOMETRY (as png_filter_kernel reports it); a rejected image produces no stream.
FDB_GLOBAL void FDB_LAUNCH_BOUNDS(UB_WARPS * 32, UB_MIN_CTAS)
    deflate_ufb_png_kernel(DeflateBatch b, const uint32_t* height, const uint32_t* stride, const uint32_t* bpp, uint32_t mode,
                           int32_t* fstatus, const UfEncTables* tables, uint32_t* next) {
    FDB_DYN_SMEM(smem_raw);
    UbSmem& s = *reinterpret_cast<UbSmem*>(smem_raw);
    for (uint32_t i = threadIdx.x; i < 512; i += blockDim.x) s.lit[i] = tables->lit[i];
    for (uint32_t i = threadIdx.x; i < 258; i += blockDim.x) s.tail_tok[i] = tables->tail_tok[i];
    for (uint32_t i = threadIdx.x; i < 14; i += blockDim.x) s.header[i] = tables->header[i];
    simt::syncthreads();
    const unsigned lane = simt::lane_id();
    UbWarp& ws = s.warp[simt::warp_in_block()];
    const uint32_t slots = gridDim.x * UB_WARPS;
    bool first = true;
    for (;;) {
        uint32_t i = blockIdx.x + gridDim.x * simt::warp_in_block();
        if (!first) {
            if (lane == 0) i = slots + simt::atomic_add(next, 1u);
            i = simt::shfl(i, 0);
        }
        first = false;
        if (i >= b.n) break;
        const uint32_t h = height[i], sd = stride[i], bp = bpp[i];
        const uint64_t n = (uint64_t)h * (1ull + sd);
        if (bp < 1 || bp > 8 || mode > 4 || n > 0xffffffffull) {
            if (lane == 0) {
                fstatus[i] = ST_PNG_BAD_GEOMETRY;
                b.out_len[i] = 0;
                b.status[i] = ST_OK;
            }
            continue;
        }
        const UbPngSrc src = {b.in_base + b.in_off[i], sd, sd + 1u, bp, mode, h, (uint32_t)n};
        int32_t st = ST_OK;
        uint64_t len = deflate_ufb_stream(s.lit, s.tail_tok, s.header, ws, src, n, b.out_base + b.out_off[i], b.out_cap[i], &st);
        if (lane == 0) {
            fstatus[i] = ST_OK;
            b.out_len[i] = len;
            b.status[i] = st;
        }
    }
}

}  // namespace fdb
