// png_filter.cuh -- K8 / K9: the PNG row filters either side of the zlib path.
//
// Not part of image-rs/fdeflate (the `png` crate does this step around its fdeflate calls); the algorithm is
// the PNG specification's section 9: a filtered image is h rows of (1 filter-type byte + stride bytes), filter
// types 0..4 = None, Sub, Up, Average, Paeth, "left" = the byte bpp positions back, bytes outside the image
// are 0.  Checked against oracle/png_filter_oracle.c, which is pinned by golden PNG files from two independent
// encoders (tests/golden/png).
//
// K8 png_unfilter_kernel -- reconstruction is a recurrence (left, up, up-left), so one image is decoded by
// ONE WARP as a diagonal wavefront: with bpp bytes per pixel, lane (r, c) owns byte channel c of row y0 + r of
// a block of R = 32 / bpp rows, and row r runs one pixel behind row r - 1.  At step t lane (r, c) handles pixel
// x = t - r: `left` is the lane's own previous result, `up` is what lane (r - 1, c) produced one step earlier
// (one shuffle), `up-left` is the `up` of the step before -- the recurrence lives in registers; only the first
// row of a block reads its `up` row from memory (written by the block before).  All five predictors are
// computed branch-free and selected by the row's type.
//
// K9 png_filter_kernel -- the forward direction has no recurrence (every predictor reads the RAW image): one
// CTA per image, one warp per row.  mode 0..4 = that type on every row; mode 5 = per row the type with the
// smallest sum of |signed filtered byte| (PNG 12.8), lowest type on ties: two passes over the row, the first
// accumulates the five sums.
#pragma once
#include "simt.h"
#include "fdb_common.h"
#include <type_traits>

namespace fdb {

struct PngBatch {
    const uint8_t* in_base;
    const uint64_t* in_off;   // [n]
    uint8_t* out_base;
    const uint64_t* out_off;  // [n]
    const uint32_t* height;   // [n] rows
    const uint32_t* stride;   // [n] bytes per raw row
    const uint32_t* bpp;      // [n] bytes per complete pixel, 1..8
    int32_t* status;          // [n] ST_OK / ST_PNG_BAD_FILTER_TYPE / ST_PNG_BAD_GEOMETRY
    uint32_t n;
    uint32_t mode;            // K9 only
};

FDB_DEVICE uint32_t png_paeth(uint32_t a, uint32_t b, uint32_t c) {  // PNG 9.4
    const int32_t p = (int32_t)a + (int32_t)b - (int32_t)c;
    int32_t pa = p - (int32_t)a, pb = p - (int32_t)b, pc = p - (int32_t)c;
    pa = pa < 0 ? -pa : pa;
    pb = pb < 0 ? -pb : pb;
    pc = pc < 0 ? -pc : pc;
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
FDB_DEVICE uint32_t png_predict(uint32_t type, uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t avg = (a + b) >> 1, pae = png_paeth(a, b, c);
    uint32_t p = 0;
    p = type == 1 ? a : p;
    p = type == 2 ? b : p;
    p = type == 3 ? avg : p;
    p = type == 4 ? pae : p;
    return p;
}

// four byte channels at once (one RGBA pixel per 32-bit word)
FDB_DEVICE uint32_t png_add4(uint32_t x, uint32_t y) {  // per-byte x + y mod 256
    return ((x & 0x7f7f7f7fu) + (y & 0x7f7f7f7fu)) ^ ((x ^ y) & 0x80808080u);
}
FDB_DEVICE uint32_t png_sub4(uint32_t x, uint32_t y) {  // per-byte x - y mod 256
    return ((x | 0x80808080u) - (y & 0x7f7f7f7fu)) ^ ((x ^ ~y) & 0x80808080u);
}
FDB_DEVICE uint32_t png_abs4(uint32_t f) {  // per byte: f < 128 ? f : 256 - f   (128 stays 128)
    const uint32_t neg = (f >> 7) & 0x01010101u;  // 1 in the bytes that are "negative"
    const uint32_t mask = neg * 0xffu;            // 0xff in those bytes
    return png_add4(f ^ mask, neg);               // two's complement of the negative bytes
}
FDB_DEVICE uint32_t png_avg4(uint32_t a, uint32_t b) {  // per-byte floor((a + b) / 2)
    return (a & b) + (((a ^ b) & 0xfefefefeu) >> 1);
}
FDB_DEVICE uint32_t png_paeth4(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r = 0;
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
        const uint32_t sh = 8u * k;
        r |= png_paeth((a >> sh) & 0xffu, (b >> sh) & 0xffu, (c >> sh) & 0xffu) << sh;
    }
    return r;
}
// PAETH = false: no row of the warp's block uses the Paeth predictor (it is three quarters of the arithmetic)
template <bool PAETH>
FDB_DEVICE uint32_t png_predict4(uint32_t type, uint32_t a, uint32_t b, uint32_t c) {
    uint32_t p = 0;
    p = type == 1 ? a : p;
    p = type == 2 ? b : p;
    p = type == 3 ? png_avg4(a, b) : p;
    if (PAETH) p = type == 4 ? png_paeth4(a, b, c) : p;
    return p;
}

static const int PNG_UNFILTER_WARPS = 4;
static const uint32_t PNG_AHEAD = 8;  // steps whose memory reads are issued together
static const uint32_t PNG_ROW_WORDS = 34;                  // chunk buffer: 32 pixel words per row + 2 of padding
static const uint32_t PNG_BUF_WORDS = 32 * PNG_ROW_WORDS;  // one chunk of 32 rows

FDB_GLOBAL void FDB_LAUNCH_BOUNDS(PNG_UNFILTER_WARPS * 32, 6) png_unfilter_kernel(PngBatch b, uint32_t* next) {
    FDB_SHARED uint32_t chunk_buf[PNG_UNFILTER_WARPS][2 * PNG_BUF_WORDS];  // RGBA fast path: two chunks per warp
    const unsigned lane = simt::lane_id();
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = simt::atomic_add(next, 1u);
        i = simt::shfl(i, 0);
        if (i >= b.n) break;
        const uint32_t h = b.height[i], stride = b.stride[i], bpp = b.bpp[i];
        if (bpp < 1 || bpp > 8) {
            if (lane == 0) b.status[i] = ST_PNG_BAD_GEOMETRY;
            continue;
        }
        const uint8_t* in = b.in_base + b.in_off[i];
        uint8_t* out = b.out_base + b.out_off[i];
        if (bpp == 4 && (stride & 3u) == 0 && ((uintptr_t)out & 3u) == 0) {
            // RGBA8 rows: LANE = ROW of a block of 32 rows, one whole pixel (a 32-bit word) per lane and step; row r
            // runs one pixel behind row r - 1, `up` is one shuffle away, `left` and `up-left` are last step's values.
            // A lane walking its own row would touch 32 different lines per load / store (measured: bound by L1
            // wavefronts), so the rows travel through shared memory in CHUNKS of 32 pixels: a chunk is loaded with
            // one coalesced 128-byte read per row (the byte-misaligned filtered row becomes aligned pixel words on the
            // way: two aligned words and a funnel shift), unfiltered in place, and leaves with one coalesced
            // 128-byte store per row.  In phase p (steps 32p .. 32p+31) the rows are spread over chunks p-1 and p, so
            // two chunk buffers alternate.  Row stride 34 words: lane r reads word 34 r + (t - r) -> bank (r + t) % 32.
            uint32_t* const buf = &chunk_buf[simt::warp_in_block()][0];
            const simt::saddr buf_s = simt::smem_addr(buf);
            const uint32_t W = stride >> 2, NC = (W + 31) >> 5;
            const uint64_t pitch = 1ull + stride;
            int32_t st = ST_OK;
            for (uint32_t y0 = 0; y0 < h && st == ST_OK; y0 += 32) {
                const uint32_t y = y0 + lane;
                const bool row_on = y < h;
                const uint32_t type = row_on ? (uint32_t)simt::ldg8(in + y * pitch) : 0u;
                if (simt::any(type > 4u)) {
                    st = ST_PNG_BAD_FILTER_TYPE;
                    break;
                }
                const uint32_t rows = h - y0 < 32u ? h - y0 : 32u;
                const uint32_t* arow = (const uint32_t*)(out + (uint64_t)(y0 - 1) * stride);  // (lane 0, y0 > 0)
                auto flush_chunk = [&](uint32_t c) {
                    const uint32_t x = 32u * c + lane;
                    for (uint32_t r = 0; r < rows; r++) {
                        const uint32_t v = simt::lds32(buf_s + 4u * ((c & 1u) * PNG_BUF_WORDS + r * PNG_ROW_WORDS + lane));
                        if (x < W) ((uint32_t*)(out + (uint64_t)(y0 + r) * stride))[x] = v;
                    }
                };
                auto load_chunk = [&](uint32_t c) {
                    const uint32_t x = 32u * c + lane;
#pragma unroll 4
                    for (uint32_t r = 0; r < rows; r++) {
                        const uint8_t* srow = in + (uint64_t)(y0 + r) * pitch + 1;
                        const uint32_t m = (uint32_t)((uintptr_t)srow & 3u);
                        const uint32_t* sw = (const uint32_t*)(srow - m);
                        // word x holds the first bytes of pixel x; the last pixel of a row also needs word W unless aligned
                        const uint32_t lo = x < W ? simt::ldg32(sw + x) : 0u;
                        const uint32_t hi = (x < W && m != 0) ? simt::ldg32(sw + x + 1) : 0u;
                        simt::sts32(buf_s + 4u * ((c & 1u) * PNG_BUF_WORDS + r * PNG_ROW_WORDS + lane), simt::funnel_r(lo, hi, 8u * m));
                    }
                };
                auto run_block = [&](auto paeth_tag) {
                    constexpr bool PAETH = decltype(paeth_tag)::value;
                    uint32_t cur = 0, up = 0;
                    for (uint32_t p = 0; p <= NC; p++) {
                        simt::syncwarp();
                        if (p >= 2) flush_chunk(p - 2);
                        if (p < NC) load_chunk(p);
                        simt::syncwarp();
                        for (uint32_t s0 = 0; s0 < 32; s0 += PNG_AHEAD) {
                            uint32_t av[PNG_AHEAD];
    #pragma unroll
                            for (uint32_t k = 0; k < PNG_AHEAD; k++) {
                                const uint32_t x = 32u * p + s0 + k;  // (lane 0: x = t)
                                av[k] = (lane == 0 && y0 > 0 && x < W) ? arow[x] : 0u;  // (stored by this warp one block ago)
                            }
    #pragma unroll
                            for (uint32_t k = 0; k < PNG_AHEAD; k++) {
                                const uint32_t x = 32u * p + s0 + k - lane;  // (wraps for t < lane: then x >= W)
                                const bool on = row_on && x < W;
                                const simt::saddr cell = buf_s + 4u * (((x >> 5) & 1u) * PNG_BUF_WORDS + lane * PNG_ROW_WORDS + (x & 31u));
                                const uint32_t f = on ? simt::lds32(cell) : 0u;
                                const uint32_t from_lane = simt::shfl_up(cur, 1);
                                const uint32_t upleft = up;
                                const uint32_t upv = lane == 0 ? av[k] : from_lane;
                                up = on ? upv : 0u;
                                const uint32_t v = png_add4(f, png_predict4<PAETH>(type, cur, up, upleft));
                                cur = on ? v : 0u;
                                if (on) simt::sts32(cell, v);
                            }
                        }
                    }
                };
                if (simt::any(type == 4u))
                    run_block(std::true_type{});
                else
                    run_block(std::false_type{});
                simt::syncwarp();
                if (NC >= 1) flush_chunk(NC - 1);
                simt::syncwarp();  // the next block's first row reads this block's last row
            }
            if (lane == 0) b.status[i] = st;
            continue;
        }
        const uint32_t R = 32u / bpp;              // rows per block
        const uint32_t r = lane / bpp, c = lane - r * bpp;
        const bool lane_on = r < R;
        const uint32_t W = (stride + bpp - 1) / bpp;  // pixel steps per row
        const uint64_t in_pitch = 1ull + stride;
        int32_t st = ST_OK;
        for (uint32_t y0 = 0; y0 < h && st == ST_OK; y0 += R) {
            const uint32_t y = y0 + r;
            const bool row_on = lane_on && y < h;
            const uint32_t type = row_on ? (uint32_t)simt::ldg8(in + y * in_pitch) : 0u;
            if (simt::any(type > 4u)) {
                st = ST_PNG_BAD_FILTER_TYPE;
                break;
            }
            const uint8_t* src = in + y * in_pitch + 1 + c;
            uint8_t* dst = out + (uint64_t)y * stride + c;
            const uint8_t* above = out + (uint64_t)(y0 - 1) * stride + c;  // (row 0 of the block only, y0 > 0)
            uint32_t cur = 0, up = 0;
            const uint32_t steps = W + R - 1;
            // The bytes a step reads from memory (its filtered byte, and for the block's first row the byte above) do
            // not depend on the recurrence, so they are fetched PNG_AHEAD steps at a time before the dependent chain of
            // those steps runs: the loads' latency overlaps instead of adding up step by step.
            for (uint32_t t0 = 0; t0 < steps; t0 += PNG_AHEAD) {
                uint32_t fv[PNG_AHEAD], av[PNG_AHEAD];
#pragma unroll
                for (uint32_t k = 0; k < PNG_AHEAD; k++) {
                    const uint32_t x = t0 + k - r;  // (wraps for t < r: then x >= W)
                    const uint32_t xb = x * bpp;
                    const bool on = row_on && x < W && xb + c < stride;
                    fv[k] = on ? (uint32_t)simt::ldg8(src + xb) : 0u;
                    av[k] = (on && r == 0 && y0 > 0) ? (uint32_t)above[xb] : 0u;  // (stored by this warp one block ago)
                }
#pragma unroll
                for (uint32_t k = 0; k < PNG_AHEAD; k++) {
                    const uint32_t x = t0 + k - r;
                    const uint32_t xb = x * bpp;
                    const bool on = row_on && x < W && xb + c < stride;
                    const uint32_t from_lane = simt::shfl_up(cur, bpp);  // row r - 1, same pixel, produced one step ago
                    const uint32_t upleft = up;
                    const uint32_t upv = r == 0 ? av[k] : from_lane;
                    up = on ? upv : 0u;
                    const uint32_t v = (fv[k] + png_predict(type, cur, up, upleft)) & 0xffu;
                    cur = on ? v : 0u;
                    if (on) dst[xb] = (uint8_t)v;
                }
            }
            simt::syncwarp();  // the next block's first row reads this block's last row
        }
        if (lane == 0) b.status[i] = st;
    }
}

static const int PNG_FILTER_WARPS = 8;

FDB_GLOBAL void FDB_LAUNCH_BOUNDS(PNG_FILTER_WARPS * 32, 4) png_filter_kernel(PngBatch b, uint32_t* next) {
    FDB_SHARED uint32_t cur_img;
    const unsigned lane = simt::lane_id(), warp = simt::warp_in_block();
    for (;;) {
        if (threadIdx.x == 0) cur_img = simt::atomic_add(next, 1u);
        simt::syncthreads();
        const uint32_t i = cur_img;
        simt::syncthreads();
        if (i >= b.n) break;
        const uint32_t h = b.height[i], stride = b.stride[i], bpp = b.bpp[i];
        if (bpp < 1 || bpp > 8 || b.mode > 5) {
            if (threadIdx.x == 0) b.status[i] = ST_PNG_BAD_GEOMETRY;
            continue;
        }
        const uint8_t* in = b.in_base + b.in_off[i];
        uint8_t* out = b.out_base + b.out_off[i];
        for (uint32_t y = warp; y < h; y += PNG_FILTER_WARPS) {
            const uint8_t* cur = in + (uint64_t)y * stride;
            const uint8_t* up = cur - stride;  // (y > 0 only)
            uint8_t* dst = out + (uint64_t)y * (1ull + stride);
            uint32_t type = b.mode;
            const bool words = bpp == 4 && (stride & 3u) == 0 && ((uintptr_t)in & 3u) == 0;  // RGBA rows, aligned words
            if (b.mode == 5) {
                uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0;
                if (words) {
                    const uint32_t* cw = (const uint32_t*)cur;
                    const uint32_t* uw = (const uint32_t*)up;
                    for (uint32_t p = lane; p < (stride >> 2); p += 32) {
                        const uint32_t v = simt::ldg32(cw + p);
                        const uint32_t a = p ? simt::ldg32(cw + p - 1) : 0u;
                        const uint32_t bb = y ? simt::ldg32(uw + p) : 0u;
                        const uint32_t cc = (y && p) ? simt::ldg32(uw + p - 1) : 0u;
                        // |signed byte| summed over the four channels: per-byte abs, then a dot product with ones
                        s0 = simt::dp4a_u(png_abs4(v), 0x01010101u, s0);
                        s1 = simt::dp4a_u(png_abs4(png_sub4(v, a)), 0x01010101u, s1);
                        s2 = simt::dp4a_u(png_abs4(png_sub4(v, bb)), 0x01010101u, s2);
                        s3 = simt::dp4a_u(png_abs4(png_sub4(v, png_avg4(a, bb))), 0x01010101u, s3);
                        s4 = simt::dp4a_u(png_abs4(png_sub4(v, png_paeth4(a, bb, cc))), 0x01010101u, s4);
                    }
                }
                for (uint32_t x = words ? stride : lane; x < stride; x += 32) {
                    const uint32_t v = simt::ldg8(cur + x);
                    const uint32_t a = x >= bpp ? (uint32_t)simt::ldg8(cur + x - bpp) : 0u;
                    const uint32_t bb = y ? (uint32_t)simt::ldg8(up + x) : 0u;
                    const uint32_t cc = (y && x >= bpp) ? (uint32_t)simt::ldg8(up + x - bpp) : 0u;
                    const uint32_t f0 = v, f1 = (v - a) & 0xffu, f2 = (v - bb) & 0xffu, f3 = (v - ((a + bb) >> 1)) & 0xffu,
                                   f4 = (v - png_paeth(a, bb, cc)) & 0xffu;
                    s0 += f0 < 128u ? f0 : 256u - f0;
                    s1 += f1 < 128u ? f1 : 256u - f1;
                    s2 += f2 < 128u ? f2 : 256u - f2;
                    s3 += f3 < 128u ? f3 : 256u - f3;
                    s4 += f4 < 128u ? f4 : 256u - f4;
                }
                s0 = simt::reduce_add(s0);
                s1 = simt::reduce_add(s1);
                s2 = simt::reduce_add(s2);
                s3 = simt::reduce_add(s3);
                s4 = simt::reduce_add(s4);
                uint32_t best = s0;
                type = 0;
                if (s1 < best) { best = s1; type = 1; }
                if (s2 < best) { best = s2; type = 2; }
                if (s3 < best) { best = s3; type = 3; }
                if (s4 < best) { best = s4; type = 4; }
            }
            if (lane == 0) dst[0] = (uint8_t)type;
            if (words) {
                // four bytes (one RGBA pixel) per lane from aligned words: the pixel, the one to its left, the one
                // above and the one above-left
                const uint32_t* cw = (const uint32_t*)cur;
                const uint32_t* uw = (const uint32_t*)up;
                for (uint32_t p = lane; p < (stride >> 2); p += 32) {
                    const uint32_t v = simt::ldg32(cw + p);
                    const uint32_t a = p ? simt::ldg32(cw + p - 1) : 0u;
                    const uint32_t bb = y ? simt::ldg32(uw + p) : 0u;
                    const uint32_t cc = (y && p) ? simt::ldg32(uw + p - 1) : 0u;
                    uint8_t* d = dst + 1 + 4 * p;
                    // (the row's type is uniform over the warp: only its predictor is evaluated)
                    const uint32_t pred = type == 4 ? png_paeth4(a, bb, cc) : png_predict4<false>(type, a, bb, cc);
                    const uint32_t f = png_sub4(v, pred);
#pragma unroll
                    for (uint32_t k = 0; k < 4; k++) d[k] = (uint8_t)(f >> (8u * k));
                }
                continue;
            }
            for (uint32_t x = lane; x < stride; x += 32) {
                const uint32_t v = simt::ldg8(cur + x);
                const uint32_t a = x >= bpp ? (uint32_t)simt::ldg8(cur + x - bpp) : 0u;
                const uint32_t bb = y ? (uint32_t)simt::ldg8(up + x) : 0u;
                const uint32_t cc = (y && x >= bpp) ? (uint32_t)simt::ldg8(up + x - bpp) : 0u;
                dst[1 + x] = (uint8_t)(v - png_predict(type, a, bb, cc));
            }
        }
        if (threadIdx.x == 0) b.status[i] = ST_OK;
    }
}

}  // namespace fdb
