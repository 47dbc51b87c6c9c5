/*
 * png_filter_oracle.c -- CPU ORACLE for the PNG row filters (the step either side of
 * the zlib path in a PNG codec).  TEST INFRASTRUCTURE ONLY, like fdeflate_oracle.c.
 *
 * This step is NOT part of image-rs/fdeflate (it lives in the `png` crate, which is not under
 * /root/reference), so there is no reference file:line to follow; the algorithm is the published one,
 * PNG (Third Edition) / ISO/IEC 15948 section 9 "Filtering": 9.2 filter types 0..4 (None, Sub, Up, Average,
 * Paeth), 9.3 "filter byte x uses the byte bpp positions to its left, bytes outside the image are 0",
 * 9.4 the Paeth predictor.  Parity pin: tests/test_png.py decodes golden PNG files written by two independent
 * encoders (Pillow, OpenCV/libpng; tools/make_png_golden.py) and compares the unfiltered pixels with the
 * pixels those libraries decode.  The filter direction is pinned by unfilter(filter(x)) == x for every mode
 * and by Pillow decoding the files our encoder writes.
 *
 * Layout: a filtered image is h rows of (1 filter-type byte + stride bytes); a raw image is h rows of
 * stride bytes; bpp = bytes per complete pixel, rounded up to 1 (PNG 9.2).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

static inline uint8_t paeth(uint8_t a, uint8_t b, uint8_t c) { /* PNG 9.4 */
    int p = (int)a + (int)b - (int)c;
    int pa = abs(p - (int)a), pb = abs(p - (int)b), pc = abs(p - (int)c);
    if (pa <= pb && pa <= pc) return a;
    if (pb <= pc) return b;
    return c;
}

static inline uint8_t predictor(int type, uint8_t a /*left*/, uint8_t b /*up*/, uint8_t c /*up-left*/) {
    switch (type) {
        case 1: return a;
        case 2: return b;
        case 3: return (uint8_t)(((int)a + (int)b) >> 1);
        case 4: return paeth(a, b, c);
        default: return 0;
    }
}

/* returns 0, or 1 + the index of the first row whose filter type is not 0..4 */
size_t fdo_png_unfilter(uint8_t *raw, const uint8_t *filtered, uint32_t h, uint32_t stride, uint32_t bpp) {
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t *src = filtered + (size_t)y * (1 + (size_t)stride);
        uint8_t *cur = raw + (size_t)y * stride;
        const uint8_t *up = y ? cur - stride : NULL;
        const int type = src[0];
        if (type > 4) return (size_t)y + 1;
        for (uint32_t x = 0; x < stride; x++) {
            uint8_t a = x >= bpp ? cur[x - bpp] : 0;
            uint8_t b = up ? up[x] : 0;
            uint8_t c = (up && x >= bpp) ? up[x - bpp] : 0;
            cur[x] = (uint8_t)(src[1 + x] + predictor(type, a, b, c));
        }
    }
    return 0;
}

static void filter_row(uint8_t *dst, int type, const uint8_t *cur, const uint8_t *up, uint32_t stride, uint32_t bpp) {
    for (uint32_t x = 0; x < stride; x++) {
        uint8_t a = x >= bpp ? cur[x - bpp] : 0;
        uint8_t b = up ? up[x] : 0;
        uint8_t c = (up && x >= bpp) ? up[x - bpp] : 0;
        dst[x] = (uint8_t)(cur[x] - predictor(type, a, b, c));
    }
}

/* mode 0..4: that filter type on every row; mode 5: per row the type with the smallest sum of |signed
 * filtered byte| (the heuristic PNG 12.8 recommends), lowest type number on ties */
void fdo_png_filter(uint8_t *filtered, const uint8_t *raw, uint32_t h, uint32_t stride, uint32_t bpp, uint32_t mode) {
    uint8_t *tmp = (uint8_t *)malloc(stride ? stride : 1);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t *cur = raw + (size_t)y * stride;
        const uint8_t *up = y ? cur - stride : NULL;
        uint8_t *dst = filtered + (size_t)y * (1 + (size_t)stride);
        int type = (int)mode;
        if (mode >= 5) {
            uint64_t best = ~0ull;
            type = 0;
            for (int t = 0; t < 5; t++) {
                filter_row(tmp, t, cur, up, stride, bpp);
                uint64_t sum = 0;
                for (uint32_t x = 0; x < stride; x++) sum += tmp[x] < 128 ? tmp[x] : 256u - tmp[x];
                if (sum < best) {
                    best = sum;
                    type = t;
                }
            }
        }
        dst[0] = (uint8_t)type;
        filter_row(dst + 1, type, cur, up, stride, bpp);
    }
    free(tmp);
}
