// synth.cuh -- K7: deterministic synthetic PNG-filtered RGBA tiles (benchmark / test input only).
//
// Recipe (integer-only so host and device agree bit for bit):
//   tile seed   s = splitmix64(seed + tile_index)
//   channel c in {R,G,B}: raw = (a_c*x + b_c*y + ((x*y) >> 6) + noise) & 0xff, a_c,b_c in [0,3] per tile,
//                         noise = hash(s, pixel, c) % 5 - 2;  A = 255
//   two constant-colour rectangles, 96 px x 64 rows (clipped to the tile), placed per tile
//   PNG filtering: row 0 Sub (type 1), other rows Paeth (type 4), bpp = 4;
//   row = 1 filter-type byte + 4*width residual bytes.
// Residuals come out as a two-sided geometric distribution around 0 with long zero runs inside the
// rectangles -- the distribution the reference's fixed Huffman code was trained for
// (reference src/tables.rs:3-6).
#pragma once
#include "simt.h"
#include "fdb_common.h"

namespace fdb {

FDB_HD uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

struct TileParams {
    uint64_t s;
    uint32_t a[3], b[3];
    uint32_t rx[2], ry[2];
    uint32_t col[2];  // RGBA packed, A = 255
};

FDB_HD TileParams tile_params(uint64_t seed, uint64_t tile, uint32_t width, uint32_t height) {
    TileParams p;
    p.s = splitmix64(seed + tile);
    uint64_t h = splitmix64(p.s ^ 0x1234567ull);
    for (int c = 0; c < 3; c++) {
        p.a[c] = (uint32_t)(h >> (4 * c)) & 3u;
        p.b[c] = (uint32_t)(h >> (4 * c + 2)) & 3u;
    }
    for (int k = 0; k < 2; k++) {
        uint64_t r = splitmix64(p.s ^ (0xABCDEFull + (uint64_t)k));
        p.rx[k] = (uint32_t)(r & 0xffffu) % width;
        p.ry[k] = (uint32_t)((r >> 16) & 0xffffu) % height;
        p.col[k] = (uint32_t)(r >> 32) | 0xff000000u;
    }
    return p;
}

FDB_HD uint32_t raw_channel(const TileParams& p, uint32_t x, uint32_t y, uint32_t c, uint32_t width) {
    if (c == 3) return 255u;
    for (int k = 0; k < 2; k++)
        if (x >= p.rx[k] && x < p.rx[k] + 96u && y >= p.ry[k] && y < p.ry[k] + 64u) return (p.col[k] >> (8u * c)) & 0xffu;
    uint64_t h = splitmix64(p.s + ((uint64_t)(y * width + x) * 4u + c) * 0x9E3779B97F4A7C15ull);
    int32_t noise = (int32_t)(h % 5u) - 2;
    return (uint32_t)((int32_t)(p.a[c] * x + p.b[c] * y + ((x * y) >> 6)) + noise) & 0xffu;
}

FDB_HD uint32_t paeth(uint32_t a, uint32_t b, uint32_t c) {
    int32_t pa = (int32_t)b - (int32_t)c, pb = (int32_t)a - (int32_t)c;
    int32_t pc = pa + pb;
    pa = pa < 0 ? -pa : pa;
    pb = pb < 0 ? -pb : pb;
    pc = pc < 0 ? -pc : pc;
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// filtered byte of channel c of pixel (x, y)
FDB_HD uint8_t filtered_byte(const TileParams& p, uint32_t x, uint32_t y, uint32_t c, uint32_t width) {
    uint32_t cur = raw_channel(p, x, y, c, width);
    uint32_t left = x ? raw_channel(p, x - 1, y, c, width) : 0u;
    if (y == 0) return (uint8_t)(cur - left);
    uint32_t up = raw_channel(p, x, y - 1, c, width);
    uint32_t ul = x ? raw_channel(p, x - 1, y - 1, c, width) : 0u;
    return (uint8_t)(cur - paeth(left, up, ul));
}

FDB_HD void synth_pixel(uint8_t* tile_out, const TileParams& p, uint32_t x, uint32_t y, uint32_t width) {
    uint8_t* row = tile_out + (uint64_t)y * (1u + 4u * width);
    if (x == 0) row[0] = y == 0 ? 1 : 4;
    for (uint32_t c = 0; c < 4; c++) row[1 + 4 * x + c] = filtered_byte(p, x, y, c, width);
}

FDB_GLOBAL void synth_tiles_kernel(uint8_t* out, uint64_t first_tile, uint64_t n_tiles, uint32_t width,
                                   uint32_t height, uint64_t seed) {
    const uint64_t per_tile = (uint64_t)width * height;
    const uint64_t total = n_tiles * per_tile;
    const uint64_t tile_bytes = (uint64_t)height * (1u + 4u * width);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t t = i / per_tile;
        uint32_t pix = (uint32_t)(i % per_tile);
        TileParams p = tile_params(seed, first_tile + t, width, height);
        synth_pixel(out + t * tile_bytes, p, pix % width, pix / width, width);
    }
}

}  // namespace fdb
