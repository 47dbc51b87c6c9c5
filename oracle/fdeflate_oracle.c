/*
 * fdeflate_oracle.c -- CPU ORACLE (test infrastructure only; see fdeflate_oracle.h).
 *
 * A plain-C restatement of image-rs/fdeflate 0.4.0-dev.  Every function cites the reference
 * file:line it follows (paths relative to the reference checkout).  The streaming state machine,
 * the bit reservoir, the fast and the careful decode loops and the table builder are restated
 * one-for-one so that corner cases (truncation, chunked input, full output buffers) behave like
 * the reference, not like "some inflate".
 */
#define _GNU_SOURCE
#include "fdeflate_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------
 * Constant tables (src/tables.rs)
 * ---------------------------------------------------------------------------------------- */

/* tables.rs:7-20 HUFFMAN_LENGTHS: the fixed PNG-trained code used by the ultra-fast encoder. */
static const uint8_t HUFFMAN_LENGTHS[286] = {
     2,  3,  4,  5,  5,  6,  6,  7,  7,  7,  8,  8,  8,  8,  8,  9,  9,  9,  9,  9,  9,  9, 10, 10, 10, 10,
    10, 10, 10, 10, 10, 11, 11, 11, 11, 11, 11, 11, 11, 11, 11, 11, 11, 12, 12, 12, 12, 12, 12, 12, 12, 12,
    12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12,
    12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12,
    12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12,
    12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12,
    12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12,
    12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12,
    12, 12, 12, 12, 12, 12, 11, 11, 11, 11, 11, 11, 11, 11, 11, 11, 10, 11, 10, 10, 10, 10, 10, 10, 10, 10,
    10,  9,  9,  9,  9,  9,  8,  9,  8,  8,  8,  8,  8,  7,  7,  7,  6,  6,  6,  5,  4,  3, 12, 12, 12,  9,
     9, 11, 10, 11, 11, 10, 11, 11, 11, 11, 11, 11, 12, 11, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12, 12,  9,
};

/* tables.rs:63-65 */
static const uint8_t CLCL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
/* tables.rs:68-70 */
static const uint8_t LEN_SYM_TO_LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2,
                                                 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
/* tables.rs:73-76 */
static const uint16_t LEN_SYM_TO_LEN_BASE[29] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23, 27,
                                                 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
/* tables.rs:79-82 */
static const uint8_t DIST_SYM_TO_DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2,  3,  3,  4,  4,  5,  5,  6,
                                                   6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
/* tables.rs:85-88 */
static const uint16_t DIST_SYM_TO_DIST_BASE[30] = {1,   2,   3,   4,   5,   7,    9,    13,   17,   25,
                                                   33,  49,  65,  97,  129, 193,  257,  385,  513,  769,
                                                   1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};

/* decompress.rs:61-63 */
#define LITERAL_ENTRY 0x8000u
#define EXCEPTIONAL_ENTRY 0x4000u
#define SECONDARY_TABLE_ENTRY 0x2000u

/* ultrafast.rs:82-86 */
static const uint8_t UF_HEADER[54] = {
    120, 1,   237, 192, 3,   160, 36,  89,  150, 198, 241, 255, 119, 238, 141, 200, 204, 167,
    114, 75,  99,  174, 109, 219, 182, 109, 219, 182, 109, 219, 182, 109, 105, 140, 158, 150,
    74,  175, 158, 50,  51,  34,  238, 249, 118, 183, 106, 122, 166, 135, 59,  107, 213, 15,
};

static uint16_t HUFFMAN_CODES[286];       /* tables.rs:22-25 via lib.rs:103-127 */
static uint16_t LENGTH_TO_SYMBOL[256];    /* tables.rs:28-43 (derived from the deflate spec) */
static uint8_t LENGTH_TO_LEN_EXTRA[256];  /* tables.rs:46-55 */
static uint32_t LITLEN_TABLE_ENTRIES[288]; /* tables.rs:99-122 */
static uint32_t DISTANCE_TABLE_ENTRIES[32]; /* tables.rs:130-140 */
static uint32_t FIXED_LITLEN_TABLE[512];  /* tables.rs:142-195, rebuilt here from FIXED_CODE_LENGTHS
                                             (the literal values are a KAT in tests/golden/) */
static uint32_t FIXED_DIST_TABLE[32];     /* tables.rs:197-202 */

static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static int build_tables_impl(size_t hlit, const uint8_t *code_lengths, uint32_t *litlen_table,
                             uint32_t *dist_table, uint16_t *secondary, size_t *secondary_len,
                             uint16_t *dist_secondary, size_t *dist_secondary_len, uint16_t *eof_code,
                             uint16_t *eof_mask, uint8_t *eof_bits);

static uint16_t reverse_bits16(uint16_t v) {
    uint16_t r = 0;
    for (int i = 0; i < 16; i++)
        if (v & (1u << i)) r |= (uint16_t)(1u << (15 - i));
    return r;
}

/* lib.rs:103-127 compute_codes: canonical codes, bit-reversed so they can be OR-ed LSB-first. */
static int compute_codes(const uint8_t *lengths, size_t n, uint16_t *codes) {
    uint32_t code = 0;
    for (unsigned len = 1; len <= 16; len++) {
        for (size_t i = 0; i < n; i++) {
            if (lengths[i] == len) {
                codes[i] = (uint16_t)(reverse_bits16((uint16_t)code) >> (16 - len));
                code += 1;
            }
        }
        code <<= 1;
    }
    return code == (2u << 16);
}

static void init_tables(void) {
    if (!compute_codes(HUFFMAN_LENGTHS, 286, HUFFMAN_CODES)) abort();

    /* tables.rs:28-55: length -> (symbol, extra bits). Index = length - 3.  Length 258 is symbol 285. */
    for (int sym = 0; sym < 29; sym++) {
        int base = LEN_SYM_TO_LEN_BASE[sym];
        int n = 1 << LEN_SYM_TO_LEN_EXTRA[sym];
        for (int j = 0; j < n; j++) {
            int len = base + j;
            if (len > 258) continue;
            if (sym == 27 && j == 31) continue; /* 258 is coded by symbol 285, see decompress.rs:1203 */
            LENGTH_TO_SYMBOL[len - 3] = (uint16_t)(257 + sym);
            LENGTH_TO_LEN_EXTRA[len - 3] = LEN_SYM_TO_LEN_EXTRA[sym];
        }
    }

    /* tables.rs:99-122 */
    for (int i = 0; i < 288; i++) LITLEN_TABLE_ENTRIES[i] = EXCEPTIONAL_ENTRY;
    for (int i = 0; i < 256; i++) LITLEN_TABLE_ENTRIES[i] = ((uint32_t)i << 16) | LITERAL_ENTRY | (1u << 8);
    for (int i = 257; i < 286; i++)
        LITLEN_TABLE_ENTRIES[i] =
            ((uint32_t)LEN_SYM_TO_LEN_BASE[i - 257] << 16) | ((uint32_t)LEN_SYM_TO_LEN_EXTRA[i - 257] << 8);
    /* tables.rs:130-140 */
    for (int i = 0; i < 32; i++) DISTANCE_TABLE_ENTRIES[i] = 0;
    for (int i = 0; i < 30; i++)
        DISTANCE_TABLE_ENTRIES[i] = ((uint32_t)DIST_SYM_TO_DIST_BASE[i] << 16) |
                                    ((uint32_t)DIST_SYM_TO_DIST_EXTRA[i] << 8) | LITERAL_ENTRY;

    /* tables.rs:142-202: the reference hard-codes these; decompress.rs:1218-1233 asserts they equal
     * build_tables(288, FIXED_CODE_LENGTHS).  We build them (tables.rs:207-232 lengths) and the test
     * suite checks the result against the literal values kept in tests/golden/fixed_tables.json. */
    uint8_t lengths[320];
    int i = 0;
    for (; i < 144; i++) lengths[i] = 8;
    for (; i < 256; i++) lengths[i] = 9;
    for (; i < 280; i++) lengths[i] = 7;
    for (; i < 288; i++) lengths[i] = 8;
    for (; i < 320; i++) lengths[i] = 5;
    static uint32_t lit[4096], dist[512];
    static uint16_t sec[FDO_SECONDARY_CAP], dsec[FDO_SECONDARY_CAP];
    size_t nsec = 0, ndsec = 0;
    uint16_t eof_code, eof_mask;
    uint8_t eof_bits;
    if (build_tables_impl(288, lengths, lit, dist, sec, &nsec, dsec, &ndsec, &eof_code, &eof_mask, &eof_bits) !=
        FDO_OK)
        abort();
    memcpy(FIXED_LITLEN_TABLE, lit, sizeof FIXED_LITLEN_TABLE);
    memcpy(FIXED_DIST_TABLE, dist, sizeof FIXED_DIST_TABLE);
}

static void ensure_init(void) { pthread_once(&g_once, init_tables); }

const uint8_t *fdo_huffman_lengths(void) { return HUFFMAN_LENGTHS; }
const uint16_t *fdo_huffman_codes(void) {
    ensure_init();
    return HUFFMAN_CODES;
}
const uint8_t *fdo_ultrafast_header(void) { return UF_HEADER; }
const uint32_t *fdo_litlen_table_entries(void) {
    ensure_init();
    return LITLEN_TABLE_ENTRIES;
}
const uint32_t *fdo_distance_table_entries(void) {
    ensure_init();
    return DISTANCE_TABLE_ENTRIES;
}
const uint16_t *fdo_length_to_symbol(void) {
    ensure_init();
    return LENGTH_TO_SYMBOL;
}
const uint8_t *fdo_length_to_len_extra(void) {
    ensure_init();
    return LENGTH_TO_LEN_EXTRA;
}

/* ------------------------------------------------------------------------------------------
 * adler32 -- RFC 1950 section 8.2 / 9; the reference calls simd_adler32::Adler32::{new,write,finish}
 * (decompress.rs:145,311,318,332; ultrafast.rs:72,95,176).
 * ---------------------------------------------------------------------------------------- */
uint32_t fdo_adler32(uint32_t adler, const uint8_t *data, size_t len) {
    /* Same arithmetic as the byte loop (a += d; b += a), regrouped per block of 32 bytes:
     * b += 32 a + sum (32 - i) d[i]; a += sum d[i] -- two reductions gcc vectorises, so that the CPU baseline is not
     * held back by a scalar checksum (the reference uses the SIMD crate simd-adler32, Cargo.toml:21). */
    uint32_t a = adler & 0xffff, b = adler >> 16;
    while (len > 0) {
        size_t n = len < 5536 ? len : 5536; /* multiple of 32 below 5552, the largest n that cannot overflow 32 bits */
        len -= n;
        while (n >= 32) {
            uint32_t s = 0, w = 0;
            for (unsigned i = 0; i < 32; i++) {
                s += data[i];
                w += (32u - i) * data[i];
            }
            b += 32u * a + w;
            a += s;
            data += 32;
            n -= 32;
        }
        while (n--) {
            a += *data++;
            b += a;
        }
        a %= 65521;
        b %= 65521;
    }
    return (b << 16) | a;
}

/* ------------------------------------------------------------------------------------------
 * huffman.rs
 * ---------------------------------------------------------------------------------------- */

/* huffman.rs:5-15 */
static uint16_t next_codeword(uint16_t codeword, uint16_t table_size) {
    if (codeword == (uint16_t)(table_size - 1)) return codeword;
    uint16_t x = (uint16_t)(codeword ^ (table_size - 1));
    unsigned lz = (unsigned)__builtin_clz((unsigned)x) - 16; /* u16::leading_zeros, x != 0 */
    unsigned adv = 15 - lz;
    uint16_t bit = (uint16_t)(1u << adv);
    codeword &= (uint16_t)(bit - 1);
    codeword |= bit;
    return codeword;
}

static unsigned ilog2_sz(size_t v) { return 63u - (unsigned)__builtin_clzll((unsigned long long)v); }

/* huffman.rs:18-184 */
int fdo_build_table(const uint8_t *lengths, size_t n_lengths, const uint32_t *entries, size_t n_entries,
                    uint16_t *codes, uint32_t *primary_table, size_t primary_size,
                    uint16_t *secondary_table, size_t *secondary_len, int is_distance_table,
                    int double_literal) {
    /* :28-31 histogram */
    size_t histogram[16] = {0};
    for (size_t i = 0; i < n_lengths; i++) histogram[lengths[i]] += 1;

    /* :34-37 */
    size_t max_length = 15;
    while (max_length > 1 && histogram[max_length] == 0) max_length -= 1;

    /* :40-59 zero / one symbol distance codes */
    if (is_distance_table) {
        if (max_length == 0) {
            for (size_t i = 0; i < primary_size; i++) primary_table[i] = 0;
            *secondary_len = 0;
            return 1;
        } else if (max_length == 1 && histogram[1] == 1) {
            size_t symbol = 0;
            while (lengths[symbol] != 1) symbol++;
            codes[symbol] = 0;
            uint32_t entry = (symbol < n_entries ? entries[symbol] : ((uint32_t)symbol << 16)) | 1;
            for (size_t i = 0; i < primary_size; i += 2) {
                primary_table[i] = entry;
                if (i + 1 < primary_size) primary_table[i + 1] = 0;
            }
            return 1;
        }
    }

    /* :63-75 offsets + completeness (Kraft) check */
    size_t offsets[16] = {0};
    size_t codespace_used = 0;
    offsets[1] = histogram[0];
    for (size_t i = 1; i < max_length; i++) {
        offsets[i + 1] = offsets[i] + histogram[i];
        codespace_used = (codespace_used << 1) + histogram[i];
    }
    codespace_used = (codespace_used << 1) + histogram[max_length];
    if (codespace_used != ((size_t)1 << max_length)) return 0;

    /* :78-84 counting sort */
    size_t next_index[16];
    memcpy(next_index, offsets, sizeof offsets);
    size_t sorted_symbols[288] = {0};
    for (size_t symbol = 0; symbol < n_lengths; symbol++) {
        uint8_t length = lengths[symbol];
        sorted_symbols[next_index[length]] = symbol;
        next_index[length] += 1;
    }

    uint16_t codeword = 0;
    size_t i = histogram[0];

    /* :90-136 primary table */
    size_t primary_table_bits = ilog2_sz(primary_size);
    size_t primary_table_mask = ((size_t)1 << primary_table_bits) - 1;
    for (size_t length = 1; length <= primary_table_bits; length++) {
        size_t current_table_end = (size_t)1 << length;

        for (size_t k = 0; k < histogram[length]; k++) {
            size_t symbol = sorted_symbols[i];
            i += 1;
            primary_table[codeword] =
                (symbol < n_entries ? entries[symbol] : ((uint32_t)symbol << 16)) | (uint32_t)length;
            codes[symbol] = codeword;
            codeword = next_codeword(codeword, (uint16_t)current_table_end);
        }

        if (double_literal) { /* :110-130 */
            for (size_t len1 = 1; len1 < length; len1++) {
                size_t len2 = length - len1;
                for (size_t s1 = offsets[len1]; s1 < next_index[len1]; s1++) {
                    for (size_t s2 = offsets[len2]; s2 < next_index[len2]; s2++) {
                        size_t sym1 = sorted_symbols[s1];
                        size_t sym2 = sorted_symbols[s2];
                        if (sym1 < 256 && sym2 < 256) {
                            uint16_t codeword1 = codes[sym1];
                            uint16_t codeword2 = codes[sym2];
                            uint16_t cw = (uint16_t)(codeword1 | (codeword2 << len1));
                            uint32_t entry =
                                ((uint32_t)sym1 << 16) | ((uint32_t)sym2 << 24) | LITERAL_ENTRY | (2u << 8);
                            primary_table[cw] = entry | (uint32_t)length;
                        }
                    }
                }
            }
        }

        /* :133-135 double the table */
        if (length < primary_table_bits)
            memcpy(primary_table + current_table_end, primary_table, current_table_end * sizeof(uint32_t));
    }

    /* :139-181 secondary table */
    size_t slen = 0;
    if (max_length > primary_table_bits) {
        size_t subtable_start = 0;
        size_t subtable_prefix = (size_t)-1;
        for (size_t length = primary_table_bits + 1; length <= max_length; length++) {
            size_t subtable_size = (size_t)1 << (length - primary_table_bits);
            uint32_t overflow_bits_mask = (uint32_t)subtable_size - 1;
            for (size_t k = 0; k < histogram[length]; k++) {
                if (((size_t)codeword & primary_table_mask) != subtable_prefix) {
                    subtable_prefix = (size_t)codeword & primary_table_mask;
                    subtable_start = slen;
                    primary_table[subtable_prefix] = ((uint32_t)subtable_start << 16) | EXCEPTIONAL_ENTRY |
                                                     SECONDARY_TABLE_ENTRY | overflow_bits_mask;
                    if (subtable_start + subtable_size > FDO_SECONDARY_CAP) abort();
                    for (size_t z = slen; z < subtable_start + subtable_size; z++) secondary_table[z] = 0;
                    slen = subtable_start + subtable_size;
                }
                size_t symbol = sorted_symbols[i];
                i += 1;
                codes[symbol] = codeword;
                secondary_table[subtable_start + ((size_t)codeword >> primary_table_bits)] =
                    (uint16_t)(((uint16_t)symbol << 4) | (uint16_t)length);
                codeword = next_codeword(codeword, (uint16_t)(1u << length));
            }

            /* :171-179 extend the subtable if longer codes share the prefix */
            if (length < max_length && ((size_t)codeword & primary_table_mask) == subtable_prefix) {
                size_t cur = slen - subtable_start;
                if (slen + cur > FDO_SECONDARY_CAP) abort();
                memcpy(secondary_table + slen, secondary_table + subtable_start, cur * sizeof(uint16_t));
                slen += cur;
                size_t new_size = slen - subtable_start;
                uint32_t mask = (uint32_t)new_size - 1;
                primary_table[subtable_prefix] =
                    ((uint32_t)subtable_start << 16) | EXCEPTIONAL_ENTRY | SECONDARY_TABLE_ENTRY | mask;
            }
        }
    }
    *secondary_len = slen;
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * decompress.rs
 * ---------------------------------------------------------------------------------------- */

enum state { ST_ZLIB_HEADER, ST_BLOCK_HEADER, ST_CODE_LENGTH_CODES, ST_CODE_LENGTHS, ST_COMPRESSED_DATA,
             ST_UNCOMPRESSED_DATA, ST_CHECKSUM, ST_DONE }; /* :84-93 */

enum queued_kind { Q_NONE, Q_RLE, Q_BACKREF }; /* :1066-1070 */

typedef struct {
    uint64_t buffer;
    uint8_t nbits;
} bitbuffer; /* :1022-1025 */

typedef struct {
    const uint8_t *ptr;
    size_t len;
} slice;

struct fdo_decompressor { /* :96-113 */
    /* CompressedBlock :71-81 */
    uint32_t litlen_table[4096];
    uint16_t secondary_table[FDO_SECONDARY_CAP];
    size_t secondary_len;
    uint32_t dist_table[512];
    uint16_t dist_secondary_table[FDO_SECONDARY_CAP];
    size_t dist_secondary_len;
    uint16_t eof_code, eof_mask;
    uint8_t eof_bits;
    /* BlockHeader :50-59 */
    size_t hlit, hdist, hclen, num_lengths_read;
    uint32_t table[128];
    uint8_t code_lengths[320];

    uint16_t uncompressed_bytes_left;
    bitbuffer bits;
    int queued_kind;
    uint8_t q_data;
    size_t q_dist, q_length;
    int last_block, fixed_table;
    int state;
    uint32_t checksum;
    int ignore_adler32;
};

/* :1035-1052 */
static void fill_buffer(bitbuffer *b, slice *input) {
    if (input->len >= 8) {
        uint8_t bits = b->nbits & 63;
        uint64_t v;
        memcpy(&v, input->ptr, 8); /* little-endian host assumed (x86-64 / aarch64) */
        b->buffer |= v << bits;
        size_t adv = (size_t)((63 - bits) / 8);
        input->ptr += adv;
        input->len -= adv;
        bits |= 56;
        b->nbits = bits;
    } else {
        size_t room = (size_t)((63 - b->nbits) / 8);
        size_t nbytes = input->len < room ? input->len : room;
        uint8_t input_data[8] = {0};
        memcpy(input_data, input->ptr, nbytes);
        uint64_t v;
        memcpy(&v, input_data, 8);
        b->buffer |= (b->nbits < 64) ? (v << b->nbits) : 0; /* checked_shl(..).unwrap_or(0) */
        b->nbits += (uint8_t)(nbytes * 8);
        input->ptr += nbytes;
        input->len -= nbytes;
    }
}
/* :1054-1057 */
static uint64_t peek_bits(const bitbuffer *b, uint8_t nbits) { return b->buffer & (((uint64_t)1 << nbits) - 1); }
/* :1059-1063 */
static void consume_bits(bitbuffer *b, uint8_t nbits) {
    b->buffer >>= nbits;
    b->nbits -= nbits;
}

/* :561-606 CompressedBlock::build_tables */
static int build_tables_impl(size_t hlit, const uint8_t *code_lengths, uint32_t *litlen_table,
                             uint32_t *dist_table, uint16_t *secondary, size_t *secondary_len,
                             uint16_t *dist_secondary, size_t *dist_secondary_len, uint16_t *eof_code,
                             uint16_t *eof_mask, uint8_t *eof_bits) {
    /* :563-566 */
    if (code_lengths[256] == 0) return FDO_BAD_LITERAL_LENGTH_HUFFMAN_TREE;

    uint16_t codes[288] = {0};
    *secondary_len = 0;
    /* :570-580 (note: the reference reports BadCodeLengthHuffmanTree here) */
    if (!fdo_build_table(code_lengths, hlit, LITLEN_TABLE_ENTRIES, 288, codes, litlen_table, 4096, secondary,
                         secondary_len, 0, 1))
        return FDO_BAD_CODE_LENGTH_HUFFMAN_TREE;

    /* :582-584 */
    *eof_code = codes[256];
    *eof_mask = (uint16_t)((1u << code_lengths[256]) - 1);
    *eof_bits = code_lengths[256];

    /* :587-603 */
    const uint8_t *lengths = code_lengths + 288;
    int all_zero = 1;
    for (int i = 0; i < 32; i++)
        if (lengths[i]) all_zero = 0;
    if (all_zero) {
        memset(dist_table, 0, 512 * sizeof(uint32_t));
    } else {
        uint16_t dist_codes[32] = {0};
        if (!fdo_build_table(lengths, 32, DISTANCE_TABLE_ENTRIES, 32, dist_codes, dist_table, 512,
                             dist_secondary, dist_secondary_len, 1, 0))
            return FDO_BAD_DISTANCE_HUFFMAN_TREE;
    }
    return FDO_OK;
}

fdo_decompressor *fdo_decompressor_new(void) { /* :123-151 */
    ensure_init();
    fdo_decompressor *d = (fdo_decompressor *)calloc(1, sizeof *d);
    if (!d) return NULL;
    d->state = ST_ZLIB_HEADER;
    d->checksum = 1; /* Adler32::new() */
    return d;
}
void fdo_decompressor_free(fdo_decompressor *d) { free(d); }
void fdo_decompressor_ignore_adler32(fdo_decompressor *d) { d->ignore_adler32 = 1; } /* :154-156 */
int fdo_decompressor_is_done(const fdo_decompressor *d) { return d->state == ST_DONE; } /* :340-342 */

/* :344-438 */
static int read_block_header(fdo_decompressor *d, slice *remaining_input) {
    for (;;) { /* the reference recurses once after skipping empty fixed blocks (:393) */
        fill_buffer(&d->bits, remaining_input);
        if (d->bits.nbits < 10) return FDO_OK;

        uint64_t start = peek_bits(&d->bits, 3);
        d->last_block = (start & 1) != 0;
        switch (start >> 1) {
        case 0: { /* :353-370 stored */
            uint8_t align_bits = (uint8_t)((d->bits.nbits - 3) % 8);
            uint8_t header_bits = (uint8_t)(3 + 32 + align_bits);
            if (d->bits.nbits < header_bits) return FDO_OK;
            uint16_t len = (uint16_t)(peek_bits(&d->bits, (uint8_t)(align_bits + 19)) >> (align_bits + 3));
            uint16_t nlen = (uint16_t)(peek_bits(&d->bits, header_bits) >> (align_bits + 19));
            if (nlen != (uint16_t)~len) return FDO_INVALID_UNCOMPRESSED_BLOCK_LENGTH;
            d->state = ST_UNCOMPRESSED_DATA;
            d->uncompressed_bytes_left = len;
            consume_bits(&d->bits, header_bits);
            return FDO_OK;
        }
        case 1: { /* :371-414 fixed */
            consume_bits(&d->bits, 3);
            if (peek_bits(&d->bits, 7) == 0) { /* :377-394 empty block */
                consume_bits(&d->bits, 7);
                if (d->last_block) {
                    d->state = ST_CHECKSUM;
                    return FDO_OK;
                }
                while (d->bits.nbits >= 10 && peek_bits(&d->bits, 10) == 2) {
                    consume_bits(&d->bits, 10);
                    fill_buffer(&d->bits, remaining_input);
                }
                continue; /* return self.read_block_header(remaining_input) */
            }
            if (!d->fixed_table) { /* :397-410 */
                d->fixed_table = 1;
                for (int c = 0; c < 4096; c += 512) memcpy(d->litlen_table + c, FIXED_LITLEN_TABLE, 512 * 4);
                for (int c = 0; c < 512; c += 32) memcpy(d->dist_table + c, FIXED_DIST_TABLE, 32 * 4);
                d->eof_bits = 7;
                d->eof_code = 0;
                d->eof_mask = 0x7f;
            }
            d->state = ST_COMPRESSED_DATA;
            return FDO_OK;
        }
        case 2: { /* :415-434 dynamic */
            if (d->bits.nbits < 17) return FDO_OK;
            d->hlit = (size_t)(peek_bits(&d->bits, 8) >> 3) + 257;
            d->hdist = (size_t)(peek_bits(&d->bits, 13) >> 8) + 1;
            d->hclen = (size_t)(peek_bits(&d->bits, 17) >> 13) + 4;
            if (d->hlit > 286) return FDO_INVALID_HLIT;
            if (d->hdist > 30) return FDO_INVALID_HDIST;
            consume_bits(&d->bits, 17);
            d->state = ST_CODE_LENGTH_CODES;
            d->fixed_table = 0;
            return FDO_OK;
        }
        default: /* :435 */
            return FDO_INVALID_BLOCK_TYPE;
        }
    }
}

/* :440-477 */
static int read_code_length_codes(fdo_decompressor *d, slice *remaining_input) {
    fill_buffer(&d->bits, remaining_input);
    if ((size_t)d->bits.nbits + remaining_input->len * 8 < 3 * d->hclen) return FDO_OK;

    uint8_t code_length_lengths[19] = {0};
    for (size_t i = 0; i < d->hclen; i++) {
        code_length_lengths[CLCL_ORDER[i]] = (uint8_t)peek_bits(&d->bits, 3);
        consume_bits(&d->bits, 3);
        if (i == 17) fill_buffer(&d->bits, remaining_input);
    }

    uint16_t codes[19] = {0};
    uint16_t dummy_secondary[8];
    size_t dummy_len = 0;
    if (!fdo_build_table(code_length_lengths, 19, NULL, 0, codes, d->table, 128, dummy_secondary, &dummy_len, 0,
                         0))
        return FDO_BAD_CODE_LENGTH_HUFFMAN_TREE;

    d->state = ST_CODE_LENGTHS;
    d->num_lengths_read = 0;
    return FDO_OK;
}

/* :479-555 */
static int read_code_lengths(fdo_decompressor *d, slice *remaining_input) {
    size_t total_lengths = d->hlit + d->hdist;
    while (d->num_lengths_read < total_lengths) {
        fill_buffer(&d->bits, remaining_input);
        if (d->bits.nbits < 7) return FDO_OK;

        uint64_t code = peek_bits(&d->bits, 7);
        uint32_t entry = d->table[code];
        uint8_t length = (uint8_t)(entry & 0x7);
        uint8_t symbol = (uint8_t)(entry >> 16);

        if (symbol <= 15) {
            d->code_lengths[d->num_lengths_read] = symbol;
            d->num_lengths_read += 1;
            consume_bits(&d->bits, length);
        } else {
            size_t base_repeat;
            uint8_t extra_bits;
            if (symbol == 16) {
                base_repeat = 3;
                extra_bits = 2;
            } else if (symbol == 17) {
                base_repeat = 3;
                extra_bits = 3;
            } else {
                base_repeat = 11;
                extra_bits = 7;
            }
            if (d->bits.nbits < length + extra_bits) return FDO_OK;

            uint8_t value = 0;
            if (symbol == 16) {
                if (d->num_lengths_read == 0) return FDO_INVALID_CODE_LENGTH_REPEAT;
                value = d->code_lengths[d->num_lengths_read - 1];
            }
            size_t repeat = (size_t)(peek_bits(&d->bits, (uint8_t)(length + extra_bits)) >> length) + base_repeat;
            if (d->num_lengths_read + repeat > total_lengths) return FDO_INVALID_CODE_LENGTH_REPEAT;
            for (size_t i = 0; i < repeat; i++) d->code_lengths[d->num_lengths_read + i] = value;
            d->num_lengths_read += repeat;
            consume_bits(&d->bits, (uint8_t)(length + extra_bits));
        }
    }

    /* :541-549 */
    memmove(d->code_lengths + 288, d->code_lengths + d->hlit, total_lengths - d->hlit);
    for (size_t i = d->hlit; i < 288; i++) d->code_lengths[i] = 0;
    for (size_t i = 288 + d->hdist; i < 320; i++) d->code_lengths[i] = 0;

    int st = build_tables_impl(d->hlit, d->code_lengths, d->litlen_table, d->dist_table, d->secondary_table,
                               &d->secondary_len, d->dist_secondary_table, &d->dist_secondary_len, &d->eof_code,
                               &d->eof_mask, &d->eof_bits);
    if (st != FDO_OK) return st;
    d->state = ST_COMPRESSED_DATA;
    return FDO_OK;
}

enum block_status { MORE_DATA_PRESENT, REACHED_END_OF_BLOCK };

/* Apply a back-reference exactly like :792-829 / :969-1006.  Returns 1 if the output filled up and
 * the remainder was queued (caller must `break`). */
static int do_copy(fdo_decompressor *d, uint8_t *output, size_t output_len, size_t *output_index, size_t length,
                   size_t dist) {
    size_t oi = *output_index;
    size_t room = output_len - oi;
    size_t copy_length = length < room ? length : room;
    if (dist == 1) {
        uint8_t last = output[oi - 1];
        memset(output + oi, last, copy_length);
        if (length - copy_length != 0) {
            d->queued_kind = Q_RLE;
            d->q_data = last;
            d->q_length = length - copy_length;
            *output_index = output_len;
            return 1;
        }
    } else if (oi + length + 15 <= output_len) {
        size_t start = oi - dist;
        memmove(output + oi, output + start, 16);
        if (length > 16 || dist < 16) {
            size_t step = dist < 16 ? dist : 16;
            for (size_t i = step; i < length; i += step) memmove(output + oi + i, output + start + i, 16);
        }
    } else {
        if (dist < copy_length) {
            for (size_t i = 0; i < copy_length; i++) output[oi + i] = output[oi + i - dist];
        } else {
            memmove(output + oi, output + oi - dist, copy_length);
        }
        if (length - copy_length != 0) {
            d->queued_kind = Q_BACKREF;
            d->q_dist = dist;
            d->q_length = length - copy_length;
            *output_index = output_len;
            return 1;
        }
    }
    *output_index = oi + copy_length;
    return 0;
}

/* :611-1018 CompressedBlock::read_compressed */
static int read_compressed(fdo_decompressor *d, slice *remaining_input, uint8_t *output, size_t output_len,
                           size_t *output_index_io, int *block_status) {
    const uint64_t litlen_table_mask = 4096 - 1;
    const unsigned litlen_table_bits = 12;
    const uint64_t dist_table_mask = 512 - 1;
    const unsigned dist_table_bits = 9;
    bitbuffer *bb = &d->bits;
    size_t output_index = *output_index_io;

    /* ---- fast loop :645-830 ---- */
    fill_buffer(bb, remaining_input);
    uint32_t litlen_entry = d->litlen_table[bb->buffer & litlen_table_mask];
    while (output_index + 8 <= output_len && remaining_input->len >= 8) {
        uint64_t bits;
        uint8_t litlen_code_bits = (uint8_t)litlen_entry;
        if (litlen_entry & LITERAL_ENTRY) {
            uint32_t litlen_entry2 = d->litlen_table[(bb->buffer >> litlen_code_bits) & litlen_table_mask];
            uint8_t litlen_code_bits2 = (uint8_t)litlen_entry2;
            uint32_t litlen_entry3 =
                d->litlen_table[(bb->buffer >> (litlen_code_bits + litlen_code_bits2)) & litlen_table_mask];
            uint8_t litlen_code_bits3 = (uint8_t)litlen_entry3;
            uint32_t litlen_entry4 =
                d->litlen_table[(bb->buffer >> (litlen_code_bits + litlen_code_bits2 + litlen_code_bits3)) &
                                litlen_table_mask];

            size_t advance_output_bytes = (litlen_entry & 0xf00) >> 8;
            output[output_index] = (uint8_t)(litlen_entry >> 16);
            output[output_index + 1] = (uint8_t)(litlen_entry >> 24);
            output_index += advance_output_bytes;

            if (litlen_entry2 & LITERAL_ENTRY) {
                size_t advance_output_bytes2 = (litlen_entry2 & 0xf00) >> 8;
                output[output_index] = (uint8_t)(litlen_entry2 >> 16);
                output[output_index + 1] = (uint8_t)(litlen_entry2 >> 24);
                output_index += advance_output_bytes2;

                if (litlen_entry3 & LITERAL_ENTRY) {
                    size_t advance_output_bytes3 = (litlen_entry3 & 0xf00) >> 8;
                    output[output_index] = (uint8_t)(litlen_entry3 >> 16);
                    output[output_index + 1] = (uint8_t)(litlen_entry3 >> 24);
                    output_index += advance_output_bytes3;

                    litlen_entry = litlen_entry4;
                    consume_bits(bb, (uint8_t)(litlen_code_bits + litlen_code_bits2 + litlen_code_bits3));
                    fill_buffer(bb, remaining_input);
                    continue;
                } else {
                    consume_bits(bb, (uint8_t)(litlen_code_bits + litlen_code_bits2));
                    litlen_entry = litlen_entry3;
                    litlen_code_bits = litlen_code_bits3;
                    fill_buffer(bb, remaining_input);
                    bits = bb->buffer;
                }
            } else {
                consume_bits(bb, litlen_code_bits);
                bits = bb->buffer;
                litlen_entry = litlen_entry2;
                litlen_code_bits = litlen_code_bits2;
                if (bb->nbits < 48) fill_buffer(bb, remaining_input);
            }
        } else {
            bits = bb->buffer;
        }

        /* :709-748 13+ bit literal, back-reference or EOF */
        uint32_t length_base;
        uint8_t length_extra_bits;
        if ((litlen_entry & EXCEPTIONAL_ENTRY) == 0) {
            length_base = litlen_entry >> 16;
            length_extra_bits = (uint8_t)(litlen_entry >> 8);
        } else if (litlen_entry & SECONDARY_TABLE_ENTRY) {
            uint32_t secondary_table_index =
                (litlen_entry >> 16) + ((uint32_t)(bits >> litlen_table_bits) & (litlen_entry & 0xff));
            uint16_t secondary_entry = d->secondary_table[secondary_table_index];
            uint16_t litlen_symbol = secondary_entry >> 4;
            uint8_t code_bits2 = (uint8_t)(secondary_entry & 0xf);
            if (litlen_symbol <= 255) {
                consume_bits(bb, code_bits2);
                litlen_entry = d->litlen_table[bb->buffer & litlen_table_mask];
                fill_buffer(bb, remaining_input);
                output[output_index] = (uint8_t)litlen_symbol;
                output_index += 1;
                continue;
            } else if (litlen_symbol == 256) {
                consume_bits(bb, code_bits2);
                *output_index_io = output_index;
                *block_status = REACHED_END_OF_BLOCK;
                return FDO_OK;
            } else {
                length_base = LEN_SYM_TO_LEN_BASE[litlen_symbol - 257];
                length_extra_bits = LEN_SYM_TO_LEN_EXTRA[litlen_symbol - 257];
                litlen_code_bits = code_bits2;
            }
        } else if (litlen_code_bits == 0) {
            return FDO_INVALID_LITERAL_LENGTH_CODE;
        } else {
            consume_bits(bb, litlen_code_bits);
            *output_index_io = output_index;
            *block_status = REACHED_END_OF_BLOCK;
            return FDO_OK;
        }
        bits >>= litlen_code_bits;

        uint64_t length_extra_mask = ((uint64_t)1 << length_extra_bits) - 1;
        size_t length = (size_t)length_base + (size_t)(bits & length_extra_mask);
        bits >>= length_extra_bits;

        uint32_t dist_entry = d->dist_table[bits & dist_table_mask];
        uint16_t dist_base;
        uint8_t dist_extra_bits, dist_code_bits;
        if (dist_entry & LITERAL_ENTRY) {
            dist_base = (uint16_t)(dist_entry >> 16);
            dist_extra_bits = (uint8_t)(dist_entry >> 8) & 0xf;
            dist_code_bits = (uint8_t)dist_entry;
        } else if ((dist_entry >> 8) == 0) {
            return FDO_INVALID_DISTANCE_CODE;
        } else {
            uint32_t secondary_table_index =
                (dist_entry >> 16) + ((uint32_t)(bits >> dist_table_bits) & (dist_entry & 0xff));
            uint16_t secondary_entry = d->dist_secondary_table[secondary_table_index];
            size_t dist_symbol = secondary_entry >> 4;
            if (dist_symbol >= 30) return FDO_INVALID_DISTANCE_CODE;
            dist_base = DIST_SYM_TO_DIST_BASE[dist_symbol];
            dist_extra_bits = DIST_SYM_TO_DIST_EXTRA[dist_symbol];
            dist_code_bits = (uint8_t)(secondary_entry & 0xf);
        }
        bits >>= dist_code_bits;

        size_t dist = (size_t)dist_base + (size_t)(bits & (((uint64_t)1 << dist_extra_bits) - 1));
        if (dist > output_index) return FDO_DISTANCE_TOO_FAR_BACK;

        consume_bits(bb, (uint8_t)(litlen_code_bits + length_extra_bits + dist_code_bits + dist_extra_bits));
        fill_buffer(bb, remaining_input);
        litlen_entry = d->litlen_table[bb->buffer & litlen_table_mask];

        if (do_copy(d, output, output_len, &output_index, length, dist)) break;
    }

    /* ---- careful loop :836-1007 ---- */
    for (;;) {
        fill_buffer(bb, remaining_input);
        if (output_index == output_len) break;

        uint64_t bits = bb->buffer;
        uint32_t entry = d->litlen_table[bits & litlen_table_mask];
        uint8_t litlen_code_bits = (uint8_t)entry;

        if (entry & LITERAL_ENTRY) {
            size_t advance_output_bytes = (entry & 0xf00) >> 8;
            if (bb->nbits < litlen_code_bits) {
                break;
            } else if (output_index + 1 < output_len) {
                output[output_index] = (uint8_t)(entry >> 16);
                output[output_index + 1] = (uint8_t)(entry >> 24);
                output_index += advance_output_bytes;
                consume_bits(bb, litlen_code_bits);
                continue;
            } else if (output_index + advance_output_bytes == output_len) {
                output[output_index] = (uint8_t)(entry >> 16);
                output_index += 1;
                consume_bits(bb, litlen_code_bits);
                break;
            } else {
                output[output_index] = (uint8_t)(entry >> 16);
                d->queued_kind = Q_RLE;
                d->q_data = (uint8_t)(entry >> 24);
                d->q_length = 1;
                output_index += 1;
                consume_bits(bb, litlen_code_bits);
                break;
            }
        }

        uint32_t length_base;
        uint8_t length_extra_bits;
        if ((entry & EXCEPTIONAL_ENTRY) == 0) {
            length_base = entry >> 16;
            length_extra_bits = (uint8_t)(entry >> 8);
        } else if (entry & SECONDARY_TABLE_ENTRY) {
            uint32_t secondary_table_index =
                (entry >> 16) + ((uint32_t)(bits >> litlen_table_bits) & (entry & 0xff));
            uint16_t secondary_entry = d->secondary_table[secondary_table_index];
            uint16_t litlen_symbol = secondary_entry >> 4;
            uint8_t code_bits2 = (uint8_t)(secondary_entry & 0xf);

            if (bb->nbits < code_bits2) {
                break;
            } else if (litlen_symbol < 256) {
                consume_bits(bb, code_bits2);
                output[output_index] = (uint8_t)litlen_symbol;
                output_index += 1;
                continue;
            } else if (litlen_symbol == 256) {
                consume_bits(bb, code_bits2);
                *output_index_io = output_index;
                *block_status = REACHED_END_OF_BLOCK;
                return FDO_OK;
            }
            length_base = LEN_SYM_TO_LEN_BASE[litlen_symbol - 257];
            length_extra_bits = LEN_SYM_TO_LEN_EXTRA[litlen_symbol - 257];
            litlen_code_bits = code_bits2;
        } else if (litlen_code_bits == 0) {
            return FDO_INVALID_LITERAL_LENGTH_CODE;
        } else {
            if (bb->nbits < litlen_code_bits) break;
            consume_bits(bb, litlen_code_bits);
            *output_index_io = output_index;
            *block_status = REACHED_END_OF_BLOCK;
            return FDO_OK;
        }
        bits >>= litlen_code_bits;

        uint64_t length_extra_mask = ((uint64_t)1 << length_extra_bits) - 1;
        size_t length = (size_t)length_base + (size_t)(bits & length_extra_mask);
        bits >>= length_extra_bits;

        uint32_t dist_entry = d->dist_table[bits & dist_table_mask];
        uint16_t dist_base;
        uint8_t dist_extra_bits, dist_code_bits;
        if (dist_entry & LITERAL_ENTRY) {
            dist_base = (uint16_t)(dist_entry >> 16);
            dist_extra_bits = (uint8_t)(dist_entry >> 8) & 0xf;
            dist_code_bits = (uint8_t)dist_entry;
        } else if (bb->nbits > (unsigned)litlen_code_bits + length_extra_bits + dist_table_bits) {
            if ((dist_entry >> 8) == 0) return FDO_INVALID_DISTANCE_CODE;
            uint32_t secondary_table_index =
                (dist_entry >> 16) + ((uint32_t)(bits >> dist_table_bits) & (dist_entry & 0xff));
            uint16_t secondary_entry = d->dist_secondary_table[secondary_table_index];
            size_t dist_symbol = secondary_entry >> 4;
            if (dist_symbol >= 30) return FDO_INVALID_DISTANCE_CODE;
            dist_base = DIST_SYM_TO_DIST_BASE[dist_symbol];
            dist_extra_bits = DIST_SYM_TO_DIST_EXTRA[dist_symbol];
            dist_code_bits = (uint8_t)(secondary_entry & 0xf);
        } else {
            break;
        }
        bits >>= dist_code_bits;

        size_t dist = (size_t)dist_base + (size_t)(bits & (((uint64_t)1 << dist_extra_bits) - 1));
        unsigned total_bits = (unsigned)litlen_code_bits + length_extra_bits + dist_code_bits + dist_extra_bits;

        if (bb->nbits < total_bits) {
            break;
        } else if (dist > output_index) {
            return FDO_DISTANCE_TOO_FAR_BACK;
        }
        consume_bits(bb, (uint8_t)total_bits);

        if (do_copy(d, output, output_len, &output_index, length, dist)) break;
    }

    /* :1009-1015 EOB peek */
    if (d->queued_kind == Q_NONE && bb->nbits >= 15 &&
        ((uint16_t)peek_bits(bb, 15) & d->eof_mask) == d->eof_code) {
        consume_bits(bb, d->eof_bits);
        *output_index_io = output_index;
        *block_status = REACHED_END_OF_BLOCK;
        return FDO_OK;
    }

    *output_index_io = output_index;
    *block_status = MORE_DATA_PRESENT;
    return FDO_OK;
}

/* :179-337 Decompressor::read */
int fdo_decompressor_read(fdo_decompressor *d, const uint8_t *input, size_t input_len, uint8_t *output,
                          size_t output_len, size_t output_position, size_t *consumed, size_t *produced) {
    if (d->state == ST_DONE) {
        *consumed = 0;
        *produced = 0;
        return FDO_OK;
    }
    if (output_position > output_len) abort(); /* assert!, :189 */

    slice remaining_input = {input, input_len};
    size_t output_index = output_position;

    /* :194-219 drain queued output */
    if (d->queued_kind != Q_NONE) {
        int kind = d->queued_kind;
        d->queued_kind = Q_NONE;
        size_t length = d->q_length;
        size_t room = output_len - output_index;
        size_t n = length < room ? length : room;
        if (kind == Q_RLE) {
            memset(output + output_index, d->q_data, n);
        } else {
            for (size_t i = 0; i < n; i++) output[output_index + i] = output[output_index + i - d->q_dist];
        }
        output_index += n;
        if (length - n != 0) {
            d->queued_kind = kind;
            d->q_length = length - n;
            *consumed = 0;
            *produced = n;
            return FDO_OK; /* note: the reference returns before updating the checksum here (:203,:215) */
        }
    }

    /* :222-329 */
    int last_state = -1;
    while (last_state != d->state) {
        last_state = d->state;
        int st = FDO_OK;
        switch (d->state) {
        case ST_ZLIB_HEADER: { /* :226-244 */
            fill_buffer(&d->bits, &remaining_input);
            if (d->bits.nbits < 16) goto out_of_loop;
            uint64_t input0 = peek_bits(&d->bits, 8);
            uint64_t input1 = (peek_bits(&d->bits, 16) >> 8) & 0xff;
            if ((input0 & 0x0f) != 0x08 || (input0 & 0xf0) > 0x70 || (input1 & 0x20) != 0 ||
                ((input0 << 8) | input1) % 31 != 0)
                return FDO_BAD_ZLIB_HEADER;
            consume_bits(&d->bits, 16);
            d->state = ST_BLOCK_HEADER;
            break;
        }
        case ST_BLOCK_HEADER:
            st = read_block_header(d, &remaining_input);
            if (st != FDO_OK) return st;
            break;
        case ST_CODE_LENGTH_CODES:
            st = read_code_length_codes(d, &remaining_input);
            if (st != FDO_OK) return st;
            break;
        case ST_CODE_LENGTHS:
            st = read_code_lengths(d, &remaining_input);
            if (st != FDO_OK) return st;
            break;
        case ST_COMPRESSED_DATA: { /* :254-270 */
            int block_status = MORE_DATA_PRESENT;
            st = read_compressed(d, &remaining_input, output, output_len, &output_index, &block_status);
            if (st != FDO_OK) return st;
            if (block_status == REACHED_END_OF_BLOCK) d->state = d->last_block ? ST_CHECKSUM : ST_BLOCK_HEADER;
            break;
        }
        case ST_UNCOMPRESSED_DATA: { /* :271-305 */
            while (d->bits.nbits > 0 && d->uncompressed_bytes_left > 0 && output_index < output_len) {
                output[output_index] = (uint8_t)peek_bits(&d->bits, 8);
                consume_bits(&d->bits, 8);
                output_index += 1;
                d->uncompressed_bytes_left -= 1;
            }
            if (d->bits.nbits == 0) d->bits.buffer = 0;

            size_t copy_bytes = d->uncompressed_bytes_left;
            if (remaining_input.len < copy_bytes) copy_bytes = remaining_input.len;
            if (output_len - output_index < copy_bytes) copy_bytes = output_len - output_index;
            memcpy(output + output_index, remaining_input.ptr, copy_bytes);
            remaining_input.ptr += copy_bytes;
            remaining_input.len -= copy_bytes;
            output_index += copy_bytes;
            d->uncompressed_bytes_left = (uint16_t)(d->uncompressed_bytes_left - copy_bytes);

            if (d->uncompressed_bytes_left == 0) d->state = d->last_block ? ST_CHECKSUM : ST_BLOCK_HEADER;
            break;
        }
        case ST_CHECKSUM: { /* :306-326 */
            fill_buffer(&d->bits, &remaining_input);
            uint8_t align_bits = d->bits.nbits % 8;
            if (d->bits.nbits >= 32 + align_bits) {
                d->checksum = fdo_adler32(d->checksum, output + output_position, output_index - output_position);
                if (align_bits != 0) consume_bits(&d->bits, align_bits);
                if (!d->ignore_adler32 &&
                    __builtin_bswap32((uint32_t)peek_bits(&d->bits, 32)) != d->checksum)
                    return FDO_WRONG_CHECKSUM;
                d->state = ST_DONE;
                consume_bits(&d->bits, 32);
                goto out_of_loop;
            }
            break;
        }
        default:
            abort(); /* State::Done => unreachable!() */
        }
    }
out_of_loop:

    /* :331-333 */
    if (!d->ignore_adler32 && d->state != ST_DONE)
        d->checksum = fdo_adler32(d->checksum, output + output_position, output_index - output_position);

    *consumed = input_len - remaining_input.len;
    *produced = output_index - output_position;
    return FDO_OK;
}

/* :1111-1144 decompress_to_vec_bounded, with the Vec replaced by a caller buffer of capacity maxlen. */
int fdo_inflate_into(const uint8_t *input, size_t input_len, uint8_t *out, size_t maxlen, uint32_t flags,
                     size_t *out_len, size_t *consumed_out) {
    fdo_decompressor *d = fdo_decompressor_new();
    if (!d) abort();
    if (flags & FDO_FLAG_IGNORE_ADLER32) fdo_decompressor_ignore_adler32(d);
    size_t visible = maxlen < 1024 ? maxlen : 1024;
    size_t input_index = 0, output_index = 0;
    uint8_t dummy = 0;
    uint8_t *outp = out ? out : &dummy;
    int result;
    for (;;) {
        size_t consumed = 0, produced = 0;
        int st = fdo_decompressor_read(d, input + input_index, input_len - input_index, outp, visible, output_index,
                                       &consumed, &produced);
        if (st != FDO_OK) {
            result = st;
            break;
        }
        input_index += consumed;
        output_index += produced;
        if (fdo_decompressor_is_done(d)) {
            result = FDO_OK;
            break;
        } else if (output_index == maxlen) {
            result = FDO_OUTPUT_TOO_LARGE;
            break;
        } else if (output_index == visible) {
            size_t grown = output_index + 32 * 1024;
            visible = grown < maxlen ? grown : maxlen;
            continue;
        } else if (input_index == input_len) {
            result = FDO_INSUFFICIENT_INPUT;
            break;
        } else {
            abort(); /* unreachable!("Read() call violated post-condition") */
        }
    }
    fdo_decompressor_free(d);
    if (out_len) *out_len = output_index;
    if (consumed_out) *consumed_out = input_index;
    return result;
}

/* :1079-1087 decompress_to_vec */
int fdo_decompress_to_vec(const uint8_t *input, size_t input_len, uint32_t flags, uint8_t **out_p,
                          size_t *out_len_p) {
    fdo_decompressor *d = fdo_decompressor_new();
    if (!d) abort();
    if (flags & FDO_FLAG_IGNORE_ADLER32) fdo_decompressor_ignore_adler32(d);
    size_t cap = 1024;
    uint8_t *output = (uint8_t *)calloc(cap, 1);
    size_t input_index = 0, output_index = 0;
    int result;
    for (;;) {
        size_t consumed = 0, produced = 0;
        int st = fdo_decompressor_read(d, input + input_index, input_len - input_index, output, cap, output_index,
                                       &consumed, &produced);
        if (st != FDO_OK) {
            result = st;
            break;
        }
        input_index += consumed;
        output_index += produced;
        if (fdo_decompressor_is_done(d)) {
            result = FDO_OK;
            break;
        } else if (output_index == cap) {
            size_t ncap = output_index + 32 * 1024;
            output = (uint8_t *)realloc(output, ncap);
            memset(output + cap, 0, ncap - cap);
            cap = ncap;
            continue;
        } else if (input_index == input_len) {
            result = FDO_INSUFFICIENT_INPUT;
            break;
        } else {
            abort();
        }
    }
    fdo_decompressor_free(d);
    if (result != FDO_OK) {
        free(output);
        *out_p = NULL;
        *out_len_p = 0;
    } else {
        *out_p = output;
        *out_len_p = output_index;
    }
    return result;
}

void fdo_free(void *p) { free(p); }

/* public wrapper around build_tables (decompress.rs:561-606) for the table KATs */
int fdo_build_tables(size_t hlit, const uint8_t *code_lengths, uint32_t *litlen_table, uint32_t *dist_table,
                     uint16_t *secondary, size_t *secondary_len, uint16_t *dist_secondary,
                     size_t *dist_secondary_len) {
    ensure_init();
    uint16_t eof_code, eof_mask;
    uint8_t eof_bits;
    return build_tables_impl(hlit, code_lengths, litlen_table, dist_table, secondary, secondary_len, dist_secondary,
                             dist_secondary_len, &eof_code, &eof_mask, &eof_bits);
}

/* ------------------------------------------------------------------------------------------
 * compress/ultrafast.rs
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t buffer;
    uint8_t nbits;
    uint8_t *out;
    size_t cap, pos;
    int overflow;
} uf_writer;

static void uf_write_all(uf_writer *w, const uint8_t *p, size_t n) {
    if (w->pos + n > w->cap) {
        w->overflow = 1;
        return;
    }
    memcpy(w->out + w->pos, p, n);
    w->pos += n;
}

/* ultrafast.rs:16-29 */
static void uf_write_bits(uf_writer *w, uint64_t bits, uint8_t nbits) {
    w->buffer |= bits << w->nbits;
    w->nbits += nbits;
    if (w->nbits >= 64) {
        uint8_t le[8];
        memcpy(le, &w->buffer, 8);
        uf_write_all(w, le, 8);
        w->nbits -= 64;
        unsigned sh = (unsigned)(nbits - w->nbits);
        w->buffer = sh < 64 ? bits >> sh : 0; /* checked_shr(..).unwrap_or(0) */
    }
}

/* ultrafast.rs:31-43 */
static void uf_flush(uf_writer *w) {
    if (w->nbits % 8 != 0) uf_write_bits(w, 0, (uint8_t)(8 - w->nbits % 8));
    if (w->nbits > 0) {
        uint8_t le[8];
        memcpy(le, &w->buffer, 8);
        uf_write_all(w, le, w->nbits / 8);
        w->buffer = 0;
        w->nbits = 0;
    }
}

/* ultrafast.rs:45-67 */
static void uf_write_run(uf_writer *w, uint32_t run) {
    uf_write_bits(w, HUFFMAN_CODES[0], HUFFMAN_LENGTHS[0]);
    run -= 1;
    while (run >= 258) {
        uf_write_bits(w, HUFFMAN_CODES[285], (uint8_t)(HUFFMAN_LENGTHS[285] + 1));
        run -= 258;
    }
    if (run > 4) {
        unsigned sym = LENGTH_TO_SYMBOL[run - 3];
        uf_write_bits(w, HUFFMAN_CODES[sym], HUFFMAN_LENGTHS[sym]);
        uint8_t len_extra = LENGTH_TO_LEN_EXTRA[run - 3];
        uint64_t extra = (uint64_t)((run - 3) & ((1u << len_extra) - 1));
        uf_write_bits(w, extra, (uint8_t)(len_extra + 1));
    } else {
        uf_write_bits(w, 0, (uint8_t)(run * HUFFMAN_LENGTHS[0]));
    }
}

static unsigned tz_bytes(uint64_t v) { return (unsigned)__builtin_ctzll(v) / 8; }
static unsigned lz_bytes(uint64_t v) { return (unsigned)__builtin_clzll(v) / 8; }

size_t fdo_ultrafast_bound(size_t n) { return 54 + (n * 12 + 7) / 8 + 2 + 4 + 8; }

/* UltraFastCompressor, call by call (ultrafast.rs:9-181).  The zero-run counter and the 8-byte chunking are LOCAL to
 * a write_data call (:97-99: `let mut run = 0; let mut chunks = data.chunks_exact(8)`), so the bytes produced depend on
 * how the input is cut into calls; the bit buffer and the checksum carry over. */
struct fdo_ultrafast {
    uf_writer w;
    uint32_t checksum;
};

/* new (:70-79) + write_headers (:81-91) */
fdo_ultrafast *fdo_ultrafast_new(uint8_t *out, size_t out_cap) {
    ensure_init();
    fdo_ultrafast *c = (fdo_ultrafast *)calloc(1, sizeof *c);
    if (!c) return NULL;
    c->w.out = out;
    c->w.cap = out_cap;
    c->checksum = 1; /* Adler32::new() */
    uf_write_all(&c->w, UF_HEADER, 53);
    uf_write_bits(&c->w, UF_HEADER[53], 5);
    return c;
}

/* write_data (:94-167) */
void fdo_ultrafast_write_data(fdo_ultrafast *c, const uint8_t *data, size_t n) {
    uf_writer *w = &c->w;
    c->checksum = fdo_adler32(c->checksum, data, n); /* :95 */
    uint32_t run = 0;
    size_t nchunks = n / 8;
    for (size_t k8 = 0; k8 < nchunks; k8++) {
        const uint8_t *chunk = data + 8 * k8;
        uint64_t ichunk;
        memcpy(&ichunk, chunk, 8);

        if (ichunk == 0) {
            run += 8;
            continue;
        } else if (run > 0) {
            uint32_t run_extra = tz_bytes(ichunk);
            uf_write_run(w, run + run_extra);
            run = 0;
            if (run_extra > 0) {
                run = lz_bytes(ichunk);
                for (unsigned k = run_extra; k < 8 - run; k++)
                    uf_write_bits(w, HUFFMAN_CODES[chunk[k]], HUFFMAN_LENGTHS[chunk[k]]);
                continue;
            }
        }

        uint32_t run_start = lz_bytes(ichunk);
        if (run_start > 0) {
            for (unsigned k = 0; k < 8 - run_start; k++)
                uf_write_bits(w, HUFFMAN_CODES[chunk[k]], HUFFMAN_LENGTHS[chunk[k]]);
            run = run_start;
            continue;
        }

        uint8_t n0 = HUFFMAN_LENGTHS[chunk[0]], n1 = HUFFMAN_LENGTHS[chunk[1]];
        uint8_t n2 = HUFFMAN_LENGTHS[chunk[2]], n3 = HUFFMAN_LENGTHS[chunk[3]];
        uint64_t bits = (uint64_t)HUFFMAN_CODES[chunk[0]] | ((uint64_t)HUFFMAN_CODES[chunk[1]] << n0) |
                        ((uint64_t)HUFFMAN_CODES[chunk[2]] << (n0 + n1)) |
                        ((uint64_t)HUFFMAN_CODES[chunk[3]] << (n0 + n1 + n2));
        uf_write_bits(w, bits, (uint8_t)(n0 + n1 + n2 + n3));

        uint8_t n4 = HUFFMAN_LENGTHS[chunk[4]], n5 = HUFFMAN_LENGTHS[chunk[5]];
        uint8_t n6 = HUFFMAN_LENGTHS[chunk[6]], n7 = HUFFMAN_LENGTHS[chunk[7]];
        uint64_t bits2 = (uint64_t)HUFFMAN_CODES[chunk[4]] | ((uint64_t)HUFFMAN_CODES[chunk[5]] << n4) |
                         ((uint64_t)HUFFMAN_CODES[chunk[6]] << (n4 + n5)) |
                         ((uint64_t)HUFFMAN_CODES[chunk[7]] << (n4 + n5 + n6));
        uf_write_bits(w, bits2, (uint8_t)(n4 + n5 + n6 + n7));
    }
    if (run > 0) uf_write_run(w, run);
    for (size_t k = nchunks * 8; k < n; k++) uf_write_bits(w, HUFFMAN_CODES[data[k]], HUFFMAN_LENGTHS[data[k]]);
}

/* finish (:170-181): end of block, pad to a byte, checksum big-endian.  Returns the stream length (0 = the
 * buffer was too small) and frees the compressor. */
size_t fdo_ultrafast_finish(fdo_ultrafast *c) {
    uf_writer *w = &c->w;
    uf_write_bits(w, HUFFMAN_CODES[256], HUFFMAN_LENGTHS[256]);
    uf_flush(w);
    uint8_t be[4] = {(uint8_t)(c->checksum >> 24), (uint8_t)(c->checksum >> 16), (uint8_t)(c->checksum >> 8),
                     (uint8_t)c->checksum};
    uf_write_all(w, be, 4);
    size_t r = w->overflow ? 0 : w->pos;
    free(c);
    return r;
}

/* compress/mod.rs:313-317: new + write_data(whole input) + finish */
size_t fdo_compress_ultra_fast(const uint8_t *data, size_t n, uint8_t *out, size_t out_cap) {
    fdo_ultrafast *c = fdo_ultrafast_new(out, out_cap);
    if (!c) return 0;
    fdo_ultrafast_write_data(c, data, n);
    return fdo_ultrafast_finish(c);
}

/* ------------------------------------------------------------------------------------------
 * compress/mod.rs level 0 ("stored"): Compressor::new(w, 0, true) (:69-101), one write_data
 * (:126-156 -> CompressorInner::Uncompressed :241-268 with Flush::None), finish (:194-214 ->
 * :234-238 empty-final-fixed-block shortcut or :254-266 final stored block).
 * ---------------------------------------------------------------------------------------- */
size_t fdo_stored_bound(size_t n) { return 2 + 5 * (n / 65535 + 1) + n + 4; }

size_t fdo_compress_stored(const uint8_t *data, size_t n, uint8_t *out, size_t out_cap) {
    if (out_cap < fdo_stored_bound(n)) return 0;
    size_t pos = 0;
    out[pos++] = 0x78; /* :71 */
    out[pos++] = 0x01;
    uint32_t checksum = fdo_adler32(1, data, n); /* :137-139 */

    /* write_data -> compress(.., Flush::None) :241-268 */
    const uint8_t *input = data;
    size_t left = n;
    while (left > 65535) {
        out[pos++] = 0x00; /* write_bits(0,3) + flush() */
        out[pos++] = 0xff;
        out[pos++] = 0xff;
        out[pos++] = 0x00;
        out[pos++] = 0x00;
        memcpy(out + pos, input, 65535);
        pos += 65535;
        input += 65535;
        left -= 65535;
    }
    if (left == 65535) { /* :254: input.len() == STORED_BLOCK_MAX_SIZE with Flush::None => non-final block */
        out[pos++] = 0x00;
        out[pos++] = 0xff;
        out[pos++] = 0xff;
        out[pos++] = 0x00;
        out[pos++] = 0x00;
        memcpy(out + pos, input, 65535);
        pos += 65535;
        input += 65535;
        left = 0;
    }
    /* finish -> compress(.., Flush::Finish) on the unwritten remainder */
    if (left == 0) { /* :234-238 write_bits(3, 10): empty final fixed block */
        out[pos++] = 0x03;
        out[pos++] = 0x00;
    } else { /* :255-265 */
        out[pos++] = 0x01;
        out[pos++] = (uint8_t)(left & 0xff);
        out[pos++] = (uint8_t)(left >> 8);
        out[pos++] = (uint8_t)(~left & 0xff);
        out[pos++] = (uint8_t)((~left >> 8) & 0xff);
        memcpy(out + pos, input, left);
        pos += left;
    }
    out[pos++] = (uint8_t)(checksum >> 24); /* :208-210 */
    out[pos++] = (uint8_t)(checksum >> 16);
    out[pos++] = (uint8_t)(checksum >> 8);
    out[pos++] = (uint8_t)checksum;
    return pos;
}

/* ------------------------------------------------------------------------------------------
 * Multi-threaded batch drivers: one stream per task (the reference itself has no threads; this is
 * how a caller would spread independent streams over host cores).  Used as the CPU baseline.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int kind; /* 0 inflate, 1 ultra-fast compress, 2 synthetic tiles */
    const uint8_t *in_base;
    const uint64_t *in_off, *in_len;
    uint8_t *out_base;
    const uint64_t *out_off, *out_cap;
    uint64_t *out_len;
    int32_t *status;
    size_t n;
    uint32_t flags;
    /* kind 2 */
    uint64_t first_tile, seed;
    uint32_t width, height;
} batch_job;

static void synth_tile(uint8_t *out, uint64_t seed, uint64_t tile, uint32_t width, uint32_t height);

static void batch_item(const batch_job *j, size_t i) {
    if (j->kind == 0) {
        size_t olen = 0, consumed = 0;
        int st = fdo_inflate_into(j->in_base + j->in_off[i], (size_t)j->in_len[i], j->out_base + j->out_off[i],
                                  (size_t)j->out_cap[i], j->flags, &olen, &consumed);
        j->out_len[i] = olen;
        if (j->status) j->status[i] = st;
    } else if (j->kind == 1) {
        j->out_len[i] = fdo_compress_ultra_fast(j->in_base + j->in_off[i], (size_t)j->in_len[i],
                                                j->out_base + j->out_off[i], (size_t)j->out_cap[i]);
    } else {
        synth_tile(j->out_base + i * (size_t)j->height * (1u + 4u * (size_t)j->width), j->seed, j->first_tile + i,
                   j->width, j->height);
    }
}

/* A persistent pool of worker threads (created on first use, one set per thread count): a batch call posts its job,
 * wakes the workers and waits; the timed region of the CPU baseline therefore holds no thread creation. */
#define POOL_MAX 512
static struct {
    pthread_mutex_t mu;
    pthread_cond_t wake, done;
    pthread_t th[POOL_MAX];
    int nthreads;            /* workers alive */
    int want;                /* workers that take part in the current job */
    const batch_job *job;
    size_t next;             /* next item of the current job */
    unsigned long generation;
    int running;             /* workers still inside the current job */
} g_pool = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, {0}, 0, 0, NULL, 0, 0, 0};

static void *pool_worker(void *arg) {
    const int id = (int)(intptr_t)arg;
    unsigned long seen = 0;
    pthread_mutex_lock(&g_pool.mu);
    for (;;) {
        while (g_pool.generation == seen) pthread_cond_wait(&g_pool.wake, &g_pool.mu);
        seen = g_pool.generation;
        if (id >= g_pool.want) continue;
        const batch_job *j = g_pool.job;
        for (;;) {
            size_t i = g_pool.next;
            size_t end = i + 4 < j->n ? i + 4 : j->n;
            g_pool.next = end;
            if (i >= j->n) break;
            pthread_mutex_unlock(&g_pool.mu);
            for (; i < end; i++) batch_item(j, i);
            pthread_mutex_lock(&g_pool.mu);
        }
        if (--g_pool.running == 0) pthread_cond_signal(&g_pool.done);
    }
    return NULL;
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static double run_batch(batch_job *j, int nthreads) {
    ensure_init();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > POOL_MAX) nthreads = POOL_MAX;
    pthread_mutex_lock(&g_pool.mu);
    while (g_pool.nthreads < nthreads) {
        if (pthread_create(&g_pool.th[g_pool.nthreads], NULL, pool_worker, (void *)(intptr_t)g_pool.nthreads) != 0) break;
        pthread_detach(g_pool.th[g_pool.nthreads]);
        g_pool.nthreads++;
    }
    if (nthreads > g_pool.nthreads) nthreads = g_pool.nthreads;
    if (nthreads < 1) { /* no worker could be created: run inline */
        pthread_mutex_unlock(&g_pool.mu);
        double t0 = now_s();
        for (size_t i = 0; i < j->n; i++) batch_item(j, i);
        return now_s() - t0;
    }
    double t0 = now_s();
    g_pool.job = j;
    g_pool.next = 0;
    g_pool.want = nthreads;
    g_pool.running = nthreads;
    g_pool.generation++;
    pthread_cond_broadcast(&g_pool.wake);
    while (g_pool.running > 0) pthread_cond_wait(&g_pool.done, &g_pool.mu);
    double t1 = now_s();
    pthread_mutex_unlock(&g_pool.mu);
    return t1 - t0;
}

double fdo_inflate_batch(const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len, uint8_t *out_base,
                         const uint64_t *out_off, const uint64_t *out_cap, uint64_t *out_len, int32_t *status,
                         size_t n, uint32_t flags, int nthreads) {
    batch_job j = {0, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, status, n, flags, 0, 0, 0, 0};
    return run_batch(&j, nthreads);
}

double fdo_compress_ultra_fast_batch(const uint8_t *in_base, const uint64_t *in_off, const uint64_t *in_len,
                                     uint8_t *out_base, const uint64_t *out_off, const uint64_t *out_cap,
                                     uint64_t *out_len, size_t n, int nthreads) {
    batch_job j = {1, in_base, in_off, in_len, out_base, out_off, out_cap, out_len, NULL, n, 0, 0, 0, 0, 0};
    return run_batch(&j, nthreads);
}

/* ------------------------------------------------------------------------------------------
 * Synthetic PNG-filtered RGBA tiles: the benchmark / test input of SURVEY.md 8d, restated here so that the CPU legs
 * of bench.py need nothing but this file (the product library has the same generator in csrc/synth.cuh; the two are
 * compared byte for byte in tests/test_oracle.py).  Integer-only, counter-based:
 *   tile seed s = splitmix64(seed + tile); channel c of pixel (x, y) = (a_c x + b_c y + ((x y) >> 6) + noise) & 0xff
 *   with a_c, b_c in [0, 3] per tile and noise = hash % 5 - 2; A = 255; two constant-colour rectangles of 96 x 64;
 *   row 0 Sub-filtered (type 1), the others Paeth (type 4); a row is 1 type byte + 4 * width residual bytes.
 * ---------------------------------------------------------------------------------------- */
static uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
typedef struct {
    uint64_t s;
    uint32_t a[3], b[3], rx[2], ry[2], col[2];
} tile_par;
static uint32_t synth_raw(const tile_par *p, uint32_t x, uint32_t y, uint32_t c, uint32_t width) {
    if (c == 3) return 255u;
    for (int k = 0; k < 2; k++)
        if (x >= p->rx[k] && x < p->rx[k] + 96u && y >= p->ry[k] && y < p->ry[k] + 64u) return (p->col[k] >> (8u * c)) & 0xffu;
    uint64_t h = sm64(p->s + ((uint64_t)(y * width + x) * 4u + c) * 0x9E3779B97F4A7C15ull);
    int32_t noise = (int32_t)(h % 5u) - 2;
    return (uint32_t)((int32_t)(p->a[c] * x + p->b[c] * y + ((x * y) >> 6)) + noise) & 0xffu;
}
static uint32_t synth_paeth(uint32_t a, uint32_t b, uint32_t c) {
    int32_t pa = (int32_t)b - (int32_t)c, pb = (int32_t)a - (int32_t)c;
    int32_t pc = pa + pb;
    pa = pa < 0 ? -pa : pa;
    pb = pb < 0 ? -pb : pb;
    pc = pc < 0 ? -pc : pc;
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
static void synth_tile(uint8_t *out, uint64_t seed, uint64_t tile, uint32_t width, uint32_t height) {
    tile_par p;
    p.s = sm64(seed + tile);
    uint64_t h = sm64(p.s ^ 0x1234567ull);
    for (int c = 0; c < 3; c++) {
        p.a[c] = (uint32_t)(h >> (4 * c)) & 3u;
        p.b[c] = (uint32_t)(h >> (4 * c + 2)) & 3u;
    }
    for (int k = 0; k < 2; k++) {
        uint64_t r = sm64(p.s ^ (0xABCDEFull + (uint64_t)k));
        p.rx[k] = (uint32_t)(r & 0xffffu) % width;
        p.ry[k] = (uint32_t)((r >> 16) & 0xffffu) % height;
        p.col[k] = (uint32_t)(r >> 32) | 0xff000000u;
    }
    for (uint32_t y = 0; y < height; y++) {
        uint8_t *row = out + (size_t)y * (1u + 4u * (size_t)width);
        row[0] = y == 0 ? 1 : 4;
        for (uint32_t x = 0; x < width; x++)
            for (uint32_t c = 0; c < 4; c++) {
                uint32_t cur = synth_raw(&p, x, y, c, width);
                uint32_t left = x ? synth_raw(&p, x - 1, y, c, width) : 0u;
                uint32_t pred = left;
                if (y != 0) {
                    uint32_t up = synth_raw(&p, x, y - 1, c, width);
                    uint32_t ul = x ? synth_raw(&p, x - 1, y - 1, c, width) : 0u;
                    pred = synth_paeth(left, up, ul);
                }
                row[1 + 4 * x + c] = (uint8_t)(cur - pred);
            }
    }
}
void fdo_synth_tiles(uint8_t *out, uint64_t first_tile, uint64_t n_tiles, uint32_t width, uint32_t height,
                     uint64_t seed, int nthreads) {
    batch_job j = {2, NULL, NULL, NULL, out, NULL, NULL, NULL, NULL, (size_t)n_tiles, 0, first_tile, seed, width, height};
    run_batch(&j, nthreads);
}

int fdo_hardware_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
