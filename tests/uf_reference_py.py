"""A second, independent restatement of the reference's UltraFastCompressor (src/compress/ultrafast.rs:16-181) in
pure Python, transliterated from the Rust (not from oracle/fdeflate_oracle.c): arbitrary-precision integer as the bit
sink instead of the 64-bit buffer, codes from tests/uf_craft.py (computed from HUFFMAN_LENGTHS by its own canonical-code
routine, src/lib.rs:103-127), the length tables from RFC 1951 3.2.5.  tests/test_oracle.py diffs it against the C oracle
over >= 10^4 random and adversarial inputs and call patterns: the strongest pin on the compressed bytes that exists
without a Rust toolchain (the reference's own tests only round-trip).  Test infrastructure only."""
from __future__ import annotations

import zlib

import uf_craft

CODES, LENGTHS, HEADER = uf_craft.CODES, uf_craft.LENGTHS, uf_craft.HEADER


def _length_symbol(length: int) -> tuple[int, int, int]:
    """RFC 1951 3.2.5 -> (symbol, extra bits, extra value) for a match length 3..258
    (what LENGTH_TO_SYMBOL / LENGTH_TO_LEN_EXTRA of src/tables.rs:28-55 tabulate, indexed by length - 3)"""
    if length == 258:
        return 285, 0, 0
    base, sym = 3, 257
    for extra in (0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5):
        if length < base + (1 << extra):
            return sym, extra, length - base
        base += 1 << extra
        sym += 1
    raise ValueError(length)


class UltraFastCompressorPy:
    def __init__(self):  # new (:70-79) + write_headers (:81-91)
        self.acc = int.from_bytes(HEADER[:53], "little")
        self.nbits = 53 * 8
        self.checksum = 1
        self._bits(HEADER[53], 5)

    def _bits(self, v: int, n: int):  # write_bits (:16-29), LSB first
        self.acc |= (v & ((1 << n) - 1)) << self.nbits
        self.nbits += n

    def _lit(self, b: int):
        self._bits(CODES[b], LENGTHS[b])

    def _run(self, run: int):  # write_run (:45-67)
        self._lit(0)
        run -= 1
        while run >= 258:
            self._bits(CODES[285], LENGTHS[285] + 1)  # code, then the 1-bit distance code 0
            run -= 258
        if run > 4:
            # LENGTH_TO_SYMBOL[run - 3], LENGTH_TO_LEN_EXTRA[run - 3]; extra = (run - 3) & BITMASKS[len_extra]
            sym, len_extra, _ = _length_symbol(run)
            self._bits(CODES[sym], LENGTHS[sym])
            self._bits((run - 3) & ((1 << len_extra) - 1), len_extra + 1)
        else:
            self._bits(0, run * LENGTHS[0])

    def write_data(self, data: bytes):  # :94-167
        self.checksum = zlib.adler32(data, self.checksum)
        run = 0
        n8 = len(data) // 8 * 8
        for i in range(0, n8, 8):
            chunk = data[i:i + 8]
            ichunk = int.from_bytes(chunk, "little")
            if ichunk == 0:
                run += 8
                continue
            elif run > 0:
                run_extra = ((ichunk & -ichunk).bit_length() - 1) // 8  # trailing_zeros / 8
                self._run(run + run_extra)
                run = 0
                if run_extra > 0:
                    run = (64 - ichunk.bit_length()) // 8  # leading_zeros / 8
                    for b in chunk[run_extra:8 - run]:
                        self._lit(b)
                    continue
            run_start = (64 - ichunk.bit_length()) // 8
            if run_start > 0:
                for b in chunk[:8 - run_start]:
                    self._lit(b)
                run = run_start
                continue
            for b in chunk:  # (:134-152 packs them four at a time; same bits)
                self._lit(b)
        if run > 0:
            self._run(run)
        for b in data[n8:]:
            self._lit(b)

    def finish(self) -> bytes:  # :170-181
        self._lit(256)
        nbytes = (self.nbits + 7) // 8
        return self.acc.to_bytes(nbytes, "little") + self.checksum.to_bytes(4, "big")


def compress_calls(calls) -> bytes:
    c = UltraFastCompressorPy()
    for d in calls:
        c.write_data(bytes(d))
    return c.finish()
