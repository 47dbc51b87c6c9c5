"""Hand-crafted streams in the ultra-fast format (constant header + arbitrary tokens).

The reference's encoder only ever emits literals and distance-1 runs of ZEROS, but a decoder must
accept any token sequence under that header.  These helpers build such streams (non-zero runs, a
match as the first token, the invalid distance code, ...) together with their expected output.
Test infrastructure only.
"""
from __future__ import annotations

import json
import zlib
from pathlib import Path

GOLD = Path(__file__).resolve().parent / "golden"
_T = json.loads((GOLD / "reference_tables.json").read_text())
LENGTHS = _T["HUFFMAN_LENGTHS"]
HEADER = bytes(_T["ULTRAFAST_HEADER"])
LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195,
            227, 258]
LEN_EXTRA = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]


def _codes(lengths):
    code = 0
    out = [0] * len(lengths)
    for ln in range(1, 16):
        for i, l in enumerate(lengths):
            if l == ln:
                out[i] = int(format(code, f"0{ln}b")[::-1], 2)
                code += 1
        code <<= 1
    return out


CODES = _codes(LENGTHS)


class BitWriter:
    def __init__(self):
        self.acc = 0
        self.n = 0

    def put(self, v, nbits):
        self.acc |= (v & ((1 << nbits) - 1)) << self.n
        self.n += nbits

    def bytes(self):
        return self.acc.to_bytes((self.n + 7) // 8, "little")


def encode(tokens, dist_bit=0, adler=None, with_eob=True):
    """tokens: ints 0..255 = literal; ('m', length) = distance-1 match of 3..258 bytes.
    Returns (stream_bytes, expected_output_bytes or None if the stream is invalid)."""
    w = BitWriter()
    w.put(int.from_bytes(HEADER[:53], "little"), 53 * 8)
    w.put(HEADER[53], 5)
    out = bytearray()
    valid = dist_bit == 0 and with_eob and adler is None
    for t in tokens:
        if isinstance(t, int):
            w.put(CODES[t], LENGTHS[t])
            out.append(t)
        else:
            length = t[1]
            sym = max(i for i, b in enumerate(LEN_BASE) if b <= length) if length < 258 else 28
            w.put(CODES[257 + sym], LENGTHS[257 + sym])
            w.put(length - LEN_BASE[sym], LEN_EXTRA[sym])
            w.put(dist_bit, 1)
            if not out:
                valid = False
            else:
                out += bytes([out[-1]]) * length
    if with_eob:
        w.put(CODES[256], LENGTHS[256])
    body = w.bytes()
    a = zlib.adler32(bytes(out)) if adler is None else adler
    return body + a.to_bytes(4, "big"), (bytes(out) if valid else None)
