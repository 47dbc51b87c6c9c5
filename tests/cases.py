"""Seeded test-case generators shared by the emulator (CPU) and GPU parity tests."""
from __future__ import annotations

import os
import random
import zlib
from pathlib import Path

import uf_craft

GOLD = Path(__file__).resolve().parent / "golden" / "reference_vectors"


def golden_streams():
    """(name, bytes) of the reference's own test vectors: fuzz corpus + tests/*.zz."""
    out = []
    d = GOLD / "fuzz_corpus_inflate"
    for f in sorted(os.listdir(d)):
        out.append((f"corpus/{f}", (d / f).read_bytes()))
    for k in (1, 2, 3):
        name = f"input-chunking-sensitivity-example{k}.zz"
        out.append((name, (GOLD / name).read_bytes()))
    return out


def payload(rng: random.Random, kind: int, n: int) -> bytes:
    if kind == 0:
        return bytes(rng.getrandbits(8) for _ in range(n))
    if kind == 1:
        return bytes(rng.choice(b"abcde") for _ in range(n))
    if kind == 2:  # periodic: overlapping and long-distance matches
        p = rng.choice([1, 2, 3, 4, 7, 8, 15, 16, 17, 31, 32, 33, 40, 258, 1000, 32768])
        base = bytes(rng.getrandbits(8) for _ in range(min(p, max(n, 1))))
        return (base * (n // len(base) + 1))[:n]
    if kind == 3:
        return bytes(n)
    if kind == 4:  # skewed alphabet: long litlen codes (13..15 bits) with zlib level >= 1
        return bytes(min(255, int(rng.expovariate(0.05))) for _ in range(n))
    return bytes(min(255, int(rng.expovariate(0.3))) for _ in range(n))


def zlib_stream(rng: random.Random, data: bytes) -> bytes:
    level = rng.choice([0, 1, 6, 9])
    strat = rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY, zlib.Z_FILTERED])
    wbits = rng.choice([9, 12, 15])
    co = zlib.compressobj(level, zlib.DEFLATED, wbits, rng.choice([1, 8, 9]), strat)
    step = rng.choice([1024, 4096, 16384, 10 ** 9])
    fl = rng.choice([None, zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH, zlib.Z_PARTIAL_FLUSH])
    out = b""
    for i in range(0, max(len(data), 1), step):
        out += co.compress(data[i:i + step])
        if fl is not None:
            out += co.flush(fl)
    return out + co.flush()


def mixed_zlib_cases(seed: int, count: int, sizes):
    """Valid, truncated, corrupted and capacity-limited zlib streams. -> list of (stream, cap)."""
    rng = random.Random(seed)
    cases = []
    for _ in range(count):
        n = rng.choice(sizes)
        data = payload(rng, rng.randrange(6), n)
        z = zlib_stream(rng, data)
        cases.append((z, len(data)))
        cases.append((z, max(0, len(data) - rng.choice([1, 2, 100]))))
        cases.append((z, len(data) + rng.choice([1, 1000])))
        if len(z) > 2:
            cases.append((z[: rng.randrange(len(z))], len(data) + 10))
            cases.append((z[: rng.randrange(len(z))], rng.randrange(len(data) + 1)))
            for _ in range(2):
                c = bytearray(z)
                for _ in range(rng.randrange(1, 4)):
                    c[rng.randrange(len(c))] ^= 1 << rng.randrange(8)
                cases.append((bytes(c), len(data) + rng.choice([0, 10, 5000])))
    return cases


def sparse_bytes(rng: random.Random, n: int) -> bytes:
    """zero runs of awkward lengths (around 258 multiples and 8-byte chunk edges) mixed with literals"""
    out = bytearray()
    while len(out) < n:
        k = rng.random()
        if k < 0.3:
            out += bytes(rng.choice([1, 3, 7, 8, 9, 15, 16, 17, 100, 257, 258, 259, 260, 515, 516, 517, 518, 519,
                                     1000, 2064, 5000, 20000]))
        elif k < 0.7:
            out += bytes(rng.choice([0, 1, 255, 2, 254, 0, 0, 3, 253]) for _ in range(rng.randrange(1, 200)))
        else:
            out += bytes(rng.getrandbits(8) for _ in range(rng.randrange(1, 100)))
    return bytes(out[:n])


def compress_inputs(seed: int, count: int, sizes):
    rng = random.Random(seed)
    fixed = [b"", b"Hello world!", bytes(7), bytes(8), bytes(9) + b"\x01", bytes(2048), bytes([5]) * 2048,
             bytes([128]) * 2048, bytes([254]) * 2048, bytes(600) + b"\x01" + bytes(515), b"\x01" + bytes(259),
             bytes(259), bytes(260), bytes(517), bytes(1000) + b"ab" + bytes(1030)]
    for n in (1, 15, 16, 17, 511, 512, 513, 1023, 1025):
        fixed.append(bytes(rng.choice([0, 0, 0, 1, 2, 255, 7]) for _ in range(n)))
        fixed.append(bytes(rng.getrandbits(8) for _ in range(n)))
    return fixed + [sparse_bytes(rng, rng.choice(sizes)) for _ in range(count)]


def crafted_uf_cases(seed: int, sizes=(0, 1, 2, 10, 100, 1000, 5000)):
    """streams with the ultra-fast header but token sequences the reference encoder never emits"""
    rng = random.Random(seed)

    def rand_tokens(n):
        t = []
        for _ in range(n):
            r = rng.random()
            if r < 0.15 and t:
                t.append(("m", rng.choice([3, 4, 5, 10, 11, 12, 18, 19, 66, 67, 130, 131, 257, 258,
                                           rng.randrange(3, 259)])))
            elif r < 0.5:
                t.append(rng.choice([0, 0, 1, 255, 2]))
            else:
                t.append(rng.getrandbits(8))
        return t

    cases = []
    for n in sizes:
        for _ in range(2):
            cases.append(uf_craft.encode(rand_tokens(n)))
    cases.append(uf_craft.encode([("m", 10), 5, 6]))                       # DistanceTooFarBack
    cases.append(uf_craft.encode([5, ("m", 258)] * 50, dist_bit=1))        # InvalidDistanceCode
    cases.append(uf_craft.encode([7] + [("m", 258)] * 400))                # long non-zero run
    cases.append(uf_craft.encode([0] + [("m", 258)] * 400 + [9]))
    cases.append(uf_craft.encode(rand_tokens(3000), with_eob=False))       # no end of block
    cases.append(uf_craft.encode(rand_tokens(500), adler=12345))           # wrong checksum
    return cases


def damaged(rng: random.Random, s: bytes, cap: int):
    """truncations, bit flips and trailing garbage of one stream -> list of (stream, cap)"""
    out = [(s[: rng.randrange(len(s))], cap), (s[: -rng.choice([1, 2, 3, 4, 5])], cap),
           (s + bytes(rng.getrandbits(8) for _ in range(rng.randrange(1, 20))), cap)]
    b = bytearray(s)
    b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
    out.append((bytes(b), cap + 100))
    if len(s) > 60:
        b = bytearray(s)
        b[rng.randrange(54, len(b))] ^= 1 << rng.randrange(8)
        out.append((bytes(b), cap + 100))
    return out


def uf_row_cases(seed: int):
    """Ultra-fast-format streams that stress the single-pass decoder's lane rows (inflate_uf.cuh): rows full of
    2-bit literals, zero runs kept in the row, runs that become gaps, more gaps than a lane's list holds, run
    chains, and mixtures.  -> (mixed, gap_heavy): lists of (stream, expected output); all stay on the fast path."""
    rng = random.Random(seed)
    lit = lambda: rng.choice([1, 2, 3, 254, 255, 7, 9, 130, rng.getrandbits(8) | 1])
    fast, overflow = [], []
    fast.append(uf_craft.encode([0] * 6000))                                   # 128 literals per lane
    fast.append(uf_craft.encode([rng.choice([0, 0, 0, 1, 255]) for _ in range(9000)]))
    t = []
    for _ in range(1500):                                                     # short runs: zeros inside the rows
        t += [lit() for _ in range(rng.randrange(0, 4))] + [0, ("m", rng.randrange(3, 41))]
    fast.append(uf_craft.encode(t))
    t = []
    for _ in range(600):                                                      # a few long runs per lane: gaps
        t += [lit() for _ in range(rng.randrange(10, 40))] + [0, ("m", rng.choice([150, 200, 257, 258]))]
        if rng.random() < 0.3:
            t += [("m", 258)] * rng.randrange(1, 4) + [("m", rng.randrange(3, 259))]
    fast.append(uf_craft.encode(t))
    t = []
    for _ in range(2500):                                                     # anything goes
        r = rng.random()
        if r < 0.25:
            t += [0, ("m", rng.choice([3, 4, 8, 12, 13, 31, 64, 100, 130, 258, rng.randrange(3, 259)]))]
        elif r < 0.3:
            t += [0] + [("m", 258)] * rng.randrange(1, 30)
        else:
            t += [rng.choice([0, 0, 1, 255, lit()]) for _ in range(rng.randrange(1, 30))]
    fast.append(uf_craft.encode(t))
    fast.append(uf_craft.encode([0] + [("m", 258)] * 3000 + [1, 2, 3]))          # one gap spanning many segments
    # a long run behind every other literal: as many gaps as a lane can ever see (each takes a word of its row)
    overflow.append(uf_craft.encode([x for _ in range(400) for x in (lit(), 0, ("m", 258))]))
    overflow.append(uf_craft.encode([x for _ in range(900) for x in (0, ("m", 258), 1, 0, ("m", 200), 0)]))
    return fast, overflow
