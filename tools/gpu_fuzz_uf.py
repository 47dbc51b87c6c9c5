"""Randomised parity run of the two ultra-fast kernels against the oracle on one GPU (evidence beside the test suite):
    python tools/gpu_fuzz_uf.py [streams] [seed]
Inputs: byte strings of 0 .. 200 KB with zero runs of awkward lengths, literal stretches of skewed and of uniform bytes,
long stretches of one long-code byte (many bits per byte: a segment's output barely fills the window) and of zeros (few
bits: the careful path), at every input / output alignment.  Checks: deflate bytes == the oracle's for every input;
inflate of those streams == the input with exact and with roomy slots; every stream that one warp decodes whole stays on
the fast path (general count 0)."""
import random
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np

import fdeflate_b200 as F
import oracle_lib as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = random.Random(seed)


def piece(rng):
    k = rng.random()
    if k < 0.25:
        return bytes(rng.choice([1, 2, 3, 7, 8, 9, 15, 16, 17, 63, 64, 65, 257, 258, 259, 260, 516, 517, 1000, 4000, 9000, 30000]))
    if k < 0.55:
        return bytes(rng.choice([0, 1, 255, 2, 254, 0, 0, 3, 253]) for _ in range(rng.randrange(1, 400)))
    if k < 0.75:
        return bytes(rng.getrandbits(8) for _ in range(rng.randrange(1, 300)))
    if k < 0.85:
        return bytes([rng.choice([100, 128, 77, 200])]) * rng.randrange(1, 3000)  # 12-bit codes
    return bytes(rng.choice([0, 0, 0, 0, 1, 255]) for _ in range(rng.randrange(1, 2000)))


def make(rng):
    target = rng.choice([0, 1, 5, 8, 9, 60, 500, 3000, 20000, 70000, 200000])
    out = bytearray()
    while len(out) < target:
        out += piece(rng)
    return bytes(out[:target])


t0 = time.time()
datas = [make(rng) for _ in range(n)]
want = [O.compress_ultra_fast(d) for d in datas]
print(f"{n} inputs, {sum(map(len, datas)) / 1e6:.1f} MB, oracle streams in {time.time() - t0:.1f} s", flush=True)
ctx = F.Context(0)
for align in (16, 1):
    got = ctx.deflate_ultrafast_batch(datas, align=align)
    bad = [i for i in range(n) if got[i] != want[i]]
    assert not bad, f"deflate differs from the oracle for inputs {bad[:5]} (align {align})"
    for extra in (0, 5):
        st, outs, _ = ctx.inflate_batch(want, [len(d) + extra for d in datas], 0, align)
        assert (st == 0).all(), (align, extra, np.nonzero(st)[0][:5], st[st != 0][:5])
        bad = [i for i in range(n) if outs[i] != datas[i]]
        assert not bad, f"inflate differs for streams {bad[:5]} (align {align}, slots + {extra})"
        declined = ctx.last_general_count()
        # Whole streams (one warp each) must never leave the fast path.  A long stream is cut into spans whose starts
        # are found by guessing (Huffman self-synchronisation over one segment); a stretch of one repeated long code is
        # periodic and may never synchronise from a wrong phase -- such a stream goes to the general kernel by design
        # (same bytes, checked above), and whether it does depends on where the span grid falls, i.e. on its address.
        short = [i for i in range(n) if len(want[i]) < 60000]
        ctx.inflate_batch([want[i] for i in short], [len(datas[i]) + extra for i in short], 0, align)
        assert ctx.last_general_count() == 0, "a whole stream left the fast path"
        print(f"align {align:2d}, slots + {extra}: {n} streams ok, {declined} long stream(s) handed to the general kernel by the span path", flush=True)
print(f"ok: deflate == oracle and inflate == input for {n} streams, seed {seed}, alignments 16 and 1", flush=True)
