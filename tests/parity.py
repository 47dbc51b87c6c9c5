"""Parity checkers: the device path (through the C ABI) against the CPU oracle, bit for bit."""
from __future__ import annotations

import zlib

import numpy as np

import oracle_lib as O
from fdeflate_b200 import STATUS_NAMES


def check_inflate(ctx, cases, flags=0, align=16, expect_general=None):
    """cases: list of (stream, cap).  Status must match the oracle; bytes must match when the oracle
    returns Ok or OutputTooLarge (the only cases in which the reference exposes output)."""
    streams = [c[0] for c in cases]
    caps = [c[1] for c in cases]
    st, outs, cons = ctx.inflate_batch(streams, caps, flags, align=align)
    problems = []
    for i, (s, cap) in enumerate(cases):
        est, eout, econs = O.inflate_into(s, cap, flags & 1)
        if st[i] != est:
            problems.append(f"#{i}: status {STATUS_NAMES[st[i]]} != oracle {STATUS_NAMES[est]} (in={len(s)} cap={cap})")
        elif est in (0, 17) and outs[i] != eout:
            a = np.frombuffer(outs[i], np.uint8)
            b = np.frombuffer(eout, np.uint8)
            m = min(a.size, b.size)
            d = np.nonzero(a[:m] != b[:m])[0]
            problems.append(f"#{i}: bytes differ (len {a.size} vs {b.size}, first diff {d[:3]})")
        elif est == 0 and not (cons[i] <= econs <= cons[i] + 7):
            # ours = exact end of the zlib stream; the reference's read() also counts the bytes its
            # 64-bit reservoir prefetched past the trailer (src/decompress.rs:1035-1052)
            problems.append(f"#{i}: consumed {cons[i]} vs oracle {econs}")
    if expect_general is not None:
        g = ctx.last_general_count()
        if g != expect_general:
            problems.append(f"general-kernel count {g} != {expect_general}")
    assert not problems, f"{len(problems)} of {len(cases)} mismatches:\n" + "\n".join(problems[:10])
    return st


def check_deflate_ultrafast(ctx, inputs, align=16):
    outs = ctx.deflate_ultrafast_batch(inputs, align=align)
    for i, (d, o) in enumerate(zip(inputs, outs)):
        ref = O.compress_ultra_fast(d)
        assert o == ref, f"#{i}: ultra-fast output differs from the oracle (n={len(d)}, {len(o)} vs {len(ref)} bytes)"
        assert zlib.decompress(o) == d
    return outs


def check_deflate_stored(ctx, inputs, align=16):
    outs = ctx.deflate_stored_batch(inputs, align=align)
    for i, (d, o) in enumerate(zip(inputs, outs)):
        assert o == O.compress_stored(d), f"#{i}: stored output differs from the oracle (n={len(d)})"
        assert zlib.decompress(o) == d
    return outs
