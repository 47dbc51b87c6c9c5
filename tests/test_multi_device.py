"""The device set behind the C ABI (fdb_multi_*): a batch is partitioned by stream over the contexts
of the set and every share runs the ordinary host-buffer call; results must equal the single-context results and the
oracle's, whatever the set.  On CPU the set is two contexts on the emulator's one device (world = 2); with -m gpu it
is two contexts on GPU 0 and, when the box has several GPUs, one context on each."""
import random
import zlib

import numpy as np
import pytest

import cases
import parity
from fdeflate_b200 import MultiContext
from fdeflate_b200.shard import partition_lpt


def _suite(multi, oracle, world):
    rng = random.Random(4)
    mixed = cases.mixed_zlib_cases(17, 6, [0, 100, 5000, 20000])
    uf = [(oracle.compress_ultra_fast(cases.sparse_bytes(rng, n)), n) for n in (0, 1, 300, 4000, 70000, 9, 2500, 130000)]
    batch = mixed + uf
    rng.shuffle(batch)
    parity.check_inflate(multi, batch, 0)
    # byte-balanced (cost = in_len + out_cap), every stream on exactly one device: contiguous runs of streams when they
    # balance within 3 % of the LPT partition of shard.py, else that partition itself
    owner = multi.last_partition(len(batch))
    cost = np.array([max(len(s) + c, 1) for s, c in batch])
    lpt = partition_lpt(cost, world)
    lpt_max = max(int(cost[p].sum()) for p in lpt)
    loads = [int(cost[owner == d].sum()) for d in range(world)]
    assert sum(loads) == int(cost.sum())
    if all(owner[i] <= owner[i + 1] for i in range(len(batch) - 1)):
        assert max(loads) <= lpt_max * 1.04
    else:
        for d in range(world):
            assert sorted(np.nonzero(owner == d)[0].tolist()) == lpt[d].tolist()
    inputs = cases.compress_inputs(7, 10, [100, 3000, 50000])
    parity.check_deflate_ultrafast(multi, inputs)
    parity.check_deflate_stored(multi, inputs[:12])
    loads = np.bincount(multi.last_partition(12), minlength=world)
    assert loads.sum() == 12
    # one long stream among short ones: contiguous runs cannot balance, the greedy partition is used and the shares
    # interleave in the caller's buffers (every device copies back exactly its own slots)
    skew = [(zlib.compress(cases.sparse_bytes(rng, 3000), 6), 3000) for _ in range(9)]
    skew.insert(5, (oracle.compress_ultra_fast(cases.sparse_bytes(rng, 200000)), 200000))
    skew.insert(6, (zlib.compress(bytes(9), 6), 9))
    parity.check_inflate(multi, skew, 0)
    owner = multi.last_partition(len(skew))
    if world == 2:  # (with many devices the long stream gets a device of its own either way)
        assert not all(owner[i] <= owner[i + 1] for i in range(len(skew) - 1))
    # fewer streams than devices, and none at all
    parity.check_deflate_ultrafast(multi, inputs[:1])
    assert multi.inflate_batch([], [])[1] == []
    # unaligned slots
    parity.check_inflate(multi, batch[:20], 0, align=1)


@pytest.mark.emul
def test_device_set_world_2_on_emulator(emul_lib, oracle):
    m = MultiContext([0, 0], emul_lib)
    try:
        _suite(m, oracle, 2)
    finally:
        m.close()
    with pytest.raises(Exception):
        MultiContext([0, 7], emul_lib)  # no such device


@pytest.mark.gpu
def test_device_set_on_gpu(oracle):
    import torch

    m = MultiContext([0, 0])
    try:
        _suite(m, oracle, 2)
    finally:
        m.close()
    n = torch.cuda.device_count()
    if n > 1:
        m = MultiContext(list(range(n)))
        try:
            _suite(m, oracle, n)
        finally:
            m.close()
