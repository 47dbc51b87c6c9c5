"""Multi-GPU path on CPU: world_size-2 gloo processes, each running the kernels on the test-only
SIMT emulator.  Checks the byte-balanced partition and that sharded results equal the oracle's and
the unsharded ones (no collective on the data path; only results are gathered)."""
import os
import random
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

import cases

HERE = Path(__file__).resolve().parent


def test_partition_lpt_is_balanced_and_deterministic():
    from fdeflate_b200.shard import partition_lpt

    rng = random.Random(1)
    costs = [int(2 ** rng.uniform(16, 24)) for _ in range(500)]  # 64 KB .. 16 MB, BASELINE config 5
    for world in (1, 2, 4, 8):
        parts = partition_lpt(costs, world)
        assert sorted(np.concatenate(parts).tolist()) == list(range(500))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs)
        assert all((a == b).all() for a, b in zip(parts, partition_lpt(costs, world)))
    assert [p.tolist() for p in partition_lpt([5, 5, 5, 5], 2)] == [[0, 2], [1, 3]]
    assert [p.size for p in partition_lpt([], 4)] == [0, 0, 0, 0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(HERE))
    sys.path.insert(0, str(HERE.parent))
    import torch.distributed as dist

    import oracle_lib as O
    from fdeflate_b200 import Context, NativeLib
    from fdeflate_b200.shard import shard_deflate_ultrafast, shard_inflate

    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = Context(0, NativeLib(HERE / "emul" / "libfdb_emul.so"))
    rng = random.Random(3)
    datas = [cases.sparse_bytes(rng, n) for n in (10, 3000, 500, 20000, 7, 9000, 100, 4000, 12000)]
    comp = shard_deflate_ultrafast(ctx, datas, rank, world)
    streams = [O.compress_ultra_fast(d) for d in datas]
    streams += [c[0] for c in cases.mixed_zlib_cases(9, 2, [100, 2000])]
    caps = [len(d) for d in datas] + [c[1] for c in cases.mixed_zlib_cases(9, 2, [100, 2000])]
    st, payload, out_len = shard_inflate(ctx, streams, caps, rank, world)
    ok = True
    if rank == 0:
        ok = comp == streams[: len(datas)]
        for i, (s, c) in enumerate(zip(streams, caps)):
            est, eout, _ = O.inflate_into(s, c)
            ok = ok and st[i] == est and (est not in (0, 17) or (payload[i] == eout and out_len[i] == len(eout)))
    q.put((rank, bool(ok), st.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_batch_world_size_2(emul_lib, oracle):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == res[1][2]  # every rank ends up with the same status vector
