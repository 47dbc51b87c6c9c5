#!/usr/bin/env python3
"""bench.py -- batch inflate & ultra-fast deflate GB/s (uncompressed) on N B200s.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A step = one pass of the hot path over one batch per GPU: batch INFLATE of `tiles` ultra-fast zlib
streams (BASELINE configs[1]) followed by batch ULTRA-FAST DEFLATE of the same `tiles` 256x256 RGBA
PNG-filtered tiles (configs[2]).  `value` counts uncompressed bytes through both directions, inputs
and outputs resident in HBM.  `e2e` is the same step through the host-buffer C-ABI calls
(fdb_inflate_batch / fdb_deflate_ultrafast_batch) from pinned host memory, copies included.
Streams shard by index across ranks with no collective (weak scaling: every GPU gets `tiles` streams);
torch.distributed (NCCL) is used only for the barrier and the max-over-ranks of the device time.

The same invocation also runs BASELINE configs[4] at one GPU's share per GPU (`sweep` in the JSON line): ragged
streams of 64 KiB..16 MiB, byte-balanced over the ranks by fdeflate_b200/shard.py, long streams cut into spans /
segments, device-resident, both directions.

`--impl reference` times the reference algorithm on the host cores (the C oracle, which is a
line-by-line restatement of the Rust crate: no Rust toolchain exists in this image, so the crate
itself cannot be built) on a bounded sample of the same workload.  That arm loads nothing but the oracle:
its input comes from the oracle's own tile generator.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "batch inflate & ultra-fast deflate GB/s (uncompressed)"
UNIT = "GB/s"
TILE_W = TILE_H = 256
TILE_BYTES = TILE_H * (1 + 4 * TILE_W)  # 262400


def config_dict(tiles: int, n_gpus: int, extra: dict | None = None) -> dict:
    d = {
        "workload": "BASELINE configs[1]+[2]: per GPU, batch inflate of %d ultra-fast zlib streams of 256x256 RGBA "
                    "PNG-filtered tiles, then batch ultra-fast deflate of the same %d tiles" % (tiles, tiles),
        "streams_per_gpu": tiles,
        "tile_bytes": TILE_BYTES,
        "uncompressed_bytes_per_gpu_step": 2 * tiles * TILE_BYTES,
        "sharding": "by stream, no collective" if n_gpus > 1 else "single GPU",
        "cache": "inputs exceed L2 (%.2f GB per GPU per step vs 126 MB)" % (tiles * TILE_BYTES * 1.4 / 1e9),
    }
    if extra:
        d.update(extra)
    return d


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU while the timed region runs"""

    def __init__(self, index: int, period_s: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.window = [0.0, float("inf")]  # only samples taken inside [t0, t1] (perf_counter) are reported
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
        }
        while not self._stop_evt.is_set():
            try:
                mhz = int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, r))
            except Exception:
                pass
            time.sleep(self.period)
        self._names = names

    def stop(self) -> dict:
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        inside = [x for x in self.samples if self.window[0] <= x[0] <= self.window[1]]
        for _, _, r in inside:
            for k, bit in getattr(self, "_names", {}).items():
                if r & bit:
                    self.reasons.add(k)
        s = sorted(x[1] for x in inside)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# CPU legs: the oracle (reference algorithm restated in C), all host threads, bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_sample(sample_tiles: int, seed_tile: int = 0):
    """host tiles + their ultra-fast streams for the CPU legs: generated by the oracle's own tile generator (the same
    bytes as the product's, tests/test_oracle.py) and compressed by the oracle itself -- the product library is not
    loaded on this path"""
    import numpy as np

    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as O

    try:
        O.lib(native=True)
        native = True
    except Exception:
        O.lib(native=False)
        native = False
    threads = O.hardware_threads()
    tiles = O.synth_tiles(seed_tile, sample_tiles, TILE_W, TILE_H, 2024, threads, native).reshape(-1)
    bound = (54 + TILE_BYTES * 12 // 8 + 16 + 15) // 16 * 16
    n = sample_tiles
    in_off = np.arange(n, dtype=np.uint64) * TILE_BYTES
    in_len = np.full(n, TILE_BYTES, dtype=np.uint64)
    c_off = np.arange(n, dtype=np.uint64) * bound
    c_cap = np.full(n, bound, dtype=np.uint64)
    comp = np.zeros(n * bound, dtype=np.uint8)
    _, c_len = O.compress_ultra_fast_batch(tiles, in_off, in_len, comp, c_off, c_cap, threads, native)
    out = np.zeros(n * TILE_BYTES, dtype=np.uint8)
    return dict(O=O, native=native, threads=threads, tiles=tiles, comp=comp, out=out, in_off=in_off, in_len=in_len,
                c_off=c_off, c_cap=c_cap, c_len=c_len, n=n)


def cpu_step(s) -> tuple[float, float]:
    """one CPU step on the sample: (inflate seconds, deflate seconds)"""
    O = s["O"]
    t_inf, out_len, status = O.inflate_batch(s["comp"], s["c_off"], s["c_len"], s["out"], s["in_off"], s["in_len"], 0,
                                             s["threads"], s["native"])
    assert (status == 0).all() and (out_len == TILE_BYTES).all()
    comp2 = s.setdefault("comp2", s["comp"].copy())
    t_def, c_len2 = O.compress_ultra_fast_batch(s["tiles"], s["in_off"], s["in_len"], comp2, s["c_off"], s["c_cap"],
                                                s["threads"], s["native"])
    assert (c_len2 == s["c_len"]).all()
    return t_inf, t_def


def cpu_baseline_measure(budget_s: float = 12.0, sample_tiles: int = 256) -> dict:
    s = cpu_sample(sample_tiles)
    cpu_step(s)  # warm-up
    t_inf = t_def = 0.0
    reps = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget_s or reps < 2:
        a, b = cpu_step(s)
        t_inf += a
        t_def += b
        reps += 1
    nbytes = reps * s["n"] * TILE_BYTES
    return {
        "value": round(2 * nbytes / (t_inf + t_def) / 1e9, 3), "unit": UNIT, "cores": s["threads"], "kind": "port",
        "sample": "%d tiles (%.0f MB uncompressed) x %d passes, inflate + ultra-fast deflate, C oracle%s, "
                  "one stream per task on %d threads" % (s["n"], s["n"] * TILE_BYTES / 1e6, reps,
                                                        " -march=native" if s["native"] else "", s["threads"]),
        "inflate_gbs": round(nbytes / t_inf / 1e9, 3), "deflate_gbs": round(nbytes / t_def / 1e9, 3),
    }


def reference_arm(args, rank: int, world: int):
    if rank != 0:
        return
    sample_tiles = min(args.tiles, 512)
    s = cpu_sample(sample_tiles)
    for _ in range(max(args.warmup, 1)):
        cpu_step(s)
    t_inf = t_def = 0.0
    for _ in range(args.steps):
        a, b = cpu_step(s)
        t_inf += a
        t_def += b
    nbytes = args.steps * s["n"] * TILE_BYTES
    value = 2 * nbytes / (t_inf + t_def) / 1e9
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * (t_inf + t_def) / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_dict(args.tiles, args.gpus, {
            "reference_sample_streams": s["n"],
            "reference_sample": "each timed step of this arm runs %d of the %d streams (a bounded sample of the same "
                                "workload; GB/s does not depend on the count)" % (s["n"], args.tiles)}),
        "cpu_baseline": {
            "value": round(value, 3), "unit": UNIT, "cores": s["threads"], "kind": "port",
            "sample": "each step = %d tiles (%.0f MB uncompressed): inflate + ultra-fast deflate with the C oracle%s "
                      "(line-by-line restatement of the Rust crate; no Rust toolchain in this image), one stream per "
                      "task on %d host threads" % (s["n"], s["n"] * TILE_BYTES / 1e6,
                                                   " built -march=native" if s["native"] else "", s["threads"]),
        },
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "inflate_gbs": round(nbytes / t_inf / 1e9, 3), "deflate_gbs": round(nbytes / t_def / 1e9, 3),
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# BASELINE configs[4]: the stream-sharded sweep (ragged streams of 64 KiB..16 MiB), one GPU's share per GPU
# ------------------------------------------------------------------------------------------------
SWEEP_W = 1024
SWEEP_ROW = 1 + 4 * SWEEP_W


def sweep_sizes(total_streams: int):
    """the batch every rank derives for itself: stream g has sweep_heights[g] rows of 1024 RGBA pixels (log-uniform
    64 KiB..16 MiB, seeded), its content is tile 1000 + g of the synthetic generator"""
    import numpy as np

    rng = np.random.default_rng(5)
    sizes = np.exp(rng.uniform(np.log(64 << 10), np.log(16 << 20), total_streams))
    heights = np.maximum(1, (sizes / SWEEP_ROW).astype(np.int64))
    return heights, heights * SWEEP_ROW


def sweep(ctx, args, rank: int, world: int, dev, dist, stream, peak: float) -> dict | None:
    import numpy as np
    import torch

    import fdeflate_b200 as F
    from fdeflate_b200.shard import partition_lpt

    per = args.sweep_streams
    if per <= 0:
        return None
    total_n = per * world
    heights, lens_all = sweep_sizes(total_n)
    parts = partition_lpt(lens_all, world)  # byte-balanced, deterministic: every rank computes the same partition
    mine = parts[rank]
    n = int(mine.size)
    lens = lens_all[mine]
    i64 = torch.int64
    offs = np.zeros(n, dtype=np.int64)
    offs[1:] = np.cumsum((lens[:-1] + 15) & ~15)
    total = int(offs[-1] + lens[-1])
    raw = torch.empty(total + 16, dtype=torch.uint8, device=dev)
    for k in range(n):
        ctx.synth_tiles_device(raw.data_ptr() + int(offs[k]), 1000 + int(mine[k]), 1, SWEEP_W, int(heights[mine[k]]), 5, stream)
    bounds = np.array([ctx.ultrafast_bound(int(l)) for l in lens], dtype=np.int64)
    coffs = np.zeros(n, dtype=np.int64)
    coffs[1:] = np.cumsum(bounds[:-1])
    comp = torch.empty(int(coffs[-1] + bounds[-1]), dtype=torch.uint8, device=dev)
    out = torch.empty(total + 16, dtype=torch.uint8, device=dev)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_off, d_len, d_coff, d_ccap = T(offs), T(lens), T(coffs), T(bounds)
    c_len = torch.zeros(n, dtype=i64, device=dev)
    c_st = torch.zeros(n, dtype=torch.int32, device=dev)
    o_len = torch.zeros(n, dtype=i64, device=dev)
    o_st = torch.zeros(n, dtype=torch.int32, device=dev)
    ctx.set_split_large(True)  # long streams: spans (inflate) / segments (deflate, by the batch policy)

    def deflate():
        ctx.deflate_ultrafast_device(raw.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), comp.data_ptr(), d_coff.data_ptr(),
                                     d_ccap.data_ptr(), c_len.data_ptr(), c_st.data_ptr(), n, stream)

    def inflate():
        ctx.inflate_device(comp.data_ptr(), d_coff.data_ptr(), c_len.data_ptr(), out.data_ptr(), d_off.data_ptr(), d_len.data_ptr(),
                           o_len.data_ptr(), 0, o_st.data_ptr(), n, F.FLAG_SPLIT_LARGE, stream)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(f, reps):
        f()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            f()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / reps

    launches0 = ctx.launch_count
    ms_def = timed(deflate, args.sweep_steps)
    # a device-pointer call does not see the stream sizes: the caller sizes the pool of lane records the span-by-span
    # inflate keeps between its two passes (1/8 of the compressed bytes + one partial span per stream)
    ctx.set_split_scratch(int(c_len.sum()) // 8 + n * 8192 + (1 << 20))
    ms_inf = timed(inflate, args.sweep_steps)
    launches = ctx.launch_count - launches0
    spans = ctx.last_split_spans(stream)
    assert int(c_st.abs().sum()) == 0 and int(o_st.abs().sum()) == 0, "sweep: a stream failed"
    assert ctx.last_general_count(stream) == 0, "sweep: ultra-fast streams left the fast path"
    assert torch.equal(o_len, d_len)
    pick = np.random.default_rng().choice(n, size=min(n, 24), replace=False)  # a fresh sample every run
    for k in pick:
        a = int(offs[k])
        assert torch.equal(out[a:a + int(lens[k])], raw[a:a + int(lens[k])]), "sweep: inflate(deflate(x)) != x"
    # ... and against the oracle (the checker), on the shortest of the sampled streams
    checked = 0
    try:
        sys.path.insert(0, str(ROOT / "tests"))
        import oracle_lib as O

        for k in sorted(pick, key=lambda k: lens[k])[:3]:
            a, c0, cl = int(offs[k]), int(coffs[k]), int(c_len[k])
            src = raw[a:a + int(lens[k])].cpu().numpy().tobytes()
            z = comp[c0:c0 + cl].cpu().numpy().tobytes()
            assert z == O.compress_ultra_fast(src), "sweep: deflate output differs from the oracle"
            assert O.inflate_into(z, len(src))[:2] == (0, src), "sweep: the oracle inflates the stream differently"
            checked += 1
    except ImportError:
        pass
    ctx.set_split_large(False)
    ctx.set_split_scratch(256 << 20)
    unc = int(lens.sum())
    cbytes = int(c_len.sum())
    t = torch.tensor([ms_inf, ms_def], dtype=torch.float64, device=dev)
    b = torch.tensor([float(unc), float(cbytes)], dtype=torch.float64, device=dev)
    mx = torch.tensor([float(unc)], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    ms_inf, ms_def = [float(x) for x in t.cpu()]
    unc_all, comp_all = [float(x) for x in b.cpu()]
    del raw, comp, out
    torch.cuda.empty_cache()
    alg = (unc_all + comp_all) / world  # per GPU and launch set, either direction
    return {
        "workload": "BASELINE configs[4] at one GPU's share per GPU (weak scaling): %d ultra-fast streams of 64 KiB..16 MiB "
                    "(log-uniform) per GPU, %d in all, byte-balanced over the ranks (LPT, fdeflate_b200/shard.py), inflate "
                    "with long streams cut into spans + ultra-fast deflate, device-resident" % (per, total_n),
        "streams_per_gpu": per, "streams_total": total_n,
        "uncompressed_bytes_total": int(unc_all), "compressed_ratio": round(comp_all / unc_all, 4),
        "balance_max_over_mean": round(float(mx.cpu()[0]) / (unc_all / world), 4),
        "value": round(2 * unc_all / ((ms_inf + ms_def) / 1e3) / 1e9, 2), "unit": UNIT,
        "inflate_gbs": round(unc_all / (ms_inf / 1e3) / 1e9, 2), "deflate_gbs": round(unc_all / (ms_def / 1e3) / 1e9, 2),
        "ms_inflate": round(ms_inf, 3), "ms_deflate": round(ms_def, 3), "steps": args.sweep_steps,
        "roofline_frac_per_gpu": {"inflate": round(alg / (ms_inf / 1e3) / 1e9 / peak, 4),
                                  "deflate": round(alg / (ms_def / 1e3) / 1e9 / peak, 4)},
        "spans_rank0": int(spans), "gpu_launches": int(launches),
        "verified": "every status Ok, lengths equal, %d random streams per rank byte-equal after inflate(deflate(x)), "
                    "%d of them equal to the oracle's deflate bytes and inflated identically by the oracle" % (len(pick), checked),
    }


def link_ceiling(dev, dist) -> dict:
    """what the host link of this box gives with every rank copying at once: 1 GiB pinned, both directions together"""
    import torch

    n = 1 << 30
    h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_a.fill_(1)
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.zeros(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)

    both()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    reps = 4
    t0 = time.perf_counter()
    for _ in range(reps):
        both()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rate = torch.tensor([reps * n / dt / 1e9], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(rate, op=dist.ReduceOp.SUM)
    return {"gbs_per_direction_all_gpus": round(float(rate.cpu()[0]), 1),
            "how": "1 GiB pinned host buffers, cudaMemcpyAsync both directions at once on every rank, %d repetitions" % reps}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def ours(args, rank: int, local_rank: int, world: int):
    import numpy as np
    import torch

    import fdeflate_b200 as F

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    ctx = F.Context(local_rank)  # raises without the CUDA library / a GPU: there is no fallback
    n = args.tiles
    stream = torch.cuda.current_stream().cuda_stream
    i64 = torch.int64

    # ---- set-up (untimed): synthetic tiles on the device, their ultra-fast streams ----
    tiles = torch.empty(n * TILE_BYTES, dtype=torch.uint8, device=dev)
    ctx.synth_tiles_device(tiles.data_ptr(), rank * n, n, TILE_W, TILE_H, 2024, stream)
    bound = ctx.ultrafast_bound(TILE_BYTES)
    t_off = torch.arange(n, dtype=i64, device=dev) * TILE_BYTES
    t_len = torch.full((n,), TILE_BYTES, dtype=i64, device=dev)
    c_off = torch.arange(n, dtype=i64, device=dev) * bound
    c_cap = torch.full((n,), bound, dtype=i64, device=dev)
    comp = torch.zeros(n * bound, dtype=torch.uint8, device=dev)
    c_len = torch.zeros(n, dtype=i64, device=dev)
    c_st = torch.zeros(n, dtype=torch.int32, device=dev)
    ctx.deflate_ultrafast_device(tiles.data_ptr(), t_off.data_ptr(), t_len.data_ptr(), comp.data_ptr(),
                                 c_off.data_ptr(), c_cap.data_ptr(), c_len.data_ptr(), c_st.data_ptr(), n, stream)
    torch.cuda.synchronize()
    assert int((c_st != 0).sum()) == 0
    comp_bytes = int(c_len.sum())

    out = torch.empty(n * TILE_BYTES, dtype=torch.uint8, device=dev)
    o_len = torch.zeros(n, dtype=i64, device=dev)
    o_st = torch.zeros(n, dtype=torch.int32, device=dev)
    comp2 = torch.zeros(n * bound, dtype=torch.uint8, device=dev)
    c_len2 = torch.zeros(n, dtype=i64, device=dev)
    c_st2 = torch.zeros(n, dtype=torch.int32, device=dev)

    def step_device(ev=None):
        if ev:
            ev[0].record()
        ctx.inflate_device(comp.data_ptr(), c_off.data_ptr(), c_len.data_ptr(), out.data_ptr(), t_off.data_ptr(),
                           t_len.data_ptr(), o_len.data_ptr(), 0, o_st.data_ptr(), n, 0, stream)
        if ev:
            ev[1].record()
        ctx.deflate_ultrafast_device(tiles.data_ptr(), t_off.data_ptr(), t_len.data_ptr(), comp2.data_ptr(),
                                     c_off.data_ptr(), c_cap.data_ptr(), c_len2.data_ptr(), c_st2.data_ptr(), n, stream)
        if ev:
            ev[2].record()

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)  # runs from before the warm-up; only its samples inside the timed region count
    sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = ctx.launch_count
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    sampler.window[0] = time.perf_counter()
    for k in range(args.steps):
        step_device(evs[k])
    barrier()
    sampler.window[1] = time.perf_counter()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    ms_inf = sum(e[0].elapsed_time(e[1]) for e in evs)
    ms_def = sum(e[1].elapsed_time(e[2]) for e in evs)
    ms_total = evs[0][0].elapsed_time(evs[-1][2])

    # ---- verification of the timed work ----
    assert int((o_st != 0).sum()) == 0 and int((c_st2 != 0).sum()) == 0
    assert ctx.last_general_count(stream) == 0, "ultra-fast streams left the fast path"
    assert torch.equal(out, tiles), "inflate(deflate(x)) != x"
    assert torch.equal(c_len2, c_len) and torch.equal(comp2, comp), "deflate output not reproducible"
    # ... and a fresh random sample of the timed work against the oracle (the checker): the deflate bytes are the
    # reference encoder's, and the reference's inflate reads them back to the tile
    oracle_checked = 0
    try:
        sys.path.insert(0, str(ROOT / "tests"))
        import oracle_lib as O

        for i in np.random.default_rng().choice(n, size=min(n, 8), replace=False):
            tile = tiles[int(i) * TILE_BYTES:(int(i) + 1) * TILE_BYTES].cpu().numpy().tobytes()
            z = comp2[int(i) * bound:int(i) * bound + int(c_len2[int(i)])].cpu().numpy().tobytes()
            assert z == O.compress_ultra_fast(tile), "deflate output differs from the oracle (stream %d)" % int(i)
            assert O.inflate_into(z, TILE_BYTES)[:2] == (0, tile), "the oracle inflates stream %d differently" % int(i)
            assert out[int(i) * TILE_BYTES:(int(i) + 1) * TILE_BYTES].cpu().numpy().tobytes() == tile
            oracle_checked += 1
    except ImportError:
        pass

    # ---- end to end through the host-buffer C ABI: pinned host memory, copies inside the timed region ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_tiles = torch.empty(n * TILE_BYTES, dtype=torch.uint8, pin_memory=True)
    h_tiles.copy_(tiles)
    h_comp = torch.zeros(n * bound, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n * TILE_BYTES, dtype=torch.uint8, pin_memory=True)
    hn = lambda t: t.numpy()
    h_t_off, h_t_len = hn(t_off.cpu()).astype(np.uint64), hn(t_len.cpu()).astype(np.uint64)
    h_c_off, h_c_cap = hn(c_off.cpu()).astype(np.uint64), hn(c_cap.cpu()).astype(np.uint64)
    np_tiles, np_comp, np_out = hn(h_tiles), hn(h_comp), hn(h_out)

    def step_host_serial():
        clen, st = ctx.deflate_ultrafast_packed(np_tiles, h_t_off, h_t_len, np_comp, h_c_off, h_c_cap)
        olen, _, st2 = ctx.inflate_packed(np_comp, h_c_off, clen, np_out, h_t_off, h_t_len, 0)
        return clen, st, olen, st2

    clen, st, olen, st2 = step_host_serial()  # warm-up (allocates the context's staging buffers)
    assert (st == 0).all() and (st2 == 0).all() and (olen == TILE_BYTES).all()
    assert bool((h_out == h_tiles).all())

    # The two directions of a step are independent batches (as in the device-resident step above), and
    # their PCIe traffic is complementary: deflate moves 1.07 GB up and 0.44 GB down, inflate the other
    # way round.  Two contexts driven from two host threads (contexts are concurrent, include/fdeflate_b200.h)
    # keep both directions of the full-duplex link busy.  --e2e-serial runs one call after the other.
    ctx2 = F.Context(local_rank)
    h_comp2 = torch.zeros(n * bound, dtype=torch.uint8, pin_memory=True)
    np_comp2 = hn(h_comp2)
    from concurrent.futures import ThreadPoolExecutor

    pool = ThreadPoolExecutor(2)

    # the inflate leg reads its streams packed back to back (16-byte aligned), the way compressed tiles are
    # kept at rest; the deflate leg writes into slots of the worst-case size, the only layout a caller can
    # prepare before the sizes are known
    p_off = np.zeros(n, dtype=np.uint64)
    p_off[1:] = np.cumsum((clen[:-1] + np.uint64(15)) & ~np.uint64(15))
    h_packed = torch.zeros(int(p_off[-1] + clen[-1]) + 16, dtype=torch.uint8, pin_memory=True)
    np_packed = hn(h_packed)
    for i in range(n):
        np_packed[int(p_off[i]):int(p_off[i]) + int(clen[i])] = np_comp[int(h_c_off[i]):int(h_c_off[i]) + int(clen[i])]

    def step_host_overlapped():
        fa = pool.submit(ctx2.deflate_ultrafast_packed, np_tiles, h_t_off, h_t_len, np_comp2, h_c_off, h_c_cap)
        fb = pool.submit(ctx.inflate_packed, np_packed, p_off, clen, np_out, h_t_off, h_t_len, 0)
        clen2, st_a = fa.result()
        olen_b, _, st_b = fb.result()
        return clen2, st_a, olen_b, st_b

    step_host = step_host_serial if args.e2e_serial else step_host_overlapped
    h_out.zero_()
    clen2, st_a, olen_b, st_b = step_host()  # warm-up of the second context, and the check of this mode
    assert (st_a == 0).all() and (st_b == 0).all() and (olen_b == TILE_BYTES).all() and (clen2 == clen).all()
    assert bool((h_out == h_tiles).all())
    if not args.e2e_serial:
        for i in range(0, n, max(1, n // 64)):  # (bytes past a stream's length are unspecified)
            a0 = int(h_c_off[i])
            assert bool((h_comp2[a0:a0 + int(clen[i])] == h_comp[a0:a0 + int(clen[i])]).all()), "overlapped deflate output differs"
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    pool.shutdown()
    h2d = int(n * TILE_BYTES + int(clen.sum()) + 2 * 4 * 8 * n)
    d2h = int(n * TILE_BYTES + int(clen.sum()) + 2 * (8 + 8 + 4) * n)
    del ctx2, h_comp2, h_packed, h_out, h_comp, h_tiles
    link = link_ceiling(dev, dist)
    peak, peak_src = measured_peak()
    sweep_res = sweep(ctx, args, rank, world, dev, dist, stream, peak)

    # ---- reduce over ranks (max time), whole-job throughput ----
    t = torch.tensor([ms_total, ms_inf, ms_def, e2e_s], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_inf, ms_def, e2e_s = [float(x) for x in t.cpu()]
    unc = n * TILE_BYTES
    value = world * 2 * unc * args.steps / (ms_total / 1e3) / 1e9
    inflate_gbs = world * unc * args.steps / (ms_inf / 1e3) / 1e9
    deflate_gbs = world * unc * args.steps / (ms_def / 1e3) / 1e9
    e2e_value = world * 2 * unc * e2e_steps / e2e_s / 1e9

    if rank == 0:
        alg_bytes = unc + comp_bytes  # per launch, both kernels: read once + written once
        inf_ms, def_ms = ms_inf / args.steps, ms_def / args.steps
        dominant = "inflate_uf_kernel" if inf_ms >= def_ms else "deflate_ufb_kernel"
        dom_ms = max(inf_ms, def_ms)
        traffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get(dominant)
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_dict(n, world, {"workloads": ["configs[1]+[2] tiles (the headline: value, roofline, e2e)"] +
                                             (["configs[4] sweep (key `sweep`)"] if sweep_res else [])}),
            "inflate_gbs": round(inflate_gbs, 2), "deflate_gbs": round(deflate_gbs, 2),
            "compressed_ratio": round(comp_bytes / unc, 4),
            "roofline": {
                "bound": "hbm", "kernel": dominant, "achieved": round(alg_bytes / (dom_ms / 1e3) / 1e9, 1),
                "peak": peak, "unit": "GB/s", "frac": round(alg_bytes / (dom_ms / 1e3) / 1e9 / peak, 4),
                "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "per_kernel": {
                    "inflate_uf_kernel": {"ms": round(inf_ms, 4), "frac": round(alg_bytes / (inf_ms / 1e3) / 1e9 / peak, 4)},
                    "deflate_ufb_kernel": {"ms": round(def_ms, 4), "frac": round(alg_bytes / (def_ms / 1e3) / 1e9 / peak, 4)},
                },
            },
            "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "fdb_deflate_ultrafast_batch + fdb_inflate_batch, pinned host buffers, " +
                           ("one call after the other" if args.e2e_serial else "the two calls issued concurrently on two contexts"),
                    # the host link bounds this number: bytes that must cross it per step / what the link moves with all
                    # ranks copying both ways at once (measured in this run)
                    "link": link,
                    "frac_of_link": round((world * max(h2d, d2h) / 1e9 / link["gbs_per_direction_all_gpus"]) / (e2e_s / e2e_steps), 3)
                    if link["gbs_per_direction_all_gpus"] > 0 else None},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "verified": "inflate(deflate(x)) == x on all streams; re-encode byte-identical; fast path on 100%% of streams; "
                        "%d random streams of the timed work equal to the oracle's bytes in both directions" % oracle_checked,
        }
        if sweep_res:
            line["sweep"] = sweep_res
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_measure()
            except Exception as e:  # the GPU numbers stand on their own
                line["cpu_baseline"] = {"error": repr(e)}
        emit(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """the one JSON line, on the process's original stdout"""
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries the one JSON line and nothing else: whatever libraries print to file descriptor 1 (NCCL's
    # version banner, for one) is sent to stderr, and the line itself goes to a duplicate of the original descriptor
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--tiles", type=int, default=4096, help="streams per GPU")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep-streams", type=int, default=8192, help="configs[4]: streams per GPU (0 = skip the sweep)")
    ap.add_argument("--sweep-steps", type=int, default=2)
    ap.add_argument("--e2e-serial", action="store_true", help="end-to-end leg: deflate call, then inflate call (no overlap)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
    else:
        ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
