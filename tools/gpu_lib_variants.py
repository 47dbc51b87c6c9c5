"""A/B of several builds of the library on one box, device-resident, each in its own process:
    python tools/gpu_lib_variants.py ab/a.so ab/b.so ...          (driver)
    python tools/gpu_lib_variants.py --one ab/a.so                (one build: inflate + deflate of the bench tiles)
Every build's outputs are verified (inflate output == tiles, deflate output of all builds must hash the same)."""
import hashlib
import subprocess
import sys

sys.path.insert(0, ".")


def one(path):
    import torch
    import fdeflate_b200 as F

    n, TB = 4096, 262400
    lib = F.NativeLib(path)
    ctx = F.Context(0, lib)
    dev = torch.device("cuda:0")
    i64 = torch.int64
    s = torch.cuda.current_stream().cuda_stream
    tiles = torch.empty(n * TB, dtype=torch.uint8, device=dev)
    ctx.synth_tiles_device(tiles.data_ptr(), 0, n, 256, 256, 2024, s)
    bound = ctx.ultrafast_bound(TB)
    t_off = torch.arange(n, dtype=i64, device=dev) * TB
    t_len = torch.full((n,), TB, dtype=i64, device=dev)
    c_off = torch.arange(n, dtype=i64, device=dev) * bound
    c_cap = torch.full((n,), bound, dtype=i64, device=dev)
    comp = torch.zeros(n * bound, dtype=torch.uint8, device=dev)
    c_len = torch.zeros(n, dtype=i64, device=dev)
    c_st = torch.zeros(n, dtype=torch.int32, device=dev)
    out = torch.zeros(n * TB, dtype=torch.uint8, device=dev)
    o_len = torch.zeros(n, dtype=i64, device=dev)
    o_st = torch.zeros(n, dtype=torch.int32, device=dev)
    d = lambda: ctx.deflate_ultrafast_device(tiles.data_ptr(), t_off.data_ptr(), t_len.data_ptr(), comp.data_ptr(), c_off.data_ptr(),
                                             c_cap.data_ptr(), c_len.data_ptr(), c_st.data_ptr(), n, s)
    f = lambda: ctx.inflate_device(comp.data_ptr(), c_off.data_ptr(), c_len.data_ptr(), out.data_ptr(), t_off.data_ptr(),
                                   t_len.data_ptr(), o_len.data_ptr(), 0, o_st.data_ptr(), n, 0, s)
    d(); f(); torch.cuda.synchronize()
    ok_inf = bool(torch.equal(out, tiles)) and int((o_st != 0).sum()) == 0
    gen = ctx.last_general_count(s)
    h = hashlib.sha256(comp.cpu().numpy().tobytes()).hexdigest()[:16]
    res = {}
    for name, fn in (("inflate", f), ("deflate", d)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 10)
        res[name] = best
    print(f"{path}: inflate {res['inflate']:.4f} ms = {n*TB/res['inflate']/1e6:.1f} GB/s | deflate {res['deflate']:.4f} ms = "
          f"{n*TB/res['deflate']/1e6:.1f} GB/s | inflate ok={ok_inf} general={gen} comp sha={h} ratio={float(c_len.sum())/(n*TB):.4f}", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--one":
        one(sys.argv[2])
    else:
        for rnd in range(2):
            for p in sys.argv[1:]:
                r = subprocess.run([sys.executable, __file__, "--one", p], capture_output=True, text=True, timeout=600)
                print((r.stdout.strip() or r.stderr.strip()[-600:]), flush=True)
