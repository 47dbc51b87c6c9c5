"""End-to-end leg only (host buffers through the C ABI), with the pipeline timeline (FDB_TRACE) summarised.
usage: python tools/gpu_e2e_trace.py [tiles] [chunk_MiB] [serial|overlap]"""
import os, sys, time
sys.path.insert(0, ".")
trace = "gpurun_out/e2e_trace.txt"
os.makedirs("gpurun_out", exist_ok=True)
if os.path.exists(trace): os.remove(trace)
os.environ["FDB_TRACE"] = trace
import numpy as np, torch
import fdeflate_b200 as F
from concurrent.futures import ThreadPoolExecutor
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mode = sys.argv[3] if len(sys.argv) > 3 else "overlap"
TB = 262400
ctx, ctx2 = F.Context(0), F.Context(0)
if chunk:
    ctx.set_pipeline_chunk(chunk << 20); ctx2.set_pipeline_chunk(chunk << 20)
dev = torch.device("cuda:0")
tiles = torch.empty(n * TB, dtype=torch.uint8, device=dev)
ctx.synth_tiles_device(tiles.data_ptr(), 0, n, 256, 256, 2024, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
bound = ctx.ultrafast_bound(TB)
h_tiles = torch.empty(n * TB, dtype=torch.uint8, pin_memory=True); h_tiles.copy_(tiles)
h_comp = torch.zeros(n * bound, dtype=torch.uint8, pin_memory=True)
h_comp2 = torch.zeros(n * bound, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n * TB, dtype=torch.uint8, pin_memory=True)
t_off = np.arange(n, dtype=np.uint64) * TB; t_len = np.full(n, TB, dtype=np.uint64)
c_off = np.arange(n, dtype=np.uint64) * bound; c_cap = np.full(n, bound, dtype=np.uint64)
clen, st = ctx.deflate_ultrafast_packed(h_tiles.numpy(), t_off, t_len, h_comp.numpy(), c_off, c_cap)
assert (st == 0).all()
pool = ThreadPoolExecutor(2)
def step():
    if mode == "serial":
        ctx.deflate_ultrafast_packed(h_tiles.numpy(), t_off, t_len, h_comp2.numpy(), c_off, c_cap)
        ctx.inflate_packed(h_comp.numpy(), c_off, clen, h_out.numpy(), t_off, t_len, 0)
    else:
        fa = pool.submit(ctx2.deflate_ultrafast_packed, h_tiles.numpy(), t_off, t_len, h_comp2.numpy(), c_off, c_cap)
        fb = pool.submit(ctx.inflate_packed, h_comp.numpy(), c_off, clen, h_out.numpy(), t_off, t_len, 0)
        fa.result(); fb.result()
step(); step()
assert bool((h_out == h_tiles).all())
open(trace, "w").close()
t0 = time.perf_counter(); step(); dt = time.perf_counter() - t0
print(f"{mode} chunk {chunk or 128} MiB: step {dt*1e3:.2f} ms -> {2*n*TB/dt/1e9:.1f} GB/s; comp max/mean = {clen.max()/clen.mean():.3f}")
rows = []
for line in open(trace):
    f = line.split()
    rows.append((int(f[3]), int(f[5].split('/')[0]), *[float(x) for x in (f[11], f[12], f[14], f[16], f[17])]))
base = min(r[2] for r in rows)
for kind in (0, 1):
    rk = [r for r in rows if r[0] == kind]
    print("kind", kind, "(inflate)" if kind == 0 else "(deflate)", len(rk), "chunks")
    for r in rk[:6] + rk[-3:]:
        print("  chunk %2d  h2d %7.2f-%7.2f (%5.2f)  kern_end %7.2f (+%5.2f)  d2h %7.2f-%7.2f (%5.2f)" % (
            r[1], r[2]-base, r[3]-base, r[3]-r[2], r[4]-base, r[4]-r[3], r[5]-base, r[6]-base, r[6]-r[5]))
    print("  sums: h2d %.2f ms  kern %.2f ms  d2h %.2f ms  span %.2f ms" % (
        sum(r[3]-r[2] for r in rk), sum(r[4]-r[3] for r in rk), sum(r[6]-r[5] for r in rk), max(r[6] for r in rk)-min(r[2] for r in rk)))
