"""Host-link ceiling of the box with ALL GPUs copying at once (nobody had measured it).
One process per GPU, 1 GiB pinned buffers, large contiguous cudaMemcpyAsync copies, three phases that start at the
same wall-clock instant in every process: host->device only, device->host only, both directions together.
    python tools/gpu_pcie_all.py [n_gpus] [seconds_per_phase]        -> one JSON line (per GPU and aggregate GB/s)
The aggregate of the third phase is what bounds the end-to-end number of bench.py at that GPU count."""
import ctypes
import json
import os
import subprocess
import sys
import time


def worker(gpu: int, t_start: float, secs: float):
    import torch

    torch.cuda.set_device(gpu)
    n = 1 << 30
    rt = ctypes.CDLL("libcudart.so.12")
    h_src = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_dst = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_src.fill_(1)
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    chunk = 64 << 20

    def h2d(o):
        rt.cudaMemcpyAsync(ctypes.c_void_p(d_a.data_ptr() + o), ctypes.c_void_p(h_src.data_ptr() + o), ctypes.c_size_t(chunk), 1,
                           ctypes.c_void_p(s1.cuda_stream))

    def d2h(o):
        rt.cudaMemcpyAsync(ctypes.c_void_p(h_dst.data_ptr() + o), ctypes.c_void_p(d_b.data_ptr() + o), ctypes.c_size_t(chunk), 2,
                           ctypes.c_void_p(s2.cuda_stream))

    res = {}
    for k, (name, fns) in enumerate((("h2d", (h2d,)), ("d2h", (d2h,)), ("both", (h2d, d2h)))):
        t0 = t_start + k * (secs + 1.5)
        for f in fns:  # warm
            f(0)
        torch.cuda.synchronize()
        while time.time() < t0:
            time.sleep(0.0005)
        moved = 0
        begin = time.perf_counter()
        while time.time() < t0 + secs:
            for o in range(0, n, chunk):
                for f in fns:
                    f(o)
            torch.cuda.synchronize()
            moved += n
        dt = time.perf_counter() - begin
        res[name] = moved / dt / 1e9  # per direction
    print(json.dumps({"gpu": gpu, **res}), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--worker":
        worker(int(sys.argv[2]), float(sys.argv[3]), float(sys.argv[4]))
        sys.exit(0)
    n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    secs = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
    out = {}
    for group in sorted({1, n_gpus}):
        t_start = time.time() + 25.0  # every process imports torch and allocates before the first phase
        procs = [subprocess.Popen([sys.executable, __file__, "--worker", str(g), str(t_start), str(secs)], stdout=subprocess.PIPE, text=True)
                 for g in range(group)]
        rows = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
        agg = {k: round(sum(r[k] for r in rows), 1) for k in ("h2d", "d2h", "both")}
        out[f"{group}_gpus"] = {"aggregate_gbs_per_direction": agg, "per_gpu": [{k: round(v, 1) if k != "gpu" else v for k, v in r.items()} for r in rows]}
    out["cpus"] = os.cpu_count()
    print(json.dumps(out))
