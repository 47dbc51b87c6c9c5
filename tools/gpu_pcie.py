"""PCIe measurements on the GPU box (what bounds the end-to-end number), one mode per question:
    python tools/gpu_pcie.py link     pinned H2D, D2H and both at once, whole buffer and in 64 / 8 MiB chunks
    python tools/gpu_pcie.py shapes   duplex ceiling with the copy shapes of the pipeline: contiguous vs 2-D, default vs
                                      write-combined pinned memory; prints the box's topology
    python tools/gpu_pcie.py mix      duplex with the exact copy mix of the bench's end-to-end step (no kernels): per chunk,
                                      H2D 52 MB contiguous (tiles) + H2D 22 MB contiguous (packed streams); D2H 54 MB
                                      contiguous (inflated tiles) + D2H 2-D 205 rows x 133 KB of pitch 393 KB (deflated
                                      streams); 20 chunks, one stream per direction; optionally an event after every copy
(`tools/gpu_pcie_all.py` runs the link measurement on several GPUs at once, one process each.)"""
import ctypes
import os
import sys
import time

import torch


def link():
    n = 1 << 30
    h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(h2d, d2h, reps=5, chunk=None):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            if chunk is None:
                if h2d:
                    with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
            else:
                for o in range(0, n, chunk):
                    if h2d:
                        with torch.cuda.stream(s1): d_a[o:o+chunk].copy_(h_a[o:o+chunk], non_blocking=True)
                    if d2h:
                        with torch.cuda.stream(s2): h_b[o:o+chunk].copy_(d_b[o:o+chunk], non_blocking=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
        return n / dt / 1e9
    run(True, True, 1)
    print("H2D alone  %.1f GB/s" % run(True, False))
    print("D2H alone  %.1f GB/s" % run(False, True))
    print("both       %.1f GB/s each way" % run(True, True))
    print("both, 64 MiB chunks %.1f GB/s each way" % run(True, True, chunk=64 << 20))
    print("both, 8 MiB chunks  %.1f GB/s each way" % run(True, True, chunk=8 << 20))


def shapes():
    rt = ctypes.CDLL("libcudart.so.12")
    n = 1 << 30
    def host_alloc(nbytes, flags):
        p = ctypes.c_void_p()
        assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags)) == 0
        return p.value
    torch.cuda.init(); torch.zeros(1, device="cuda")
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(src_h, dst_h, chunk, reps=4, pitch=None, width=None):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            for o in range(0, n, chunk):
                if pitch is None:
                    rt.cudaMemcpyAsync(ctypes.c_void_p(d_a.data_ptr() + o), ctypes.c_void_p(src_h + o), ctypes.c_size_t(chunk), 1, ctypes.c_void_p(s1.cuda_stream))
                    rt.cudaMemcpyAsync(ctypes.c_void_p(dst_h + o), ctypes.c_void_p(d_b.data_ptr() + o), ctypes.c_size_t(chunk), 2, ctypes.c_void_p(s2.cuda_stream))
                else:
                    rows = chunk // pitch
                    rt.cudaMemcpy2DAsync(ctypes.c_void_p(d_a.data_ptr() + o), ctypes.c_size_t(pitch), ctypes.c_void_p(src_h + o), ctypes.c_size_t(pitch), ctypes.c_size_t(width), ctypes.c_size_t(rows), 1, ctypes.c_void_p(s1.cuda_stream))
                    rt.cudaMemcpy2DAsync(ctypes.c_void_p(dst_h + o), ctypes.c_size_t(pitch), ctypes.c_void_p(d_b.data_ptr() + o), ctypes.c_size_t(pitch), ctypes.c_size_t(width), ctypes.c_size_t(rows), 2, ctypes.c_void_p(s2.cuda_stream))
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
        moved = n if pitch is None else (n // pitch) * width
        return moved / dt / 1e9
    for name, fl in (("default", 0), ("write-combined src", 4)):
        src = host_alloc(n, fl); dst = host_alloc(n, 0)
        ctypes.memset(src, 1, n) if fl == 0 else None
        print(name, "contiguous 52 MiB chunks, both ways: %.1f GB/s each way" % run(src, dst, 52 << 20))
        print(name, "2-D rows 131072 of pitch 393216, both ways: %.1f GB/s each way (payload)" % run(src, dst, 393216 * 128, pitch=393216, width=131072))
    os.system("nvidia-smi topo -m | head -8; lscpu | grep -E 'NUMA|Socket|Model name|^CPU\\(s\\)'")


def mix():
    rt = ctypes.CDLL("libcudart.so.12")
    def host_alloc(nbytes):
        p = ctypes.c_void_p(); assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(0)) == 0; return p.value
    torch.zeros(1, device="cuda")
    N, TB, BOUND = 4096, 262400, 393680
    per = 205
    h_tiles, h_packed, h_out, h_comp = host_alloc(N * TB), host_alloc(N * 110000), host_alloc(N * TB), host_alloc(N * BOUND)
    d_in = torch.empty(N * TB, dtype=torch.uint8, device="cuda"); d_in2 = torch.empty(N * 110000, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(N * TB, dtype=torch.uint8, device="cuda"); d_comp = torch.empty(N * BOUND, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    evs = [torch.cuda.Event() for _ in range(200)]
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    def step(with_events, width):
        k = 0
        for c in range(0, N, per):
            n = min(per, N - c)
            rt.cudaMemcpyAsync(vp(d_in.data_ptr() + c * TB), vp(h_tiles + c * TB), sz(n * TB), 1, vp(s1.cuda_stream))
            if with_events: evs[k].record(s1); k += 1
            rt.cudaMemcpyAsync(vp(d_in2.data_ptr() + c * 106800), vp(h_packed + c * 106800), sz(n * 106800), 1, vp(s1.cuda_stream))
            if with_events: evs[k].record(s1); k += 1
            rt.cudaMemcpyAsync(vp(h_out + c * TB), vp(d_out.data_ptr() + c * TB), sz(n * TB), 2, vp(s2.cuda_stream))
            if with_events: evs[k].record(s2); k += 1
            rt.cudaMemcpy2DAsync(vp(h_comp + c * BOUND), sz(BOUND), vp(d_comp.data_ptr() + c * BOUND), sz(BOUND), sz(width), sz(n), 2, vp(s2.cuda_stream))
            if with_events: evs[k].record(s2); k += 1
    for with_events in (False, True):
        for width in (133000, 106800):
            step(with_events, width); torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(3): step(with_events, width)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
            up = N * TB + N * 106800; down = N * TB + N * width
            print(f"events={with_events} 2-D width {width}: {dt*1e3:.1f} ms per step; H2D {up/dt/1e9:.1f} GB/s, D2H {down/dt/1e9:.1f} GB/s; as e2e {2*N*TB/dt/1e9:.1f} GB/s")


if __name__ == "__main__":
    {"link": link, "shapes": shapes, "mix": mix}[sys.argv[1] if len(sys.argv) > 1 else "link"]()
