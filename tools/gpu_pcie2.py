"""PCIe duplex ceiling with the copy shapes of the pipeline: contiguous vs 2-D, default vs write-combined pinned."""
import ctypes, time, torch, os
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
