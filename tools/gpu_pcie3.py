"""PCIe duplex with the exact copy mix of the bench's end-to-end step (no kernels): per chunk, H2D 52 MB contiguous
(tiles) + H2D 22 MB contiguous (packed streams); D2H 54 MB contiguous (inflated tiles) + D2H 2-D 205 rows x 133 KB of
pitch 393 KB (deflated streams); 20 chunks, one stream per direction; optionally an event record after every copy."""
import ctypes, time, torch, sys
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
