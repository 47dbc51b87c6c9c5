// simt_emul.h -- TEST-ONLY cooperative-fiber SIMT emulator (g++ -DFDB_EMUL).
//
// Runs the unmodified kernel source of fdeflate_b200/csrc on the CPU: one CTA at a time, one
// ucontext fiber per CUDA thread, round-robin scheduling with a context switch at every warp- or
// block-level collective.  It checks LOGIC (indexing, scans, carries, table builds) against the
// oracle without a GPU; it cannot see data races or memory-model bugs (compute-sanitizer on the
// GPU box covers those).  This header also stubs the handful of CUDA runtime calls capi.cu uses.
// Nothing under fdeflate_b200/ loads the emulator build.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

struct uint3 {
    unsigned x, y, z;
};
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 {
    uint32_t x, y, z, w;
};
struct uint2 {
    uint32_t x, y;
};
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

namespace fdb_emul {

struct WarpState {
    uint64_t slot[32];
    uint32_t arrived = 0;
    uint32_t expected = 0;
    uint32_t phase = 0;
};

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
};

struct BlockState {
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    unsigned nthreads = 0;
    unsigned block_arrived = 0;
    unsigned block_phase = 0;
    unsigned live = 0;
    unsigned cur = 0;
    unsigned long long spins = 0;
    ucontext_t sched;
    std::function<void()> body;
};

extern BlockState* g_blk;
extern unsigned char* g_dyn_smem;

void yield();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

inline void stuck_check() {
    if (++g_blk->spins > 400000000ull) {
        fprintf(stderr, "fdb_emul: deadlock (divergent collective?) in block thread %u\n", g_blk->cur);
        abort();
    }
}

}  // namespace fdb_emul

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

#define FDB_DEVICE static inline
#define FDB_MEMBER inline
#define FDB_DEVICE_NOINLINE static
#define FDB_GLOBAL static
#define FDB_SHARED static
#define FDB_DYN_SMEM(name) unsigned char* name = fdb_emul::g_dyn_smem
#define FDB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    fdb_emul::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define FDB_LAUNCH_BOUNDS(t, b)
#define FDB_FULL 0xffffffffu
#define __restrict__

namespace simt {
static inline unsigned lane_id() { return threadIdx.x & 31u; }
static inline unsigned warp_in_block() { return threadIdx.x >> 5; }

static inline void warp_barrier() {
    using namespace fdb_emul;
    WarpState& w = g_blk->warps[threadIdx.x >> 5];
    uint32_t my = w.phase;
    w.arrived |= 1u << (threadIdx.x & 31u);
    if (w.arrived == w.expected) {
        w.arrived = 0;
        w.phase++;
        g_blk->spins = 0;
    } else {
        while (w.phase == my) {
            stuck_check();
            yield();
        }
    }
}
static inline void syncwarp() { warp_barrier(); }
static inline void syncthreads() {
    using namespace fdb_emul;
    unsigned my = g_blk->block_phase;
    g_blk->block_arrived++;
    if (g_blk->block_arrived == g_blk->nthreads) {
        g_blk->block_arrived = 0;
        g_blk->block_phase++;
        g_blk->spins = 0;
    } else {
        while (g_blk->block_phase == my) {
            stuck_check();
            yield();
        }
    }
}

template <class T>
static inline T exchange(T v, unsigned src_lane, bool valid_src) {
    using namespace fdb_emul;
    WarpState& w = g_blk->warps[threadIdx.x >> 5];
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w.slot[threadIdx.x & 31u] = raw;
    warp_barrier();
    T r = v;
    if (valid_src) {
        uint64_t got = w.slot[src_lane & 31u];
        memcpy(&r, &got, sizeof(T));
    }
    warp_barrier();
    return r;
}
static inline uint32_t shfl(uint32_t v, unsigned src) { return exchange(v, src, true); }
static inline int32_t shfl(int32_t v, unsigned src) { return exchange(v, src, true); }
static inline uint64_t shfl(uint64_t v, unsigned src) { return exchange(v, src, true); }
static inline uint32_t shfl_up(uint32_t v, unsigned d) { return exchange(v, lane_id() - d, lane_id() >= d); }
static inline uint64_t shfl_up(uint64_t v, unsigned d) { return exchange(v, lane_id() - d, lane_id() >= d); }
static inline uint32_t shfl_down(uint32_t v, unsigned d) { return exchange(v, lane_id() + d, lane_id() + d < 32); }
static inline uint64_t shfl_down(uint64_t v, unsigned d) { return exchange(v, lane_id() + d, lane_id() + d < 32); }
static inline uint32_t shfl_xor(uint32_t v, unsigned m) { return exchange(v, lane_id() ^ m, true); }
static inline uint64_t shfl_xor(uint64_t v, unsigned m) { return exchange(v, lane_id() ^ m, true); }
static inline uint32_t ballot(bool p) {
    using namespace fdb_emul;
    WarpState& w = g_blk->warps[threadIdx.x >> 5];
    w.slot[threadIdx.x & 31u] = p ? 1 : 0;
    warp_barrier();
    uint32_t r = 0;
    for (unsigned i = 0; i < 32; i++)
        if ((w.expected >> i) & 1u)
            if (w.slot[i]) r |= 1u << i;
    warp_barrier();
    return r;
}
static inline bool any(bool p) { return ballot(p) != 0; }
static inline bool all(bool p) { return ballot(!p) == 0; }
static inline uint32_t popc(uint32_t v) { return (uint32_t)__builtin_popcount(v); }
static inline uint32_t clz(uint32_t v) { return v ? (uint32_t)__builtin_clz(v) : 32u; }
static inline uint32_t ffs(uint32_t v) { return (uint32_t)__builtin_ffs((int)v); }
static inline uint32_t brev(uint32_t v) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++)
        if (v & (1u << i)) r |= 1u << (31 - i);
    return r;
}
static inline uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (s & 31u));
}
static inline uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)((v << (s & 31u)) >> 32);
}
static inline uint32_t funnel_rc(uint32_t lo, uint32_t hi, uint32_t s) {
    uint64_t v = ((uint64_t)hi << 32) | lo;
    return (uint32_t)(v >> (s > 32u ? 32u : s));
}
static inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
    uint8_t bytes[8];
    memcpy(bytes, &a, 4);
    memcpy(bytes + 4, &b, 4);
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0xf;
        uint8_t v = bytes[sel & 7];
        if (sel & 8) v = (v & 0x80) ? 0xff : 0x00;
        r |= (uint32_t)v << (8 * i);
    }
    return r;
}
static inline uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) { return prmt(a, b, s & 0x7777u); }
static inline uint32_t dp4a_u(uint32_t a, uint32_t b, uint32_t c) {
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
    return c;
}
static inline uint32_t atomic_add(uint32_t* p, uint32_t v) {
    uint32_t o = *p;
    *p = o + v;
    return o;
}
static inline uint64_t atomic_add(uint64_t* p, uint64_t v) {
    uint64_t o = *p;
    *p = o + v;
    return o;
}
static inline uint32_t atomic_or(uint32_t* p, uint32_t v) {
    uint32_t o = *p;
    *p = o | v;
    return o;
}
static inline void threadfence() {}
static inline uint32_t ldg32(const uint32_t* p) { return *p; }
static inline uint4 ldg128(const uint4* p) { return *p; }
static inline uint8_t ldg8(const uint8_t* p) { return *p; }
static inline uint8_t ldcg8(const uint8_t* p) { return *(const volatile uint8_t*)p; }
static inline void stcs128(uint4* p, uint4 v) { *p = v; }
static inline void prefetch_l2(const void*) {}
// explicit shared-window addressing: plain host pointers here
typedef uintptr_t saddr;
static inline saddr smem_addr(const void* p) { return (saddr)p; }
static inline uint32_t lds32_ro(saddr a) { return *(const uint32_t*)a; }
static inline uint32_t lds16_ro(saddr a) { return *(const uint16_t*)a; }
static inline uint2 lds64_ro(saddr a) { return *(const uint2*)a; }
static inline uint32_t lds32(saddr a) { return *(const uint32_t*)a; }
static inline uint32_t lds8(saddr a) { return *(const uint8_t*)a; }
static inline void sts8(saddr a, uint32_t v) { *(uint8_t*)a = (uint8_t)v; }
static inline void sts8_if(saddr a, uint32_t v, bool p) { if (p) sts8(a, v); }
static inline void sts32(saddr a, uint32_t v) { *(uint32_t*)a = v; }
static inline void sts32_if(saddr a, uint32_t v, bool p) { if (p) sts32(a, v); }
static inline void atoms_or(saddr a, uint32_t v) { *(uint32_t*)a |= v; }
static inline uint4 lds128(saddr a) { return *(const uint4*)a; }
// bulk asynchronous copy + mbarrier: the emulator copies when the copy is issued and counts completed phases in the
// barrier word (one copy per phase); a wait yields until the phase of the given parity has completed
static inline void mbar_init(saddr bar, uint32_t) { *(uint64_t*)bar = 0; }
static inline void mbar_fence_init() {}
static inline void mbar_expect_tx(saddr, uint32_t) {}
static inline void bulk_g2s(saddr dst, const void* src, uint32_t bytes, saddr bar) {
    memcpy((void*)dst, src, bytes);
    (*(uint64_t*)bar)++;
}
static inline void mbar_wait(saddr bar, uint32_t parity) {
    while (((*(volatile uint64_t*)bar) & 1u) == (parity & 1u)) {
        fdb_emul::stuck_check();
        fdb_emul::yield();
    }
}
}  // namespace simt

// ---- CUDA runtime stubs used by capi.cu --------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaEventDisableTiming = 2 };
enum { cudaDevAttrMultiProcessorCount = 16, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 4; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? cudaSuccess : 1; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 1; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 1; }
// (the emulator has no pinned memory: every host pointer is "unregistered", so the host-buffer calls take the copy path)
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; void* devicePointer; void* hostPointer; };
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
    a->type = cudaMemoryTypeUnregistered; a->devicePointer = nullptr; a->hostPointer = (void*)p; return 0;
}
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h,
                                            cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; r++) memmove((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) {
    if (n) memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
