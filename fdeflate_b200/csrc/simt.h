// simt.h -- thin SIMT vocabulary used by every kernel in this directory.
//
// Product build (nvcc, sm_100a): every wrapper is a one-line forward to the CUDA intrinsic.
//
// Test-only build (g++ -DFDB_EMUL): the same kernel source is compiled as plain C++ and run by
// the cooperative-fiber SIMT emulator in simt_emul.h, so that warp-level logic (shuffles, ballots,
// scans, shared-memory staging) can be exercised against the oracle on a machine without a GPU.
// The emulator build produces tests/emul/libfdb_emul.so; the package never loads it, and it is
// not a fallback: fdeflate_b200 raises if libfdeflate_b200.so (the CUDA build) is missing.
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef FDB_EMUL
#include "simt_emul.h"
#else
#include <cuda_runtime.h>

#define FDB_DEVICE __device__ __forceinline__
#define FDB_MEMBER __device__ __forceinline__
#define FDB_DEVICE_NOINLINE __device__ __noinline__
#define FDB_GLOBAL __global__
#define FDB_SHARED __shared__
#define FDB_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define FDB_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define FDB_LAUNCH_BOUNDS(t, b) __launch_bounds__(t, b)
#define FDB_FULL 0xffffffffu

namespace simt {
FDB_DEVICE unsigned lane_id() { return threadIdx.x & 31u; }
FDB_DEVICE unsigned warp_in_block() { return threadIdx.x >> 5; }
FDB_DEVICE void syncwarp() { __syncwarp(); }
FDB_DEVICE void syncthreads() { __syncthreads(); }
FDB_DEVICE uint32_t shfl(uint32_t v, unsigned src) { return __shfl_sync(FDB_FULL, v, src); }
FDB_DEVICE int32_t shfl(int32_t v, unsigned src) { return __shfl_sync(FDB_FULL, v, src); }
FDB_DEVICE uint64_t shfl(uint64_t v, unsigned src) {
    return (uint64_t)__shfl_sync(FDB_FULL, (unsigned long long)v, src);
}
FDB_DEVICE uint32_t shfl_up(uint32_t v, unsigned d) { return __shfl_up_sync(FDB_FULL, v, d); }
FDB_DEVICE uint64_t shfl_up(uint64_t v, unsigned d) {
    return (uint64_t)__shfl_up_sync(FDB_FULL, (unsigned long long)v, d);
}
FDB_DEVICE uint32_t shfl_down(uint32_t v, unsigned d) { return __shfl_down_sync(FDB_FULL, v, d); }
FDB_DEVICE uint64_t shfl_down(uint64_t v, unsigned d) {
    return (uint64_t)__shfl_down_sync(FDB_FULL, (unsigned long long)v, d);
}
FDB_DEVICE uint32_t shfl_xor(uint32_t v, unsigned m) { return __shfl_xor_sync(FDB_FULL, v, m); }
FDB_DEVICE uint64_t shfl_xor(uint64_t v, unsigned m) {
    return (uint64_t)__shfl_xor_sync(FDB_FULL, (unsigned long long)v, m);
}
FDB_DEVICE uint32_t ballot(bool p) { return __ballot_sync(FDB_FULL, p); }
FDB_DEVICE bool any(bool p) { return __any_sync(FDB_FULL, p) != 0; }
FDB_DEVICE bool all(bool p) { return __all_sync(FDB_FULL, p) != 0; }
FDB_DEVICE uint32_t popc(uint32_t v) { return (uint32_t)__popc(v); }
FDB_DEVICE uint32_t clz(uint32_t v) { return (uint32_t)__clz((int)v); }
FDB_DEVICE uint32_t ffs(uint32_t v) { return (uint32_t)__ffs((int)v); }  // 1-based, 0 if none
FDB_DEVICE uint32_t brev(uint32_t v) { return __brev(v); }
FDB_DEVICE uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }
// ((hi:lo) << (s mod 32)) >> 32
FDB_DEVICE uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_l(lo, hi, s); }
// clamped variant: a shift of 32 or more returns hi (funnel_r takes the shift mod 32)
FDB_DEVICE uint32_t funnel_rc(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_rc(lo, hi, s); }
FDB_DEVICE uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
// full prmt.b32: selector nibble bit 3 replicates the sign of the selected byte (__byte_perm masks it off)
FDB_DEVICE uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
    return r;
}
FDB_DEVICE uint32_t dp4a_u(uint32_t a, uint32_t b, uint32_t c) { return __dp4a(a, b, c); }
FDB_DEVICE uint32_t atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
FDB_DEVICE uint64_t atomic_add(uint64_t* p, uint64_t v) {
    return (uint64_t)atomicAdd((unsigned long long*)p, (unsigned long long)v);
}
FDB_DEVICE uint32_t atomic_or(uint32_t* p, uint32_t v) { return atomicOr(p, v); }
FDB_DEVICE void threadfence() { __threadfence(); }
FDB_DEVICE uint32_t ldg32(const uint32_t* p) { return __ldg(p); }
FDB_DEVICE uint4 ldg128(const uint4* p) { return __ldg(p); }
FDB_DEVICE uint8_t ldg8(const uint8_t* p) { return __ldg(p); }
// L2-coherent byte load (bypasses L1): data another lane of this warp stored earlier in the kernel
FDB_DEVICE uint8_t ldcg8(const uint8_t* p) { return __ldcg(p); }
// hint: bring the 128-byte line at p into L2 (no register, no dependency)
FDB_DEVICE void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// streaming (evict-first) 16-byte store: output bytes are written once and never re-read by us
FDB_DEVICE void stcs128(uint4* p, uint4 v) { __stcs(p, v); }

// ---- explicit shared-window addressing --------------------------------------------------------
// Hot decode loops keep 32-bit shared-memory addresses in registers and issue ld/st.shared directly;
// with generic pointers the compiler re-derives the shared window base inside the loop.
//   *_ro : constant tables (filled before the CTA barrier); the compiler may schedule these freely.
//   others: staging rows / output windows that change between warp barriers; they are ordered with
//           the surrounding C++ accesses by the "memory" clobber.
typedef uint32_t saddr;
FDB_DEVICE saddr smem_addr(const void* p) {
    saddr a = (saddr)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));  // opaque: keep the address in a register instead of re-deriving it per use
    return a;
}
FDB_DEVICE uint32_t lds32_ro(saddr a) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
FDB_DEVICE uint32_t lds16_ro(saddr a) {
    uint32_t v;
    asm("{\n\t.reg .u16 t;\n\tld.shared.u16 t, [%1];\n\tcvt.u32.u16 %0, t;\n\t}" : "=r"(v) : "r"(a));
    return v;
}
FDB_DEVICE uint2 lds64_ro(saddr a) {
    uint2 v;
    asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
FDB_DEVICE uint32_t lds32(saddr a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
FDB_DEVICE uint32_t lds8(saddr a) {
    uint32_t v;
    asm volatile("{\n\t.reg .u16 t;\n\tld.shared.u8 t, [%1];\n\tcvt.u32.u16 %0, t;\n\t}" : "=r"(v) : "r"(a) : "memory");
    return v;
}
FDB_DEVICE void sts8(saddr a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// predicated byte store: never a branch (the compiler turns `if (p) sts8(..)` into a divergent region with its
// BSSY / BSYNC pair when several of them nest)
FDB_DEVICE void sts8_if(saddr a, uint32_t v, bool p) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.u8 [%0], %1;\n\t}" ::"r"(a), "r"(v), "r"((uint32_t)p) : "memory");
}
FDB_DEVICE void sts32(saddr a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
FDB_DEVICE void sts32_if(saddr a, uint32_t v, bool p) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.u32 [%0], %1;\n\t}" ::"r"(a), "r"(v), "r"((uint32_t)p) : "memory");
}
// OR into a shared-memory word (no return value): bytes of one word that belong to different lanes
FDB_DEVICE void atoms_or(saddr a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
FDB_DEVICE uint4 lds128(saddr a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}

// ---- bulk asynchronous copy global -> shared, completion on an mbarrier (sm_90+; SASS: UBLKCP / SYNCS) ----------
// One thread arms the barrier with the byte count and issues the copy; the copy engine moves the bytes without any
// register or load/store-unit traffic; every consumer waits on the barrier's phase parity.  src, dst: 16-byte aligned,
// bytes: a multiple of 16.
FDB_DEVICE void mbar_init(saddr bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(arrivals) : "memory");
}
FDB_DEVICE void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
FDB_DEVICE void mbar_expect_tx(saddr bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
FDB_DEVICE void bulk_g2s(saddr dst, const void* src, uint32_t bytes, saddr bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
FDB_DEVICE void mbar_wait(saddr bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
}  // namespace simt
#endif

namespace simt {
// lanes strictly below me
FDB_DEVICE uint32_t lanemask_lt() { return (1u << lane_id()) - 1u; }

// inclusive warp prefix sums
FDB_DEVICE uint32_t scan_incl_add(uint32_t v) {
#pragma unroll
    for (unsigned d = 1; d < 32; d <<= 1) {
        uint32_t t = shfl_up(v, d);
        if (lane_id() >= d) v += t;
    }
    return v;
}
FDB_DEVICE uint64_t scan_incl_add(uint64_t v) {
#pragma unroll
    for (unsigned d = 1; d < 32; d <<= 1) {
        uint64_t t = shfl_up(v, d);
        if (lane_id() >= d) v += t;
    }
    return v;
}
FDB_DEVICE uint64_t reduce_add(uint64_t v) {
#pragma unroll
    for (unsigned d = 16; d > 0; d >>= 1) v += shfl_xor(v, d);
    return v;
}
FDB_DEVICE uint32_t reduce_add(uint32_t v) {
#pragma unroll
    for (unsigned d = 16; d > 0; d >>= 1) v += shfl_xor(v, d);
    return v;
}
}  // namespace simt
