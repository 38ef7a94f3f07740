// Shared host/device helpers for libtcs_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <atomic>
#include <cstdint>
#include <cstdio>

#include "tcs_b200.h"

namespace tcs {

// ---- host-side error plumbing ----------------------------------------------------------------
void set_error(const char* fmt, ...);

#define TCS_REQUIRE(cond, code, ...)                    \
    do {                                                \
        if (!(cond)) {                                  \
            ::tcs::set_error(__VA_ARGS__);              \
            return (code);                              \
        }                                               \
    } while (0)

// Check the launch (not the execution): no host sync happens inside the library.
#define TCS_CHECK_LAUNCH(what)                                                      \
    do {                                                                            \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) {                                                   \
            ::tcs::set_error("%s: %s", (what), cudaGetErrorString(e__));            \
            return static_cast<int>(e__);                                           \
        }                                                                           \
    } while (0)

#define TCS_CHECK_CUDA(expr)                                                        \
    do {                                                                            \
        cudaError_t e__ = (expr);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            ::tcs::set_error("%s: %s", #expr, cudaGetErrorString(e__));             \
            return static_cast<int>(e__);                                           \
        }                                                                           \
    } while (0)

// Function attributes (dynamic shared memory opt-in, carve-out hints) apply per DEVICE: run `body` the first time the
// enclosing call site is reached on each device of the process.  Two racing threads may both run it (idempotent).
#define TCS_ONCE_PER_DEVICE(...)                                                    \
    do {                                                                            \
        static std::atomic<bool> done__[64];                                        \
        int dev__ = 0;                                                              \
        TCS_CHECK_CUDA(cudaGetDevice(&dev__));                                      \
        if (!done__[dev__ & 63].load(std::memory_order_acquire)) {                  \
            __VA_ARGS__                                                             \
            done__[dev__ & 63].store(true, std::memory_order_release);              \
        }                                                                           \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

int num_sms();  // cached SM count of the current device
// Shared-memory carve-out (percent of the unified L1/shared array) to request for a kernel: the tuned default,
// or the value of environment variable `env` when set (development aid).  Kernels that stage tiles in shared
// memory but stream through L1 starve when resident CTAs take the whole array as shared memory.
int carveout_percent(const char* env, int tuned_default);

// ---- device helpers ----------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// x / c for a loop-invariant c, correctly rounded (bit-identical to IEEE division) in three FP
// instructions: with rc = RN(1/c), q0 = RN(x*rc), r = x - q0*c (exact in an FMA), q = RN(q0 + r*rc)
// (Markstein).  Checked against x / c for every c = 1..4096 and 8e7 numerators on the host.
__device__ __forceinline__ float div_by_const(float x, float c, float rc) {
    const float q0 = __fmul_rn(x, rc);
    const float r = __fmaf_rn(-q0, c, x);
    return __fmaf_rn(r, rc, q0);
}

// Programmatic dependent launch.  A kernel launched with launch_pdl() may have its CTAs scheduled while the previous kernel of the
// stream is still draining; pdl_wait_then_release() at its top blocks until that kernel has completed and its writes are visible
// (so nothing changes semantically) and lets the NEXT programmatic launch begin.  In a kernel launched the ordinary way both
// instructions are no-ops.  Worth 1-2 us per kernel boundary, i.e. something for chains of short kernels.
__device__ __forceinline__ void pdl_wait_then_release() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... Exp, typename... Act>
inline cudaError_t launch_pdl_if(bool pdl, void (*kernel)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Act&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;            // without the attribute this is an ordinary stream-ordered launch
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<Exp>(args)...);
}
template <typename... Exp, typename... Act>
inline cudaError_t launch_pdl(void (*kernel)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Act&&... args) {
    return launch_pdl_if(true, kernel, grid, block, smem, s, static_cast<Act&&>(args)...);
}

// Packed pairs of fp32 (sm_100: FADD2 / FMUL2 / FFMA2 — one issue slot for two independent lanes).  Round-to-nearest, nothing
// contracted: each lane's result is bit-identical to the scalar __fadd_rn / __fmul_rn / __fmaf_rn.
struct f32x2 { uint64_t v; };
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f32x2 pk2(float a) { return pk2(a, a); }
__device__ __forceinline__ float lo2(f32x2 x) { return __uint_as_float((uint32_t)(x.v & 0xffffffffull)); }   // register aliasing, no instruction
__device__ __forceinline__ float hi2(f32x2 x) { return __uint_as_float((uint32_t)(x.v >> 32)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }

// Explicit shared-space accesses by 32-bit shared address.  Pointers into dynamically carved shared memory
// lose their address space in the compiler's eyes and turn into generic LD/ST (slower, and tracked on the
// long scoreboard); these keep them LDS/STS.  volatile: they stay ordered with the mbarrier waits around them.
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 lds_v4_f32(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_v4_f32(uint32_t a, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_v4_u32(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// Streaming (read-once / write-once) global accesses: keep them out of L1.
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream_f1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
// Read-only load that keeps its place in program order (volatile asm): used to issue a batch of independent
// loads back to back where the compiler would otherwise sink them next to their uses.
__device__ __forceinline__ float ldg_ordered_f1(const float* p) {
    float r;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
// Streaming 16-byte load that asks L2 for a 64-byte granule instead of the default 128-byte line: for scattered
// sub-line reads (the lookup's tap windows) it cuts the DRAM traffic by a third (109 -> 76 MB per launch).
__device__ __forceinline__ float4 ldg_stream64_f4(const float4* p) {
    float4 r;
#if defined(TCS_LOOKUP_L2HINT) && TCS_LOOKUP_L2HINT
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
#else
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
#endif
    return r;
}
__device__ __forceinline__ void stg_stream_f1(float* p, float v) {
#if defined(TCS_LOOKUP_L2HINT) && TCS_LOOKUP_L2HINT == 1
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.f32 [%0], %1, %2;" :: "l"(p), "f"(v), "l"(pol) : "memory");
#else
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
#endif
}

#endif  // __CUDACC__

}  // namespace tcs
