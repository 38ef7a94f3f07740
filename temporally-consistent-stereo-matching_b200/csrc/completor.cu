// Input stems of the disparity completion network (SURVEY.md section 8f, rank 3's other half), one kernel:
//   disp_f4 = conv_disp_stem(disp)      1 -> 64 -> 64   (1x1, ReLU between)            ref: core/update.py:312-314,375
//   cost_f4 = conv_cost_stem(cost)      1 -> 32 -> 32                                   ref: core/update.py:315-317,376
//   mask_f4 = conv_mask_stem(mask)      1 -> 32 -> 32                                   ref: core/update.py:318-320,377
//   x4_disp = conv_disp_fuse(cat(...))  128 -> 128 -> 64 (1x1, ReLU between)            ref: core/update.py:321-323,378
// i.e. a per-pixel 3 -> 64 multi-layer perceptron (30 848 multiply-adds per pixel).  The reference runs it as 8 cuDNN 1x1
// convolutions, 4 ReLUs and a cat: 13 launches and 1.7 KB of intermediates per pixel through HBM for 12 bytes in and
// 256 bytes out.  Here: thread = pixel, every weight is read from shared memory as a warp-wide broadcast (LDS.128, one
// wavefront, four FMAs per load), the first layers' outputs and the 128-wide hidden vector live in the thread's own column
// of a shared staging tile (no barrier: nobody else touches it), accumulators in registers (128 at the widest layer; the
// CTA is alone on its SM because the weights take 122 KB, so registers are plentiful).  fp32 FMAs in the order "inputs
// ascending"; parity with torch's fp32 convolutions 1e-5 relative (a different, equally valid summation order).
#include "tcs_common.cuh"

namespace tcs {
namespace completor {

constexpr int kD = 64, kC = 32, kM = 32;       // stem widths
constexpr int kCat = kD + kC + kM;             // 128
constexpr int kHid = 128, kOut = 64;
constexpr int kThreads = 128;

// packed weights (floats), every matrix transposed to [input][output] so that consecutive outputs are one float4
constexpr int oW1d = 0, oB1d = oW1d + kD, oW2d = oB1d + kD, oB2d = oW2d + kD * kD;
constexpr int oW1c = oB2d + kD, oB1c = oW1c + kC, oW2c = oB1c + kC, oB2c = oW2c + kC * kC;
constexpr int oW1m = oB2c + kC, oB1m = oW1m + kM, oW2m = oB1m + kM, oB2m = oW2m + kM * kM;
constexpr int oW3 = oB2m + kM, oB3 = oW3 + kCat * kHid, oW4 = oB3 + kHid, oB4 = oW4 + kHid * kOut;
constexpr int kWeightFloats = oB4 + kOut;      // 31 296
constexpr int kSmemBytes = (kWeightFloats + kCat * kThreads) * 4;   // 125 184 + 65 536

// out[o] (+)= sum_i w_t[i][o] * x_i for one input value x_i: kN outputs, weights broadcast from shared memory.
template <int kN>
__device__ __forceinline__ void axpy_row(float (&acc)[kN], uint32_t w_row, float x) {
#pragma unroll
    for (int o = 0; o < kN; o += 4) {
        const float4 w = lds_v4_f32(w_row + 4u * o);
        acc[o] = fmaf(w.x, x, acc[o]);
        acc[o + 1] = fmaf(w.y, x, acc[o + 1]);
        acc[o + 2] = fmaf(w.z, x, acc[o + 2]);
        acc[o + 3] = fmaf(w.w, x, acc[o + 3]);
    }
}

// One stem: scalar -> kN (ReLU) -> kN, written to the thread's staging column at rows [row0, row0 + kN).
template <int kN>
__device__ __forceinline__ void stem(uint32_t wbase, int oW1, int oB1, int oW2, int oB2, float x, uint32_t stage, int row0) {
    float acc[kN];
#pragma unroll
    for (int o = 0; o < kN; ++o) acc[o] = lds_f32(wbase + 4u * (oB2 + o));
#pragma unroll 4
    for (int i = 0; i < kN; ++i) {
        const float h = fmaxf(fmaf(lds_f32(wbase + 4u * (oW1 + i)), x, lds_f32(wbase + 4u * (oB1 + i))), 0.0f);
        axpy_row<kN>(acc, wbase + 4u * (oW2 + i * kN), h);
    }
#pragma unroll
    for (int o = 0; o < kN; ++o) asm volatile("st.shared.f32 [%0], %1;" :: "r"(stage + 4u * kThreads * (row0 + o)), "f"(acc[o]) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
completor_stems_kernel(const float* __restrict__ disp, const float* __restrict__ cost, const float* __restrict__ mask,
                       const float* __restrict__ weights, float* __restrict__ out, int HW, long long npix) {
    extern __shared__ float smem[];
    for (int i = threadIdx.x; i < kWeightFloats / 4; i += kThreads)
        reinterpret_cast<float4*>(smem)[i] = __ldg(reinterpret_cast<const float4*>(weights) + i);
    __syncthreads();
    const uint32_t wbase = smem_u32(smem);
    const uint32_t stage = smem_u32(smem + kWeightFloats) + 4u * threadIdx.x;       // entry k of this thread: stage + 4*kThreads*k
    for (long long p0 = (long long)blockIdx.x * kThreads; p0 < npix; p0 += (long long)gridDim.x * kThreads) {
        const long long p = p0 + threadIdx.x;
        if (p >= npix) continue;                                                    // no barrier below: threads are independent
        stem<kD>(wbase, oW1d, oB1d, oW2d, oB2d, __ldg(disp + p), stage, 0);
        stem<kC>(wbase, oW1c, oB1c, oW2c, oB2c, __ldg(cost + p), stage, kD);
        stem<kM>(wbase, oW1m, oB1m, oW2m, oB2m, __ldg(mask + p), stage, kD + kC);
        // ---- conv_disp_fuse[0] + ReLU: 128 -> 128, all accumulators in registers, then back into the staging column
        {
            float acc[kHid];
#pragma unroll
            for (int o = 0; o < kHid; ++o) acc[o] = lds_f32(wbase + 4u * (oB3 + o));
#pragma unroll 2
            for (int i = 0; i < kCat; ++i) axpy_row<kHid>(acc, wbase + 4u * (oW3 + i * kHid), lds_f32(stage + 4u * kThreads * i));
#pragma unroll
            for (int o = 0; o < kHid; ++o)
                asm volatile("st.shared.f32 [%0], %1;" :: "r"(stage + 4u * kThreads * o), "f"(fmaxf(acc[o], 0.0f)) : "memory");
        }
        // ---- conv_disp_fuse[2]: 128 -> 64, stored as 64 planes (lanes = consecutive pixels: coalesced)
        {
            float acc[kOut];
#pragma unroll
            for (int o = 0; o < kOut; ++o) acc[o] = lds_f32(wbase + 4u * (oB4 + o));
#pragma unroll 4
            for (int i = 0; i < kHid; ++i) axpy_row<kOut>(acc, wbase + 4u * (oW4 + i * kOut), lds_f32(stage + 4u * kThreads * i));
            const long long n = p / HW;
            float* o_ptr = out + (n * kOut) * HW + (p - n * HW);
#pragma unroll
            for (int o = 0; o < kOut; ++o) o_ptr[(long long)o * HW] = acc[o];
        }
    }
}

}  // namespace completor
}  // namespace tcs

extern "C" int tcs_completor_stems_weight_floats(void) { return tcs::completor::kWeightFloats; }

extern "C" int tcs_completor_stems(const float* disp, const float* cost, const float* mask, const float* weights, float* out,
                                   int N, int H, int W, void* stream) {
    using namespace tcs;
    using namespace tcs::completor;
    TCS_REQUIRE(disp && cost && mask && weights && out, TCS_E_BADARG, "tcs_completor_stems: null pointer");
    TCS_REQUIRE(N > 0 && H > 0 && W > 0, TCS_E_BADARG, "tcs_completor_stems: bad sizes");
    TCS_REQUIRE(aligned16(weights), TCS_E_ALIGN, "tcs_completor_stems: weights must be 16-byte aligned");
    TCS_ONCE_PER_DEVICE(
        TCS_CHECK_CUDA(cudaFuncSetAttribute(completor_stems_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    );
    const long long npix = (long long)N * H * W;
    const long long blocks = ceil_div_ll(npix, kThreads);
    const int grid = (int)(blocks < (long long)num_sms() ? blocks : (long long)num_sms());
    completor_stems_kernel<<<grid, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(disp, cost, mask, weights, out, H * W, npix);
    TCS_CHECK_LAUNCH("tcs_completor_stems");
    return 0;
}
