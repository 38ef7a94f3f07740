// Correlation pre-pass: L2-normalise NCHW fp32 features over channels and write them channels-last
// as 16-bit tensor-core operands (hi [+ lo residual]) and/or as fp32 (alternate-path operand).
// ref: core/corr.py:58-59 (F.normalize(fmap, dim=1): x / max(||x||_2, 1e-12)).
//
// HBM-bound transpose.  One CTA = one (b, h) row x 32 consecutive w; thread = (pixel, group of 16 channels): lanes are
// consecutive pixels, so every channel is one coalesced 128-byte warp load, and the thread's 16 channels leave as whole
// 32-byte sectors (2 x 16-byte stores per 16-bit tensor) of the channels-last row.  The values stay in registers between the
// norm and the emit; the only shared memory is the [groups][32] table of partial sums of squares, added in group order by
// every thread of the pixel (deterministic).  The first version staged the tile transposed in shared memory and spent 37
// instructions per element (53 % issue-active at 42 % DRAM, profiles/r02_prepass.md); this one spends about 10.
// Algorithmic bytes per pixel: 4C in + 2C (hi) [+ 2C lo] [+ 4C n32].
#include "tcs_common.cuh"

namespace tcs {

constexpr int kPreTileW = 32;                // pixels per CTA = lanes
#ifndef TCS_PRE_THREADS
#define TCS_PRE_THREADS 256
#endif
#ifndef TCS_PRE_GROUP
#define TCS_PRE_GROUP 16
#endif
constexpr int kPreThreads = TCS_PRE_THREADS;
constexpr int kPreWarps = kPreThreads / 32;
constexpr int kPreGroup = TCS_PRE_GROUP;     // channels per thread item (16, 32 or 64: whole sectors of the 16-bit rows)
constexpr int kPreMaxGroups = 512 / kPreGroup;
constexpr float kFp16OperandScale = 256.0f;  // unit-vector entries x 2^8 keep fp16 away from subnormals

template <bool kFp16>
__device__ __forceinline__ uint32_t pack_hi_lo(float a, float b, uint32_t& lo_pack) {
    if constexpr (kFp16) {
        a *= kFp16OperandScale;
        b *= kFp16OperandScale;
        __half ha = __float2half_rn(a), hb = __float2half_rn(b);
        __half la = __float2half_rn(a - __half2float(ha)), lb = __float2half_rn(b - __half2float(hb));
        lo_pack = (uint32_t)__half_as_ushort(la) | ((uint32_t)__half_as_ushort(lb) << 16);
        return (uint32_t)__half_as_ushort(ha) | ((uint32_t)__half_as_ushort(hb) << 16);
    } else {
        __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
        __nv_bfloat16 la = __float2bfloat16_rn(a - __bfloat162float(ha));
        __nv_bfloat16 lb = __float2bfloat16_rn(b - __bfloat162float(hb));
        lo_pack = (uint32_t)__bfloat16_as_ushort(la) | ((uint32_t)__bfloat16_as_ushort(lb) << 16);
        return (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
    }
}

// kItems: channel groups per thread (groups g = warp + kPreWarps * i): ceil(C / kPreGroup / kPreWarps) rounded up to 1, 2 or 4.
template <bool kFp16, int kItems>
__global__ void __launch_bounds__(kPreThreads)
corr_prepass_kernel(const float* __restrict__ fmap, uint32_t* __restrict__ hi, uint32_t* __restrict__ lo,
                    float* __restrict__ n32, int C, int H, int W, int kblocked) {
    __shared__ float part[kPreMaxGroups][kPreTileW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = blockIdx.x * kPreTileW + lane;
    const int h = blockIdx.y, b = blockIdx.z;
    const bool live = w < W;
    const int groups = C / kPreGroup;
    const size_t plane = (size_t)H * W;
    const float* src = fmap + ((size_t)b * C * H + h) * W + (live ? w : 0);

    // ---- every load of the thread first (kItems x 16 independent 4-byte loads, each a full line per warp)
    float x[kItems][kPreGroup];
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int g = warp + kPreWarps * i;
#pragma unroll
        for (int k = 0; k < kPreGroup; ++k)
            x[i][k] = (live && g < groups) ? ldg_stream_f1(src + (size_t)(g * kPreGroup + k) * plane) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int g = warp + kPreWarps * i;
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < kPreGroup; ++k) acc = fmaf(x[i][k], x[i][k], acc);
        if (g < groups) part[g][lane] = acc;
    }
    __syncthreads();
    if (!live) return;
    float ss = 0.0f;
    for (int g = 0; g < groups; ++g) ss += part[g][lane];          // the same order in every thread of the pixel
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    const float rden = __frcp_rn(denom);
    const size_t pix = ((size_t)b * H + h) * W + w;

    // ---- x / denom (the exact 3-instruction division by a loop-invariant denominator), split, emit
#pragma unroll
    for (int i = 0; i < kItems; ++i) {
        const int g = warp + kPreWarps * i;
        if (g >= groups) break;                                     // warp-uniform
        const int c = g * kPreGroup;
        float v[kPreGroup];
#pragma unroll
        for (int k = 0; k < kPreGroup; ++k) v[k] = div_by_const(x[i][k], denom, rden);
        if (n32 != nullptr) {
            float4* o = reinterpret_cast<float4*>(n32 + pix * C + c);
#pragma unroll
            for (int k = 0; k < kPreGroup; k += 4) o[k >> 2] = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
        }
        if (hi != nullptr) {
            uint32_t ph[kPreGroup / 2], pl[kPreGroup / 2];
#pragma unroll
            for (int k = 0; k < kPreGroup; k += 2) ph[k >> 1] = pack_hi_lo<kFp16>(v[k], v[k + 1], pl[k >> 1]);
            // pixel-major [B,H,W,C] or K-block-major [B,H,C/64,W,64]: either way 16 channels are one aligned 32-byte sector
            const size_t e = kblocked ? ((((size_t)b * H + h) * (C >> 6) + (c >> 6)) * W + w) * 64 + (c & 63) : pix * C + c;
            uint4* oh = reinterpret_cast<uint4*>(hi + (e >> 1));
#pragma unroll
            for (int q = 0; q < kPreGroup / 8; ++q) oh[q] = make_uint4(ph[4 * q], ph[4 * q + 1], ph[4 * q + 2], ph[4 * q + 3]);
            if (lo != nullptr) {
                uint4* ol = reinterpret_cast<uint4*>(lo + (e >> 1));
#pragma unroll
                for (int q = 0; q < kPreGroup / 8; ++q) ol[q] = make_uint4(pl[4 * q], pl[4 * q + 1], pl[4 * q + 2], pl[4 * q + 3]);
            }
        }
    }
}

// out[b,h,j,:] = 0.5 * (in[b,h,2j,:] + in[b,h,2j+1,:]), channels last.
__global__ void fmap_pool_w_kernel(const float4* __restrict__ in, float4* __restrict__ out,
                                   int W, int Wo, int C4, long long total) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c4 = (int)(i % C4);
    long long r = i / C4;
    const int j = (int)(r % Wo);
    const long long bh = r / Wo;
    const float4 a = in[(bh * W + 2 * j) * C4 + c4];
    const float4 b = in[(bh * W + 2 * j + 1) * C4 + c4];
    out[i] = make_float4((a.x + b.x) * 0.5f, (a.y + b.y) * 0.5f, (a.z + b.z) * 0.5f, (a.w + b.w) * 0.5f);
}

}  // namespace tcs

static int launch_prepass(const float* fmap, void* hi, void* lo, float* n32, int B, int C, int H, int W, int prec, int kblocked,
                          void* stream) {
    using namespace tcs;
    TCS_REQUIRE(fmap != nullptr && (hi != nullptr || n32 != nullptr), TCS_E_BADARG,
                "tcs_corr_prepass: fmap and at least one of hi / n32 are required");
    TCS_REQUIRE(lo == nullptr || hi != nullptr, TCS_E_BADARG, "tcs_corr_prepass: lo needs hi");
    TCS_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, TCS_E_BADARG, "tcs_corr_prepass: non-positive size");
    TCS_REQUIRE(C % 64 == 0 && C <= 512, TCS_E_SHAPE, "tcs_corr_prepass: C=%d must be a multiple of 64, <= 512", C);
    TCS_REQUIRE(H <= 65535 && B <= 65535, TCS_E_SHAPE, "tcs_corr_prepass: H and B must be <= 65535");
    TCS_REQUIRE(prec >= TCS_PREC_BF16 && prec <= TCS_PREC_FP16X3, TCS_E_BADARG, "tcs_corr_prepass: bad prec %d", prec);
    TCS_REQUIRE(aligned16(fmap) && aligned16(hi) && aligned16(lo) && aligned16(n32), TCS_E_ALIGN,
                "tcs_corr_prepass: pointers must be 16-byte aligned");
    const bool fp16 = (prec == TCS_PREC_FP16 || prec == TCS_PREC_FP16X3);
    dim3 grid(ceil_div(W, kPreTileW), H, B);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint32_t* h32 = static_cast<uint32_t*>(hi);
    uint32_t* l32 = static_cast<uint32_t*>(lo);
#define TCS_PREPASS_LAUNCH(F16, ITEMS) corr_prepass_kernel<F16, ITEMS><<<grid, kPreThreads, 0, s>>>(fmap, h32, l32, n32, C, H, W, kblocked)
    const int per_thread = ceil_div(C / kPreGroup, kPreWarps);
    static_assert(512 / kPreGroup <= 4 * kPreWarps, "prepass: at most 4 channel groups per thread");
    if (per_thread <= 1) { if (fp16) TCS_PREPASS_LAUNCH(true, 1); else TCS_PREPASS_LAUNCH(false, 1); }
    else if (per_thread <= 2) { if (fp16) TCS_PREPASS_LAUNCH(true, 2); else TCS_PREPASS_LAUNCH(false, 2); }
    else { if (fp16) TCS_PREPASS_LAUNCH(true, 4); else TCS_PREPASS_LAUNCH(false, 4); }
#undef TCS_PREPASS_LAUNCH
    TCS_CHECK_LAUNCH("tcs_corr_prepass");
    return 0;
}

extern "C" int tcs_corr_prepass(const float* fmap, void* hi, void* lo, float* n32,
                                int B, int C, int H, int W, int prec, void* stream) {
    return launch_prepass(fmap, hi, lo, n32, B, C, H, W, prec, 0, stream);
}

extern "C" int tcs_corr_prepass_kblocked(const float* fmap, void* hi, void* lo,
                                         int B, int C, int H, int W, int prec, void* stream) {
    TCS_REQUIRE(hi != nullptr, TCS_E_BADARG, "tcs_corr_prepass_kblocked: hi is required");
    return launch_prepass(fmap, hi, lo, nullptr, B, C, H, W, prec, 1, stream);
}

extern "C" int tcs_fmap_pool_w(const float* in, float* out, int B, int H, int W, int C, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(in != nullptr && out != nullptr, TCS_E_BADARG, "tcs_fmap_pool_w: null pointer");
    TCS_REQUIRE(B > 0 && H > 0 && W >= 2 && C > 0 && C % 4 == 0, TCS_E_SHAPE, "tcs_fmap_pool_w: need W >= 2 and C %% 4 == 0");
    TCS_REQUIRE(aligned16(in) && aligned16(out), TCS_E_ALIGN, "tcs_fmap_pool_w: pointers must be 16-byte aligned");
    const int Wo = W / 2, C4 = C / 4;
    const long long total = (long long)B * H * Wo * C4;
    const int threads = 256;
    fmap_pool_w_kernel<<<(unsigned)ceil_div_ll(total, threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), W, Wo, C4, total);
    TCS_CHECK_LAUNCH("tcs_fmap_pool_w");
    return 0;
}
