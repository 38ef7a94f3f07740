// Correlation pre-pass: L2-normalise NCHW fp32 features over channels and write them channels-last
// as 16-bit tensor-core operands (hi [+ lo residual]) and/or as fp32 (alternate-path operand).
// ref: core/corr.py:58-59 (F.normalize(fmap, dim=1): x / max(||x||_2, 1e-12)).
//
// HBM-bound transpose.  One CTA = one (b, h) row x 32 consecutive w.  The [C x 32] fp32 tile is read
// once with 16-byte loads (128 B per channel row), transposed through shared memory (pitch C+1 words:
// conflict-free both ways), reduced per pixel with warp shuffles and written as 512 B (16-bit) /
// 1 KB (fp32) contiguous pixels.  Algorithmic bytes per pixel: 4C in + 2C (hi) [+ 2C lo] [+ 4C n32].
#include "tcs_common.cuh"

namespace tcs {

#ifndef TCS_PRE_TILE_W
#define TCS_PRE_TILE_W 32
#endif
constexpr int kPreTileW = TCS_PRE_TILE_W;          // pixels per CTA: 32 (128-byte channel rows) or 64 (256-byte rows)
constexpr int kPreThreads = 256;
constexpr int kPreLoadLanes = kPreTileW / 4;        // threads that cover one channel row with 16-byte loads
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

template <bool kFp16>
__global__ void __launch_bounds__(kPreThreads)
corr_prepass_kernel(const float* __restrict__ fmap, uint32_t* __restrict__ hi, uint32_t* __restrict__ lo,
                    float* __restrict__ n32, int C, int H, int W, int kblocked) {
    extern __shared__ float tile[];  // [32][C + 1]
    const int pitch = C + 1;
    const int w0 = blockIdx.x * kPreTileW;
    const int h = blockIdx.y;
    const int b = blockIdx.z;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    // ---- load [C][32] (w fastest in global) -> tile[w][c]
    {
        const int w4 = (tid % kPreLoadLanes) * 4;
        const int c_off = tid / kPreLoadLanes;
        const bool vec_ok = ((W & 3) == 0) && (w0 + w4 + 3 < W);
        const size_t plane = (size_t)H * W;
        const float* src = fmap + ((size_t)b * C * H + h) * W + w0 + w4;
        for (int c = c_off; c < C; c += kPreThreads / kPreLoadLanes) {
            const float* p = src + (size_t)c * plane;
            float v[4];
            if (vec_ok) {
                float4 t = ldg_stream_f4(reinterpret_cast<const float4*>(p));
                v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = (w0 + w4 + i < W) ? __ldg(p + i) : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) tile[(w4 + i) * pitch + c] = v[i];
        }
    }
    __syncthreads();

    // ---- per pixel: norm over channels, normalise, emit.  Each warp owns 4 pixels; their reductions are
    // interleaved (independent shuffle chains) and x / denom uses the exact 3-instruction division by a
    // loop-invariant denominator.
    constexpr int kPix = kPreTileW / (kPreThreads / 32);   // 4
    const int pairs = C >> 6;  // channel pairs per lane: channels 2*lane + 64*k, +1
    float ss[kPix];
#pragma unroll
    for (int i = 0; i < kPix; ++i) {
        const float* row = tile + (warp * kPix + i) * pitch;
        float acc = 0.0f;
        for (int c = lane; c < C; c += 32) {
            const float v = row[c];
            acc = fmaf(v, v, acc);
        }
        ss[i] = acc;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < kPix; ++i) ss[i] += __shfl_xor_sync(0xffffffffu, ss[i], o);
#pragma unroll
    for (int i = 0; i < kPix; ++i) {
        const int wl = warp * kPix + i;
        const int w = w0 + wl;
        if (w >= W) break;  // warp-uniform
        const float* row = tile + wl * pitch;
        const float denom = fmaxf(sqrtf(ss[i]), 1e-12f);
        const float rden = __frcp_rn(denom);
        const size_t pix = ((size_t)b * H + h) * W + w;
        for (int k = 0; k < pairs; ++k) {
            const int c = 2 * lane + 64 * k;
            const float a = div_by_const(row[c], denom, rden);
            const float d = div_by_const(row[c + 1], denom, rden);
            if (n32 != nullptr) *reinterpret_cast<float2*>(n32 + pix * C + c) = make_float2(a, d);
            if (hi != nullptr) {
                uint32_t lo_pack;
                const uint32_t hi_pack = pack_hi_lo<kFp16>(a, d, lo_pack);
                // pixel-major [B,H,W,C] or K-block-major [B,H,C/64,W,64]: either way this warp store is one 128-byte line
                const size_t e = kblocked ? ((((size_t)b * H + h) * pairs + k) * W + w) * 64 + 2 * lane : pix * C + c;
                hi[e >> 1] = hi_pack;
                if (lo != nullptr) lo[e >> 1] = lo_pack;
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
    const size_t smem = (size_t)kPreTileW * (C + 1) * sizeof(float);
    dim3 grid(ceil_div(W, kPreTileW), H, B);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (fp16) {
        TCS_ONCE_PER_DEVICE(
            TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_prepass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPreTileW * 513 * 4));
            { const int cv = carveout_percent("TCS_CARVE_PREPASS", -1); if (cv >= 0) TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_prepass_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cv)); }
        );
        corr_prepass_kernel<true><<<grid, kPreThreads, smem, s>>>(fmap, (uint32_t*)hi, (uint32_t*)lo, n32, C, H, W, kblocked);
    } else {
        TCS_ONCE_PER_DEVICE(
            TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_prepass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPreTileW * 513 * 4));
            { const int cv = carveout_percent("TCS_CARVE_PREPASS", -1); if (cv >= 0) TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_prepass_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cv)); }
        );
        corr_prepass_kernel<false><<<grid, kPreThreads, smem, s>>>(fmap, (uint32_t*)hi, (uint32_t*)lo, n32, C, H, W, kblocked);
    }
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
