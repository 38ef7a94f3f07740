// Exact-fp32 correlation build on the CUDA cores (no tensor cores): strict-parity mode and the
// on-device cross-check of the tcgen05 path.  Same outputs as corr_build.cu.
// ref: core/corr.py:54-62 (corr) and core/corr.py:15-23 (pyramid).
//
// Per (b,h) row an "NT" SGEMM C[w1,w2] = sum_c A[w1,c] * B[w2,c] with both operands channels-last
// (K contiguous).  CTA tile 64 (w1) x 128 (w2), 256 threads, 4 x 8 accumulators per thread arranged
// as two groups of 4 consecutive columns (tx*4 and 64+tx*4) so that shared-memory reads are
// conflict-free and levels 1,2 pool inside the thread; level 3 needs one lane^1 shuffle.
#include "tcs_common.cuh"

namespace tcs {

constexpr int kFm = 64, kFn = 128, kFk = 16, kFThreads = 256;

struct BuildFp32Params {
    const float* a;
    const float* b;
    float* lvl[TCS_MAX_LEVELS];
    int W1, W2, C, num_levels, m_tiles, n_tiles;
};

__device__ __forceinline__ void store_n(float* row, int col, int limit, bool vec_ok, const float* v, int n) {
    if (n == 4 && vec_ok && col + 3 < limit) {
        *reinterpret_cast<float4*>(row + col) = make_float4(v[0], v[1], v[2], v[3]);
    } else if (n == 2 && vec_ok && col + 1 < limit) {
        *reinterpret_cast<float2*>(row + col) = make_float2(v[0], v[1]);
    } else {
        for (int i = 0; i < n; ++i)
            if (col + i < limit) row[col + i] = v[i];
    }
}

__global__ void __launch_bounds__(kFThreads)
corr_build_fp32_kernel(const BuildFp32Params p) {
    __shared__ float As[kFk][kFm + 4];
    __shared__ float Bs[kFk][kFn + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m_t = blockIdx.x % p.m_tiles, n_t = blockIdx.x / p.m_tiles;
    const int bh = blockIdx.y;
    const int m0 = m_t * kFm, n0 = n_t * kFn;
    const int C = p.C;
    const float* A = p.a + (size_t)bh * p.W1 * C;
    const float* Bm = p.b + (size_t)bh * p.W2 * C;

    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

    const int lr = tid >> 2, lk = (tid & 3) * 4;  // loader: row, k offset
    for (int k0 = 0; k0 < C; k0 += kFk) {
        {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + lr < p.W1) v = *reinterpret_cast<const float4*>(A + (size_t)(m0 + lr) * C + k0 + lk);
            As[lk][lr] = v.x; As[lk + 1][lr] = v.y; As[lk + 2][lr] = v.z; As[lk + 3][lr] = v.w;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = lr + 64 * r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + row < p.W2) v = *reinterpret_cast<const float4*>(Bm + (size_t)(n0 + row) * C + k0 + lk);
            Bs[lk][row] = v.x; Bs[lk + 1][row] = v.y; Bs[lk + 2][row] = v.z; Bs[lk + 3][row] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kFk; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue: levels 0..3
    const int W1 = p.W1, W2 = p.W2;
    const int W2_1 = W2 >> 1, W2_2 = W2 >> 2, W2_3 = W2 >> 3;
    const bool vec0 = (W2 & 3) == 0, vec1 = (W2_1 & 1) == 0, vec3ok = false;
    (void)vec3ok;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + ty * 4 + i;
        const size_t rb = (size_t)bh * W1 + row;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int col = n0 + g * 64 + tx * 4;
            const float* v = &acc[i][g * 4];
            const float l1[2] = {(v[0] + v[1]) * 0.5f, (v[2] + v[3]) * 0.5f};
            const float l2 = (l1[0] + l1[1]) * 0.5f;
            const float l2n = __shfl_xor_sync(0xffffffffu, l2, 1);  // all lanes participate
            if (row < W1) {
                store_n(p.lvl[0] + rb * W2, col, W2, vec0, v, 4);
                if (p.num_levels > 1) store_n(p.lvl[1] + rb * W2_1, col >> 1, W2_1, vec1, l1, 2);
                if (p.num_levels > 2 && (col >> 2) < W2_2) p.lvl[2][rb * W2_2 + (col >> 2)] = l2;
                if (p.num_levels > 3 && (tx & 1) == 0 && (col >> 3) < W2_3)
                    p.lvl[3][rb * W2_3 + (col >> 3)] = (l2 + l2n) * 0.5f;
            }
        }
    }
}

}  // namespace tcs

extern "C" int tcs_corr_build_fp32(const float* a_n32, const float* b_n32,
                                   float* lvl0, float* lvl1, float* lvl2, float* lvl3,
                                   int B, int H, int W1, int W2, int C, int num_levels, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(a_n32 != nullptr && b_n32 != nullptr && lvl0 != nullptr, TCS_E_BADARG, "tcs_corr_build_fp32: null pointer");
    TCS_REQUIRE(num_levels >= 1 && num_levels <= TCS_MAX_LEVELS, TCS_E_SHAPE, "tcs_corr_build_fp32: num_levels=%d not in [1,4]", num_levels);
    float* lv[4] = {lvl0, lvl1, lvl2, lvl3};
    for (int l = 0; l < num_levels; ++l)
        TCS_REQUIRE(lv[l] != nullptr && aligned16(lv[l]), TCS_E_ALIGN, "tcs_corr_build_fp32: level %d pointer null or unaligned", l);
    TCS_REQUIRE(B > 0 && H > 0 && W1 > 0 && W2 >= 8 && C > 0 && C % kFk == 0, TCS_E_SHAPE, "tcs_corr_build_fp32: bad sizes (C %% 16 == 0, W2 >= 8)");
    TCS_REQUIRE(aligned16(a_n32) && aligned16(b_n32), TCS_E_ALIGN, "tcs_corr_build_fp32: operands must be 16-byte aligned");
    TCS_REQUIRE((long long)B * H <= 65535, TCS_E_SHAPE, "tcs_corr_build_fp32: B*H must be <= 65535");
    BuildFp32Params p{};
    p.a = a_n32; p.b = b_n32;
    for (int l = 0; l < 4; ++l) p.lvl[l] = lv[l];
    p.W1 = W1; p.W2 = W2; p.C = C; p.num_levels = num_levels;
    p.m_tiles = ceil_div(W1, kFm); p.n_tiles = ceil_div(W2, kFn);
    dim3 grid(p.m_tiles * p.n_tiles, B * H);
    corr_build_fp32_kernel<<<grid, kFThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
    TCS_CHECK_LAUNCH("tcs_corr_build_fp32");
    return 0;
}
