// Cost-volume initialisation loss, the per-pixel terms.   ref: train_stereo.py:150-172 (init_loss)
//
// The reference reads the masked, transposed cost volume cv[b, w2, h, w1] = corr[b, h, w1, w2] * [w2 <= w1] (corr.py:25-31) about
// ten times: two gathers for phi(index_gt), an int64 index volume (repeat), two float comparisons over the volume, masked_fill and
// a top-k along the strided w2 axis.  Here one warp owns one pixel (b, h, w1) and reads ITS ROW of level 0 once — corr[b,h,w1,:]
// is contiguous in the pyramid, so neither the transposed copy nor the index volume ever exists:
//   phi      = frac * rho(df + 1) + (1 - frac) * rho(df),  df = floor(index_gt), rho(i) = cv[clip(i, 0, D-1)]      (:151-158)
//   cv_nm    = cv with [index_gt - 1.5, index_gt + 1.5) and every column of an unmasked pixel filled with 0          (:166-170)
//   cost_nm  = the k largest entries of cv_nm along w2, descending                                                    (:171)
// plus, for the backward, the w2 of each of the k entries (-1 when the entry is a filled or masked zero: no gradient reaches
// the volume through it, exactly as masked_fill / the [w2 <= w1] product cut it in the reference).  Ties between equal values go
// to the lowest w2 (torch.topk leaves the order of ties unspecified; the loss does not depend on it).
// Backward: one thread per pixel adds (1 - frac) g, frac g and the k top-k gradients into its own row of d(volume): no atomics.
#include "tcs_common.cuh"

namespace tcs {
namespace initloss {

constexpr int kMaxK = 8;
constexpr int kPerLane = 16;                 // W2 <= 512

__global__ void __launch_bounds__(256)
init_loss_forward_kernel(const float* __restrict__ vol, int W2p, const float* __restrict__ index_gt, const unsigned char* __restrict__ mask,
                         float* __restrict__ phi, float* __restrict__ cost_nm, int* __restrict__ idx_nm, int B, int H, int W1,
                         int W2, int k) {
    const int lane = threadIdx.x & 31;
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long npix = (long long)B * H * W1;
    if (p >= npix) return;                                         // warp-uniform
    const int w1 = (int)(p % W1);
    const long long bh = p / W1;
    const int h = (int)(bh % H), b = (int)(bh / H);
    const float* row = vol + p * W2p;
    const int D = W2;
    const float d = __ldg(index_gt + p);                           // already clipped to [0, D - 1] (:164)
    const bool m = __ldg(mask + p) != 0;
    const float low = __fsub_rn(d, 1.5f), high = __fadd_rn(d, 1.5f);

    float v[kPerLane];
    bool elig[kPerLane];                                           // a gradient can reach the volume through this entry
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
        const int w2 = lane + 32 * i;
        const bool inside = w2 < D;
        const float x = (inside && w2 <= w1) ? __ldg(row + w2) : 0.0f;                 // corr.py:28-31: zero where w2 > w1
        const bool filled = !m || ((float)w2 >= low && (float)w2 < high);              // :168-170
        elig[i] = inside && !filled && w2 <= w1;
        v[i] = inside ? (filled ? 0.0f : x) : -INFINITY;
    }
    if (lane == 0) {
        const float dff = floorf(d);
        const int df = (int)dff;
        const float frac = __fsub_rn(d, dff);
        const int i0 = min(max(df, 0), D - 1), i1 = min(max(df + 1, 0), D - 1);        // rho clips (:151-152)
        const float r0 = i0 <= w1 ? __ldg(row + i0) : 0.0f, r1 = i1 <= w1 ? __ldg(row + i1) : 0.0f;
        phi[p] = __fadd_rn(__fmul_rn(frac, r1), __fmul_rn(__fsub_rn(1.0f, frac), r0));  // :158
    }
    for (int j = 0; j < k; ++j) {
        float best = -INFINITY;
        int bi = 0x7fffffff;
        bool be = false;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) {                       // ascending w2 within the lane: the first maximum is the lowest index
            if (v[i] > best) { best = v[i]; bi = lane + 32 * i; be = elig[i]; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const bool oe = __shfl_xor_sync(0xffffffffu, (int)be, o) != 0;
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; be = oe; }
        }
#pragma unroll
        for (int i = 0; i < kPerLane; ++i)
            if (lane + 32 * i == bi) v[i] = -INFINITY;             // taken
        if (lane == 0) {
            const long long o = (((long long)b * k + j) * H + h) * W1 + w1;
            cost_nm[o] = best;                                     // D >= k is checked by the host, so best is finite
            idx_nm[o] = be ? bi : -1;
        }
    }
}

__global__ void __launch_bounds__(256)
init_loss_backward_kernel(const float* __restrict__ g_phi, const float* __restrict__ g_nm, const float* __restrict__ index_gt,
                          const int* __restrict__ idx_nm, float* __restrict__ g_vol, int B, int H, int W1, int W2, int k) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long npix = (long long)B * H * W1;
    if (p >= npix) return;
    const int w1 = (int)(p % W1);
    const long long bh = p / W1;
    const int h = (int)(bh % H), b = (int)(bh / H);
    float* row = g_vol + p * W2;                                   // zero-filled by the host wrapper; only this thread touches it
    const int D = W2;
    const float d = __ldg(index_gt + p);
    const float dff = floorf(d);
    const int df = (int)dff;
    const float frac = __fsub_rn(d, dff);
    const int i0 = min(max(df, 0), D - 1), i1 = min(max(df + 1, 0), D - 1);
    const float g = __ldg(g_phi + p);
    if (i0 <= w1) row[i0] += __fmul_rn(__fsub_rn(1.0f, frac), g);
    if (i1 <= w1) row[i1] += __fmul_rn(frac, g);
    for (int j = 0; j < k; ++j) {
        const long long o = (((long long)b * k + j) * H + h) * W1 + w1;
        const int idx = __ldg(idx_nm + o);
        if (idx >= 0) row[idx] += __ldg(g_nm + o);
    }
}

}  // namespace initloss
}  // namespace tcs

extern "C" int tcs_init_loss_forward(const float* level0, int W2_pitch, const float* index_gt, const unsigned char* mask,
                                     float* phi, float* cost_nm, int* idx_nm, int B, int H, int W1, int W2, int k, void* stream) {
    using namespace tcs;
    using namespace tcs::initloss;
    TCS_REQUIRE(level0 && index_gt && mask && phi && cost_nm && idx_nm, TCS_E_BADARG, "tcs_init_loss_forward: null pointer");
    TCS_REQUIRE(B > 0 && H > 0 && W1 > 0 && W2 > 0, TCS_E_BADARG, "tcs_init_loss_forward: bad sizes");
    TCS_REQUIRE(W2 <= 32 * kPerLane, TCS_E_SHAPE, "tcs_init_loss_forward: W2=%d above %d", W2, 32 * kPerLane);
    TCS_REQUIRE(k >= 1 && k <= kMaxK && k <= W2, TCS_E_SHAPE, "tcs_init_loss_forward: k=%d not in [1, min(%d, W2)]", k, kMaxK);
    const int W2p = W2_pitch > 0 ? W2_pitch : W2;
    TCS_REQUIRE(W2p >= W2, TCS_E_SHAPE, "tcs_init_loss_forward: row pitch %d below W2=%d", W2p, W2);
    const long long npix = (long long)B * H * W1;
    const long long blocks = ceil_div_ll(npix, 8);
    TCS_REQUIRE(blocks < 0x7fffffffLL, TCS_E_SHAPE, "tcs_init_loss_forward: too many pixels");
    init_loss_forward_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(level0, W2p, index_gt, mask, phi, cost_nm,
                                                                                              idx_nm, B, H, W1, W2, k);
    TCS_CHECK_LAUNCH("tcs_init_loss_forward");
    return 0;
}

extern "C" int tcs_init_loss_backward(const float* grad_phi, const float* grad_cost_nm, const float* index_gt, const int* idx_nm,
                                      float* grad_level0, int B, int H, int W1, int W2, int k, void* stream) {
    using namespace tcs;
    using namespace tcs::initloss;
    TCS_REQUIRE(grad_phi && grad_cost_nm && index_gt && idx_nm && grad_level0, TCS_E_BADARG, "tcs_init_loss_backward: null pointer");
    TCS_REQUIRE(B > 0 && H > 0 && W1 > 0 && W2 > 0 && k >= 1 && k <= kMaxK, TCS_E_BADARG, "tcs_init_loss_backward: bad sizes");
    const long long npix = (long long)B * H * W1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TCS_CHECK_CUDA(cudaMemsetAsync(grad_level0, 0, (size_t)npix * W2 * sizeof(float), s));
    init_loss_backward_kernel<<<(unsigned)ceil_div_ll(npix, 256), 256, 0, s>>>(grad_phi, grad_cost_nm, index_gt, idx_nm, grad_level0,
                                                                                 B, H, W1, W2, k);
    TCS_CHECK_LAUNCH("tcs_init_loss_backward");
    return 0;
}
