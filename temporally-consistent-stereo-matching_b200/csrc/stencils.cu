// Per-GRU-iteration 3x3 stencils on the quarter-resolution disparity: each of them is a
// grouped conv2d with a one-hot / difference kernel plus a few elementwise ops in the reference (5-15 launches and
// several padded temporaries apiece); here one kernel each, one thread per pixel, everything in registers.
//   tcs_disp_gradient_xy      ref: core/utils/geo_utils.py:115-132 (disp2disp_gradient_xy)
//   tcs_disp_grad_candidates  ref: core/utils/geo_utils.py:73-101  (disp2disp_grad_candidates)
//   tcs_disp_propagate        ref: core/update.py:259-289          (DispRefine.propagate_disparity)
//   tcs_convex_upsample       ref: core/tc_stereo.py:75-88         (TCStereo.upsample_flow; rank 3)
// All arithmetic that the reference does on small integers (coordinate differences 0, +-1, +-2 ...) is exact in fp32,
// every product below has such a factor, and the remaining additions / divisions are single IEEE operations in the
// reference's order: the results are bit-identical to the reference's, not merely close.
#include "tcs_common.cuh"

namespace tcs {

// 8-neighbourhood in the order the reference walks it (v = row, u = column of the 3x3 kernel)   geo_utils.py:83
__constant__ int kRingV[8] = {0, 0, 0, 1, 2, 2, 2, 1};
__constant__ int kRingU[8] = {0, 1, 2, 2, 2, 1, 0, 0};

// ---- disp2disp_gradient_xy: forward differences on the replicate-padded map, and the "no big gradient" mask -------
__global__ void __launch_bounds__(256)
disp_gradient_xy_kernel(const float* __restrict__ disp, float* __restrict__ grads, unsigned char* __restrict__ edge_mask,
                        int H, int W, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int HW = H * W;
    const long long n = i / HW;
    const int p = (int)(i - n * HW);
    const int y = p / W, x = p - y * W;
    const float* d = disp + n * HW;
    const float c = d[p];
    const float gx = __fsub_rn(d[y * W + min(x + 1, W - 1)], c);      // kernel (1,2) - centre
    const float gy = __fsub_rn(d[min(y + 1, H - 1) * W + x], c);      // kernel (2,1) - centre
    grads[(n * 2 + 0) * HW + p] = gx;
    grads[(n * 2 + 1) * HW + p] = gy;
    if (edge_mask != nullptr) edge_mask[i] = (fabsf(gx) < 5.0f && fabsf(gy) < 5.0f) ? 1 : 0;   // geo_utils.py:130
}

// ---- disp2disp_grad_candidates: normals of the triangles (centre, ring[k], ring[k+2]) of the surface (x, y, disp) ---
// Level i uses the ring at distance i + 1 on the ZERO-padded map; the rings of all levels are concatenated before the
// "k + 2" pairing, so the last two entries of a level pair with the first two of the next (and wrap at the end).
constexpr int kMaxGradLevels = 4;

template <int kLevels>
__global__ void __launch_bounds__(256)
disp_grad_candidates_kernel(const float* __restrict__ disp, float* __restrict__ out, int H, int W, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int HW = H * W;
    const long long n = i / HW;
    const int p = (int)(i - n * HW);
    const int y = p / W, x = p - y * W;
    const float* d = disp + n * HW;
    const float c = d[p];
    constexpr int K = 8 * kLevels;
    float gx[K], gy[K], gd[K];                                       // statically indexed: registers
#pragma unroll
    for (int l = 0; l < kLevels; ++l) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int du = (kRingU[k] - 1) * (l + 1), dv = (kRingV[k] - 1) * (l + 1);
            const int xx = x + du, yy = y + dv;
            const float nb = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? d[yy * W + xx] : 0.0f;   // F.pad zeros
            gx[8 * l + k] = (float)du;
            gy[8 * l + k] = (float)dv;
            gd[8 * l + k] = __fsub_rn(nb, c);
        }
    }
    float* o0 = out + (n * 2 + 0) * (long long)K * HW + p;
    float* o1 = out + (n * 2 + 1) * (long long)K * HW + p;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int k2 = (k + 2) % K;                                    // torch.roll(grads, -2, dim=2)
        // torch.cross over (x, y, disp)
        const float c0 = __fsub_rn(__fmul_rn(gy[k], gd[k2]), __fmul_rn(gd[k], gy[k2]));
        const float c1 = __fsub_rn(__fmul_rn(gd[k], gx[k2]), __fmul_rn(gx[k], gd[k2]));
        const float c2 = __fsub_rn(__fmul_rn(gx[k], gy[k2]), __fmul_rn(gy[k], gx[k2]));
        o0[(long long)k * HW] = __fdiv_rn(-c0, c2);
        o1[(long long)k * HW] = __fdiv_rn(-c1, c2);
    }
}

// ---- DispRefine.propagate_disparity: the 9 plane-extrapolated disparities and the 18 |gradient differences| --------
__global__ void __launch_bounds__(256)
disp_propagate_kernel(const float* __restrict__ grad, const float* __restrict__ disp, float* __restrict__ prop,
                      float* __restrict__ matrix, int H, int W, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int HW = H * W;
    const long long n = i / HW;
    const int p = (int)(i - n * HW);
    const int y = p / W, x = p - y * W;
    const float* d = disp + n * HW;
    const float* g0 = grad + (n * 2 + 0) * HW;
    const float* g1 = grad + (n * 2 + 1) * HW;
    const float gcx = g0[p], gcy = g1[p];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int v = k / 3, u = k - 3 * v;                            // update.py:221: row-major 3x3
        const int xx = x + u - 1, yy = y + v - 1;
        const bool in = xx >= 0 && xx < W && yy >= 0 && yy < H;
        const float m = d[min(max(yy, 0), H - 1) * W + min(max(xx, 0), W - 1)];      // replicate padding
        const float gx = in ? g0[yy * W + xx] : 0.0f, gy = in ? g1[yy * W + xx] : 0.0f;   // zero padding
        const float cx = (float)(1 - u), cy = (float)(1 - v);          // centre - neighbour coordinates
        // disparity_map_prop + grad_x * dx + grad_y * dy, left to right                       update.py:283
        prop[(n * 9 + k) * HW + p] = __fadd_rn(__fadd_rn(m, __fmul_rn(gx, cx)), __fmul_rn(gy, cy));
        matrix[(n * 18 + k) * HW + p] = fabsf(__fsub_rn(gcx, gx));
        matrix[(n * 18 + 9 + k) * HW + p] = fabsf(__fsub_rn(gcy, gy));
    }
}

// ---- TCStereo.upsample_flow: convex combination of the 3x3 coarse neighbours ------------
// One thread per (coarse pixel, sub-row): lanes are consecutive x, blockIdx.y is the sub-row, so every logit load is
// coalesced and all 9 f of them are in flight at once; max / exp / sum / divide and the weighted sum in registers; the
// f sub-pixels of the row leave as one contiguous store per lane (f = 4: a 16-byte store, 512 bytes per warp).
template <int kF>
__global__ void __launch_bounds__(128)
convex_upsample_kernel(const float* __restrict__ flow, const float* __restrict__ mask, float* __restrict__ out,
                       int D, int H, int W, int scale, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int si = blockIdx.y;
    const int HW = H * W;
    const long long n = i / HW;
    const int p = (int)(i - n * HW);
    const int y = p / W, x = p - y * W;
    const float* mk = mask + n * 9 * kF * kF * (long long)HW + p;
    float w[kF][9];
#pragma unroll
    for (int sj = 0; sj < kF; ++sj)
#pragma unroll
        for (int k = 0; k < 9; ++k) w[sj][k] = ldg_stream_f1(mk + (long long)((k * kF + si) * kF + sj) * HW);
#pragma unroll
    for (int sj = 0; sj < kF; ++sj) {                                   // softmax(mask - max) over k   tc_stereo.py:80-81
        float mx = w[sj][0];
#pragma unroll
        for (int k = 1; k < 9; ++k) mx = fmaxf(mx, w[sj][k]);
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            w[sj][k] = expf(__fsub_rn(w[sj][k], mx));
            s = __fadd_rn(s, w[sj][k]);
        }
#pragma unroll
        for (int k = 0; k < 9; ++k) w[sj][k] = __fdiv_rn(w[sj][k], s);
    }
    for (int d = 0; d < D; ++d) {
        const float* fl = flow + (n * D + d) * HW;
        float r[kF];
#pragma unroll
        for (int sj = 0; sj < kF; ++sj) r[sj] = 0.0f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;               // F.unfold(.., [3,3], padding=1): zeros outside
            const float v = (xx >= 0 && xx < W && yy >= 0 && yy < H) ? fl[yy * W + xx] : 0.0f;
            const float u = scale ? __fmul_rn((float)kF, v) : v;
#pragma unroll
            for (int sj = 0; sj < kF; ++sj) r[sj] = __fadd_rn(r[sj], __fmul_rn(w[sj][k], u));
        }
        float* row = out + ((n * D + d) * (long long)(kF * H) + (long long)y * kF + si) * (kF * W) + (long long)x * kF;
        if (kF == 4) {
            *reinterpret_cast<float4*>(row) = make_float4(r[0], r[1], r[2], r[3]);
        } else {
#pragma unroll
            for (int sj = 0; sj < kF; ++sj) row[sj] = r[sj];
        }
    }
}

}  // namespace tcs

extern "C" int tcs_convex_upsample(const float* flow, const float* mask, float* out, int N, int D, int H, int W,
                                   int factor, int scale, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(flow && mask && out, TCS_E_BADARG, "tcs_convex_upsample: null pointer");
    TCS_REQUIRE(N > 0 && D > 0 && H > 0 && W > 0 && (long long)H * W * 16 < 0x7fffffffLL, TCS_E_BADARG, "tcs_convex_upsample: bad sizes");
    TCS_REQUIRE(factor == 2 || factor == 4 || factor == 8, TCS_E_SHAPE, "tcs_convex_upsample: factor=%d must be 2, 4 or 8", factor);
    TCS_REQUIRE(aligned16(out), TCS_E_ALIGN, "tcs_convex_upsample: out must be 16-byte aligned");
    const long long total = (long long)N * H * W;
    const unsigned blocks = (unsigned)ceil_div_ll(total, 128);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (factor) {
        case 2: convex_upsample_kernel<2><<<dim3(blocks, 2), 128, 0, s>>>(flow, mask, out, D, H, W, scale, total); break;
        case 4: convex_upsample_kernel<4><<<dim3(blocks, 4), 128, 0, s>>>(flow, mask, out, D, H, W, scale, total); break;
        default: convex_upsample_kernel<8><<<dim3(blocks, 8), 128, 0, s>>>(flow, mask, out, D, H, W, scale, total); break;
    }
    TCS_CHECK_LAUNCH("tcs_convex_upsample");
    return 0;
}

extern "C" int tcs_disp_gradient_xy(const float* disp, float* grads, unsigned char* edge_mask, int N, int H, int W, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(disp && grads, TCS_E_BADARG, "tcs_disp_gradient_xy: null pointer");
    TCS_REQUIRE(N > 0 && H > 0 && W > 0 && (long long)H * W < 0x7fffffffLL, TCS_E_BADARG, "tcs_disp_gradient_xy: bad sizes");
    const long long total = (long long)N * H * W;
    disp_gradient_xy_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(disp, grads, edge_mask, H, W, total);
    TCS_CHECK_LAUNCH("tcs_disp_gradient_xy");
    return 0;
}

extern "C" int tcs_disp_grad_candidates(const float* disp, float* out, int N, int H, int W, int levels, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(disp && out, TCS_E_BADARG, "tcs_disp_grad_candidates: null pointer");
    TCS_REQUIRE(N > 0 && H > 0 && W > 0 && (long long)H * W < 0x7fffffffLL, TCS_E_BADARG, "tcs_disp_grad_candidates: bad sizes");
    TCS_REQUIRE(levels >= 1 && levels <= kMaxGradLevels, TCS_E_SHAPE, "tcs_disp_grad_candidates: level=%d must be 1..%d", levels, kMaxGradLevels);
    const long long total = (long long)N * H * W;
    const unsigned blocks = (unsigned)ceil_div_ll(total, 256);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (levels) {
        case 1: disp_grad_candidates_kernel<1><<<blocks, 256, 0, s>>>(disp, out, H, W, total); break;
        case 2: disp_grad_candidates_kernel<2><<<blocks, 256, 0, s>>>(disp, out, H, W, total); break;
        case 3: disp_grad_candidates_kernel<3><<<blocks, 256, 0, s>>>(disp, out, H, W, total); break;
        default: disp_grad_candidates_kernel<4><<<blocks, 256, 0, s>>>(disp, out, H, W, total); break;
    }
    TCS_CHECK_LAUNCH("tcs_disp_grad_candidates");
    return 0;
}

extern "C" int tcs_disp_propagate(const float* grad, const float* disp, float* prop, float* matrix, int N, int H, int W, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(grad && disp && prop && matrix, TCS_E_BADARG, "tcs_disp_propagate: null pointer");
    TCS_REQUIRE(N > 0 && H > 0 && W > 0 && (long long)H * W < 0x7fffffffLL, TCS_E_BADARG, "tcs_disp_propagate: bad sizes");
    const long long total = (long long)N * H * W;
    disp_propagate_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(grad, disp, prop, matrix, H, W, total);
    TCS_CHECK_LAUNCH("tcs_disp_propagate");
    return 0;
}
