// Relative camera pose on the device: out = T2 * inv(T1) for batched world2cam 4x4 matrices.
// ref: core/utils/geo_utils.py:148-155 (cal_relative_transformation = matmul(T2, linalg.inv(T1))), called twice per
// temporal frame (core/tc_stereo.py:127,159).  torch.linalg.inv is a batched LU that reads its `info` back on the host
// (a device->host sync per call); this is one tiny launch with none.
#include "tcs_common.cuh"

namespace tcs {

// One thread per matrix.  Gauss-Jordan with partial pivoting in fp64, the product in fp64, one rounding to fp32 at
// the end: the result is the correctly rounded T2 * inv(T1) for any well-conditioned T1 (the reference's fp32 LU +
// fp32 matmul agree with it to a few ulp; parity gate 1e-5 rel + 1e-6 abs).  A singular T1 yields NaN/Inf rows exactly
// as a division by a zero pivot does; no status is reported from the device (the C-ABI never syncs).
__global__ void relative_pose_kernel(const float* __restrict__ T1, const float* __restrict__ T2, float* __restrict__ out, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double a[4][8];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            a[r][c] = (double)T1[b * 16 + r * 4 + c];
            a[r][4 + c] = (r == c) ? 1.0 : 0.0;
        }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int piv = k;
        double best = fabs(a[k][k]);
#pragma unroll
        for (int r = k + 1; r < 4; ++r) {
            const double v = fabs(a[r][k]);
            if (v > best) { best = v; piv = r; }
        }
#pragma unroll
        for (int r = k + 1; r < 4; ++r)
            if (r == piv) {
#pragma unroll
                for (int c = 0; c < 8; ++c) { const double t = a[k][c]; a[k][c] = a[r][c]; a[r][c] = t; }
            }
        const double inv = 1.0 / a[k][k];
#pragma unroll
        for (int c = 0; c < 8; ++c) a[k][c] *= inv;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (r == k) continue;
            const double f = a[r][k];
#pragma unroll
            for (int c = 0; c < 8; ++c) a[r][c] = fma(-f, a[k][c], a[r][c]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) s = fma((double)T2[b * 16 + r * 4 + k], a[k][4 + c], s);
            out[b * 16 + r * 4 + c] = (float)s;
        }
}

}  // namespace tcs

extern "C" int tcs_relative_pose(const float* T1, const float* T2, float* out, int B, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(T1 != nullptr && T2 != nullptr && out != nullptr, TCS_E_BADARG, "tcs_relative_pose: null pointer");
    TCS_REQUIRE(B > 0, TCS_E_BADARG, "tcs_relative_pose: B=%d", B);
    relative_pose_kernel<<<ceil_div(B, 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(T1, T2, out, B);
    TCS_CHECK_LAUNCH("tcs_relative_pose");
    return 0;
}
