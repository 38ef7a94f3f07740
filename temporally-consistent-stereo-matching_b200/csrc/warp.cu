// Temporal step: carry the previous frame's disparity, features and hidden states into the current view.
//   tcs_warp_forward     geometry -> soft-splat (vector red.global.add) -> normalise + mask + matching cost
//                        ref: core/utils/geo_utils.py:158-198 (warp), core/utils/splatting/softsplat.py:232-274,
//                        :284-335 (softsplat / softsplat_out), core/tc_stereo.py:139-140 (cost)
//   tcs_backward_grid    ref: core/utils/geo_utils.py:201-236 (get_backward_grid)
//   tcs_bilinear_sample  ref: core/utils/utils.py:82-97 (bilinear_sampler -> F.grid_sample)
//   tcs_grid_halve       ref: core/tc_stereo.py:163 (0.5 * F.interpolate(grid, 0.5, bilinear, align_corners))
//
// The splat is where the bytes are.  The reference launches one thread per (pixel, channel) scalar, so the
// flow, the four targets and the four weights are recomputed 258 times per pixel and every atomic is a
// lone 4-byte red.  Here one warp owns a source pixel: targets and weights are computed once, the 256
// feature channels are read from a shared-memory transposed tile (coalesced NCHW loads), and each lane
// issues one 16-byte red.global.add.v4.f32 per 4 channels and target into a channels-last accumulator
// [B,H,W,C+4].  The channel order inside the accumulator is a private permutation
// (position 128*j + 4*lane + i  <->  channel lane + 32*(4*j + i)) chosen so that both the splat's
// shared-memory reads and the normalise kernel's shared-memory writes are bank-conflict free.
//
// Geometry is evaluated with explicit round-to-nearest intrinsics (no compiler-chosen contraction) in a
// fixed order so that the validity masks and the integer splat targets are reproducible bit for bit by the
// CPU oracle.
#include "tcs_common.cuh"

#include <cmath>
#include <cstdlib>

namespace tcs {

constexpr int kTileW = 32;
constexpr int kWarpThreads = 256;

// control words in the scratch header (zeroed with the sums)
constexpr int kCtrlMaxFlowX = 0, kCtrlMaxFlowY = 1, kCtrlFallback = 2;
constexpr int kGatherMaxR = 24;        // search radius (source pixels) the gather kernel is willing to scan
constexpr int kGatherMaxEntries = 48;  // distinct (dx, dy) offsets a 32-pixel target segment may draw from

struct Cam {
    float K[9], Ki[9], T[12], bf;
};

__device__ __forceinline__ Cam load_cam(const float* __restrict__ rel_T, const float* __restrict__ K,
                                        const float* __restrict__ K_inv, const float* __restrict__ baseline, int b) {
    Cam c;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        c.K[i] = __ldg(K + b * 9 + i);
        c.Ki[i] = __ldg(K_inv + b * 9 + i);
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) c.T[i] = __ldg(rel_T + b * 16 + i);
    c.bf = __fmul_rn(__ldg(baseline + b), c.K[0]);   // baseline * fx   (geo_utils.py:16)
    return c;
}

// The FMA chain a GEMM micro-kernel runs over k = 0, 1, 2 (bit-identical to torch.matmul on the CPU for
// these 3x3 / 4x4 products; the oracle restates it the same way).
__device__ __forceinline__ float dot3(const float* m, float x, float y, float z) {
    return __fmaf_rn(m[2], z, __fmaf_rn(m[1], y, __fmul_rn(m[0], x)));
}

// depth * K^-1 [x,y,1]^T, then the rigid transform (geo_utils.py:32-42, :135-145).
__device__ __forceinline__ void project(const Cam& c, float depth, float x, float y, float (&P)[3]) {
    float q[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) q[i] = __fmul_rn(depth, dot3(c.Ki + 3 * i, x, y, 1.0f));
#pragma unroll
    for (int i = 0; i < 3; ++i) P[i] = __fadd_rn(dot3(c.T + 4 * i, q[0], q[1], q[2]), c.T[4 * i + 3]);
}

__device__ __forceinline__ float finite_or_m1(float v) { return (isnan(v) || isinf(v)) ? -1.0f : v; }

// (K P)_{0,1} / z with NaN/Inf -> -1 (geo_utils.py:45-57).
__device__ __forceinline__ void reproject(const Cam& c, const float (&P)[3], float& u, float& v) {
    u = finite_or_m1(__fdiv_rn(dot3(c.K, P[0], P[1], P[2]), P[2]));
    v = finite_or_m1(__fdiv_rn(dot3(c.K + 3, P[0], P[1], P[2]), P[2]));
}

// ---- forward warp, kernel A: geometry + per-sample disparity sums ----------------------------------------
__global__ void __launch_bounds__(256)
warp_geometry_kernel(const float* __restrict__ disp, const float* __restrict__ rel_T, const float* __restrict__ K,
                     const float* __restrict__ K_inv, const float* __restrict__ baseline,
                     float* __restrict__ disp1, float* __restrict__ tx, float* __restrict__ ty,
                     float* __restrict__ valid, double* __restrict__ sums, int* __restrict__ ctrl, int H, int W) {
    const int b = blockIdx.y;
    const int HW = H * W;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double local = 0.0;
    float fmax_x = 0.0f, fmax_y = 0.0f;   // largest |target - source| of a pixel that will be splatted
    if (i < HW) {
        const Cam c = load_cam(rel_T, K, K_inv, baseline, b);
        const int y = i / W, x = i - y * W;
        const float d = __ldg(disp + (size_t)b * HW + i);
        const float depth = __fdiv_rn(c.bf, fmaxf(d, 0.001f));                 // disp2depth
        float P[3];
        project(c, depth, (float)x, (float)y, P);
        const float d1 = finite_or_m1(__fdiv_rn(c.bf, P[2]));                  // depth2disp
        const bool ok = (d1 > 0.0f) && (d1 < (float)W);                        // geo_utils.py:185
        float u, v;
        reproject(c, P, u, v);
        // flow = coords' - coords0; the splat then uses x + flow (softsplat.py:297-298)
        const float fx_ = __fadd_rn((float)x, __fsub_rn(u, (float)x));
        const float fy_ = __fadd_rn((float)y, __fsub_rn(v, (float)y));
        const size_t o = (size_t)b * HW + i;
        disp1[o] = d1;
        tx[o] = fx_;
        ty[o] = fy_;
        valid[o] = ok ? 1.0f : 0.0f;
        local = (double)d1;
        if (ok && isfinite(fx_) && isfinite(fy_)) {
            fmax_x = fabsf(fx_ - (float)x);
            fmax_y = fabsf(fy_ - (float)y);
        }
    }
    // non-negative floats order like their bit patterns: one atomicMax per warp
    fmax_x = fmaxf(fmax_x, 0.0f);
    fmax_y = fmaxf(fmax_y, 0.0f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        fmax_x = fmaxf(fmax_x, __shfl_xor_sync(0xffffffffu, fmax_x, o));
        fmax_y = fmaxf(fmax_y, __shfl_xor_sync(0xffffffffu, fmax_y, o));
    }
    __shared__ float wmax[2][8];
    if ((threadIdx.x & 31) == 0) { wmax[0][threadIdx.x >> 5] = fmax_x; wmax[1][threadIdx.x >> 5] = fmax_y; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ double wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        float mx = 0.0f, my = 0.0f;
        for (int w = 0; w < 8; ++w) { s += wsum[w]; mx = fmaxf(mx, wmax[0][w]); my = fmaxf(my, wmax[1][w]); }
        atomicAdd(sums + b, s);
        if (mx > 0.0f) atomicMax(ctrl + kCtrlMaxFlowX, __float_as_int(mx));
        if (my > 0.0f) atomicMax(ctrl + kCtrlMaxFlowY, __float_as_int(my));
    }
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" :: "l"(p), "f"(a), "f"(b) : "memory");
}

// [C][32] NCHW tile (w fastest in global) -> tile[w][c], pitch C + 1.
__device__ __forceinline__ void load_tile_transposed(const float* __restrict__ fmap, float* tile, int pitch,
                                                     int b, int h, int w0, int C, int H, int W) {
    const int tid = threadIdx.x;
    const int w4 = (tid & 7) * 4;
    const int c_off = tid >> 3;
    const bool vec_ok = ((W & 3) == 0) && (w0 + w4 + 3 < W);
    const size_t plane = (size_t)H * W;
    const float* src = fmap + ((size_t)b * C * H + h) * W + w0 + w4;
    for (int c = c_off; c < C; c += kWarpThreads / 8) {
        const float* p = src + (size_t)c * plane;
        float v[4];
        if (vec_ok) {
            const float4 t = ldg_stream_f4(reinterpret_cast<const float4*>(p));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = (w0 + w4 + i < W) ? __ldg(p + i) : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) tile[(w4 + i) * pitch + c] = v[i];
    }
}

// ---- forward warp, kernel B: soft-splat -----------------------------------------------------------------------
template <int kGroups>  // C / 128
__global__ void __launch_bounds__(kWarpThreads)
warp_splat_kernel(const float* __restrict__ fmap, const float* __restrict__ disp1, const float* __restrict__ tx,
                  const float* __restrict__ ty, const float* __restrict__ wgt, const int* __restrict__ ctrl,
                  float* __restrict__ accum, int B, int H, int W) {
    if (ctrl[kCtrlFallback] == 0) return;   // deterministic mode and the gather kernel handled this frame
    constexpr int C = kGroups * 128;
    constexpr int CP = C + 4;
    constexpr int pitch = C + 1;
    extern __shared__ float tile[];  // [32][C + 1]
    const int w0 = blockIdx.x * kTileW, h = blockIdx.y, b = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    load_tile_transposed(fmap, tile, pitch, b, h, w0, C, H, W);

    __syncthreads();

#pragma unroll 1
    for (int i = 0; i < kTileW / 8; ++i) {
        const int wl = warp * (kTileW / 8) + i;
        const int w = w0 + wl;
        if (w >= W) break;
        const size_t idx = ((size_t)b * H + h) * W + w;
        const float e = __ldg(wgt + idx);                  // valid * exp(metric); 0 for pixels that contribute nothing
        if (e == 0.0f) continue;
        const float fx_ = __ldg(tx + idx), fy_ = __ldg(ty + idx);
        const float d1 = __ldg(disp1 + idx);
        const int nwx = (int)floorf(fx_), nwy = (int)floorf(fy_);
        const int sex = nwx + 1, sey = nwy + 1;
        // softsplat.py:314-317
        const float wt[4] = {__fmul_rn(__fsub_rn((float)sex, fx_), __fsub_rn((float)sey, fy_)),    // NW
                             __fmul_rn(__fsub_rn(fx_, (float)nwx), __fsub_rn((float)sey, fy_)),    // NE
                             __fmul_rn(__fsub_rn((float)sex, fx_), __fsub_rn(fy_, (float)nwy)),    // SW
                             __fmul_rn(__fsub_rn(fx_, (float)nwx), __fsub_rn(fy_, (float)nwy))};   // SE
        const int txs[4] = {nwx, sex, nwx, sex};
        const int tys[4] = {nwy, nwy, sey, sey};
        float v[kGroups * 4];
        const float* row = tile + wl * pitch;
#pragma unroll
        for (int k = 0; k < kGroups * 4; ++k) v[k] = __fmul_rn(row[lane + 32 * k], e);
        const float tail0 = __fmul_rn(d1, e);  // the disparity channel and the normaliser channel
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (txs[t] < 0 || txs[t] >= W || tys[t] < 0 || tys[t] >= H) continue;
            float* dst = accum + (((size_t)b * H + tys[t]) * W + txs[t]) * CP;
            const float wgt = wt[t];
#pragma unroll
            for (int j = 0; j < kGroups; ++j)
                red_add_v4(dst + j * 128 + 4 * lane, __fmul_rn(v[4 * j], wgt), __fmul_rn(v[4 * j + 1], wgt),
                           __fmul_rn(v[4 * j + 2], wgt), __fmul_rn(v[4 * j + 3], wgt));
            if (lane == 0) red_add_v2(dst + C, __fmul_rn(tail0, wgt), __fmul_rn(e, wgt));
        }
    }
}

// ---- forward warp, kernel C': the same when the caller wants only the cost (tc_stereo.py:139-140 is all the model
// reads of the warped features).  No warped-feature tile and no per-channel divisions: the CURRENT features are
// transposed through shared memory instead, each warp walks its pixels with the accumulator row straight from
// global memory (scaled by one reciprocal per pixel, so the range stays that of the normalised features), and the
// three sums of the cosine are warp-reduced in a fixed order.
template <int kGroups>
__global__ void __launch_bounds__(kWarpThreads)
warp_cost_kernel(const float* __restrict__ accum, const float* __restrict__ cur_fmap, float* __restrict__ out_disp,
                 float* __restrict__ out_mask, float* __restrict__ out_cost, const int* __restrict__ ctrl, int H, int W) {
    if (ctrl[kCtrlFallback] == 0) return;   // the gather kernel handled this frame
    constexpr int C = kGroups * 128;
    constexpr int CP = C + 4;
    constexpr int kPerWarp = C / 8;
    constexpr int kPx = kTileW / 8;
    extern __shared__ float tile[];            // [C][33]: current features, channel-major
    const int w0 = blockIdx.x * kTileW, h = blockIdx.y, b = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = w0 + lane;
    const bool in_w = w < W;
    const size_t plane = (size_t)H * W;
    const size_t base = ((size_t)b * C * H + h) * W + w;
    float f[kPerWarp];
#pragma unroll
    for (int k = 0; k < kPerWarp; ++k)
        f[k] = in_w ? ldg_stream_f1(cur_fmap + base + (size_t)(warp + 8 * k) * plane) : 0.0f;
    float2 tail[kPx];
    float4 a[kPx][kGroups];
#pragma unroll
    for (int i = 0; i < kPx; ++i) {
        const int wl = warp * kPx + i;
        const bool live = w0 + wl < W;
        const float* src = accum + (((size_t)b * H + h) * W + (live ? w0 + wl : 0)) * CP;
        tail[i] = live ? *reinterpret_cast<const float2*>(src + C) : make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < kGroups; ++j)
            a[i][j] = live ? ldg_stream_f4(reinterpret_cast<const float4*>(src + j * 128 + 4 * lane)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < kPerWarp; ++k) tile[(warp + 8 * k) * 33 + lane] = f[k];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPx; ++i) {
        const int wl = warp * kPx + i;
        const bool live = w0 + wl < W;
        const float nrm = fmaxf(tail[i].y, 1e-7f);              // clip(1e-7, None)   softsplat.py:268
        const float m = (tail[i].y != 0.0f) ? 1.0f : 0.0f;      // softsplat.py:258
        const float r = __fdiv_rn(1.0f, nrm);
        float dot = 0.0f, s1 = 0.0f, sw = 0.0f;
#pragma unroll
        for (int j = 0; j < kGroups; ++j) {
            const float av[4] = {a[i][j].x, a[i][j].y, a[i][j].z, a[i][j].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {                        // accumulator position 128j + 4 lane + q  <->  channel below
                const float fv = tile[(lane + 32 * (4 * j + q)) * 33 + wl];
                const float v = __fmul_rn(av[q], r);
                dot = fmaf(fv, v, dot);
                s1 = fmaf(fv, fv, s1);
                sw = fmaf(v, v, sw);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dot += __shfl_xor_sync(0xffffffffu, dot, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            sw += __shfl_xor_sync(0xffffffffu, sw, o);
        }
        if (lane == 0 && live) {
            const size_t idx = ((size_t)b * H + h) * W + w0 + wl;
            out_disp[idx] = __fdiv_rn(tail[i].x, nrm);
            out_mask[idx] = m;
            // sum_c normalize(f)_c * normalize(v)_c  (F.normalize eps 1e-12), times the splat mask (tc_stereo.py:139-140)
            const float den = __fmul_rn(fmaxf(sqrtf(s1), 1e-12f), fmaxf(sqrtf(sw), 1e-12f));
            out_cost[idx] = __fmul_rn(__fdiv_rn(dot, den), m);
        }
    }
}

// ---- forward warp, kernel C: normalise, mask, NCHW re-layout, matching cost -------------------------------------
template <int kGroups>
__global__ void __launch_bounds__(kWarpThreads)
warp_finalize_kernel(const float* __restrict__ accum, const float* __restrict__ cur_fmap,
                     float* __restrict__ out_disp, float* __restrict__ out_fmap, float* __restrict__ out_mask,
                     float* __restrict__ out_cost, const int* __restrict__ ctrl, int H, int W) {
    if (ctrl[kCtrlFallback] == 0) return;   // the gather kernel handled this frame
    constexpr int C = kGroups * 128;
    constexpr int CP = C + 4;
    extern __shared__ float tile[];            // [C][33] then red[8][32][3] then maskv[32]
    float* red = tile + C * 33;
    float* maskv = red + 8 * 32 * 3;
    const int w0 = blockIdx.x * kTileW, h = blockIdx.y, b = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kPerWarp = C / 8;            // channels warp + 8k of phase 2

    // phase 0: the current frame's features for the cost are independent of everything else: fetch them first
    const int w = w0 + lane;
    const bool in_w = w < W;
    const size_t plane = (size_t)H * W;
    const size_t base = ((size_t)b * C * H + h) * W + w;
    const bool want_cost = (out_cost != nullptr) && (cur_fmap != nullptr);
    float f[kPerWarp];
#pragma unroll
    for (int k = 0; k < kPerWarp; ++k)
        f[k] = (want_cost && in_w) ? ldg_stream_f1(cur_fmap + base + (size_t)(warp + 8 * k) * plane) : 0.0f;

    // phase 1: accumulator (channels-last, permuted) -> normalised values in tile[c][w]
    float2 tail[kTileW / 8];
    float4 a[kTileW / 8][kGroups];
#pragma unroll
    for (int i = 0; i < kTileW / 8; ++i) {
        const int wl = warp * (kTileW / 8) + i;
        const bool live = w0 + wl < W;
        const size_t idx = ((size_t)b * H + h) * W + (live ? w0 + wl : 0);
        const float* src = accum + idx * CP;
        tail[i] = live ? *reinterpret_cast<const float2*>(src + C) : make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < kGroups; ++j)
            a[i][j] = live ? *reinterpret_cast<const float4*>(src + j * 128 + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < kTileW / 8; ++i) {
        const int wl = warp * (kTileW / 8) + i;
        const bool live = w0 + wl < W;
        const float nrm = fmaxf(tail[i].y, 1e-7f);              // clip(1e-7, None)   softsplat.py:268
        const float m = (tail[i].y != 0.0f) ? 1.0f : 0.0f;      // softsplat.py:258
#pragma unroll
        for (int j = 0; j < kGroups; ++j) {
            tile[(lane + 32 * (4 * j + 0)) * 33 + wl] = __fdiv_rn(a[i][j].x, nrm);
            tile[(lane + 32 * (4 * j + 1)) * 33 + wl] = __fdiv_rn(a[i][j].y, nrm);
            tile[(lane + 32 * (4 * j + 2)) * 33 + wl] = __fdiv_rn(a[i][j].z, nrm);
            tile[(lane + 32 * (4 * j + 3)) * 33 + wl] = __fdiv_rn(a[i][j].w, nrm);
        }
        if (lane == 0) {
            maskv[wl] = live ? m : 0.0f;
            if (live) {
                const size_t idx = ((size_t)b * H + h) * W + w0 + wl;
                out_disp[idx] = __fdiv_rn(tail[i].x, nrm);
                out_mask[idx] = m;
            }
        }
    }
    __syncthreads();

    // phase 2: NCHW stores (one 128-byte line per channel and warp) + the per-pixel dot products of the cost
    float dot = 0.0f, s1 = 0.0f, sw = 0.0f;
#pragma unroll
    for (int k = 0; k < kPerWarp; ++k) {
        const int c = warp + 8 * k;
        const float v = tile[c * 33 + lane];
        if (in_w && out_fmap != nullptr) stg_stream_f1(out_fmap + base + (size_t)c * plane, v);
        dot = fmaf(f[k], v, dot);
        s1 = fmaf(f[k], f[k], s1);
        sw = fmaf(v, v, sw);
    }
    if (!want_cost) return;
    red[(warp * 32 + lane) * 3 + 0] = dot;
    red[(warp * 32 + lane) * 3 + 1] = s1;
    red[(warp * 32 + lane) * 3 + 2] = sw;
    __syncthreads();
    if (warp == 0 && in_w) {
        dot = s1 = sw = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            dot += red[(k * 32 + lane) * 3 + 0];
            s1 += red[(k * 32 + lane) * 3 + 1];
            sw += red[(k * 32 + lane) * 3 + 2];
        }
        // sum_c normalize(f)_c * normalize(v)_c  (F.normalize eps 1e-12), times the splat mask (tc_stereo.py:139-140)
        const float den = __fmul_rn(fmaxf(sqrtf(s1), 1e-12f), fmaxf(sqrtf(sw), 1e-12f));
        out_cost[((size_t)b * H + h) * W + w] = __fmul_rn(__fdiv_rn(dot, den), maskv[lane]);
    }
}

// ---- forward warp, kernel A2: soft-splat weight of every source pixel -----------------------------------------
// wgt = valid * exp(clamp(disp' - mean, -50, 50)), 0 for pixels the splat skips (invalid, non-finite target).
// Also decides whether the gather kernel can handle the frame (flow within its search radius).
__global__ void __launch_bounds__(256)
warp_weight_kernel(const float* __restrict__ disp1, const float* __restrict__ tx, const float* __restrict__ ty,
                   float* __restrict__ valid_to_wgt, const double* __restrict__ sums, int* __restrict__ ctrl,
                   int B, int HW, int per_sample_mean) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        const float fx = __int_as_float(ctrl[kCtrlMaxFlowX]), fy = __int_as_float(ctrl[kCtrlMaxFlowY]);
        if (!(fx <= (float)(kGatherMaxR - 2)) || !(fy <= (float)(kGatherMaxR - 2))) atomicExch(ctrl + kCtrlFallback, 1);
    }
    if (i >= HW) return;
    // softsplat metric: disparity minus its mean (geo_utils.py:193); batch-global unless asked otherwise
    double s = 0.0, n;
    if (per_sample_mean) {
        s = sums[b];
        n = (double)HW;
    } else {
        for (int k = 0; k < B; ++k) s += sums[k];
        n = (double)B * HW;
    }
    const float mean = (float)(s / n);
    const size_t o = (size_t)b * HW + i;
    float w = 0.0f;
    if (valid_to_wgt[o] != 0.0f && isfinite(tx[o]) && isfinite(ty[o]))      // softsplat.py:236,300-301
        w = expf(fminf(fmaxf(__fsub_rn(disp1[o], mean), -50.0f), 50.0f));
    valid_to_wgt[o] = w;
}

// ---- forward warp, kernel B' (TCS_WARP_DETERMINISTIC): the splat as a gather ---------------------------------------------------
// A forward splat scatters every source pixel into the 4 integer neighbours of its target.  Frame-to-frame flow is
// a few pixels, so the same sum can be collected from the target's side: one warp owns 32 consecutive target pixels
// of a row, scans the (2Rx+1) x (2Ry+1) source offsets the frame's largest flow allows, keeps the offsets (dx, dy)
// from which at least one of its targets receives a non-zero weight (a handful when the flow is smooth), and then
// walks the channels: out[c] = sum_e w_e * fmap[c][y+dy_e][x+dx_e] / clip(norm).  Everything stays NCHW (lanes
// are consecutive x: coalesced loads and stores, no transposes), there is no accumulator in global memory, no
// memset, no atomics, the sum order is fixed (the result is deterministic, unlike the scatter), and the
// normalisation, the mask and the matching cost are produced in the same pass.  Frames whose flow exceeds the
// search radius, or segments that need more than kGatherMaxEntries offsets, raise the fallback flag and the scatter
// kernels redo the frame.  It is latency-bound (693-769 us against 452 us for the scatter at 540p x 8), so it is
// the opt-in deterministic mode, not the default.
constexpr int kGatherWarps = 4;
constexpr int kGatherRegEntries = 12;   // offsets whose weights are held in registers (the common, smooth-flow case)
constexpr int kGatherWinFloats = 4 * 1024;   // staged source window: 4 planes (weight, target x, target y, disp')

// sum over the entries of w_e * fmap[c][.. + off_e] for kCh consecutive channels, the first kN entries held in
// registers.  The loads are UNCONDITIONAL (an entry that does not feed this lane has offset 0 = the lane's own
// pixel, always in range, and its value is replaced by 0), so all kN * kCh of them are independent and issued
// back to back: the kernel is latency-bound and this is what buys memory-level parallelism.
template <int kN, int kCh>
__device__ __forceinline__ void gather_block(const float* __restrict__ src, size_t plane, const float (&w)[kGatherRegEntries],
                                             const int (&off)[kGatherRegEntries], float (&acc)[kCh]) {
    float v[kN][kCh];
#pragma unroll
    for (int e = 0; e < kN; ++e)
#pragma unroll
        for (int k = 0; k < kCh; ++k) v[e][k] = ldg_ordered_f1(src + (size_t)k * plane + off[e]);
#pragma unroll
    for (int e = 0; e < kN; ++e)
#pragma unroll
        for (int k = 0; k < kCh; ++k) acc[k] = fmaf((w[e] != 0.0f) ? v[e][k] : 0.0f, w[e], acc[k]);
}

__global__ void __launch_bounds__(kGatherWarps * 32)
warp_gather_kernel(const float* __restrict__ fmap, const float* __restrict__ disp1, const float* __restrict__ tx,
                   const float* __restrict__ ty, const float* __restrict__ wgt, const float* __restrict__ cur_fmap,
                   float* __restrict__ out_disp, float* __restrict__ out_fmap, float* __restrict__ out_mask,
                   float* __restrict__ out_cost, int* __restrict__ ctrl, int C, int H, int W) {
    __shared__ float s_w[kGatherWarps][kGatherMaxEntries][32];
    __shared__ int s_off[kGatherWarps][kGatherMaxEntries];
    __shared__ float s_win[kGatherWinFloats];
    if (ctrl[kCtrlFallback] != 0) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int Y0 = blockIdx.y * kGatherWarps;
    const int Y = Y0 + wib;
    const int b = blockIdx.z;
    const int X0 = blockIdx.x * 32;
    const int X = X0 + lane;
    const bool in_w = X < W;
    const int Rx = min((int)ceilf(__int_as_float(ctrl[kCtrlMaxFlowX])) + 1, kGatherMaxR);
    const int Ry = min((int)ceilf(__int_as_float(ctrl[kCtrlMaxFlowY])) + 1, kGatherMaxR);
    const size_t pbase = (size_t)b * H * W;
    float (*sw)[32] = s_w[wib];
    int* soff = s_off[wib];

    // ---- stage the CTA's source window (weight, target x, target y, disp') in shared memory when it fits
    const int win_w = 32 + 2 * Rx, win_h = kGatherWarps + 2 * Ry;
    const int win_n = win_w * win_h;
    const bool staged = 4 * win_n <= kGatherWinFloats;
    if (staged) {
        for (int i = threadIdx.x; i < win_n; i += kGatherWarps * 32) {
            const int wy = i / win_w, wx = i - wy * win_w;
            const int y = Y0 - Ry + wy, x = X0 - Rx + wx;
            float e = 0.0f, fx_ = 0.0f, fy_ = 0.0f, d1 = 0.0f;
            if (y >= 0 && y < H && x >= 0 && x < W) {
                const size_t so = pbase + (size_t)y * W + x;
                e = __ldg(wgt + so);
                fx_ = __ldg(tx + so);
                fy_ = __ldg(ty + so);
                d1 = __ldg(disp1 + so);
            }
            s_win[i] = e;
            s_win[win_n + i] = fx_;
            s_win[2 * win_n + i] = fy_;
            s_win[3 * win_n + i] = d1;
        }
    }
    __syncthreads();
    if (Y >= H) return;                          // warp-uniform

    // ---- phase A: which source offsets feed this segment, and with what weight per target
    int n_ent = 0;
    float norm = 0.0f, dsum = 0.0f;
    const float Xf = (float)X, Yf = (float)Y;
    for (int dy = -Ry; dy <= Ry; ++dy) {
        const int y = Y + dy;
        if (y < 0 || y >= H) continue;           // warp-uniform
        for (int dx = -Rx; dx <= Rx; ++dx) {
            const int x = X + dx;
            float w = 0.0f, d1 = 0.0f;
            if (in_w && x >= 0 && x < W) {
                float e, fx_, fy_;
                if (staged) {
                    const int i = (wib + Ry + dy) * win_w + (lane + Rx + dx);
                    e = s_win[i]; fx_ = s_win[win_n + i]; fy_ = s_win[2 * win_n + i]; d1 = s_win[3 * win_n + i];
                } else {
                    const size_t so = pbase + (size_t)y * W + x;
                    e = __ldg(wgt + so); fx_ = __ldg(tx + so); fy_ = __ldg(ty + so); d1 = __ldg(disp1 + so);
                }
                if (e != 0.0f) {
                    const float nwx = floorf(fx_), nwy = floorf(fy_);
                    // bilinear weights of softsplat.py:314-317, seen from the target: x part * y part
                    float wx = 0.0f, wy = 0.0f;
                    if (Xf == nwx) wx = __fsub_rn(__fadd_rn(nwx, 1.0f), fx_);
                    else if (Xf == __fadd_rn(nwx, 1.0f)) wx = __fsub_rn(fx_, nwx);
                    if (Yf == nwy) wy = __fsub_rn(__fadd_rn(nwy, 1.0f), fy_);
                    else if (Yf == __fadd_rn(nwy, 1.0f)) wy = __fsub_rn(fy_, nwy);
                    w = __fmul_rn(e, __fmul_rn(wx, wy));
                }
            }
            if (__any_sync(0xffffffffu, w != 0.0f)) {
                if (n_ent < kGatherMaxEntries) {
                    sw[n_ent][lane] = w;
                    if (lane == 0) soff[n_ent] = dy * W + dx;
                }
                ++n_ent;
                norm = __fadd_rn(norm, w);
                dsum = fmaf(d1, w, dsum);
            }
        }
    }
    if (n_ent > kGatherMaxEntries) {             // flow too irregular for the gather: let the scatter path redo the frame
        if (lane == 0) atomicExch(ctrl + kCtrlFallback, 1);
        return;
    }
    __syncwarp();
    const float nrm = fmaxf(norm, 1e-7f);                       // clip(1e-7, None)   softsplat.py:268
    const float m = (norm != 0.0f) ? 1.0f : 0.0f;               // softsplat.py:258
    const size_t po = pbase + (size_t)Y * W + X;
    if (in_w) {
        out_disp[po] = __fdiv_rn(dsum, nrm);
        out_mask[po] = m;
    }

    // ---- phase B: channels
    const size_t plane = (size_t)H * W;
    const float* src = fmap + (size_t)b * C * plane + (size_t)Y * W + X;
    const size_t obase = (size_t)b * C * plane + (size_t)Y * W + X;
    const bool want_cost = (out_cost != nullptr) && (cur_fmap != nullptr);
    float dot = 0.0f, s1 = 0.0f, s2 = 0.0f;
    float wreg[kGatherRegEntries];
    int offreg[kGatherRegEntries];
#pragma unroll
    for (int e = 0; e < kGatherRegEntries; ++e) {
        wreg[e] = (e < n_ent) ? sw[e][lane] : 0.0f;
        offreg[e] = (e < n_ent && wreg[e] != 0.0f) ? soff[e] : 0;   // lanes an entry does not feed read their own pixel
    }
    constexpr int kCh = 4;
    const bool few = n_ent <= kGatherRegEntries / 2;              // warp-uniform
    for (int c = 0; c < C; c += kCh) {
        float acc[kCh];
#pragma unroll
        for (int k = 0; k < kCh; ++k) acc[k] = 0.0f;
        const float* sc = src + (size_t)c * plane;
        if (few) gather_block<kGatherRegEntries / 2, kCh>(sc, plane, wreg, offreg, acc);
        else gather_block<kGatherRegEntries, kCh>(sc, plane, wreg, offreg, acc);
        for (int e = kGatherRegEntries; e < n_ent; ++e) {       // the rare tail beyond the register-held entries
            const float w = sw[e][lane];
            const int off = soff[e];
            if (w != 0.0f) {
#pragma unroll
                for (int k = 0; k < kCh; ++k) acc[k] = fmaf(__ldg(sc + (size_t)k * plane + off), w, acc[k]);
            }
        }
        if (in_w) {
#pragma unroll
            for (int k = 0; k < kCh; ++k) {
                const float v = __fdiv_rn(acc[k], nrm);
                if (out_fmap != nullptr) stg_stream_f1(out_fmap + obase + (size_t)(c + k) * plane, v);
                if (want_cost) {
                    const float f = ldg_stream_f1(cur_fmap + obase + (size_t)(c + k) * plane);
                    dot = fmaf(f, v, dot);
                    s1 = fmaf(f, f, s1);
                    s2 = fmaf(v, v, s2);
                }
            }
        }
    }
    if (want_cost && in_w) {
        // sum_c normalize(f)_c * normalize(v)_c  (F.normalize eps 1e-12), times the splat mask (tc_stereo.py:139-140)
        const float den = __fmul_rn(fmaxf(sqrtf(s1), 1e-12f), fmaxf(sqrtf(s2), 1e-12f));
        out_cost[po] = __fmul_rn(__fdiv_rn(dot, den), m);
    }
}

// ---- fallback: zero the scatter accumulator (only when the gather kernel gave up) ------------------------------
__global__ void __launch_bounds__(256)
warp_zero_kernel(float4* __restrict__ accum, size_t n4, const int* __restrict__ ctrl) {
    if (ctrl[kCtrlFallback] == 0) return;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) accum[i] = z;
}

// ---- get_backward_grid ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
backward_grid_kernel(const float* __restrict__ disp, const float* __restrict__ rel_T, const float* __restrict__ K,
                     const float* __restrict__ K_inv, const float* __restrict__ baseline, float* __restrict__ grid,
                     int H, int W) {
    const int b = blockIdx.y;
    const int HW = H * W;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const Cam c = load_cam(rel_T, K, K_inv, baseline, b);
    const int y = i / W, x = i - y * W;
    const float d = fmaxf(__ldg(disp + (size_t)b * HW + i), 0.01f);            // geo_utils.py:217
    const float depth = __fdiv_rn(c.bf, fmaxf(d, 0.001f));
    float P[3];
    project(c, depth, (float)x, (float)y, P);
    float u, v;
    reproject(c, P, u, v);
    if (!(P[2] > 0.0f)) { u = -1.0f; v = -1.0f; }                              // geo_utils.py:229,233
    grid[((size_t)b * 2 + 0) * HW + i] = u;
    grid[((size_t)b * 2 + 1) * HW + i] = v;
}

// ---- bilinear_sampler --------------------------------------------------------------------------------------------------
// grid_sample(bilinear, zeros, align_corners=True) on pixel coordinates, including the wrapper's normalise and
// ATen's un-normalise round trip.  One thread per (output pixel, group of 8 channels): the four weights are
// computed once, then the 32 corner loads of the group are issued back to back (memory-level parallelism),
// lanes being consecutive output pixels so that loads and stores of one channel plane coalesce.
constexpr int kSampleCh = 8;

template <bool kFull>   // kFull: C is a multiple of kSampleCh, no per-channel guards (keeps the loads batched)
__global__ void __launch_bounds__(256)
bilinear_sample_kernel(const float* __restrict__ img, const float* __restrict__ grid_xy, float* __restrict__ out,
                       int C, int Hi, int Wi, int Ho, int Wo) {
    const int b = blockIdx.z;
    const int HWo = Ho * Wo;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HWo) return;
    const float gx = __ldg(grid_xy + ((size_t)b * 2 + 0) * HWo + i);
    const float gy = __ldg(grid_xy + ((size_t)b * 2 + 1) * HWo + i);
    const float wm1 = (float)(Wi - 1), hm1 = (float)(Hi - 1);
    const float xn = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, gx), wm1), 1.0f);                  // utils.py:86
    const float yn = (Hi > 1) ? __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, gy), hm1), 1.0f) : gy;  // utils.py:87-88
    const float ix = __fmul_rn(__fmul_rn(__fadd_rn(xn, 1.0f), 0.5f), wm1);
    const float iy = __fmul_rn(__fmul_rn(__fadd_rn(yn, 1.0f), 0.5f), hm1);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float x1f = __fadd_rn(x0f, 1.0f), y1f = __fadd_rn(y0f, 1.0f);
    // NaN / huge coordinates fall out of range on every corner (float compares, then the int cast is safe);
    // an out-of-range corner gets weight 0 and a clamped (valid) address, so the loads need no predicates
    const bool xin0 = (x0f >= 0.0f) && (x0f <= wm1), xin1 = (x1f >= 0.0f) && (x1f <= wm1);
    const bool yin0 = (y0f >= 0.0f) && (y0f <= hm1), yin1 = (y1f >= 0.0f) && (y1f <= hm1);
    const int x0 = xin0 ? (int)x0f : 0, x1 = xin1 ? (int)x1f : 0;
    const int y0 = yin0 ? (int)y0f : 0, y1 = yin1 ? (int)y1f : 0;
    const bool nw = xin0 && yin0, ne = xin1 && yin0, sw = xin0 && yin1, se = xin1 && yin1;
    const float w_nw = __fmul_rn(__fsub_rn(x1f, ix), __fsub_rn(y1f, iy));
    const float w_ne = __fmul_rn(__fsub_rn(ix, x0f), __fsub_rn(y1f, iy));
    const float w_sw = __fmul_rn(__fsub_rn(x1f, ix), __fsub_rn(iy, y0f));
    const float w_se = __fmul_rn(__fsub_rn(ix, x0f), __fsub_rn(iy, y0f));
    const int o_nw = y0 * Wi + x0, o_ne = y0 * Wi + x1, o_sw = y1 * Wi + x0, o_se = y1 * Wi + x1;
    const size_t plane_i = (size_t)Hi * Wi;
    const int c_begin = blockIdx.y * kSampleCh;
    const float* src = img + ((size_t)b * C + c_begin) * plane_i;
    float* dst = out + ((size_t)b * C + c_begin) * HWo + i;
    float v[kSampleCh][4];
#pragma unroll
    for (int k = 0; k < kSampleCh; ++k) {
        const bool live = kFull || (c_begin + k < C);
        const float* s = src + (live ? (size_t)k * plane_i : 0);
        v[k][0] = ldg_ordered_f1(s + o_nw);
        v[k][1] = ldg_ordered_f1(s + o_ne);
        v[k][2] = ldg_ordered_f1(s + o_sw);
        v[k][3] = ldg_ordered_f1(s + o_se);
    }
#pragma unroll
    for (int k = 0; k < kSampleCh; ++k) {
        if (!kFull && c_begin + k >= C) break;
        // corners accumulate in the order nw, ne, sw, se; a corner outside the image contributes nothing
        float r = 0.0f;
        if (nw) r = __fmul_rn(v[k][0], w_nw);
        if (ne) r = __fadd_rn(r, __fmul_rn(v[k][1], w_ne));
        if (sw) r = __fadd_rn(r, __fmul_rn(v[k][2], w_sw));
        if (se) r = __fadd_rn(r, __fmul_rn(v[k][3], w_se));
        stg_stream_f1(dst + (size_t)k * HWo, r);
    }
}

// ---- 0.5 * interpolate(grid, 0.5, bilinear, align_corners=True) ------------------------------------------------------------
__global__ void __launch_bounds__(256)
grid_halve_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int Ho, int Wo, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int xo = (int)(i % Wo);
    const int yo = (int)((i / Wo) % Ho);
    const long long bc = i / ((long long)Wo * Ho);
    // ATen area_pixel_compute_scale with align_corners: (in - 1) / (out - 1), 0 when out == 1
    const float sh = (Ho > 1) ? __fdiv_rn((float)(H - 1), (float)(Ho - 1)) : 0.0f;
    const float sw = (Wo > 1) ? __fdiv_rn((float)(W - 1), (float)(Wo - 1)) : 0.0f;
    const float ys = __fmul_rn(sh, (float)yo), xs = __fmul_rn(sw, (float)xo);
    const int y0 = (int)ys, x0 = (int)xs;
    const int yp = (y0 < H - 1) ? 1 : 0, xp = (x0 < W - 1) ? 1 : 0;
    const float ly1 = __fsub_rn(ys, (float)y0), ly0 = __fsub_rn(1.0f, ly1);
    const float lx1 = __fsub_rn(xs, (float)x0), lx0 = __fsub_rn(1.0f, lx1);
    const float* p = in + bc * (long long)H * W + (long long)y0 * W + x0;
    const float v00 = __ldg(p), v01 = __ldg(p + xp), v10 = __ldg(p + (long long)yp * W), v11 = __ldg(p + (long long)yp * W + xp);
    const float top = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
    const float bot = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
    out[i] = __fmul_rn(0.5f, __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot)));
}

static size_t align256(size_t n) { return (n + 255) & ~static_cast<size_t>(255); }

struct WarpScratch {
    size_t sums, accum, disp1, tx, ty, valid, total;
};
static WarpScratch warp_scratch_layout(int B, int C, int H, int W) {
    WarpScratch s;
    const size_t npix = (size_t)B * H * W;
    s.sums = 0;
    s.accum = align256((size_t)B * sizeof(double) + 64);   // per-sample sums, then the control words
    s.disp1 = s.accum + align256(npix * (C + 4) * sizeof(float));
    s.tx = s.disp1 + align256(npix * sizeof(float));
    s.ty = s.tx + align256(npix * sizeof(float));
    s.valid = s.ty + align256(npix * sizeof(float));
    s.total = s.valid + align256(npix * sizeof(float));
    return s;
}

}  // namespace tcs

extern "C" long long tcs_warp_scratch_bytes(int B, int C, int H, int W) {
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0) return 0;
    return (long long)tcs::warp_scratch_layout(B, C, H, W).total;
}

extern "C" int tcs_warp_forward(const float* disp, const float* fmap, const float* rel_T, const float* K,
                                const float* K_inv, const float* baseline, const float* cur_fmap,
                                float* out_disp, float* out_fmap, float* out_mask, float* out_cost,
                                void* scratch, int B, int C, int H, int W, int flags, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(disp && fmap && rel_T && K && K_inv && baseline && out_disp && out_mask && scratch,
                TCS_E_BADARG, "tcs_warp_forward: null pointer");
    TCS_REQUIRE(out_fmap != nullptr || out_cost != nullptr, TCS_E_BADARG, "tcs_warp_forward: neither out_fmap nor out_cost requested");
    TCS_REQUIRE(out_cost == nullptr || cur_fmap != nullptr, TCS_E_BADARG, "tcs_warp_forward: out_cost needs cur_fmap");
    TCS_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, TCS_E_BADARG, "tcs_warp_forward: non-positive size");
    TCS_REQUIRE(C == 128 || C == 256 || C == 384 || C == 512, TCS_E_SHAPE, "tcs_warp_forward: C=%d must be 128, 256, 384 or 512", C);
    TCS_REQUIRE(H <= 65535 && B <= 65535, TCS_E_SHAPE, "tcs_warp_forward: H and B must be <= 65535");
    TCS_REQUIRE(aligned16(fmap) && aligned16(scratch) && aligned16(cur_fmap), TCS_E_ALIGN,
                "tcs_warp_forward: fmap, cur_fmap and scratch must be 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const WarpScratch L = warp_scratch_layout(B, C, H, W);
    char* base = static_cast<char*>(scratch);
    double* sums = reinterpret_cast<double*>(base + L.sums);
    float* accum = reinterpret_cast<float*>(base + L.accum);
    float* disp1 = reinterpret_cast<float*>(base + L.disp1);
    float* tx = reinterpret_cast<float*>(base + L.tx);
    float* ty = reinterpret_cast<float*>(base + L.ty);
    float* valid = reinterpret_cast<float*>(base + L.valid);

    int* ctrl = reinterpret_cast<int*>(base + L.sums + (size_t)B * sizeof(double));
    const int per_sample_mean = (flags & TCS_WARP_PER_SAMPLE_MEAN) ? 1 : 0;
    const bool gather = (flags & TCS_WARP_DETERMINISTIC) != 0;
    // scatter (default): zero sums + control words + accumulator and preset the "scatter" flag;
    // gather: zero only the header; the accumulator is zeroed by a kernel, and only if the gather gives up
    TCS_CHECK_CUDA(cudaMemsetAsync(base, 0, gather ? L.accum : L.disp1, s));
    if (!gather) TCS_CHECK_CUDA(cudaMemsetAsync(ctrl + kCtrlFallback, 1, 1, s));
    {
        dim3 grid(ceil_div(H * W, 256), B);
        warp_geometry_kernel<<<grid, 256, 0, s>>>(disp, rel_T, K, K_inv, baseline, disp1, tx, ty, valid, sums, ctrl, H, W);
        TCS_CHECK_LAUNCH("tcs_warp_forward(geometry)");
        warp_weight_kernel<<<grid, 256, 0, s>>>(disp1, tx, ty, valid, sums, ctrl, B, H * W, per_sample_mean);
        TCS_CHECK_LAUNCH("tcs_warp_forward(weights)");
        if (gather) {
            dim3 ggrid(ceil_div(W, 32), ceil_div(H, kGatherWarps), B);
            warp_gather_kernel<<<ggrid, kGatherWarps * 32, 0, s>>>(fmap, disp1, tx, ty, valid, cur_fmap, out_disp, out_fmap, out_mask,
                                                                   out_cost, ctrl, C, H, W);
            TCS_CHECK_LAUNCH("tcs_warp_forward(gather)");
            const size_t n4 = (size_t)B * H * W * (C + 4) / 4;
            warp_zero_kernel<<<num_sms() * 4, 256, 0, s>>>(reinterpret_cast<float4*>(accum), n4, ctrl);
            TCS_CHECK_LAUNCH("tcs_warp_forward(zero)");
        }
    }
    const dim3 grid(ceil_div(W, kTileW), H, B);
    const size_t smem_splat = (size_t)kTileW * (C + 1) * sizeof(float);
    const size_t smem_fin = ((size_t)C * 33 + 8 * 32 * 3 + 32) * sizeof(float);
    const size_t smem_cost = (size_t)C * 33 * sizeof(float);
#define TCS_WARP_CASE(G)                                                                                              \
    case G: {                                                                                                         \
        static bool attr_done = false;                                                                                \
        if (!attr_done) {                                                                                             \
            TCS_CHECK_CUDA(cudaFuncSetAttribute(warp_splat_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_splat)); \
            TCS_CHECK_CUDA(cudaFuncSetAttribute(warp_finalize_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fin)); \
            TCS_CHECK_CUDA(cudaFuncSetAttribute(warp_cost_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cost)); \
            { const int cv = carveout_percent("TCS_CARVE_SPLAT", -1); if (cv >= 0) TCS_CHECK_CUDA(cudaFuncSetAttribute(warp_splat_kernel<G>, cudaFuncAttributePreferredSharedMemoryCarveout, cv)); } \
            { const int cv = carveout_percent("TCS_CARVE_FINALIZE", -1); if (cv >= 0) TCS_CHECK_CUDA(cudaFuncSetAttribute(warp_finalize_kernel<G>, cudaFuncAttributePreferredSharedMemoryCarveout, cv)); } \
            attr_done = true;                                                                                         \
        }                                                                                                             \
        warp_splat_kernel<G><<<grid, kWarpThreads, smem_splat, s>>>(fmap, disp1, tx, ty, valid, ctrl, accum, B, H, W); \
        TCS_CHECK_LAUNCH("tcs_warp_forward(splat)");                                                                  \
        if (out_fmap == nullptr && out_cost != nullptr && cur_fmap != nullptr)                                        \
            warp_cost_kernel<G><<<grid, kWarpThreads, smem_cost, s>>>(accum, cur_fmap, out_disp, out_mask, out_cost, ctrl, H, W); \
        else                                                                                                          \
            warp_finalize_kernel<G><<<grid, kWarpThreads, smem_fin, s>>>(accum, cur_fmap, out_disp, out_fmap, out_mask, out_cost, ctrl, H, W); \
        TCS_CHECK_LAUNCH("tcs_warp_forward(finalize)");                                                               \
    } break;
    switch (C / 128) {
        TCS_WARP_CASE(1)
        TCS_WARP_CASE(2)
        TCS_WARP_CASE(3)
        TCS_WARP_CASE(4)
    }
#undef TCS_WARP_CASE
    return 0;
}

extern "C" int tcs_backward_grid(const float* disp, const float* rel_T, const float* K, const float* K_inv,
                                 const float* baseline, float* grid, int B, int H, int W, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(disp && rel_T && K && K_inv && baseline && grid, TCS_E_BADARG, "tcs_backward_grid: null pointer");
    TCS_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, TCS_E_BADARG, "tcs_backward_grid: bad sizes");
    dim3 g(ceil_div(H * W, 256), B);
    backward_grid_kernel<<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(disp, rel_T, K, K_inv, baseline, grid, H, W);
    TCS_CHECK_LAUNCH("tcs_backward_grid");
    return 0;
}

extern "C" int tcs_bilinear_sample(const float* img, const float* grid_xy, float* out,
                                   int B, int C, int Hi, int Wi, int Ho, int Wo, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(img && grid_xy && out, TCS_E_BADARG, "tcs_bilinear_sample: null pointer");
    TCS_REQUIRE(B > 0 && B <= 65535 && C > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0, TCS_E_BADARG, "tcs_bilinear_sample: bad sizes");
    const int groups = ceil_div(C, kSampleCh);
    TCS_REQUIRE(groups <= 65535, TCS_E_SHAPE, "tcs_bilinear_sample: C too large");
    TCS_REQUIRE((long long)Hi * Wi < 0x7fffffffLL, TCS_E_SHAPE, "tcs_bilinear_sample: image plane too large");
    dim3 g(ceil_div(Ho * Wo, 256), groups, B);
    if (C % kSampleCh == 0)
        bilinear_sample_kernel<true><<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(img, grid_xy, out, C, Hi, Wi, Ho, Wo);
    else
        bilinear_sample_kernel<false><<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(img, grid_xy, out, C, Hi, Wi, Ho, Wo);
    TCS_CHECK_LAUNCH("tcs_bilinear_sample");
    return 0;
}

extern "C" int tcs_grid_halve(const float* in, float* out, int B, int H, int W, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(in && out, TCS_E_BADARG, "tcs_grid_halve: null pointer");
    TCS_REQUIRE(B > 0 && H >= 2 && W >= 2, TCS_E_SHAPE, "tcs_grid_halve: need H, W >= 2");
    const int Ho = H / 2, Wo = W / 2;
    const long long total = (long long)B * 2 * Ho * Wo;
    grid_halve_kernel<<<(unsigned)ceil_div_ll(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, H, W, Ho, Wo, total);
    TCS_CHECK_LAUNCH("tcs_grid_halve");
    return 0;
}
