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
// A second formulation (TCS_WARP_DETERMINISTIC) collects the same sums from the target's side through exact,
// sorted per-target contributor lists and a transposed copy of the source features: no accumulator, no
// floating-point atomics, a fixed summation order.  The transposed copy of a frame's features can be produced
// by that frame's own cost kernel (cur_t_out) and consumed by the next frame's call (fmap_t), which is how a
// sequence runs (see the "list formulation" section below).
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
                     float* __restrict__ valid, double* __restrict__ sums, int H, int W) {
    const int b = blockIdx.y;
    const int HW = H * W;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double local = 0.0;
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
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ double wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += wsum[w];
        atomicAdd(sums + b, s);
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
                  const float* __restrict__ ty, const float* __restrict__ wgt,
                  float* __restrict__ accum, int B, int H, int W) {
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

// ---- list formulation, front end of the two normalise kernels ---------------------------------------------------
// Contributor lists (built by the kernels further down): target t owns entries[start[t] .. start[t+1]), each
// (source pixel inside the sample, e * bilinear weight), sorted by source pixel.  src_t is the previous frame's
// features transposed to pixel-major rows in the accumulator's channel permutation, so one entry is one coalesced
// 4*C-byte read for a warp.
struct ListArgs {
    const float* src_t;     // [B*H*W][C]
    const int* start;       // [B*H*W + 1]
    const int2* entries;
    const float* disp1;     // [B*H*W] reprojected disparity of the source pixels
};

// Sum of the contributions to one target (the whole warp works on it; n is warp-uniform).  tail = (sum w * disp', sum w).
// (Staging the block's offsets and entries in shared memory first was measured: 482 us instead of 355 -- the extra
// barrier behind the current-feature loads and 20 more registers cost more than the index round trips.)
// One round: kE consecutive entries of the list, all their rows in flight before the first is used.
template <int kGroups, int kE>
__device__ __forceinline__ void gather_round(const ListArgs& la, size_t sample_px0, int s0, int n, int j0, int lane,
                                             float4 (&a)[kGroups], float& norm, float& dsum) {
    constexpr int C = kGroups * 128;
    int2 en[kE];
    float4 v[kE][kGroups];
    float d1[kE];
#pragma unroll
    for (int u = 0; u < kE; ++u) {
        const bool live = j0 + u < n;
        en[u] = live ? __ldg(la.entries + s0 + j0 + u) : make_int2(0, 0);   // weight 0: contributes nothing
        const float* row = la.src_t + (sample_px0 + en[u].x) * C + 4 * lane;
#pragma unroll
        for (int j = 0; j < kGroups; ++j)
            v[u][j] = live ? __ldg(reinterpret_cast<const float4*>(row + j * 128)) : make_float4(0.f, 0.f, 0.f, 0.f);
        d1[u] = live ? __ldg(la.disp1 + sample_px0 + en[u].x) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < kE; ++u) {
        const float w = __int_as_float(en[u].y);
#pragma unroll
        for (int j = 0; j < kGroups; ++j) {
            a[j].x = fmaf(v[u][j].x, w, a[j].x);
            a[j].y = fmaf(v[u][j].y, w, a[j].y);
            a[j].z = fmaf(v[u][j].z, w, a[j].z);
            a[j].w = fmaf(v[u][j].w, w, a[j].w);
        }
        norm = __fadd_rn(norm, w);
        dsum = fmaf(d1[u], w, dsum);
    }
}

template <int kGroups>
__device__ __forceinline__ void gather_target(const ListArgs& la, size_t sample_px0, int s0, int n, int lane,
                                              float4 (&a)[kGroups], float2& tail) {
#pragma unroll
    for (int j = 0; j < kGroups; ++j) a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    float norm = 0.0f, dsum = 0.0f;
    // a list holds 4 entries on average: one round of four (4 * kGroups 16-byte loads in flight per lane), then
    // rounds of two so that a fifth or sixth entry does not pay for four predicated slots
    if (n > 0) gather_round<kGroups, 4>(la, sample_px0, s0, n, 0, lane, a, norm, dsum);
    for (int j0 = 4; j0 < n; j0 += 2) gather_round<kGroups, 2>(la, sample_px0, s0, n, j0, lane, a, norm, dsum);
    tail = make_float2(dsum, norm);
}

// NCHW features -> pixel-major rows [B*H*W][C] in the accumulator's channel permutation (position 128 j + 4 lane + q
// holds channel lane + 32 (4 j + q)): coalesced both ways through the transposed shared-memory tile.
template <int kGroups>
__global__ void __launch_bounds__(kWarpThreads)
warp_transpose_kernel(const float* __restrict__ fmap, float* __restrict__ dst, int H, int W) {
    constexpr int C = kGroups * 128;
    constexpr int pitch = C + 1;
    extern __shared__ float tile[];  // [32][C + 1]
    const int w0 = blockIdx.x * kTileW, h = blockIdx.y, b = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    load_tile_transposed(fmap, tile, pitch, b, h, w0, C, H, W);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kTileW / 8; ++i) {
        const int wl = warp * (kTileW / 8) + i;
        if (w0 + wl >= W) break;
        const float* row = tile + wl * pitch;
        float* out = dst + (((size_t)b * H + h) * W + w0 + wl) * C + 4 * lane;
#pragma unroll
        for (int j = 0; j < kGroups; ++j)
            *reinterpret_cast<float4*>(out + j * 128) = make_float4(row[lane + 32 * (4 * j + 0)], row[lane + 32 * (4 * j + 1)],
                                                                    row[lane + 32 * (4 * j + 2)], row[lane + 32 * (4 * j + 3)]);
    }
}

// ---- forward warp, kernel C': the same when the caller wants only the cost (tc_stereo.py:139-140 is all the model
// reads of the warped features).  No warped-feature tile and no per-channel divisions: the CURRENT features are
// transposed through shared memory instead, each warp walks its pixels with the accumulator row straight from
// global memory -- or, kLists, summed from the target's contributor list -- scaled by one reciprocal per pixel (so
// the range stays that of the normalised features), and the three sums of the cosine are warp-reduced in a fixed
// order.  cur_t_out (nullable) receives the current features as pixel-major rows: next frame's source.
template <int kGroups>
__device__ __forceinline__ void cost_pixel(const float4 (&a)[kGroups], float2 tail, const float* tile, int wl, bool live, int lane,
                                           size_t tpix, float* __restrict__ out_disp, float* __restrict__ out_mask,
                                           float* __restrict__ out_cost, float* __restrict__ cur_t_out) {
    constexpr int C = kGroups * 128;
    const float nrm = fmaxf(tail.y, 1e-7f);                     // clip(1e-7, None)   softsplat.py:268
    const float m = (tail.y != 0.0f) ? 1.0f : 0.0f;             // softsplat.py:258
    const float r = __fdiv_rn(1.0f, nrm);
    float dot = 0.0f, s1 = 0.0f, sw = 0.0f;
#pragma unroll
    for (int j = 0; j < kGroups; ++j) {
        const float av[4] = {a[j].x, a[j].y, a[j].z, a[j].w};
        float fq[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {                            // accumulator position 128j + 4 lane + q  <->  channel below
            const float fv = tile[(lane + 32 * (4 * j + q)) * 33 + wl];
            const float v = __fmul_rn(av[q], r);
            fq[q] = fv;
            dot = fmaf(fv, v, dot);
            s1 = fmaf(fv, fv, s1);
            sw = fmaf(v, v, sw);
        }
        // the current features are next frame's source: hand them on already transposed (pixel-major, permuted)
        if (cur_t_out != nullptr && live)     // not read again before the next frame: cache-streaming store
            asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(cur_t_out + tpix * C + j * 128 + 4 * lane),
                         "f"(fq[0]), "f"(fq[1]), "f"(fq[2]), "f"(fq[3]) : "memory");
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dot += __shfl_xor_sync(0xffffffffu, dot, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        sw += __shfl_xor_sync(0xffffffffu, sw, o);
    }
    if (lane == 0 && live) {
        out_disp[tpix] = __fdiv_rn(tail.x, nrm);
        out_mask[tpix] = m;
        // sum_c normalize(f)_c * normalize(v)_c  (F.normalize eps 1e-12), times the splat mask (tc_stereo.py:139-140)
        const float den = __fmul_rn(fmaxf(sqrtf(s1), 1e-12f), fmaxf(sqrtf(sw), 1e-12f));
        out_cost[tpix] = __fmul_rn(__fdiv_rn(dot, den), m);
    }
}

template <int kGroups, bool kLists>
__global__ void __launch_bounds__(kWarpThreads, 4)   // 64 registers; lists: 3 CTAs at 80 registers 0.379 ms, 5 at 48 spill (0.42 ms)
warp_cost_kernel(const float* __restrict__ accum, const ListArgs la, const float* __restrict__ cur_fmap, float* __restrict__ out_disp,
                 float* __restrict__ out_mask, float* __restrict__ out_cost, float* __restrict__ cur_t_out, int H, int W) {
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
    const size_t row_px = ((size_t)b * H + h) * W;
    int ls0[kPx], ln[kPx];                     // list offsets of this warp's targets: fetched first, needed last
    if (kLists) {
#pragma unroll
        for (int i = 0; i < kPx; ++i) {
            const int wl = warp * kPx + i;
            const bool live = w0 + wl < W;
            const size_t tpix = row_px + (live ? w0 + wl : 0);
            ls0[i] = __ldg(la.start + tpix);
            ln[i] = live ? __ldg(la.start + tpix + 1) - ls0[i] : 0;
        }
    }
    float f[kPerWarp];
#pragma unroll
    for (int k = 0; k < kPerWarp; ++k)
        f[k] = in_w ? ldg_stream_f1(cur_fmap + base + (size_t)(warp + 8 * k) * plane) : 0.0f;
    if (kLists) {
        // lists: one pixel at a time (gather, then its cosine), so only one accumulator row is live in registers and
        // the block's only barrier comes before the list walks, whose lengths differ from warp to warp
#pragma unroll
        for (int k = 0; k < kPerWarp; ++k) tile[(warp + 8 * k) * 33 + lane] = f[k];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < kPx; ++i) {
            const int wl = warp * kPx + i;
            const bool live = w0 + wl < W;
            float4 a[kGroups];
            float2 tail;
            gather_target<kGroups>(la, (size_t)b * plane, ls0[i], ln[i], lane, a, tail);   // n = 0 for a dead pixel
            cost_pixel<kGroups>(a, tail, tile, wl, live, lane, row_px + w0 + wl, out_disp, out_mask, out_cost, cur_t_out);
        }
        return;
    }
    float2 tail[kPx];
    float4 a[kPx][kGroups];
#pragma unroll
    for (int i = 0; i < kPx; ++i) {            // all rows of the warp's pixels in flight at once
        const int wl = warp * kPx + i;
        const bool live = w0 + wl < W;
        const float* src = accum + (row_px + (live ? w0 + wl : 0)) * CP;
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
        cost_pixel<kGroups>(a[i], tail[i], tile, wl, w0 + wl < W, lane, row_px + w0 + wl, out_disp, out_mask, out_cost, cur_t_out);
    }
}

// ---- forward warp, kernel C: normalise, mask, NCHW re-layout, matching cost -------------------------------------
template <int kGroups, bool kLists>
__global__ void __launch_bounds__(kWarpThreads)
warp_finalize_kernel(const float* __restrict__ accum, const ListArgs la, const float* __restrict__ cur_fmap,
                     float* __restrict__ out_disp, float* __restrict__ out_fmap, float* __restrict__ out_mask,
                     float* __restrict__ out_cost, int H, int W) {
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
        if (kLists) {
            if (live) {
                const int s0 = __ldg(la.start + idx);
                gather_target<kGroups>(la, (size_t)b * plane, s0, __ldg(la.start + idx + 1) - s0, lane, a[i], tail[i]);
            } else {
                tail[i] = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < kGroups; ++j) a[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            const float* src = accum + idx * CP;
            tail[i] = live ? *reinterpret_cast<const float2*>(src + C) : make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < kGroups; ++j)
                a[i][j] = live ? *reinterpret_cast<const float4*>(src + j * 128 + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
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
struct SplatCorners {
    int t[4];      // target pixel index inside the sample, -1 when outside the image or the product is zero
    float w[4];    // e * bilinear weight
};

__device__ __forceinline__ SplatCorners splat_corners(float e, float fx_, float fy_, int H, int W) {
    SplatCorners c;
    const int nwx = (int)floorf(fx_), nwy = (int)floorf(fy_);
    const int sex = nwx + 1, sey = nwy + 1;
    // softsplat.py:314-317
    const float wt[4] = {__fmul_rn(__fsub_rn((float)sex, fx_), __fsub_rn((float)sey, fy_)),    // NW
                         __fmul_rn(__fsub_rn(fx_, (float)nwx), __fsub_rn((float)sey, fy_)),    // NE
                         __fmul_rn(__fsub_rn((float)sex, fx_), __fsub_rn(fy_, (float)nwy)),    // SW
                         __fmul_rn(__fsub_rn(fx_, (float)nwx), __fsub_rn(fy_, (float)nwy))};   // SE
    const int txs[4] = {nwx, sex, nwx, sex};
    const int tys[4] = {nwy, nwy, sey, sey};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool in = txs[k] >= 0 && txs[k] < W && tys[k] >= 0 && tys[k] < H;
        c.w[k] = __fmul_rn(e, wt[k]);
        c.t[k] = (in && c.w[k] != 0.0f) ? tys[k] * W + txs[k] : -1;   // a zero product adds nothing, not even to the mask
    }
    return c;
}

// wgt = valid * exp(clamp(disp' - mean, -50, 50)), 0 for pixels the splat skips (invalid, non-finite target).
__global__ void __launch_bounds__(256)
warp_weight_kernel(const float* __restrict__ disp1, const float* __restrict__ tx, const float* __restrict__ ty,
                   float* __restrict__ valid_to_wgt, const double* __restrict__ sums,
                   int B, int HW, int per_sample_mean, int* __restrict__ cnt, int H, int W) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
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
    if (cnt != nullptr && w != 0.0f) {          // list formulation, step (1): contributions per target
        const SplatCorners c = splat_corners(w, tx[o], ty[o], H, W);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (c.t[k] >= 0) atomicAdd(cnt + (size_t)b * HW + c.t[k], 1);
    }
}

// ---- forward warp, list formulation ---------------------------------------------------------------------------
// The scatter above moves every accumulator byte through DRAM four times (memset, red read-modify-write, finalize
// read).  The same sum can be collected from the target's side without any accumulator if each target knows who
// feeds it.  So: (1) count the contributions each target receives (one int atomic per corner instead of 65 vector
// reds), (2) prefix-sum the counts into list offsets (per row, then across rows), (3) fill the lists with
// (source pixel, e * bilinear weight), (4) one pass over the targets: walk the list, gather the source features
// straight from the NCHW map (lanes are neighbouring targets, their sources are neighbours too: coalesced), divide
// by the summed weights, write mask / disparity / features and fold the matching cost in.  Any flow field is
// handled (the lists are exact, CSR) and each list is sorted by source index first, which fixes the summation order
// (the reference's atomic scatter is unordered): this is the TCS_WARP_DETERMINISTIC mode.  The gather of NCHW
// features at per-lane source pixels costs ~6 sectors per request with i.i.d. flow and the kernel is latency-bound
// (636 us against 374 us for the scatter at 540p x 8), so the scatter stays the default.
// (3) append the entries (the counters of step (1), made by warp_weight_kernel, serve as cursors and end at zero).
__global__ void __launch_bounds__(256)
warp_fill_kernel(const float* __restrict__ wgt, const float* __restrict__ tx, const float* __restrict__ ty,
                 int* __restrict__ cnt, const int* __restrict__ start, int2* __restrict__ entries, int H, int W) {
    const int b = blockIdx.y;
    const int HW = H * W;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const size_t o = (size_t)b * HW + i;
    const float e = __ldg(wgt + o);
    if (e == 0.0f) return;
    const SplatCorners c = splat_corners(e, __ldg(tx + o), __ldg(ty + o), H, W);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (c.t[k] < 0) continue;
        const size_t t = (size_t)b * HW + c.t[k];
        const int slot = atomicSub(cnt + t, 1) - 1;
        entries[(size_t)__ldg(start + t) + slot] = make_int2(i, __float_as_int(c.w[k]));
    }
}

// (2a) contributions per target row.
__global__ void __launch_bounds__(256)
warp_rowsum_kernel(const int* __restrict__ cnt, int* __restrict__ rowtot, int W) {
    const int row = blockIdx.x;
    int s = 0;
    for (int x = threadIdx.x; x < W; x += 256) s += cnt[(size_t)row * W + x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ int ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += ws[w];
        rowtot[row] = t;
    }
}

// (2b) list offsets: rows before this one (summed here: a few thousand ints) + exclusive scan inside the row.
__global__ void __launch_bounds__(256)
warp_offsets_kernel(const int* __restrict__ cnt, const int* __restrict__ rowtot, int* __restrict__ start, int W, int rows) {
    const int row = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int ws[8];
    __shared__ int carry;
    int s = 0;
    for (int r = threadIdx.x; r < row; r += 256) s += rowtot[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) ws[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += ws[w];
        carry = t;
    }
    __syncthreads();
    for (int x0 = 0; x0 < W; x0 += 256) {
        const int x = x0 + threadIdx.x;
        const int v = (x < W) ? cnt[(size_t)row * W + x] : 0;
        int inc = v;                                            // inclusive scan inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        const int base = carry;
        __syncthreads();                                        // everyone has read carry and the previous ws
        if (lane == 31) ws[warp] = inc;
        __syncthreads();
        int before = 0;
        for (int w = 0; w < warp; ++w) before += ws[w];
        if (x < W) start[(size_t)row * W + x] = base + before + inc - v;
        if (threadIdx.x == 255) carry = base + before + inc;
        __syncthreads();
    }
    if (row == rows - 1 && threadIdx.x == 0) start[(size_t)rows * W] = carry;   // one past the end
}

// (3b) TCS_WARP_DETERMINISTIC: order each list by source pixel.
__global__ void __launch_bounds__(256)
warp_sort_kernel(const int* __restrict__ start, int2* __restrict__ entries, long long npix) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npix) return;
    const int s0 = start[t], n = start[t + 1] - s0;
    int2* e = entries + s0;
    if (n <= 1) return;
    if (n <= 8) {                                               // the usual case: in registers, 19-comparator network
        int2 r[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = (i < n) ? e[i] : make_int2(0x7fffffff, 0);   // padding sinks to the end
        constexpr int kNet[19][2] = {{0, 1}, {2, 3}, {4, 5}, {6, 7}, {0, 2}, {1, 3}, {4, 6}, {5, 7}, {1, 2}, {5, 6},
                                     {0, 4}, {1, 5}, {2, 6}, {3, 7}, {2, 4}, {3, 5}, {1, 2}, {3, 4}, {5, 6}};
#pragma unroll
        for (int c = 0; c < 19; ++c) {
            int2& a = r[kNet[c][0]];
            int2& b = r[kNet[c][1]];
            if (a.x > b.x) { const int2 tmp = a; a = b; b = tmp; }     // source pixels of a list are distinct: no ties
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < n) e[i] = r[i];
        return;
    }
    for (int i = 1; i < n; ++i) {                               // insertion sort for the rare long list
        const int2 key = e[i];
        int j = i - 1;
        while (j >= 0 && e[j].x > key.x) { e[j + 1] = e[j]; --j; }
        e[j + 1] = key;
    }
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
constexpr int kSampleCh = 16;

// One (output pixel, channel group) of the gather: img_b / out_b are the sample's [C,Hi,Wi] / [C,Ho*Wo] planes.
// kBatch: channels whose 4 corner loads are issued back to back before any is used.
// Addresses: ONE 64-bit base per tensor and 32-bit element offsets (the hosts require C * H * W < 2^31), so that a load costs
// a 32-bit add and one widening multiply-add instead of five 64-bit instructions: the kernel is bound by instruction issue
// (67 % issue-active at 37 % DRAM with 57 instructions per output element before, profiles/r02_warp_kernels.md).
template <bool kFull, int kBatch = kSampleCh, int kCh = kSampleCh>   // kFull: C is a multiple of kCh, no per-channel guards (keeps the loads batched)
__device__ __forceinline__ void sample_group(const float* __restrict__ img_b, float gx, float gy, float* __restrict__ out_b,
                                             int C, int Hi, int Wi, int HWo, int i, int c_begin) {
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
    // Four corner addresses and the output address walk the channel planes as plain 64-bit BYTE addresses (one add-with-carry
    // pair each per channel).  Written as pointers or as base[index] the compiler keeps element indices and rebuilds every
    // address with a 64-bit add and a 64-bit shift-add: 31 integer instructions per channel against 11 of useful work.
    const uint64_t plane_b = (uint64_t)Hi * Wi * 4u, out_b_step = (uint64_t)HWo * 4u;
    const uint64_t src = reinterpret_cast<uint64_t>(img_b) + (uint64_t)c_begin * plane_b;
    uint64_t a_nw = src + 4u * (uint64_t)(unsigned)(y0 * Wi + x0), a_ne = src + 4u * (uint64_t)(unsigned)(y0 * Wi + x1);
    uint64_t a_sw = src + 4u * (uint64_t)(unsigned)(y1 * Wi + x0), a_se = src + 4u * (uint64_t)(unsigned)(y1 * Wi + x1);
    uint64_t a_dst = reinterpret_cast<uint64_t>(out_b) + ((uint64_t)c_begin * HWo + i) * 4u;
#pragma unroll
    for (int k0 = 0; k0 < kCh; k0 += kBatch) {
        float v[kBatch][4];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            const int k = k0 + j;
            v[j][0] = ldg_ordered_f1(reinterpret_cast<const float*>(a_nw));
            v[j][1] = ldg_ordered_f1(reinterpret_cast<const float*>(a_ne));
            v[j][2] = ldg_ordered_f1(reinterpret_cast<const float*>(a_sw));
            v[j][3] = ldg_ordered_f1(reinterpret_cast<const float*>(a_se));
            if (kFull || c_begin + k + 1 < C) {                     // a dead channel re-reads the last live one
                a_nw += plane_b; a_ne += plane_b; a_sw += plane_b; a_se += plane_b;
            }
        }
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
            const int k = k0 + j;
            if (!kFull && c_begin + k >= C) break;
            // corners accumulate in the order nw, ne, sw, se; a corner outside the image contributes nothing
            float r = 0.0f;
            if (nw) r = __fmul_rn(v[j][0], w_nw);
            if (ne) r = __fadd_rn(r, __fmul_rn(v[j][1], w_ne));
            if (sw) r = __fadd_rn(r, __fmul_rn(v[j][2], w_sw));
            if (se) r = __fadd_rn(r, __fmul_rn(v[j][3], w_se));
            stg_stream_f1(reinterpret_cast<float*>(a_dst), r);
            a_dst += out_b_step;
        }
    }
}

template <bool kFull>
__global__ void __launch_bounds__(256)
bilinear_sample_kernel(const float* __restrict__ img, const float* __restrict__ grid_xy, float* __restrict__ out,
                       int C, int Hi, int Wi, int Ho, int Wo) {
    const int b = blockIdx.z;
    const int HWo = Ho * Wo;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HWo) return;
    const float gx = __ldg(grid_xy + ((size_t)b * 2 + 0) * HWo + i);
    const float gy = __ldg(grid_xy + ((size_t)b * 2 + 1) * HWo + i);
    sample_group<kFull>(img + (size_t)b * C * Hi * Wi, gx, gy, out + (size_t)b * C * HWo, C, Hi, Wi, HWo, i, blockIdx.y * kSampleCh);
}

// ---- 0.5 * interpolate(grid, 0.5, bilinear, align_corners=True) ------------------------------------------------------------
// One element of 0.5 * F.interpolate(plane, scale_factor=0.5, 'bilinear', align_corners=True): the source position and
// the four interpolation weights of output (yo, xo) ...
struct HalvePos {
    int y0, x0, yp, xp;
    float ly0, ly1, lx0, lx1;
};
__device__ __forceinline__ HalvePos halve_pos(int H, int W, int Ho, int Wo, int yo, int xo) {
    HalvePos q;
    // ATen area_pixel_compute_scale with align_corners: (in - 1) / (out - 1), 0 when out == 1
    const float sh = (Ho > 1) ? __fdiv_rn((float)(H - 1), (float)(Ho - 1)) : 0.0f;
    const float sw = (Wo > 1) ? __fdiv_rn((float)(W - 1), (float)(Wo - 1)) : 0.0f;
    const float ys = __fmul_rn(sh, (float)yo), xs = __fmul_rn(sw, (float)xo);
    q.y0 = (int)ys; q.x0 = (int)xs;
    q.yp = (q.y0 < H - 1) ? 1 : 0; q.xp = (q.x0 < W - 1) ? 1 : 0;
    q.ly1 = __fsub_rn(ys, (float)q.y0); q.ly0 = __fsub_rn(1.0f, q.ly1);
    q.lx1 = __fsub_rn(xs, (float)q.x0); q.lx0 = __fsub_rn(1.0f, q.lx1);
    return q;
}
// ... and the value from the four source values.
__device__ __forceinline__ float halve_mix(const HalvePos& q, float v00, float v01, float v10, float v11) {
    const float top = __fadd_rn(__fmul_rn(q.lx0, v00), __fmul_rn(q.lx1, v01));
    const float bot = __fadd_rn(__fmul_rn(q.lx0, v10), __fmul_rn(q.lx1, v11));
    return __fmul_rn(0.5f, __fadd_rn(__fmul_rn(q.ly0, top), __fmul_rn(q.ly1, bot)));
}
__device__ __forceinline__ float halve_at(const float* __restrict__ plane, int H, int W, int Ho, int Wo, int yo, int xo) {
    const HalvePos q = halve_pos(H, W, Ho, Wo, yo, xo);
    const float* p = plane + (long long)q.y0 * W + q.x0;
    return halve_mix(q, __ldg(p), __ldg(p + q.xp), __ldg(p + (long long)q.yp * W), __ldg(p + (long long)q.yp * W + q.xp));
}

__global__ void __launch_bounds__(256)
grid_halve_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, int Ho, int Wo, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int xo = (int)(i % Wo);
    const int yo = (int)((i / Wo) % Ho);
    const long long bc = i / ((long long)Wo * Ho);
    out[i] = halve_at(in + bc * (long long)H * W, H, W, Ho, Wo, yo, xo);
}

// ---- the whole hidden-state warp of tc_stereo.py:159-163 in ONE launch ---------------------------------------------
// Three gathers with a grid halved between them were five launches, of which the two small levels ran at 50 % / 28 %
// of the HBM rate (tails and launch gaps, not bytes).  Here the blocks of all three levels share one grid
// (blockIdx.x runs over level 0's pixel blocks, then level 1's, then level 2's), and a thread of level 1 / 2 halves
// the backward grid for its own pixel on the fly - the same operations in the same order as grid_halve_kernel, so
// the sampled positions are bit-identical to the chained launches (level 2 re-derives the four level-1 values it
// mixes: 16 grid reads per plane, from L1/L2).
struct Hidden3Params {
    const float* net[3];
    float* out[3];
    const float* grid;
    int C[3];
    int H[3], W[3];
    int blocks[3];       // pixel blocks of each level
};

// The backward grid at level l for output pixel i: read (l = 0) or halved once / twice on the fly.  Not inlined: its
// temporaries must not stretch the register allocation of the gather that follows (180 registers when inlined).
__device__ __noinline__ float2 hidden3_grid(const float* __restrict__ g0, int l, int i, int H0, int W0, int H1, int W1, int Hl, int Wl) {
    float g[2];
    if (l == 0) {
        g[0] = __ldg(g0 + i);
        g[1] = __ldg(g0 + (size_t)H0 * W0 + i);
    } else {
        const int yo = i / Wl, xo = i - yo * Wl;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            const float* plane = g0 + (size_t)c * H0 * W0;
            if (l == 1) {
                g[c] = halve_at(plane, H0, W0, H1, W1, yo, xo);
            } else {
                const HalvePos q = halve_pos(H1, W1, Hl, Wl, yo, xo);     // position in the level-1 grid, whose values are
                const float v00 = halve_at(plane, H0, W0, H1, W1, q.y0, q.x0);             // halved from level 0 here
                const float v01 = halve_at(plane, H0, W0, H1, W1, q.y0, q.x0 + q.xp);
                const float v10 = halve_at(plane, H0, W0, H1, W1, q.y0 + q.yp, q.x0);
                const float v11 = halve_at(plane, H0, W0, H1, W1, q.y0 + q.yp, q.x0 + q.xp);
                g[c] = halve_mix(q, v00, v01, v10, v11);
            }
        }
    }
    return make_float2(g[0], g[1]);
}

#ifndef TCS_HIDDEN_CH
#define TCS_HIDDEN_CH 32
#endif
constexpr int kHiddenCh = TCS_HIDDEN_CH;   // channels per thread: the per-pixel set-up (grid halving, two exact divisions) is paid once per group

template <bool kFull>   // every level's channel count is a multiple of kHiddenCh
__global__ void __launch_bounds__(256)
warp_hidden3_kernel(const Hidden3Params p) {
    const int b = blockIdx.z;
    int bx = blockIdx.x, l = 0;
    if (bx >= p.blocks[0]) { bx -= p.blocks[0]; l = 1; }
    if (l == 1 && bx >= p.blocks[1]) { bx -= p.blocks[1]; l = 2; }
    const int Hl = l == 0 ? p.H[0] : l == 1 ? p.H[1] : p.H[2];
    const int Wl = l == 0 ? p.W[0] : l == 1 ? p.W[1] : p.W[2];
    const int Cl = l == 0 ? p.C[0] : l == 1 ? p.C[1] : p.C[2];
    const int c_begin = blockIdx.y * kHiddenCh;
    if (c_begin >= Cl) return;                                   // block-uniform
    const int HWl = Hl * Wl;
    const int i = bx * 256 + threadIdx.x;
    if (i >= HWl) return;
    const float2 gxy = hidden3_grid(p.grid + (size_t)b * 2 * p.H[0] * p.W[0], l, i, p.H[0], p.W[0], p.H[1], p.W[1], Hl, Wl);
    const float g[2] = {gxy.x, gxy.y};
    const float* net = l == 0 ? p.net[0] : l == 1 ? p.net[1] : p.net[2];
    float* out = l == 0 ? p.out[0] : l == 1 ? p.out[1] : p.out[2];
    const float* img_b = net + (size_t)b * Cl * HWl;
    float* out_b = out + (size_t)b * Cl * HWl;
#ifndef TCS_HIDDEN_BATCH
#define TCS_HIDDEN_BATCH 8
#endif
    sample_group<kFull, TCS_HIDDEN_BATCH, kHiddenCh>(img_b, g[0], g[1], out_b, Cl, Hl, Wl, HWl, i, c_begin);   // 8: 44 registers, 5 CTAs per SM
}

static size_t align256(size_t n) { return (n + 255) & ~static_cast<size_t>(255); }

struct WarpScratch {
    size_t sums, cnt, accum, disp1, tx, ty, valid, start, rowtot, entries, total;
};
static WarpScratch warp_scratch_layout(int B, int C, int H, int W) {
    WarpScratch s;
    const size_t npix = (size_t)B * H * W;
    s.sums = 0;
    s.cnt = align256((size_t)B * sizeof(double) + 64);     // per-sample sums, then the control words
    s.accum = s.cnt + align256(npix * sizeof(int));        // contributions per target (list mode), zeroed with the header
    s.disp1 = s.accum + align256(npix * (C + 4) * sizeof(float));
    s.tx = s.disp1 + align256(npix * sizeof(float));
    s.ty = s.tx + align256(npix * sizeof(float));
    s.valid = s.ty + align256(npix * sizeof(float));
    s.start = s.valid + align256(npix * sizeof(float));
    s.rowtot = s.start + align256((npix + 1) * sizeof(int));
    s.entries = s.rowtot + align256((size_t)B * H * sizeof(int));
    s.total = s.entries + align256(4 * npix * sizeof(int2));
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
                                const float* fmap_t, float* cur_t_out,
                                void* scratch, int B, int C, int H, int W, int flags, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(disp && fmap && rel_T && K && K_inv && baseline && out_disp && out_mask && scratch,
                TCS_E_BADARG, "tcs_warp_forward: null pointer");
    TCS_REQUIRE(out_fmap != nullptr || out_cost != nullptr, TCS_E_BADARG, "tcs_warp_forward: neither out_fmap nor out_cost requested");
    TCS_REQUIRE(out_cost == nullptr || cur_fmap != nullptr, TCS_E_BADARG, "tcs_warp_forward: out_cost needs cur_fmap");
    TCS_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, TCS_E_BADARG, "tcs_warp_forward: non-positive size");
    TCS_REQUIRE(C == 128 || C == 256 || C == 384 || C == 512, TCS_E_SHAPE, "tcs_warp_forward: C=%d must be 128, 256, 384 or 512", C);
    TCS_REQUIRE(H <= 65535 && B <= 65535, TCS_E_SHAPE, "tcs_warp_forward: H and B must be <= 65535");
    TCS_REQUIRE(aligned16(fmap) && aligned16(scratch) && aligned16(cur_fmap) && aligned16(fmap_t) && aligned16(cur_t_out), TCS_E_ALIGN,
                "tcs_warp_forward: fmap, cur_fmap, fmap_t, cur_t_out and scratch must be 16-byte aligned");
    TCS_REQUIRE(cur_t_out == nullptr || (out_fmap == nullptr && out_cost != nullptr), TCS_E_BADARG,
                "tcs_warp_forward: cur_t_out is produced by the cost-only call (out_fmap null, out_cost given)");
    TCS_REQUIRE(fmap_t == nullptr || (flags & TCS_WARP_DETERMINISTIC), TCS_E_BADARG,
                "tcs_warp_forward: fmap_t is consumed by the TCS_WARP_DETERMINISTIC (list) formulation only");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const WarpScratch L = warp_scratch_layout(B, C, H, W);
    char* base = static_cast<char*>(scratch);
    double* sums = reinterpret_cast<double*>(base + L.sums);
    float* accum = reinterpret_cast<float*>(base + L.accum);
    float* disp1 = reinterpret_cast<float*>(base + L.disp1);
    float* tx = reinterpret_cast<float*>(base + L.tx);
    float* ty = reinterpret_cast<float*>(base + L.ty);
    float* valid = reinterpret_cast<float*>(base + L.valid);

    const int per_sample_mean = (flags & TCS_WARP_PER_SAMPLE_MEAN) ? 1 : 0;
    const bool lists = (flags & TCS_WARP_DETERMINISTIC) != 0;
    // scatter (default): zero the sums and the accumulator; lists: the sums and the per-target counters only
    TCS_CHECK_CUDA(cudaMemsetAsync(base, 0, lists ? L.accum : L.disp1, s));
    const dim3 pgrid(ceil_div(H * W, 256), B);
    warp_geometry_kernel<<<pgrid, 256, 0, s>>>(disp, rel_T, K, K_inv, baseline, disp1, tx, ty, valid, sums, H, W);
    TCS_CHECK_LAUNCH("tcs_warp_forward(geometry)");
    warp_weight_kernel<<<pgrid, 256, 0, s>>>(disp1, tx, ty, valid, sums, B, H * W, per_sample_mean,
                                             lists ? reinterpret_cast<int*>(base + L.cnt) : nullptr, H, W);
    TCS_CHECK_LAUNCH("tcs_warp_forward(weights)");
    ListArgs la = {nullptr, nullptr, nullptr, nullptr};
    if (lists) {
        // ---- list formulation: no accumulator, no floating-point atomics, fixed summation order
        int* cnt = reinterpret_cast<int*>(base + L.cnt);
        int* start = reinterpret_cast<int*>(base + L.start);
        int* rowtot = reinterpret_cast<int*>(base + L.rowtot);
        int2* entries = reinterpret_cast<int2*>(base + L.entries);
        la.src_t = fmap_t != nullptr ? fmap_t : accum;   // else the accumulator's space holds the transposed features
        la.start = start;
        la.entries = entries;
        la.disp1 = disp1;
        const long long npix = (long long)B * H * W;
        TCS_REQUIRE(npix * 4 < 0x7fffffffLL, TCS_E_SHAPE, "tcs_warp_forward: too many pixels for 32-bit list offsets");
        warp_rowsum_kernel<<<B * H, 256, 0, s>>>(cnt, rowtot, W);
        TCS_CHECK_LAUNCH("tcs_warp_forward(row sums)");
        warp_offsets_kernel<<<B * H, 256, 0, s>>>(cnt, rowtot, start, W, B * H);
        TCS_CHECK_LAUNCH("tcs_warp_forward(offsets)");
        warp_fill_kernel<<<pgrid, 256, 0, s>>>(valid, tx, ty, cnt, start, entries, H, W);
        TCS_CHECK_LAUNCH("tcs_warp_forward(fill)");
        warp_sort_kernel<<<(unsigned)ceil_div_ll(npix, 256), 256, 0, s>>>(start, entries, npix);
        TCS_CHECK_LAUNCH("tcs_warp_forward(sort)");
    }
    const dim3 grid(ceil_div(W, kTileW), H, B);
    const size_t smem_splat = (size_t)kTileW * (C + 1) * sizeof(float);
    const size_t smem_fin = ((size_t)C * 33 + 8 * 32 * 3 + 32) * sizeof(float);
    const size_t smem_cost = (size_t)C * 33 * sizeof(float);
#define TCS_WARP_CASE(G)                                                                                              \
    case G: {                                                                                                         \
        TCS_ONCE_PER_DEVICE(                                                                                          \
            TCS_CHECK_CUDA(cudaFuncSetAttribute(warp_splat_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_splat)); \
            TCS_CHECK_CUDA(cudaFuncSetAttribute(warp_transpose_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_splat)); \
            TCS_CHECK_CUDA(cudaFuncSetAttribute(warp_finalize_kernel<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fin)); \
            TCS_CHECK_CUDA(cudaFuncSetAttribute(warp_finalize_kernel<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fin)); \
            TCS_CHECK_CUDA(cudaFuncSetAttribute(warp_cost_kernel<G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cost)); \
            TCS_CHECK_CUDA(cudaFuncSetAttribute(warp_cost_kernel<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cost)); \
            { const int cv = carveout_percent("TCS_CARVE_SPLAT", -1); if (cv >= 0) TCS_CHECK_CUDA(cudaFuncSetAttribute(warp_splat_kernel<G>, cudaFuncAttributePreferredSharedMemoryCarveout, cv)); } \
        );                                                                                                            \
        const bool cost_only = out_fmap == nullptr && out_cost != nullptr && cur_fmap != nullptr;                     \
        if (lists) {                                                                                                  \
            if (fmap_t == nullptr) {                                                                                  \
                warp_transpose_kernel<G><<<grid, kWarpThreads, smem_splat, s>>>(fmap, accum, H, W);                   \
                TCS_CHECK_LAUNCH("tcs_warp_forward(transpose)");                                                      \
            }                                                                                                         \
            if (cost_only)                                                                                            \
                warp_cost_kernel<G, true><<<grid, kWarpThreads, smem_cost, s>>>(accum, la, cur_fmap, out_disp, out_mask, out_cost, cur_t_out, H, W); \
            else                                                                                                      \
                warp_finalize_kernel<G, true><<<grid, kWarpThreads, smem_fin, s>>>(accum, la, cur_fmap, out_disp, out_fmap, out_mask, out_cost, H, W); \
        } else {                                                                                                      \
            warp_splat_kernel<G><<<grid, kWarpThreads, smem_splat, s>>>(fmap, disp1, tx, ty, valid, accum, B, H, W);  \
            TCS_CHECK_LAUNCH("tcs_warp_forward(splat)");                                                              \
            if (cost_only)                                                                                            \
                warp_cost_kernel<G, false><<<grid, kWarpThreads, smem_cost, s>>>(accum, la, cur_fmap, out_disp, out_mask, out_cost, cur_t_out, H, W); \
            else                                                                                                      \
                warp_finalize_kernel<G, false><<<grid, kWarpThreads, smem_fin, s>>>(accum, la, cur_fmap, out_disp, out_fmap, out_mask, out_cost, H, W); \
        }                                                                                                             \
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
    TCS_REQUIRE((long long)groups * kSampleCh * Hi * Wi < 0x7fffffffLL && (long long)groups * kSampleCh * Ho * Wo < 0x7fffffffLL, TCS_E_SHAPE,
                "tcs_bilinear_sample: C * H * W must stay below 2^31 (32-bit element offsets per sample)");
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

extern "C" int tcs_warp_hidden_states(const float* net0, const float* net1, const float* net2, const float* grid,
                                      float* out0, float* out1, float* out2, int B, int C0, int C1, int C2, int H, int W,
                                      void* stream) {
    using namespace tcs;
    TCS_REQUIRE(net0 && net1 && net2 && grid && out0 && out1 && out2, TCS_E_BADARG, "tcs_warp_hidden_states: null pointer");
    TCS_REQUIRE(B > 0 && B <= 65535 && C0 > 0 && C1 > 0 && C2 > 0, TCS_E_BADARG, "tcs_warp_hidden_states: bad sizes");
    TCS_REQUIRE(H >= 4 && W >= 4, TCS_E_SHAPE, "tcs_warp_hidden_states: need H, W >= 4 (three levels, each halved)");
    TCS_REQUIRE((long long)H * W < 0x7fffffffLL, TCS_E_SHAPE, "tcs_warp_hidden_states: image plane too large");
    Hidden3Params p{};
    p.net[0] = net0; p.net[1] = net1; p.net[2] = net2;
    p.out[0] = out0; p.out[1] = out1; p.out[2] = out2;
    p.grid = grid;
    p.C[0] = C0; p.C[1] = C1; p.C[2] = C2;
    p.H[0] = H; p.W[0] = W;
    for (int l = 1; l < 3; ++l) { p.H[l] = p.H[l - 1] / 2; p.W[l] = p.W[l - 1] / 2; }     // F.interpolate(scale_factor=0.5): floor
    int cmax = C0 > C1 ? C0 : C1;
    if (C2 > cmax) cmax = C2;
    const int groups = ceil_div(cmax, kHiddenCh);
    TCS_REQUIRE(groups <= 65535, TCS_E_SHAPE, "tcs_warp_hidden_states: C too large");
    TCS_REQUIRE((long long)groups * kHiddenCh * H * W < 0x7fffffffLL, TCS_E_SHAPE, "tcs_warp_hidden_states: C * H * W must stay below 2^31");
    long long bx = 0;
    for (int l = 0; l < 3; ++l) { p.blocks[l] = ceil_div(p.H[l] * p.W[l], 256); bx += p.blocks[l]; }
    dim3 g((unsigned)bx, groups, B);
    if (C0 % kHiddenCh == 0 && C1 % kHiddenCh == 0 && C2 % kHiddenCh == 0)
        warp_hidden3_kernel<true><<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    else
        warp_hidden3_kernel<false><<<g, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    TCS_CHECK_LAUNCH("tcs_warp_hidden_states");
    return 0;
}
