// Fused correlation build: fp32 NCHW feature maps in, every pyramid level out, in ONE kernel — no operand
// round trip through HBM.  ref: core/corr.py:54-62 (F.normalize + einsum) and core/corr.py:15-23 (pyramid).
//
// tcs_corr_prepass + tcs_corr_build move 2.07 GB for an algorithmic 1.0 GB at 540p x 8 sequences (the
// normalised 16-bit operands are written once and read once).  Here the operands never leave the SM:
//
//   warp 0        tcgen05.mma issuer (one thread): 128 x N x 16 UMMAs, fp32 accumulators in TMEM
//                 (2 x 256 columns: tile i+1's MMAs overlap tile i's epilogue).
//   warps 1..7    converters.  Per (b,h) image row: pass 1 reads the row of both maps (coalesced along w),
//                 accumulates sum x^2 per pixel in a fixed order and publishes 1/max(||x||, 1e-12);
//                 pass 2 re-reads the row (an L2 hit: 0.5 MB per row and CTA), scales, splits into 16-bit
//                 hi / lo and stores straight into the K-major SWIZZLE_128B layout the UMMA descriptors
//                 expect (the layout TMA would have produced), one 64-channel K block per pipeline stage;
//                 fence.proxy.async + mbarrier arrive hand the stage to the MMA thread.
//   warps 8..15   epilogue, shared with corr_build.cu (corr_epilogue.cuh).
//
// A tile is one (b,h) row x 128 w1 x all w2 (W2 <= 240 fits one UMMA N); a CTA walks whole rows so that the
// norms are computed once per row.  x / ||x|| is evaluated as x * (1 / ||x||) here (one rounding more than
// the pre-pass's exact division, far below the 16-bit split that follows).
#include "tcs_common.cuh"
#include "sm100_ptx.cuh"
#include "corr_epilogue.cuh"

namespace tcs {
namespace fused {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kMaxN = 240;                          // one N tile; 2 stages of hi+lo must fit in shared memory
constexpr int kMaxW1 = 384;
constexpr int kStages = 2;
constexpr int kATile = kBlockM * kBlockK * 2;       // 16 KB
constexpr int kBTile = kMaxN * kBlockK * 2;         // 30 KB
constexpr int kStageBytes = 2 * kATile + 2 * kBTile;    // A_hi, A_lo, B_hi, B_lo = 92 KB
constexpr int kAccStages = 2;
constexpr int kAccCols = 256;
constexpr int kTmemCols = kAccStages * kAccCols;
constexpr int kConvWarps = 7;
constexpr int kConvThreads = kConvWarps * 32;       // 224
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (1 + kConvWarps + kEpiWarps);   // 512
constexpr int kEpiStageBytes = 4096;                // per epilogue warp: [32][32] fp32, reused for level 1
constexpr int kNormSlices = 2;                      // channel slices whose partial sums are combined in order
constexpr int kNormCols = kMaxW1 + kMaxN + 16;      // A columns then B columns
constexpr int kNormBytes = kNormSlices * kNormCols * 4 + kNormCols * 4;
constexpr int kBarrierBytes = 256;
constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kEpiWarps * kEpiStageBytes + kNormBytes + kBarrierBytes;
static_assert(kSmemBytes <= 232448, "fused build: shared memory budget exceeded");

struct Params {
    const float* fmap1;
    const float* fmap2;
    float* lvl[TCS_MAX_LEVELS];
    int H, W1, W2, C, num_levels;
    int num_rows;      // B * H
    int num_m;         // ceil(W1 / 128)
    int block_n;       // W2 rounded up to 16
    int kblocks;       // C / 64
    int passes;        // 1 or 3
    int fp16;          // operand format
    uint32_t idesc;
    float out_scale;   // undoes the operand scaling in the epilogue
    float in_scale;    // 2^8 for fp16 operands (keeps unit-vector entries away from subnormals), 1 for bf16
};

template <bool kFp16>
__device__ __forceinline__ void split16(float a, float b, uint32_t& hi, uint32_t& lo) {
    if constexpr (kFp16) {
        const __half2 h = __floats2half2_rn(a, b);
        const float2 back = __half22float2(h);
        const __half2 l = __floats2half2_rn(a - back.x, b - back.y);
        hi = *reinterpret_cast<const uint32_t*>(&h);
        lo = *reinterpret_cast<const uint32_t*>(&l);
    } else {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        const float2 back = __bfloat1622float2(h);
        const __nv_bfloat162 l = __floats2bfloat162_rn(a - back.x, b - back.y);
        hi = *reinterpret_cast<const uint32_t*>(&h);
        lo = *reinterpret_cast<const uint32_t*>(&l);
    }
}

__device__ __forceinline__ void conv_barrier() {   // converters only (named barrier 1)
    asm volatile("bar.sync 1, %0;" :: "n"(kConvThreads) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// One K block of one operand: rows [w_first, w_first + rows) of fmap[b,:,h,:], channels [c0, c0 + 64) ->
// 16-bit hi / lo tiles in K-major SWIZZLE_128B layout (row r at r*128 B inside 1 KB atoms of 8 rows, its 16-byte
// chunk j stored at chunk position j ^ (r & 7)).  A work item = 32 consecutive rows (the lanes) x 8 channels:
// 8 coalesced 128-byte loads, one 16-byte shared store per tile.  Rows beyond `w_limit` are written as zeros.
template <bool kFp16>
__device__ __forceinline__ void convert_operand(const float* __restrict__ plane0, size_t plane_stride, int w_first,
                                                int w_limit, int rows, const float* __restrict__ inv, float in_scale,
                                                uint8_t* tile_hi, uint8_t* tile_lo, bool want_lo, int item0, int item_step,
                                                int lane) {
    const int groups = (rows + 31) >> 5;
    const int items = groups * 8;
    // kBatch items are in flight at once (kBatch * 8 independent loads per lane) to cover the L2 latency
    constexpr int kBatch = 4;
    for (int it0 = item0; it0 < items; it0 += item_step * kBatch) {
        float x[kBatch][8];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int it = it0 + u * item_step;
            const int g = it >> 3, oct = it & 7;
            const int r = g * 32 + lane;
            const int w = w_first + r;
            const bool live = (it < items) && (r < rows) && (w < w_limit);
            const float* src = plane0 + (size_t)((live ? oct : 0) * 8) * plane_stride + (live ? w : 0);
#pragma unroll
            for (int i = 0; i < 8; ++i) x[u][i] = ldg_ordered_f1(src + (size_t)i * plane_stride);
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int it = it0 + u * item_step;
            const int g = it >> 3, oct = it & 7;
            const int r = g * 32 + lane;                       // row inside the tile
            if (it >= items || r >= rows) continue;
            const int w = w_first + r;
            const bool live = w < w_limit;
            const float s = live ? inv[w] * in_scale : 0.0f;
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) split16<kFp16>(x[u][2 * i] * s, x[u][2 * i + 1] * s, hi[i], lo[i]);
            const uint32_t off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((oct ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(tile_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (want_lo) *reinterpret_cast<uint4*>(tile_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

template <bool kFp16>
__global__ void __launch_bounds__(kThreads, 1)
corr_build_fused_kernel(const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* epi_base = smem + kStages * kStageBytes;
    float* norm_part = reinterpret_cast<float*>(epi_base + kEpiWarps * kEpiStageBytes);   // [kNormSlices][kNormCols]
    float* inv_norm = norm_part + kNormSlices * kNormCols;                                 // [kNormCols]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(inv_norm) + kNormCols * 4);
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = bar_full + 8 * kStages;
    const uint32_t bar_tfull = bar_empty + 8 * kStages;
    const uint32_t bar_tempty = bar_tfull + 8 * kAccStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAccStages);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < kStages; ++i) {
                ptx::mbar_init(bar_full + 8 * i, kConvThreads);
                ptx::mbar_init(bar_empty + 8 * i, 1);
            }
            for (int i = 0; i < kAccStages; ++i) {
                ptx::mbar_init(bar_tfull + 8 * i, 1);
                ptx::mbar_init(bar_tempty + 8 * i, kEpiWarps * 32);
            }
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(smem_u32(tmem_slot), kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int b_tile_bytes = p.block_n * (kBlockK * 2);

    if (warp == 0) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            int iter = 0;
            for (int row = blockIdx.x; row < p.num_rows; row += gridDim.x) {
                for (int m_t = 0; m_t < p.num_m; ++m_t, ++iter) {
                    const uint32_t acc = iter & 1;
                    const uint32_t acc_phase = (iter >> 1) & 1;
                    ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                    ptx::tc_fence_after_sync();
                    const uint32_t tmem_d = tmem_base + acc * kAccCols;
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        ptx::mbar_wait(bar_full + 8 * stage, phase);
                        ptx::tc_fence_after_sync();
                        const uint32_t sa_hi = smem_u32(smem + stage * kStageBytes);
                        const uint32_t sa_lo = sa_hi + kATile;
                        const uint32_t sb_hi = sa_lo + kATile;
                        const uint32_t sb_lo = sb_hi + kBTile;
                        for (int pass = 0; pass < p.passes; ++pass) {   // hi*hi, hi*lo, lo*hi
                            const uint64_t da = ptx::make_kmajor_sw128_desc(pass == 2 ? sa_lo : sa_hi);
                            const uint64_t db = ptx::make_kmajor_sw128_desc(pass == 1 ? sb_lo : sb_hi);
#pragma unroll
                            for (int k = 0; k < kBlockK / kUmmaK; ++k)
                                ptx::umma_f16(tmem_d, da + 2 * k, db + 2 * k, p.idesc, (kb | pass | k) != 0 ? 1u : 0u);
                        }
                        ptx::umma_commit(bar_empty + 8 * stage);
                        if (kb == p.kblocks - 1) ptx::umma_commit(bar_tfull + 8 * acc);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp <= kConvWarps) {
        // ================= converters =================
        const int cw = warp - 1;                      // 0..6
        const int ct = cw * 32 + lane;                // 0..223
        const size_t plane = (size_t)p.H * p.W1;      // fmap1 channel stride
        const size_t plane2 = (size_t)p.H * p.W2;
        const int ncol = p.W1 + p.W2;                 // pass-1 columns: A then B
        const int ch_per_slice = p.C / kNormSlices;
        const bool want_lo = p.passes == 3;
        uint32_t stage = 0, phase = 0;
        for (int row = blockIdx.x; row < p.num_rows; row += gridDim.x) {
            const int b = row / p.H, h = row - b * p.H;
            const float* a_row = p.fmap1 + ((size_t)b * p.C * p.H + h) * p.W1;
            const float* b_row = p.fmap2 + ((size_t)b * p.C * p.H + h) * p.W2;
            // ---- pass 1: sum of squares per pixel.  Item = (32 columns, one channel slice); partial sums are
            // stored per slice and combined in slice order, so the result does not depend on scheduling.
            {
                const int cgroups = (ncol + 31) >> 5;
                for (int it = cw; it < cgroups * kNormSlices; it += kConvWarps) {
                    const int g = it / kNormSlices, sl = it - g * kNormSlices;
                    const int col = g * 32 + lane;
                    float acc = 0.0f;
                    if (col < ncol) {
                        const bool is_a = col < p.W1;
                        const float* src = is_a ? a_row + col : b_row + (col - p.W1);
                        const size_t ps = is_a ? plane : plane2;
                        src += (size_t)(sl * ch_per_slice) * ps;
                        for (int c = 0; c < ch_per_slice; c += 32) {   // 32 loads in flight per lane
                            float x[32];
#pragma unroll
                            for (int i = 0; i < 32; ++i) x[i] = ldg_ordered_f1(src + (size_t)(c + i) * ps);
#pragma unroll
                            for (int i = 0; i < 32; ++i) acc = fmaf(x[i], x[i], acc);
                        }
                        norm_part[sl * kNormCols + col] = acc;
                    }
                }
                conv_barrier();
                for (int col = ct; col < ncol; col += kConvThreads) {
                    float ss = norm_part[col];
#pragma unroll
                    for (int sl = 1; sl < kNormSlices; ++sl) ss += norm_part[sl * kNormCols + col];
                    inv_norm[col] = __frcp_rn(fmaxf(sqrtf(ss), 1e-12f));      // corr.py:58-59, eps of F.normalize
                }
                conv_barrier();
            }
            // ---- pass 2: one K block per stage, A tile of this M tile and the whole B row
            for (int m_t = 0; m_t < p.num_m; ++m_t) {
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    uint8_t* sa_hi = smem + stage * kStageBytes;
                    uint8_t* sa_lo = sa_hi + kATile;
                    uint8_t* sb_hi = sa_lo + kATile;
                    uint8_t* sb_lo = sb_hi + kBTile;
                    const size_t c0 = (size_t)kb * kBlockK;
                    // A: 4 row groups x 8 octets = 32 items; B: up to 8 x 8 = 64 items; interleave over the 7 warps
                    convert_operand<kFp16>(a_row + c0 * plane, plane, m_t * kBlockM, p.W1, kBlockM, inv_norm, p.in_scale,
                                           sa_hi, sa_lo, want_lo, cw, kConvWarps, lane);
                    convert_operand<kFp16>(b_row + c0 * plane2, plane2, 0, p.W2, p.block_n, inv_norm + p.W1, p.in_scale,
                                           sb_hi, sb_lo, want_lo, (cw + 4) % kConvWarps, kConvWarps, lane);
                    fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core's async proxy
                    ptx::mbar_arrive(bar_full + 8 * stage);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
            (void)b_tile_bytes;
        }
    } else {
        // ================= epilogue =================
        const int ew = warp - (1 + kConvWarps);   // 0..7
        const int quarter = warp & 3;             // TMEM lane quarter this warp may access
        float4* stage0 = reinterpret_cast<float4*>(epi_base + ew * kEpiStageBytes);
        EpilogueArgs ea;
#pragma unroll
        for (int l = 0; l < TCS_MAX_LEVELS; ++l) ea.lvl[l] = p.lvl[l];
        ea.W1 = p.W1; ea.W2 = p.W2; ea.num_levels = p.num_levels; ea.scale = p.out_scale;
        int iter = 0;
        for (int row = blockIdx.x; row < p.num_rows; row += gridDim.x) {
            for (int m_t = 0; m_t < p.num_m; ++m_t, ++iter) {
                const uint32_t acc = iter & 1;
                const uint32_t acc_phase = (iter >> 1) & 1;
                ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
                ptx::tc_fence_after_sync();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kAccCols;
                epilogue_tile<true>(ea, taddr, 0, p.W2, m_t * kBlockM + quarter * 32, (size_t)row * p.W1, ew >> 2, lane,
                                    stage0, stage0, bar_tempty + 8 * acc);
            }
        }
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace fused
}  // namespace tcs

extern "C" int tcs_corr_build_fused(const float* fmap1, const float* fmap2,
                                    float* lvl0, float* lvl1, float* lvl2, float* lvl3,
                                    int B, int H, int W1, int W2, int C, int num_levels, int prec, void* stream) {
    using namespace tcs;
    using namespace tcs::fused;
    TCS_REQUIRE(fmap1 != nullptr && fmap2 != nullptr && lvl0 != nullptr, TCS_E_BADARG, "tcs_corr_build_fused: null pointer");
    TCS_REQUIRE(prec >= TCS_PREC_BF16 && prec <= TCS_PREC_FP16X3, TCS_E_BADARG, "tcs_corr_build_fused: bad prec %d", prec);
    TCS_REQUIRE(num_levels >= 1 && num_levels <= TCS_MAX_LEVELS, TCS_E_SHAPE, "tcs_corr_build_fused: num_levels=%d not in [1,4]", num_levels);
    float* lv[4] = {lvl0, lvl1, lvl2, lvl3};
    for (int l = 0; l < num_levels; ++l)
        TCS_REQUIRE(lv[l] != nullptr && aligned16(lv[l]), TCS_E_ALIGN, "tcs_corr_build_fused: level %d pointer null or not 16-byte aligned", l);
    TCS_REQUIRE(B > 0 && H > 0 && W1 > 0 && C > 0, TCS_E_BADARG, "tcs_corr_build_fused: bad sizes");
    TCS_REQUIRE(W2 >= 8 && W2 <= kMaxN && W1 <= kMaxW1, TCS_E_SHAPE,
                "tcs_corr_build_fused: needs 8 <= W2 <= %d and W1 <= %d (got W1=%d W2=%d); use tcs_corr_prepass + tcs_corr_build", kMaxN, kMaxW1, W1, W2);
    TCS_REQUIRE((W2 >> (num_levels - 1)) >= 1, TCS_E_SHAPE, "tcs_corr_build_fused: W2 too small for %d levels", num_levels);
    TCS_REQUIRE(C % (kBlockK * 1) == 0 && C % (kNormSlices * 32) == 0, TCS_E_SHAPE, "tcs_corr_build_fused: C=%d must be a multiple of 64", C);
    const bool x3 = (prec == TCS_PREC_BF16X3 || prec == TCS_PREC_FP16X3);
    const bool fp16 = (prec == TCS_PREC_FP16 || prec == TCS_PREC_FP16X3);

    Params p{};
    p.fmap1 = fmap1; p.fmap2 = fmap2;
    for (int l = 0; l < 4; ++l) p.lvl[l] = lv[l];
    p.H = H; p.W1 = W1; p.W2 = W2; p.C = C; p.num_levels = num_levels;
    p.num_rows = B * H;
    p.num_m = ceil_div(W1, kBlockM);
    p.block_n = ceil_div(W2, 16) * 16;
    p.kblocks = C / kBlockK;
    p.passes = x3 ? 3 : 1;
    p.fp16 = fp16 ? 1 : 0;
    p.idesc = ptx::make_idesc_f16(fp16 ? 0u : 1u, kBlockM, (uint32_t)p.block_n);
    p.in_scale = fp16 ? 256.0f : 1.0f;
    p.out_scale = fp16 ? (1.0f / 65536.0f) : 1.0f;

    static bool attr_done[2] = {false, false};
    if (!attr_done[fp16 ? 1 : 0]) {
        if (fp16) TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_build_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        else TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_build_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_done[fp16 ? 1 : 0] = true;
    }
    const int grid = p.num_rows < num_sms() ? p.num_rows : num_sms();
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (fp16) corr_build_fused_kernel<true><<<grid, kThreads, kSmemBytes, s>>>(p);
    else corr_build_fused_kernel<false><<<grid, kThreads, kSmemBytes, s>>>(p);
    TCS_CHECK_LAUNCH("tcs_corr_build_fused");
    return 0;
}
