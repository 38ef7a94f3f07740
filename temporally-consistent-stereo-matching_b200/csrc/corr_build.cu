// Correlation build: per (b,h) image row, D[w1,w2] = sum_c n1[b,h,w1,c] * n2[b,h,w2,c] on the
// 5th-generation tensor cores, with every pyramid level written from the epilogue in one pass.
// ref: core/corr.py:54-62 (CorrBlock1D.corr, einsum 'aijk,aijh->ajkh') and core/corr.py:15-23 (pyramid).
//
// Structure (one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor.3d of [128 x 64] (A) and [block_n x 64] (B) 16-bit
//               K-major tiles, SWIZZLE_128B, into a 3-stage shared-memory ring (mbarrier complete_tx).
//   warp 1      tcgen05.mma issuer (one thread): UMMA 128 x block_n x 16, fp32 accumulators in TMEM,
//               two accumulator stages (2 x 256 columns) so the MMAs of tile i+1 overlap the epilogue of
//               tile i; tcgen05.commit releases smem stages and publishes finished accumulators.
//   warps 2..9  epilogue (two warps per TMEM lane quarter, taking even / odd column chunks):
//               tcgen05.ld 32 lanes x 32 columns -> registers (thread = one w1 row, so the
//               avg-pool cascade along w2 is purely intra-thread), swizzled smem transpose for L0/L1,
//               coalesced 16-byte global stores of levels 0..3.
// The *X3 precisions run three MMA passes per K block (hi*hi, hi*lo, lo*hi) into the same accumulator.
//
// Roofline: HBM-write-bound (fp32 levels: 7.5 B per 512 FLOP); see DESIGN.md.
#include "tcs_common.cuh"
#include "sm100_ptx.cuh"
#include "corr_epilogue.cuh"
#include "tma_host.cuh"

namespace tcs {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;   // 64 x 16 bit = 128 B: one SWIZZLE_128B row
constexpr int kUmmaK = 16;
constexpr int kMaxBlockN = 256;
constexpr int kStages = 3;
constexpr int kABytes = kBlockM * kBlockK * 2;     // 16 KB
constexpr int kBBytes = kMaxBlockN * kBlockK * 2;  // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;     // 48 KB
constexpr int kAccStages = 2;
constexpr int kAccCols = 256;
constexpr int kTmemCols = kAccStages * kAccCols;   // 512: the whole TMEM of the SM (1 CTA / SM)
constexpr int kEpiWarps = 8;                      // two per TMEM lane quarter: even / odd 32-column chunks
constexpr int kBuildThreads = 64 + 32 * kEpiWarps; // 320
constexpr int kEpiStageBytes = 4096 + 2048;        // per epilogue warp: L0 [32][32] + L1 [32][16] fp32
constexpr int kBarrierBytes = 256;
constexpr int kBuildSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + kEpiWarps * kEpiStageBytes + kBarrierBytes;

struct BuildParams {
    float* lvl[TCS_MAX_LEVELS];
    int W1, W2, num_levels;
    int num_m, n_tiles, block_n;
    int total_tiles;
    int kblocks;   // C / 64
    int passes;    // 1 or 3
    uint32_t idesc;
    float scale;
};

__global__ void __launch_bounds__(kBuildThreads, 1)
corr_build_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                  const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                  const BuildParams p) {
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1 KB alignment.
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* epi_base = smem + kStages * kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + kEpiWarps * kEpiStageBytes);
    // barrier slots: full[kStages], empty[kStages], tmem_full[kAccStages], tmem_empty[kAccStages]
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = bar_full + 8 * kStages;
    const uint32_t bar_tfull = bar_empty + 8 * kStages;
    const uint32_t bar_tempty = bar_tfull + 8 * kAccStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAccStages);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tm_a_hi);
        ptx::prefetch_tensormap(&tm_b_hi);
        if (p.passes == 3) {
            ptx::prefetch_tensormap(&tm_a_lo);
            ptx::prefetch_tensormap(&tm_b_lo);
        }
        for (int i = 0; i < kStages; ++i) {
            ptx::mbar_init(bar_full + 8 * i, 1);
            ptx::mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < kAccStages; ++i) {
            ptx::mbar_init(bar_tfull + 8 * i, 1);
            ptx::mbar_init(bar_tempty + 8 * i, kEpiWarps * 32);
        }
        ptx::fence_barrier_init();
    } else if (warp == 1) {
        ptx::tmem_alloc(smem_u32(tmem_slot), kTmemCols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int steps_per_tile = p.kblocks * p.passes;
    const uint32_t stage_tx_bytes = kABytes + p.block_n * (kBlockK * 2);

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const int n_t = tile % p.n_tiles;
                const int m_t = (tile / p.n_tiles) % p.num_m;
                const int bh = tile / (p.n_tiles * p.num_m);
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    for (int pass = 0; pass < p.passes; ++pass) {
                        ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                        const uint32_t sb = sa + kABytes;
                        const uint32_t full = bar_full + 8 * stage;
                        ptx::mbar_arrive_expect_tx(full, stage_tx_bytes);
                        ptx::tma_load_3d(sa, pass == 2 ? &tm_a_lo : &tm_a_hi, full, kb * kBlockK, m_t * kBlockM, bh);
                        ptx::tma_load_3d(sb, pass == 1 ? &tm_b_lo : &tm_b_hi, full, kb * kBlockK, n_t * p.block_n, bh);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            int iter = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++iter) {
                const uint32_t acc = iter & 1;
                const uint32_t acc_phase = (iter >> 1) & 1;
                ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                ptx::tc_fence_after_sync();
                const uint32_t tmem_d = tmem_base + acc * kAccCols;
                for (int s = 0; s < steps_per_tile; ++s) {
                    ptx::mbar_wait(bar_full + 8 * stage, phase);
                    ptx::tc_fence_after_sync();
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes);
                    const uint64_t da = ptx::make_kmajor_sw128_desc(sa);
                    const uint64_t db = ptx::make_kmajor_sw128_desc(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                        // advance 16 elements = 32 B along K inside the swizzle atom: +2 in 16-byte units
                        ptx::umma_f16(tmem_d, da + 2 * k, db + 2 * k, p.idesc, (s | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(bar_empty + 8 * stage);  // smem stage reusable once these MMAs retire
                    if (s == steps_per_tile - 1) ptx::umma_commit(bar_tfull + 8 * acc);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ================= epilogue =================
        const int ew = warp - 2;            // staging buffer index
        const int quarter = warp & 3;       // TMEM lane quarter this warp may access
        const uint32_t stage0 = smem_u32(epi_base + ew * kEpiStageBytes);   // [32 rows][8 float4]
        const uint32_t stage1 = stage0 + 4096;                              // [32 rows][4 float4]
        EpilogueArgs ea;
#pragma unroll
        for (int l = 0; l < TCS_MAX_LEVELS; ++l) ea.lvl[l] = p.lvl[l];
        ea.W1 = p.W1; ea.W2 = p.W2; ea.num_levels = p.num_levels; ea.scale = p.scale;
        int iter = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++iter) {
            const int n_t = tile % p.n_tiles;
            const int m_t = (tile / p.n_tiles) % p.num_m;
            const int bh = tile / (p.n_tiles * p.num_m);
            const uint32_t acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const int n0 = n_t * p.block_n;
            const int n_end = min(n0 + p.block_n, p.W2);   // exclusive column limit of this tile at level 0
            ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
            ptx::tc_fence_after_sync();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kAccCols;
            epilogue_tile<false>(ea, taddr, n0, n_end, m_t * kBlockM + quarter * 32, (size_t)bh * p.W1, ew >> 2, lane,
                                 stage0, stage1, bar_tempty + 8 * acc);
        }
    }

    // ---- teardown: everyone done with TMEM before it is released
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- host side ------------------------------------------------------------------------------------

// Operand [BH, W, C] 16-bit, channels contiguous; box = [1, box_w, 64], 128 B swizzle.
static int make_operand_map(CUtensorMap* tm, const void* base, int BH, int W, int C, int box_w, bool fp16) {
    EncodeTiledFn enc = get_encode_fn();
    TCS_REQUIRE(enc != nullptr, TCS_E_DRIVER, "tcs_corr_build: cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)BH};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2};
    cuuint32_t box[3] = {(cuuint32_t)kBlockK, (cuuint32_t)box_w, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TCS_REQUIRE(r == CUDA_SUCCESS, TCS_E_DRIVER, "tcs_corr_build: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return 0;
}

}  // namespace tcs

extern "C" int tcs_corr_build(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo,
                              float* lvl0, float* lvl1, float* lvl2, float* lvl3,
                              int B, int H, int W1, int W2, int C, int num_levels, int prec, int W2_pitch, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(a_hi != nullptr && b_hi != nullptr && lvl0 != nullptr, TCS_E_BADARG, "tcs_corr_build: null operand / level 0");
    // Row pitch of level 0 (level l: pitch >> l).  With a pitch beyond W2 the right operand's rows past W2 are TMA zero fill, so
    // the extra columns of every level come out as exact zeros (W2 % 8 == 0: no pooled entry mixes real and padding columns).
    const int W2_valid = W2;
    if (W2_pitch > 0 && W2_pitch != W2) {
        TCS_REQUIRE(W2_pitch > W2 && W2_pitch % 16 == 0 && W2 % 8 == 0, TCS_E_SHAPE,
                    "tcs_corr_build: a row pitch (%d) other than W2 (%d) needs W2 %% 8 == 0 and a pitch that is a multiple of 16", W2_pitch, W2);
        W2 = W2_pitch;
    }
    TCS_REQUIRE(prec >= TCS_PREC_BF16 && prec <= TCS_PREC_FP16X3, TCS_E_BADARG, "tcs_corr_build: bad prec %d", prec);
    const bool x3 = (prec == TCS_PREC_BF16X3 || prec == TCS_PREC_FP16X3);
    const bool fp16 = (prec == TCS_PREC_FP16 || prec == TCS_PREC_FP16X3);
    TCS_REQUIRE(!x3 || (a_lo != nullptr && b_lo != nullptr), TCS_E_BADARG, "tcs_corr_build: the X3 modes need the lo operands");
    TCS_REQUIRE(num_levels >= 1 && num_levels <= TCS_MAX_LEVELS, TCS_E_SHAPE, "tcs_corr_build: num_levels=%d not in [1,4]", num_levels);
    float* lv[4] = {lvl0, lvl1, lvl2, lvl3};
    for (int l = 0; l < num_levels; ++l)
        TCS_REQUIRE((lv[l] != nullptr || (l & 1)) && aligned16(lv[l]), TCS_E_ALIGN,
                    "tcs_corr_build: level %d pointer null or not 16-byte aligned (only the odd levels may be omitted)", l);
    TCS_REQUIRE(B > 0 && H > 0 && W1 > 0 && W2 >= 8 && C > 0, TCS_E_BADARG, "tcs_corr_build: bad sizes");
    TCS_REQUIRE((W2 >> (num_levels - 1)) >= 1, TCS_E_SHAPE, "tcs_corr_build: W2 too small for %d levels", num_levels);
    TCS_REQUIRE(C % kBlockK == 0, TCS_E_SHAPE, "tcs_corr_build: C=%d must be a multiple of 64", C);
    TCS_REQUIRE(aligned16(a_hi) && aligned16(a_lo) && aligned16(b_hi) && aligned16(b_lo), TCS_E_ALIGN,
                "tcs_corr_build: operands must be 16-byte aligned");
    TCS_REQUIRE((long long)B * H <= 0x7fffffffLL / 1024, TCS_E_SHAPE, "tcs_corr_build: B*H too large");

    BuildParams p{};
    for (int l = 0; l < 4; ++l) p.lvl[l] = lv[l];
    p.W1 = W1; p.W2 = W2; p.num_levels = num_levels;
    p.num_m = ceil_div(W1, kBlockM);
    p.n_tiles = ceil_div(W2, kMaxBlockN);
    // UMMA N must be a multiple of 16; with several N tiles every tile must also start on a 32-column
    // boundary so that the level-3 float4 stores (one per 32 level-0 columns) stay 16-byte aligned.
    p.block_n = ceil_div(ceil_div(W2, p.n_tiles), p.n_tiles > 1 ? 32 : 16) * (p.n_tiles > 1 ? 32 : 16);
    const long long total = (long long)B * H * p.num_m * p.n_tiles;
    TCS_REQUIRE(total < 0x7fffffffLL, TCS_E_SHAPE, "tcs_corr_build: too many tiles");
    p.total_tiles = (int)total;
    p.kblocks = C / kBlockK;
    p.passes = x3 ? 3 : 1;
    p.idesc = ptx::make_idesc_f16(fp16 ? 0u : 1u, kBlockM, (uint32_t)p.block_n);
    p.scale = fp16 ? (1.0f / 65536.0f) : 1.0f;  // undo the 2^8 operand scaling of both sides

    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    int rc;
    if ((rc = make_operand_map(&ta_hi, a_hi, B * H, W1, C, kBlockM, fp16)) != 0) return rc;
    if ((rc = make_operand_map(&tb_hi, b_hi, B * H, W2_valid, C, p.block_n, fp16)) != 0) return rc;
    if (x3) {
        if ((rc = make_operand_map(&ta_lo, a_lo, B * H, W1, C, kBlockM, fp16)) != 0) return rc;
        if ((rc = make_operand_map(&tb_lo, b_lo, B * H, W2_valid, C, p.block_n, fp16)) != 0) return rc;
    } else {
        ta_lo = ta_hi;
        tb_lo = tb_hi;
    }

    TCS_ONCE_PER_DEVICE(
        TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBuildSmemBytes));
    );
    const int grid = (int)((total < (long long)num_sms()) ? total : (long long)num_sms());
    corr_build_kernel<<<grid, kBuildThreads, kBuildSmemBytes, static_cast<cudaStream_t>(stream)>>>(ta_hi, ta_lo, tb_hi, tb_lo, p);
    TCS_CHECK_LAUNCH("tcs_corr_build");
    return 0;
}
