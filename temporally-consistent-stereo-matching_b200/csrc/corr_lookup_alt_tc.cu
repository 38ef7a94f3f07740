// Alternate (on-the-fly) lookup on the tensor cores: the same result as tcs_corr_lookup without a pyramid in HBM.
// New capability (the reference has no such path); contract = CorrBlock1D.__call__, ref: core/corr.py:33-52, sampling
// arithmetic of core/utils/utils.py:82-97.
//
// Per tile = 128 consecutive w1 of one (b,h) image row, the part of the row's correlation block that the tile's
// coordinates can touch — the BAND of level-0 columns [8(min f3 - 5), 8(max f3 + 7)), f3 = floor(coords / 8), i.e.
// the union of the rows' level-3 windows, which contain the windows of every finer level — is built into TMEM exactly
// as the correlation build does it (TMA-fed K-major 16-bit operands, tcgen05.mma kind::f16, fp32 accumulators, the
// hi/lo split precisions as three passes), and the 4 x 9 taps are sampled in the epilogue.  The volume never exists:
//   warp 0      TMA producer: per 64-channel K block the [128 x 64] A and [256 | 64 x 64] B boxes (SWIZZLE_128B, contiguous
//               in the K-block-major operands) of the hi part, then of the lo part, each into one slot of a 4-slot ring
//   warp 1      MMA issuer: UMMA 128 x N x 16 with N = the band width (multiple of 16, <= 256 per chunk; a band wider
//               than 256 columns is covered by several chunks whose tap contributions add), two accumulator stages
//   warps 2..9  epilogue, two per TMEM lane quarter (even / odd 32-column blocks; afterwards two levels' taps each):
//               thread = one pixel.  tcgen05.ld 32 columns at a time (only the
//               blocks the warp's own pixels can touch), avg-pool cascade (a+b)*0.5 along w2 in registers (the
//               expression of corr.py:21-23 and of the build epilogue, so the levels are bit-identical to the pyramid's),
//               each value is dropped into the pixel's private 12-entry window of its level in shared memory when it
//               falls inside, and after the last chunk the 36 taps are interpolated from the windows with grid_sample's
//               exact normalise / un-normalise round trip and stored as full 128-byte lines of each tap plane.
// All three roles derive the band from the coordinates with the same warp-collective function, so they agree on the
// number of chunks without communicating.
//
// Roofline: per call the operands stream once from HBM (4 x B*H*W*C*2 bytes in the x3 modes) and the MMAs are
// 2*B*H*W1*band*C flops per pass; at 1080p (2 sequences) 535 MB and 3 x ~36 GFLOP: both near 80-90 us.
#include "tcs_common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.cuh"

#include <climits>
#include <cstdlib>

namespace tcs {
namespace alt {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                          // 64 x 16 bit = 128 B: one SWIZZLE_128B row = one L2 line
constexpr int kUmmaK = 16;
constexpr int kMaxN = 256;
constexpr int kSmallN = 64;
constexpr int kStages = 4;
constexpr int kABytes = kBlockM * kBlockK * 2;       // 16 KB
constexpr int kBBytes = kMaxN * kBlockK * 2;         // 32 KB
// One ring slot = one K block of the A and B tiles of ONE operand part (hi or lo).  The split precisions take two
// consecutive slots per K block (hi, then lo) and run hi*hi, hi*lo, lo*hi from them, so every part is fetched once
// (a stage per pass fetches the hi parts twice: a third more L2 -> shared traffic).  The operands are K-BLOCK-MAJOR,
// [B,H,C/64,W,64] (tcs_corr_prepass_kblocked), so that each box is one contiguous run of whole 128-byte lines; boxes
// cut out of pixel-major [B,H,W,C] rows (128 B out of every 512) left this kernel waiting on TMA at 27 % tensor-pipe
// activity with DRAM at 43 %.
constexpr int kStageBytes = kABytes + kBBytes;       // 48 KB
constexpr int kAccStages = 2;
constexpr int kAccCols = 256;
constexpr int kTmemCols = kAccStages * kAccCols;     // 512
constexpr int kEpiWarps = 8;                         // two per TMEM lane quarter: even / odd 32-column blocks, then two levels each
constexpr int kThreads = 64 + 32 * kEpiWarps;        // 320
constexpr int kWinEntries = 12;                      // per level: entries f_l - 5 .. f_l + 6
constexpr int kWinBytes = 4 * kWinEntries * 32 * 4;  // per lane quarter: [4 levels][12][32 pixels] fp32 = 6 KB
constexpr int kBarrierBytes = 256;
constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 4 * kWinBytes + kBarrierBytes;

struct Params {
    const float* coords;
    long long coords_bstride;
    float* out;
    int H, W1, W2, HW;
    int num_m, total_tiles;
    uint32_t num_m_mul, num_m_shr, H_mul, H_shr;   // tile -> (bh, m_t) and bh -> (b, h) without integer division
    int kblocks, passes;
    uint32_t ab_format;   // 0 fp16, 1 bf16
    float scale;
    int debug;            // development aid (TCS_ALT_DEBUG): 1 = epilogue only waits and releases, 2 = no MMAs are issued
};

// n / d for 0 <= n < 2^31 as one multiply-high and a shift (Granlund & Montgomery; constants from fast_divisor()).
__device__ __forceinline__ int fast_div(int n, int d, uint32_t mul, uint32_t shr) {
    return d == 1 ? n : (int)(__umulhi((uint32_t)n, mul) >> shr);
}

struct Band {
    int lo;        // first level-0 column (multiple of 8, >= 0)
    int hi;        // exclusive end (multiple of 8)
    int nchunks;   // 0 when no pixel of the tile touches the row
};

__device__ __forceinline__ float sane_coord(float c) { return (fabsf(c) <= 1.0e9f) ? c : 1.0e9f; }

// floor(c / 2^l) clamped so that a far-away coordinate keeps a finite, harmless window position
__device__ __forceinline__ int level_floor(float c, int l, int Wl) {
    const float cl = c * (1.0f / (float)(1 << l));
    return (int)fminf(fmaxf(floorf(cl), -16.0f), (float)(Wl + 16));
}

// Level-0 column range a pixel can touch (its level-3 window, which contains the windows of the finer levels), or an
// empty range when every tap is zero padding.  grid_sample's normalise / un-normalise round trip moves a tap by less
// than 2e-4 px on levels up to 1024 wide, so unless the fractional part of coords/8 comes within 2^-9 of 0 or 1 the
// level-3 taps' floors are f3 - 4 .. f3 + 4 and the entries f3 - 4 .. f3 + 5 suffice (80 columns); otherwise one more
// entry on that side may be read (96 columns).  80 instead of 96 is what lets a 128-pixel tile with up to ~40 px of
// disparity spread fit ONE 256-column chunk.
__device__ __forceinline__ void pixel_range(float c, int W2, int& lo, int& hi) {
    const float c3 = c * 0.125f;
    const float fl = floorf(c3), frac = c3 - fl;
    const int f3 = (int)fminf(fmaxf(fl, -16.0f), (float)((W2 >> 3) + 16));
    const bool wide = W2 > 8192;
    lo = 8 * (f3 - ((wide || !(frac >= 1.0f / 512.0f)) ? 5 : 4));
    hi = 8 * (f3 + ((wide || !(frac <= 1.0f - 1.0f / 512.0f)) ? 7 : 6));
    if (hi <= 0 || lo >= W2) { lo = INT_MAX; hi = INT_MIN; }
}

// The coordinates of a tile's 128 pixels, 4 per lane (rows m0 + lane + 32 j; a far-away value for rows past W1).  Every
// role loads them ONE TILE AHEAD (the loads are issued in program order and only waited for when first used), so the
// global-memory latency of the band computation is off every role's per-tile critical path.
struct TileCoords { float c[4]; };
__device__ __forceinline__ void load_tile_coords(const Params& p, int tile, int lane, TileCoords& tc) {
    const int bh = fast_div(tile, p.num_m, p.num_m_mul, p.num_m_shr);
    const int m_t = tile - bh * p.num_m;
    const int b = fast_div(bh, p.H, p.H_mul, p.H_shr), h = bh - b * p.H;
    const float* crow = p.coords + (long long)b * p.coords_bstride + (long long)h * p.W1;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int row = m_t * kBlockM + lane + 32 * j;
        tc.c[j] = 1.0e9f;
        if (row < p.W1) tc.c[j] = ldg_ordered_f1(crow + row);
    }
}

// Warp-collective: the band of a tile from its coordinates.
__device__ __forceinline__ Band tile_band(const TileCoords& tc, int W2) {
    int lo = INT_MAX, hi = INT_MIN;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int l, h;
        pixel_range(sane_coord(tc.c[j]), W2, l, h);
        lo = min(lo, l);
        hi = max(hi, h);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    Band b;
    if (hi <= lo) { b.lo = 0; b.hi = 0; b.nchunks = 0; return b; }
    b.lo = max(lo, 0);
    b.hi = min(hi, (W2 + 15) & ~15);
    b.nchunks = (b.hi - b.lo + kMaxN - 1) / kMaxN;
    return b;
}
__device__ __forceinline__ int chunk_cols(const Band& b, int k) {       // UMMA N of chunk k: multiple of 16, <= 256
    const int w = min(b.hi - (b.lo + kMaxN * k), kMaxN);
    return (w + 15) & ~15;
}

// grid_sample position of tap xk on a level of width wm1 + 1 (same roundings as corr_lookup.cu's sample_pos_fast).
__device__ __forceinline__ float sample_pos_fast(float xk, float wm1, float rc, float hwm1) {
    const float xg = __fmaf_rn(2.0f, div_by_const(xk, wm1, rc), -1.0f);
    return __fmul_rn(__fadd_rn(xg, 1.0f), hwm1);
}

template <bool kX3>
__global__ void __launch_bounds__(kThreads, 1)
corr_lookup_alt_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                          const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                          const __grid_constant__ CUtensorMap tm_bs_hi, const __grid_constant__ CUtensorMap tm_bs_lo,
                          const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* win_base = smem + kStages * kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(win_base + 4 * kWinBytes);
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
        ptx::prefetch_tensormap(&tm_bs_hi);
        if (kX3) {
            ptx::prefetch_tensormap(&tm_a_lo);
            ptx::prefetch_tensormap(&tm_b_lo);
            ptx::prefetch_tensormap(&tm_bs_lo);
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

    if (warp == 0) {
        // ================= TMA producer =================
        // (single-thread loops: every instruction here is on the critical path of the ring, so the per-slot work is an
        // address add, the barrier handshake and the two bulk-tensor issues)
        uint32_t stage = 0, phase = 0;
        const uint32_t smem0 = smem_u32(smem);
        TileCoords cur, nxt;
        if ((int)blockIdx.x < p.total_tiles) load_tile_coords(p, blockIdx.x, lane, cur);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int bh = fast_div(tile, p.num_m, p.num_m_mul, p.num_m_shr);
            const int m0 = (tile - bh * p.num_m) * kBlockM;
            nxt = cur;
            if (tile + (int)gridDim.x < p.total_tiles) load_tile_coords(p, tile + gridDim.x, lane, nxt);
            const Band bd = tile_band(cur, p.W2);
            cur = nxt;
            if (lane == 0) {
                for (int k = 0; k < bd.nchunks; ++k) {
                    const bool small = chunk_cols(bd, k) <= kSmallN;
                    const uint32_t tx = kABytes + (small ? kSmallN : kMaxN) * (kBlockK * 2);
                    const int col0 = bd.lo + kMaxN * k;
                    const CUtensorMap* tb_hi = small ? &tm_bs_hi : &tm_b_hi;
                    const CUtensorMap* tb_lo = small ? &tm_bs_lo : &tm_b_lo;
                    for (int kb = 0; kb < p.kblocks; ++kb) {
#pragma unroll
                        for (int part = 0; part < (kX3 ? 2 : 1); ++part) {           // hi, then lo
                            ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                            const uint32_t sa = smem0 + stage * kStageBytes;
                            const uint32_t full = bar_full + 8 * stage;
                            if (++stage == kStages) { stage = 0; phase ^= 1; }
                            if (p.debug & 8) { ptx::mbar_arrive(full); continue; }
                            ptx::mbar_arrive_expect_tx(full, tx);
                            ptx::tma_load_4d(sa, part ? &tm_a_lo : &tm_a_hi, full, 0, m0, kb, bh);
                            ptx::tma_load_4d(sa + kABytes, part ? tb_lo : tb_hi, full, 0, col0, kb, bh);
                        }
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        uint32_t stage = 0, phase = 0;
        int iter = 0;
        // descriptor of slot 0's A tile; slot s is + s * (kStageBytes >> 4), its B tile + (kABytes >> 4) more (the start
        // address field holds (addr & 0x3ffff) >> 4 and shared addresses stay below 256 KB, so the adds never carry out)
        const uint64_t desc0 = ptx::make_kmajor_sw128_desc(smem_u32(smem));
        TileCoords cur, nxt;
        if ((int)blockIdx.x < p.total_tiles) load_tile_coords(p, blockIdx.x, lane, cur);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            nxt = cur;
            if (tile + (int)gridDim.x < p.total_tiles) load_tile_coords(p, tile + gridDim.x, lane, nxt);
            const Band bd = tile_band(cur, p.W2);
            cur = nxt;
            if (lane == 0) {
                for (int k = 0; k < bd.nchunks; ++k, ++iter) {
                    const uint32_t acc = iter & 1;
                    const uint32_t acc_phase = (iter >> 1) & 1;
                    const uint32_t idesc = ptx::make_idesc_f16(p.ab_format, kBlockM, (uint32_t)chunk_cols(bd, k));
                    ptx::mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                    ptx::tc_fence_after_sync();
                    const uint32_t tmem_d = tmem_base + acc * kAccCols;
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        // slot of the hi part, and (split precisions) the next slot with the lo part
                        const uint32_t s_hi = stage, ph_hi = phase;
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                        uint32_t s_lo = s_hi, ph_lo = ph_hi;
                        if (kX3) {
                            s_lo = stage; ph_lo = phase;
                            if (++stage == kStages) { stage = 0; phase ^= 1; }
                        }
                        const uint64_t a_hi = desc0 + (uint64_t)(s_hi * (kStageBytes >> 4)), b_hi = a_hi + (kABytes >> 4);
                        const uint64_t a_lo = desc0 + (uint64_t)(s_lo * (kStageBytes >> 4)), b_lo = a_lo + (kABytes >> 4);
                        const uint32_t first = kb != 0 ? 1u : 0u;          // the chunk's very first MMA overwrites the accumulator
                        ptx::mbar_wait(bar_full + 8 * s_hi, ph_hi);
                        ptx::tc_fence_after_sync();
                        if (!(p.debug & 2)) {                               // hi*hi, hi*lo, lo*hi: the order of tcs_corr_build
                            ptx::umma_f16(tmem_d, a_hi, b_hi, idesc, first);
#pragma unroll
                            for (int kk = 1; kk < kBlockK / kUmmaK; ++kk) ptx::umma_f16(tmem_d, a_hi + 2 * kk, b_hi + 2 * kk, idesc, 1u);
                        }
                        if (kX3) {
                            ptx::mbar_wait(bar_full + 8 * s_lo, ph_lo);
                            ptx::tc_fence_after_sync();
                            if (!(p.debug & 2)) {
#pragma unroll
                                for (int kk = 0; kk < kBlockK / kUmmaK; ++kk) ptx::umma_f16(tmem_d, a_hi + 2 * kk, b_lo + 2 * kk, idesc, 1u);
#pragma unroll
                                for (int kk = 0; kk < kBlockK / kUmmaK; ++kk) ptx::umma_f16(tmem_d, a_lo + 2 * kk, b_hi + 2 * kk, idesc, 1u);
                            }
                        }
                        ptx::umma_commit(bar_empty + 8 * s_hi);
                        if (kX3) ptx::umma_commit(bar_empty + 8 * s_lo);
                        if (kb == p.kblocks - 1) ptx::umma_commit(bar_tfull + 8 * acc);
                    }
                }
            } else {
                iter += bd.nchunks;
            }
            iter = __shfl_sync(0xffffffffu, iter, 0);
        }
    } else {
        // ================= epilogue: thread = pixel =================
        const int quarter = warp & 3;                         // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                     // which of the quarter's two warps: block parity, then level pair
        const uint32_t win = smem_u32(win_base + quarter * kWinBytes) + 4u * lane;   // entry (l, i): win + 128 * (12 l + i)
        const uint32_t pair_bar = 1 + quarter;                // named barrier of the quarter's two warps
        const int W2 = p.W2;
        int iter = 0;
        TileCoords cur, nxt;
        if ((int)blockIdx.x < p.total_tiles) load_tile_coords(p, blockIdx.x, lane, cur);
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int bh = fast_div(tile, p.num_m, p.num_m_mul, p.num_m_shr);
            const int m_t = tile - bh * p.num_m;
            const int b = fast_div(bh, p.H, p.H_mul, p.H_shr), h = bh - b * p.H;
            nxt = cur;
            if (tile + (int)gridDim.x < p.total_tiles) load_tile_coords(p, tile + gridDim.x, lane, nxt);
            const Band bd = tile_band(cur, W2);
            const int row = m_t * kBlockM + quarter * 32 + lane;
            const bool in_row = row < p.W1;
            const float c0 = sane_coord(quarter == 0 ? cur.c[0] : quarter == 1 ? cur.c[1] : quarter == 2 ? cur.c[2] : cur.c[3]);   // 1e9 past W1
            cur = nxt;
            int my_lo, my_hi;
            pixel_range(c0, W2, my_lo, my_hi);
            if (!in_row) { my_lo = INT_MAX; my_hi = INT_MIN; }
            int wfirst[4];                                     // in-row index of window entry 0, per level
#pragma unroll
            for (int l = 0; l < 4; ++l) wfirst[l] = level_floor(c0, l, W2 >> l) - 5;

            for (int k = 0; k < bd.nchunks; ++k, ++iter) {
                const uint32_t acc = iter & 1;
                const uint32_t acc_phase = (iter >> 1) & 1;
                const int col0 = bd.lo + kMaxN * k;
                const int nblk = (chunk_cols(bd, k) + 31) >> 5;
                ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
                ptx::tc_fence_after_sync();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kAccCols;
                for (int blk = half; blk < ((p.debug & 1) ? 0 : nblk); blk += 2) {
                    const int cg = col0 + 32 * blk;           // first level-0 column of the block (multiple of 8)
                    if (!__any_sync(0xffffffffu, my_lo < cg + 32 && my_hi > cg)) continue;   // nobody here touches it
                    float v[32];
                    ptx::tmem_ld_32x32(taddr + 32 * blk, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] *= p.scale;
                    // ---- level 0
                    {
                        const int d0 = cg - wfirst[0];        // window index of v[0]
                        if (__any_sync(0xffffffffu, d0 > -32 && d0 < kWinEntries)) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const unsigned i = (unsigned)(d0 + j);
                                if (i < (unsigned)kWinEntries) asm volatile("st.shared.f32 [%0], %1;" :: "r"(win + 128u * i), "f"(v[j]) : "memory");
                            }
                        }
                    }
                    float l1[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) l1[j] = (v[2 * j] + v[2 * j + 1]) * 0.5f;
                    {
                        const int d1 = (cg >> 1) - wfirst[1];
                        if (__any_sync(0xffffffffu, d1 > -16 && d1 < kWinEntries)) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const unsigned i = (unsigned)(d1 + j);
                                if (i < (unsigned)kWinEntries) asm volatile("st.shared.f32 [%0], %1;" :: "r"(win + 128u * (kWinEntries + i)), "f"(l1[j]) : "memory");
                            }
                        }
                    }
                    float l2[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) l2[j] = (l1[2 * j] + l1[2 * j + 1]) * 0.5f;
                    {
                        const int d2 = (cg >> 2) - wfirst[2];
                        if (__any_sync(0xffffffffu, d2 > -8 && d2 < kWinEntries)) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const unsigned i = (unsigned)(d2 + j);
                                if (i < (unsigned)kWinEntries) asm volatile("st.shared.f32 [%0], %1;" :: "r"(win + 128u * (2 * kWinEntries + i)), "f"(l2[j]) : "memory");
                            }
                        }
                    }
                    {
                        const int d3 = (cg >> 3) - wfirst[3];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const unsigned i = (unsigned)(d3 + j);
                            const float l3 = (l2[2 * j] + l2[2 * j + 1]) * 0.5f;
                            if (i < (unsigned)kWinEntries) asm volatile("st.shared.f32 [%0], %1;" :: "r"(win + 128u * (3 * kWinEntries + i)), "f"(l3) : "memory");
                        }
                    }
                }
                // this warp has read everything it needs from the accumulator: hand it back to the MMA warp
                ptx::tc_fence_before_sync();
                ptx::mbar_arrive(bar_tempty + 8 * acc);
            }

            // ---- the taps from the pixel's windows, filled by both warps of the quarter: this warp takes two levels
            asm volatile("bar.sync %0, 64;" :: "r"(pair_bar) : "memory");
            if (in_row && !(p.debug & 4)) {
                float* o = p.out + ((long long)b * 36 + 18 * half) * p.HW + (long long)h * p.W1 + row;
#pragma unroll
                for (int ll = 0; ll < 2; ++ll) {
                    const int l = 2 * half + ll;
                    const int Wl = W2 >> l;
                    const float wm1 = (float)(Wl - 1), rc = __frcp_rn(wm1), hwm1 = __fmul_rn(0.5f, wm1);
                    const float cl = c0 * (1.0f / (float)(1 << l));            // coords / 2^l (exact), corr.py:43
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const float ix = sample_pos_fast(__fadd_rn((float)(t - 4), cl), wm1, rc, hwm1);
                        const float x0f = floorf(ix);
                        const bool in0 = (x0f >= 0.0f) && (x0f <= wm1);         // x0 inside the level
                        const bool in1 = (x0f >= -1.0f) && (x0f < wm1);         // x0 + 1 inside the level
                        const int x0 = (in0 || in1) ? (int)x0f : 0;
                        const float w_hi = in1 ? __fsub_rn(ix, x0f) : 0.0f;
                        const float w_lo = in0 ? __fsub_rn(__fadd_rn(x0f, 1.0f), ix) : 0.0f;
                        const int i0 = min(max(x0 - wfirst[l], 0), kWinEntries - 2);
                        const float v0 = lds_f32(win + 128u * (l * kWinEntries + i0));
                        const float v1 = lds_f32(win + 128u * (l * kWinEntries + i0 + 1));
                        // a zero weight stands for "outside the level" (zeros padding) or "nothing was built there":
                        // the product must be 0 whatever the slot holds
                        const float a0 = (w_lo != 0.0f) ? __fmul_rn(v0, w_lo) : 0.0f;
                        const float r = (w_hi != 0.0f) ? fmaf(v1, w_hi, a0) : a0;
                        stg_stream_f1(o, r);
                        o += p.HW;
                    }
                }
            }
            asm volatile("bar.sync %0, 64;" :: "r"(pair_bar) : "memory");   // the partner is done reading before the next tile's values land
        }
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// (mul, shr) with n / d == __umulhi(n, mul) >> shr for every 0 <= n < 2^31, d >= 2.
static void fast_divisor(int d, uint32_t* mul, uint32_t* shr) {
    if (d <= 1) { *mul = 0; *shr = 0; return; }
    int lg = 0;
    while ((1LL << lg) < d) ++lg;
    const int pbits = 31 + lg;
    *mul = (uint32_t)(((1ULL << pbits) + (unsigned long long)d - 1) / (unsigned long long)d);
    *shr = (uint32_t)(pbits - 32);
}

// K-block-major operand [BH, C/64, W, 64] 16-bit; box = [1, 1, box_w, 64] (contiguous), 128 B swizzle, zero fill outside.
static int make_operand_map(CUtensorMap* tm, const void* base, int BH, int W, int C, int box_w, bool fp16) {
    EncodeTiledFn enc = get_encode_fn();
    TCS_REQUIRE(enc != nullptr, TCS_E_DRIVER, "tcs_corr_lookup_alt_tc: cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)kBlockK, (cuuint64_t)W, (cuuint64_t)(C / kBlockK), (cuuint64_t)BH};
    cuuint64_t strides[3] = {(cuuint64_t)kBlockK * 2, (cuuint64_t)W * kBlockK * 2, (cuuint64_t)W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)box_w, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TCS_REQUIRE(r == CUDA_SUCCESS, TCS_E_DRIVER, "tcs_corr_lookup_alt_tc: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return 0;
}

}  // namespace alt
}  // namespace tcs

extern "C" int tcs_corr_lookup_alt_tc(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo,
                                      const float* coords, long long coords_bstride, float* out,
                                      int B, int H, int W1, int W2, int C, int prec, void* stream) {
    using namespace tcs;
    using namespace tcs::alt;
    TCS_REQUIRE(a_hi != nullptr && b_hi != nullptr && coords != nullptr && out != nullptr, TCS_E_BADARG,
                "tcs_corr_lookup_alt_tc: null operand / coords / out");
    TCS_REQUIRE(prec >= TCS_PREC_BF16 && prec <= TCS_PREC_FP16X3, TCS_E_BADARG, "tcs_corr_lookup_alt_tc: bad prec %d", prec);
    const bool x3 = (prec == TCS_PREC_BF16X3 || prec == TCS_PREC_FP16X3);
    const bool fp16 = (prec == TCS_PREC_FP16 || prec == TCS_PREC_FP16X3);
    TCS_REQUIRE(!x3 || (a_lo != nullptr && b_lo != nullptr), TCS_E_BADARG, "tcs_corr_lookup_alt_tc: the X3 modes need the lo operands");
    TCS_REQUIRE(B > 0 && H > 0 && W1 > 0 && C > 0, TCS_E_BADARG, "tcs_corr_lookup_alt_tc: bad sizes");
    TCS_REQUIRE(W2 >= 16, TCS_E_SHAPE, "tcs_corr_lookup_alt_tc: W2=%d must be >= 16 (4 levels, each at least 2 wide)", W2);
    TCS_REQUIRE(C % kBlockK == 0, TCS_E_SHAPE, "tcs_corr_lookup_alt_tc: C=%d must be a multiple of 64", C);
    TCS_REQUIRE(aligned16(a_hi) && aligned16(a_lo) && aligned16(b_hi) && aligned16(b_lo), TCS_E_ALIGN,
                "tcs_corr_lookup_alt_tc: operands must be 16-byte aligned");
    TCS_REQUIRE((long long)B * H <= 0x7fffffffLL / 1024, TCS_E_SHAPE, "tcs_corr_lookup_alt_tc: B*H too large");

    Params p{};
    p.coords = coords; p.coords_bstride = coords_bstride; p.out = out;
    p.H = H; p.W1 = W1; p.W2 = W2; p.HW = H * W1;
    p.num_m = ceil_div(W1, kBlockM);
    const long long total = (long long)B * H * p.num_m;
    TCS_REQUIRE(total < 0x7fffffffLL, TCS_E_SHAPE, "tcs_corr_lookup_alt_tc: too many tiles");
    p.total_tiles = (int)total;
    fast_divisor(p.num_m, &p.num_m_mul, &p.num_m_shr);
    fast_divisor(H, &p.H_mul, &p.H_shr);
    p.kblocks = C / kBlockK;
    p.passes = x3 ? 3 : 1;
    p.ab_format = fp16 ? 0u : 1u;
    p.scale = fp16 ? (1.0f / 65536.0f) : 1.0f;
    { const char* e = getenv("TCS_ALT_DEBUG"); p.debug = e ? atoi(e) : 0; }

    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo, tbs_hi, tbs_lo;
    int rc;
    if ((rc = make_operand_map(&ta_hi, a_hi, B * H, W1, C, kBlockM, fp16)) != 0) return rc;
    if ((rc = make_operand_map(&tb_hi, b_hi, B * H, W2, C, kMaxN, fp16)) != 0) return rc;
    if ((rc = make_operand_map(&tbs_hi, b_hi, B * H, W2, C, kSmallN, fp16)) != 0) return rc;
    if (x3) {
        if ((rc = make_operand_map(&ta_lo, a_lo, B * H, W1, C, kBlockM, fp16)) != 0) return rc;
        if ((rc = make_operand_map(&tb_lo, b_lo, B * H, W2, C, kMaxN, fp16)) != 0) return rc;
        if ((rc = make_operand_map(&tbs_lo, b_lo, B * H, W2, C, kSmallN, fp16)) != 0) return rc;
    } else {
        ta_lo = ta_hi; tb_lo = tb_hi; tbs_lo = tbs_hi;
    }
    TCS_ONCE_PER_DEVICE(
        TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_lookup_alt_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_lookup_alt_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    );
    const int grid = (int)((total < (long long)num_sms()) ? total : (long long)num_sms());
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (x3) corr_lookup_alt_tc_kernel<true><<<grid, kThreads, kSmemBytes, s>>>(ta_hi, ta_lo, tb_hi, tb_lo, tbs_hi, tbs_lo, p);
    else corr_lookup_alt_tc_kernel<false><<<grid, kThreads, kSmemBytes, s>>>(ta_hi, ta_lo, tb_hi, tb_lo, tbs_hi, tbs_lo, p);
    TCS_CHECK_LAUNCH("tcs_corr_lookup_alt_tc");
    return 0;
}
