// Fused correlation build: fp32 NCHW feature maps in, every pyramid level out, in ONE kernel — no operand
// round trip through HBM.  ref: core/corr.py:54-62 (F.normalize + einsum) and core/corr.py:15-23 (pyramid).
//
// tcs_corr_prepass + tcs_corr_build move 2.07 GB for an algorithmic 1.0 GB at 540p x 8 sequences (the
// normalised 16-bit operands are written once and read once).  Here the operands never leave the SM:
//
//   warp 0        tcgen05.mma issuer (one thread): 128 x N x 16 UMMAs, fp32 accumulators in TMEM
//                 (2 x 256 columns: tile i+1's MMAs overlap tile i's epilogue).
//   warp 1        TMA producer (one thread): cp.async.bulk.tensor.4d boxes [16 channels][128 | N pixels] of
//                 the RAW fp32 maps into a 4-slot staging ring (mbarrier complete_tx).  Asynchronous bulk
//                 loads keep ~92 KB in flight per SM with no registers, which a handful of converter warps
//                 issuing ordinary loads cannot.
//   warps 2..11   converters (320 threads).  Per (b,h) image row the staged boxes go by twice:
//                 pass 1 accumulates sum x^2 per pixel (each thread owns fixed pixels, channels in order, so
//                 the result is deterministic) and publishes 1/max(||x||, 1e-12);
//                 pass 2 (the second read is an L2 hit) scales, splits into 16-bit hi / lo and stores into
//                 the K-major SWIZZLE_64B tile layout the UMMA descriptors expect, one 32-channel K block
//                 per pipeline stage; fence.proxy.async + mbarrier arrive hand the stage to the MMA thread.
//   warps 12..15  epilogue, shared with corr_build.cu (corr_epilogue.cuh); optional TMA-store variant below.
//
// The sizes above are those of ONE CTA per SM (TCS_FUSED_CTAS=1).  The default is TWO half-size CTAs per SM (one operand stage,
// one accumulator stage, 2 + 2 staging slots, 8 converter warps, 448 threads, 114 KB each; work dealt tile by tile; the second
// half of the grid starts 5 us late): while one CTA of the pair is in its DRAM-bound norm pass the other converts, multiplies and
// stores.  0.255 against 0.277 ms at 540p x 8 and never slower on the shapes that take this kernel.
// A tile is one (b,h) row x 128 w1 x all w2 (W2 <= 240 fits one UMMA N); a one-per-SM CTA walks whole rows so that the
// norms are computed once per row.  x / ||x|| is evaluated as x * (1 / ||x||) here (one rounding more than
// the pre-pass's exact division, far below the 16-bit split that follows).
#include "tcs_common.cuh"
#include "sm100_ptx.cuh"
#include "corr_epilogue.cuh"
#include "tma_host.cuh"
#include <cstdlib>

namespace tcs {
namespace fused {

constexpr int kBlockM = 128;
constexpr int kBlockK = 32;                         // channels per pipeline stage: 64-byte K-major rows
constexpr int kUmmaK = 16;
constexpr int kMaxN = 240;                          // one N tile
constexpr int kMaxW1 = 256;
#ifndef TCS_FUSED_CTAS
#define TCS_FUSED_CTAS 2                            // CTAs per SM: 2 = half-size CTAs whose norm and convert passes interleave on the SM
#endif
constexpr int kCtasPerSm = TCS_FUSED_CTAS;
constexpr int kStages = kCtasPerSm == 2 ? 1 : 2;    // operand tile stages
constexpr int kSlotK = 16;                          // channels per staging slot (two slots feed one operand stage)
#ifndef TCS_FUSED_SLOTS
#define TCS_FUSED_SLOTS (TCS_FUSED_CTAS == 2 ? 2 : 5)
#endif
constexpr int kSlots = TCS_FUSED_SLOTS;             // fp32 staging slots: three in flight while one is converted
constexpr int kATile = kBlockM * kBlockK * 2;       // 8 KB
constexpr int kBTile = kMaxN * kBlockK * 2;         // 15 KB
constexpr int kStageBytes = 2 * kATile + 2 * kBTile;    // A_hi, A_lo, B_hi, B_lo = 46 KB
constexpr int kABox = kSlotK * kBlockM * 4;         // 8 KB of fp32
constexpr int kBBox = kSlotK * kMaxN * 4;           // 15 KB of fp32 (box width = block_n <= 240)
constexpr int kSlotBytes = kABox + kBBox;           // 23 KB
constexpr int kXSlots = kCtasPerSm == 2 ? 2 : 4;    // norm pass only: the (then idle) operand stages serve as extra staging slots
constexpr int kAccStages = kCtasPerSm == 2 ? 1 : 2;
constexpr int kAccCols = 256;
constexpr int kTmemCols = kAccStages * kAccCols;
#ifndef TCS_FUSED_CONV_WARPS
#define TCS_FUSED_CONV_WARPS (TCS_FUSED_CTAS == 2 ? 8 : 10)
#endif
constexpr int kConvWarps = TCS_FUSED_CONV_WARPS;
constexpr int kConvThreads = kConvWarps * 32;       // 320
constexpr int kEpiWarps = 4;
constexpr int kThreads = 32 * (2 + kConvWarps + kEpiWarps);   // 512
constexpr int kEpiStageBytes = 4096;                // per epilogue warp: [32][32] fp32 (the TMA-store variant needs 6144)
constexpr int kInvBytes = (kMaxW1 + 256) * 4;       // 1/norm of the A pixels, then of the B pixels
constexpr int kBarrierBytes = 256;
constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kSlots * kSlotBytes + kEpiWarps * kEpiStageBytes + kInvBytes + kBarrierBytes;
static_assert(kSmemBytes <= 232448 / kCtasPerSm - (kCtasPerSm - 1) * 1024, "fused build: shared memory budget exceeded");
static_assert(kXSlots * kSlotBytes <= kStages * kStageBytes, "extra norm-pass slots live inside the operand stages");
constexpr int kRing1 = kSlots + kXSlots;            // staging ring of the norm pass
static_assert(8 * (2 * kRing1 + 2 * kStages + 2 * kAccStages + 1) + 4 <= kBarrierBytes, "barrier area");
static_assert(2 * kABox <= kSlotBytes, "a slot must hold two A boxes (norm pass of the later M tiles)");
static_assert(kStageBytes % 1024 == 0 && kSlotBytes % 512 == 0 && kATile % 512 == 0 && kBTile % 512 == 0, "tile alignment");

struct Params {
    float* lvl[TCS_MAX_LEVELS];
    int H, W1, W2, C, num_levels;
    int num_rows;      // B * H
    int num_m;         // ceil(W1 / 128)
    int block_n;       // W2 rounded up to 16 (UMMA N and the B box width)
    int kblocks;       // C / 32 (operand stages per tile)
    int passes;        // 1 or 3
    uint32_t idesc;
    float out_scale;   // undoes the operand scaling in the epilogue
    float in_scale;    // 2^8 for fp16 operands (keeps unit-vector entries away from subnormals), 1 for bf16
    int tma_store;     // levels 0 and 1 leave through TMA (needs W2 % 8 == 0)
    int tile_mode;     // work items are single M tiles (unit u = k * grid + cta), not whole rows
    int stagger_ns;    // two CTAs per SM: the second half of the grid starts this much later, so that the pair's passes interleave
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

__device__ __forceinline__ unsigned long long ptx_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void conv_barrier() {   // converters only (named barrier 1)
    asm volatile("bar.sync 1, %0;" :: "n"(kConvThreads) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// One staged fp32 box [16 channels][box_w pixels] (half of a 32-channel operand stage) -> 16-bit hi / lo tiles in K-major SWIZZLE_64B layout: tile row
// r (a pixel) holds its 32 channels in 64 B; rows form 512-byte atoms of 8; the 16-byte chunk j of row r sits at
// chunk position j ^ ((r >> 1) & 3).  A work item = 32 consecutive rows (the lanes) x 8 channels: 8 conflict-free
// shared loads, one 16-byte shared store per tile.
template <bool kFp16>
__device__ __forceinline__ void convert_box(uint32_t box, int box_w, int rows, uint32_t inv, float in_scale, uint32_t tile_hi,
                                            uint32_t tile_lo, bool want_lo, int half, int item0, int item_step, int lane) {
    const int items = ((rows + 31) >> 5) * 2;              // the box holds 16 channels = 2 octets: chunks 2*half + {0,1}
    constexpr int kIlp = 1;                                // items in flight per lane (2 measured slower: 426 vs 402 us)
    for (int it0 = item0; it0 < items; it0 += kIlp * item_step) {
        float x[kIlp][8], sc[kIlp];
        int rr[kIlp], oc[kIlp];
        bool live[kIlp];
#pragma unroll
        for (int u = 0; u < kIlp; ++u) {
            const int it = it0 + u * item_step;
            const int g = it >> 1, oct_local = it & 1;
            rr[u] = g * 32 + lane;                         // row inside the tile == pixel inside the box
            oc[u] = 2 * half + oct_local;
            live[u] = (it < items) && (rr[u] < rows);
            const int r = live[u] ? rr[u] : 0;
            const uint32_t src = box + 4u * ((live[u] ? oct_local * 8 : 0) * box_w + r);
#pragma unroll
            for (int i = 0; i < 8; ++i) x[u][i] = lds_f32(src + 4u * (i * box_w));
            sc[u] = lds_f32(inv + 4u * r);
        }
#pragma unroll
        for (int u = 0; u < kIlp; ++u) {
            if (!live[u]) continue;
            const float s = sc[u] * in_scale;
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) split16<kFp16>(x[u][2 * i] * s, x[u][2 * i + 1] * s, hi[i], lo[i]);
            const int r = rr[u];
            const uint32_t off = (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 64u + (uint32_t)((oc[u] ^ ((r >> 1) & 3)) << 4);
            sts_v4_u32(tile_hi + off, hi[0], hi[1], hi[2], hi[3]);
            if (want_lo) sts_v4_u32(tile_lo + off, lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

// Epilogue of the fused kernel: levels 0 and 1 leave through TMA.  The thread-owns-a-row registers are staged in
// exactly the layouts SWIZZLE_128B ([32 rows][128 B], chunk ^ (row & 7)) and SWIZZLE_64B ([32 rows][64 B],
// chunk ^ ((row >> 1) & 3)) describe, so one elected lane replaces 12 predicated store instructions per lane and
// chunk with two bulk tensor stores; rows beyond W1 and columns beyond W2 are clipped by the tensor maps.
template <int kChunkStride>
__device__ __forceinline__ void epilogue_tile_tma(const EpilogueArgs& p, const CUtensorMap* tm_l0, const CUtensorMap* tm_l1,
                                                  uint32_t taddr, int n_end, int row0, int bh, int parity, int lane,
                                                  uint32_t stage0, uint32_t stage1, uint32_t bar_tempty) {
    const int W1 = p.W1, W2 = p.W2;
    const int W2_2 = W2 >> 2, W2_3 = W2 >> 3;
    const bool vec2 = (W2_2 & 3) == 0, vec3 = (W2_3 & 3) == 0;
    const float scale = p.scale;
    const size_t rbase = (size_t)bh * W1;
    const int n_chunks = (n_end + 31) >> 5;
    const int ch_last = parity + ((n_chunks - 1 - parity) / kChunkStride) * kChunkStride;
    if (parity >= n_chunks) {
        ptx::tc_fence_before_sync();
        ptx::mbar_arrive(bar_tempty);
    }
    for (int ch = parity; ch < n_chunks; ch += kChunkStride) {
        float v[32];
        ptx::tmem_ld_32x32(taddr + ch * 32, v);
        if (ch == ch_last) {
            ptx::tc_fence_before_sync();
            ptx::mbar_arrive(bar_tempty);
        }
        const int cg = ch * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= scale;
        float l1[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) l1[j] = (v[2 * j] + v[2 * j + 1]) * 0.5f;
        // the previous chunk's bulk stores must have finished reading the staging boxes
        if (lane == 0) ptx::tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int s = 0; s < 8; ++s)
            sts_v4_f32(stage0 + 16u * (lane * 8 + (s ^ (lane & 7))), make_float4(v[4 * s], v[4 * s + 1], v[4 * s + 2], v[4 * s + 3]));
        if (p.num_levels > 1) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
                sts_v4_f32(stage1 + 16u * (lane * 4 + (s ^ ((lane >> 1) & 3))), make_float4(l1[4 * s], l1[4 * s + 1], l1[4 * s + 2], l1[4 * s + 3]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && row0 < W1) {
            ptx::tma_store_3d(tm_l0, stage0, cg, row0, bh);
            if (p.num_levels > 1) ptx::tma_store_3d(tm_l1, stage1, cg >> 1, row0, bh);
            ptx::tma_store_commit();
        }
        // ---- levels 2 and 3: 32 B / 16 B per row, written straight from the owning thread
        if (p.num_levels > 2) {
            float l2[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) l2[j] = (l1[2 * j] + l1[2 * j + 1]) * 0.5f;
            const int row = row0 + lane;
            if (row < W1) {
                float* r2 = p.lvl[2] + (rbase + row) * W2_2;
                const int lim2 = min(n_end >> 2, W2_2);
                store4(r2, (cg >> 2), lim2, vec2, make_float4(l2[0], l2[1], l2[2], l2[3]));
                store4(r2, (cg >> 2) + 4, lim2, vec2, make_float4(l2[4], l2[5], l2[6], l2[7]));
                if (p.num_levels > 3 && p.lvl[3] != nullptr) {
                    float* r3 = p.lvl[3] + (rbase + row) * W2_3;
                    const int lim3 = min(n_end >> 3, W2_3);
                    store4(r3, (cg >> 3), lim3, vec3,
                           make_float4((l2[0] + l2[1]) * 0.5f, (l2[2] + l2[3]) * 0.5f,
                                       (l2[4] + l2[5]) * 0.5f, (l2[6] + l2[7]) * 0.5f));
                }
            }
        }
    }
}

// Work list of a CTA: whole image rows (all M tiles, so the row's B norms are computed once), interleaved over the
// grid; when the last, partial wave has no more tiles than there are CTAs it is dealt out tile by tile instead, so
// the tail costs one tile rather than one row.
__device__ __forceinline__ bool work_item(const Params& p, int k, int& row, int& m_lo, int& m_hi) {
    const int grid = (int)gridDim.x, cta = (int)blockIdx.x;
    if (p.tile_mode) {
        const int u = k * grid + cta;
        if (u >= p.num_rows * p.num_m) return false;
        row = u / p.num_m;
        m_lo = u - row * p.num_m; m_hi = m_lo + 1;
        return true;
    }
    const int waves = p.num_rows / grid;
    m_lo = 0; m_hi = p.num_m;
    if (k < waves) { row = k * grid + cta; return true; }
    const int rem = p.num_rows - waves * grid;
    if (k > waves || rem == 0) return false;
    if (rem * p.num_m <= grid) {
        if (cta >= rem * p.num_m) return false;
        row = waves * grid + cta / p.num_m;
        m_lo = cta - (cta / p.num_m) * p.num_m; m_hi = m_lo + 1;
        return true;
    }
    row = waves * grid + cta;
    return cta < rem;
}

template <bool kFp16>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
corr_build_fused_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                        const __grid_constant__ CUtensorMap tm_l0, const __grid_constant__ CUtensorMap tm_l1, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t* slot_base = smem + kStages * kStageBytes;
    uint8_t* epi_base = slot_base + kSlots * kSlotBytes;
    float* inv_a = reinterpret_cast<float*>(epi_base + kEpiWarps * kEpiStageBytes);   // [kMaxW1]
    float* inv_b = inv_a + kMaxW1;                                                     // [256]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(inv_a) + kInvBytes);
    const uint32_t bar_sfull = smem_u32(bars);                    // staging slot filled by TMA
    const uint32_t bar_sempty = bar_sfull + 8 * kRing1;           // staging slot drained by the converters
    const uint32_t bar_full = bar_sempty + 8 * kRing1;            // operand stage written by the converters
    const uint32_t bar_empty = bar_full + 8 * kStages;            // operand stage consumed by the MMAs
    const uint32_t bar_tfull = bar_empty + 8 * kStages;
    const uint32_t bar_tempty = bar_tfull + 8 * kAccStages;
    const uint32_t bar_rowdone = bar_tempty + 8 * kAccStages;     // every MMA of a row has retired: its operand stages are idle
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kRing1 + 2 * kStages + 2 * kAccStages + 1);
    // staging slot j: the dedicated ring, then (norm pass only) the operand stages cut into slot-sized pieces
    const uint32_t slot0_addr = smem_u32(slot_base), stage0_addr = smem_u32(smem);
    auto slot_addr = [&](int j) -> uint32_t {
        return j < kSlots ? slot0_addr + (uint32_t)j * kSlotBytes : stage0_addr + (uint32_t)(j - kSlots) * kSlotBytes;
    };

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < kRing1; ++i) {
                ptx::mbar_init(bar_sfull + 8 * i, 1);
                ptx::mbar_init(bar_sempty + 8 * i, kConvThreads);
            }
            ptx::mbar_init(bar_rowdone, 1);
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
    } else if (warp == 1 && lane == 0) {
        ptx::prefetch_tensormap(&tm_a);
        ptx::prefetch_tensormap(&tm_b);
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int a_box_bytes = kSlotK * kBlockM * 4;
    const int b_box_bytes = kSlotK * p.block_n * 4;

    if (warp == 0) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            int iter = 0;
            int row, m_lo, m_hi;
            for (int k = 0; work_item(p, k, row, m_lo, m_hi); ++k) {
                for (int m_t = m_lo; m_t < m_hi; ++m_t, ++iter) {
                    const uint32_t acc = iter % kAccStages;
                    const uint32_t acc_phase = (iter / kAccStages) & 1;
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
                            const uint64_t da = ptx::make_kmajor_sw64_desc(pass == 2 ? sa_lo : sa_hi);
                            const uint64_t db = ptx::make_kmajor_sw64_desc(pass == 1 ? sb_lo : sb_hi);
#pragma unroll
                            for (int k = 0; k < kBlockK / kUmmaK; ++k)
                                ptx::umma_f16(tmem_d, da + 2 * k, db + 2 * k, p.idesc, (kb | pass | k) != 0 ? 1u : 0u);
                        }
                        ptx::umma_commit(bar_empty + 8 * stage);
                        if (kb == p.kblocks - 1) {
                            ptx::umma_commit(bar_tfull + 8 * acc);
                            if (m_t == m_hi - 1) ptx::umma_commit(bar_rowdone);
                        }
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= TMA producer =================
        // Per row: pass 1 (A tile of every M tile, B only with the first), then pass 2 (A tile + B for every M tile).
        if (lane == 0) {
            uint32_t bits = 0;                         // per-slot phase parity (the two passes use rings of different length)
            if (p.stagger_ns > 0 && (int)blockIdx.x >= ((int)gridDim.x + 1) / 2) {
                const unsigned long long t0 = ptx_globaltimer();
                while (ptx_globaltimer() - t0 < (unsigned long long)p.stagger_ns) __nanosleep(200);
            }
            int rows_done = 0;
            int row, m_lo, m_hi;
            for (; work_item(p, rows_done, row, m_lo, m_hi); ++rows_done) {
                const int b = row / p.H, h = row - b * p.H;
                for (int pass = 0; pass < 2; ++pass) {
                    const int ring = pass == 0 ? kRing1 : kSlots;
                    int j = 0;
                    bool stages_idle = rows_done == 0;
                    for (int m_t = m_lo; m_t < m_hi; ++m_t) {
                        const bool with_b = (pass == 1) || (m_t == m_lo);
                        // a slot without a B box has room for two A boxes: fewer, fuller round trips in the norm pass
                        const int per_slot = with_b ? 1 : 2;
                        for (int kb = 0; kb < 2 * p.kblocks; kb += per_slot) {      // 16-channel boxes
                            if (j >= kSlots && !stages_idle) {   // the previous row's MMAs still read the operand stages
                                ptx::mbar_wait(bar_rowdone, (rows_done - 1) & 1);
                                stages_idle = true;
                            }
                            ptx::mbar_wait(bar_sempty + 8 * j, ((bits >> j) & 1) ^ 1);
                            const uint32_t dst = slot_addr(j);
                            const uint32_t full = bar_sfull + 8 * j;
                            ptx::mbar_arrive_expect_tx(full, with_b ? a_box_bytes + b_box_bytes : 2 * a_box_bytes);
                            ptx::tma_load_4d(dst, &tm_a, full, m_t * kBlockM, h, kb * kSlotK, b);
                            if (with_b) ptx::tma_load_4d(dst + kABox, &tm_b, full, 0, h, kb * kSlotK, b);
                            else ptx::tma_load_4d(dst + kABox, &tm_a, full, m_t * kBlockM, h, (kb + 1) * kSlotK, b);
                            bits ^= 1u << j;
                            if (++j == ring) j = 0;
                        }
                    }
                }
            }
        }
    } else if (warp < 2 + kConvWarps) {
        // ================= converters =================
        const int cw = warp - 2;                      // 0..9
        const int ct = cw * 32 + lane;                // 0..319
        const bool want_lo = p.passes == 3;
        const int bw = p.block_n;                     // B box width (pixels)
        uint32_t bits = 0;                            // per-slot phase parity of the staging slots (as in the producer)
        uint32_t stage = 0, phase = 0;                // operand stages
        int row, m_lo, m_hi;
        for (int k = 0; work_item(p, k, row, m_lo, m_hi); ++k) {
            // ---- pass 1: sum of squares per pixel; a thread owns fixed pixels and adds channels in order
            float ss_b0 = 0.0f, ss_b1 = 0.0f;
            int j = 0;
            for (int m_t = m_lo; m_t < m_hi; ++m_t) {
                float ss_a = 0.0f;
                const int per_slot = m_t == m_lo ? 1 : 2;    // M tiles after the first come two A boxes to a slot
                for (int kb = 0; kb < 2 * p.kblocks; kb += per_slot) {
                    ptx::mbar_wait(bar_sfull + 8 * j, (bits >> j) & 1);
                    const uint32_t abox = slot_addr(j);
                    const uint32_t bbox = abox + kABox;
                    if (ct < kBlockM) {
#pragma unroll
                        for (int c = 0; c < kSlotK; ++c) { const float v = lds_f32(abox + 4u * (c * kBlockM + ct)); ss_a = fmaf(v, v, ss_a); }
                        if (m_t != m_lo) {
#pragma unroll
                            for (int c = kSlotK; c < 2 * kSlotK; ++c) { const float v = lds_f32(abox + 4u * (c * kBlockM + ct)); ss_a = fmaf(v, v, ss_a); }
                        }
                    }
                    if (m_t == m_lo) {
                        if (ct < bw) {
#pragma unroll
                            for (int c = 0; c < kSlotK; ++c) { const float v = lds_f32(bbox + 4u * (c * bw + ct)); ss_b0 = fmaf(v, v, ss_b0); }
                        }
                        if (ct + kConvThreads < bw) {
#pragma unroll
                            for (int c = 0; c < kSlotK; ++c) { const float v = lds_f32(bbox + 4u * (c * bw + ct + kConvThreads)); ss_b1 = fmaf(v, v, ss_b1); }
                        }
                    }
                    ptx::mbar_arrive(bar_sempty + 8 * j);
                    bits ^= 1u << j;
                    if (++j == kRing1) j = 0;
                }
                if (ct < kBlockM) inv_a[m_t * kBlockM + ct] = __frcp_rn(fmaxf(sqrtf(ss_a), 1e-12f));   // corr.py:58-59
            }
            if (ct < bw) inv_b[ct] = __frcp_rn(fmaxf(sqrtf(ss_b0), 1e-12f));
            if (ct + kConvThreads < bw) inv_b[ct + kConvThreads] = __frcp_rn(fmaxf(sqrtf(ss_b1), 1e-12f));
            conv_barrier();   // also: every converter is done reading the norm-pass slots that live in the operand stages
            j = 0;
            // ---- pass 2: one 32-channel K block per stage: the A tile of this M tile and the whole B row
            for (int m_t = m_lo; m_t < m_hi; ++m_t) {
                for (int kb = 0; kb < p.kblocks; ++kb) {
                    ptx::mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t sa_hi = smem_u32(smem + stage * kStageBytes);
                    const uint32_t sa_lo = sa_hi + kATile;
                    const uint32_t sb_hi = sa_lo + kATile;
                    const uint32_t sb_lo = sb_hi + kBTile;
                    for (int half = 0; half < 2; ++half) {       // two 16-channel slots fill one 32-channel stage
                        ptx::mbar_wait(bar_sfull + 8 * j, (bits >> j) & 1);
                        const uint32_t abox = slot_addr(j);
                        const uint32_t bbox = abox + kABox;
                        {
                            // A: 4 row groups x 2 octets = 8 items; B: up to 8 x 2 = 16 items; 24 items over 6 warps
                            convert_box<kFp16>(abox, kBlockM, kBlockM, smem_u32(inv_a + m_t * kBlockM), p.in_scale, sa_hi, sa_lo, want_lo, half, cw, kConvWarps, lane);
                            convert_box<kFp16>(bbox, bw, bw, smem_u32(inv_b), p.in_scale, sb_hi, sb_lo, want_lo, half, (cw + 8) % kConvWarps, kConvWarps, lane);
                        }
                        ptx::mbar_arrive(bar_sempty + 8 * j);
                        bits ^= 1u << j;
                        if (++j == kSlots) j = 0;
                    }
                    fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core's async proxy
                    ptx::mbar_arrive(bar_full + 8 * stage);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
            conv_barrier();   // inv_a / inv_b are rewritten by the next row's pass 1
        }
    } else {
        // ================= epilogue =================
        const int ew = warp - (2 + kConvWarps);   // 0..7
        const int quarter = warp & 3;             // TMEM lane quarter this warp may access
        const uint32_t stage0 = smem_u32(epi_base + ew * kEpiStageBytes);
        EpilogueArgs ea;
#pragma unroll
        for (int l = 0; l < TCS_MAX_LEVELS; ++l) ea.lvl[l] = p.lvl[l];
        ea.W1 = p.W1; ea.W2 = p.W2; ea.num_levels = p.num_levels; ea.scale = p.out_scale;
        int iter = 0;
        int row, m_lo, m_hi;
        for (int k = 0; work_item(p, k, row, m_lo, m_hi); ++k) {
            for (int m_t = m_lo; m_t < m_hi; ++m_t, ++iter) {
                const uint32_t acc = iter % kAccStages;
                const uint32_t acc_phase = (iter / kAccStages) & 1;
                ptx::mbar_wait(bar_tfull + 8 * acc, acc_phase);
                ptx::tc_fence_after_sync();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kAccCols;
                if (p.tma_store)
                    epilogue_tile_tma<kEpiWarps / 4>(ea, &tm_l0, &tm_l1, taddr, p.W2, m_t * kBlockM + quarter * 32, row, ew >> 2, lane,
                                                     stage0, stage0 + 4096, bar_tempty + 8 * acc);
                else
                    epilogue_tile<true, kEpiWarps / 4>(ea, taddr, 0, p.W2, m_t * kBlockM + quarter * 32, (size_t)row * p.W1, ew >> 2, lane,
                                                       stage0, stage0, bar_tempty + 8 * acc);
            }
        }
        if (p.tma_store && lane == 0) ptx::tma_store_wait_all();   // the bulk stores read shared memory: finish before exit
    }

    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// fp32 feature map [B, C, H, W] as a 4-D tensor (w, h, c, b); box = [box_w pixels, 1 row, 16 channels, 1 sample].
static int make_fmap_map(CUtensorMap* tm, const float* base, int B, int C, int H, int W, int box_w) {
    EncodeTiledFn enc = get_encode_fn();
    TCS_REQUIRE(enc != nullptr, TCS_E_DRIVER, "tcs_corr_build_fused: cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
    cuuint32_t box[4] = {(cuuint32_t)box_w, 1, (cuuint32_t)kSlotK, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TCS_REQUIRE(r == CUDA_SUCCESS, TCS_E_DRIVER, "tcs_corr_build_fused: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return 0;
}

// Pyramid level [BH, W1, Wl] fp32 as a 3-D tensor (wl, w1, bh); box = [box_w columns, 32 rows, 1].
static int make_level_map(CUtensorMap* tm, float* base, int BH, int W1, int Wl, int box_w, CUtensorMapSwizzle swz) {
    EncodeTiledFn enc = get_encode_fn();
    TCS_REQUIRE(enc != nullptr, TCS_E_DRIVER, "tcs_corr_build_fused: cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {(cuuint64_t)Wl, (cuuint64_t)W1, (cuuint64_t)BH};
    cuuint64_t strides[2] = {(cuuint64_t)Wl * 4, (cuuint64_t)W1 * Wl * 4};
    cuuint32_t box[3] = {(cuuint32_t)box_w, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TCS_REQUIRE(r == CUDA_SUCCESS, TCS_E_DRIVER, "tcs_corr_build_fused: cuTensorMapEncodeTiled(level) failed (CUresult %d)", (int)r);
    return 0;
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
        TCS_REQUIRE((lv[l] != nullptr || (l & 1)) && aligned16(lv[l]), TCS_E_ALIGN,
                    "tcs_corr_build_fused: level %d pointer null or not 16-byte aligned (only the odd levels may be omitted)", l);
    TCS_REQUIRE(B > 0 && H > 0 && W1 > 0 && C > 0, TCS_E_BADARG, "tcs_corr_build_fused: bad sizes");
    TCS_REQUIRE(W2 >= 8 && W2 <= kMaxN && W1 <= kMaxW1 && W1 % 4 == 0 && W2 % 4 == 0, TCS_E_SHAPE,
                "tcs_corr_build_fused: needs 8 <= W2 <= %d, W1 <= %d, both multiples of 4 (got W1=%d W2=%d); use tcs_corr_prepass + tcs_corr_build",
                kMaxN, kMaxW1, W1, W2);
    TCS_REQUIRE((W2 >> (num_levels - 1)) >= 1, TCS_E_SHAPE, "tcs_corr_build_fused: W2 too small for %d levels", num_levels);
    TCS_REQUIRE(C % kBlockK == 0, TCS_E_SHAPE, "tcs_corr_build_fused: C=%d must be a multiple of 32", C);
    TCS_REQUIRE(aligned16(fmap1) && aligned16(fmap2), TCS_E_ALIGN, "tcs_corr_build_fused: feature maps must be 16-byte aligned");
    const bool x3 = (prec == TCS_PREC_BF16X3 || prec == TCS_PREC_FP16X3);
    const bool fp16 = (prec == TCS_PREC_FP16 || prec == TCS_PREC_FP16X3);

    Params p{};
    for (int l = 0; l < 4; ++l) p.lvl[l] = lv[l];
    p.H = H; p.W1 = W1; p.W2 = W2; p.C = C; p.num_levels = num_levels;
    p.num_rows = B * H;
    p.num_m = ceil_div(W1, kBlockM);
    p.block_n = ceil_div(W2, 16) * 16;
    p.kblocks = C / kBlockK;
    p.passes = x3 ? 3 : 1;
    p.idesc = ptx::make_idesc_f16(fp16 ? 0u : 1u, kBlockM, (uint32_t)p.block_n);
    p.in_scale = fp16 ? 256.0f : 1.0f;
    p.out_scale = fp16 ? (1.0f / 65536.0f) : 1.0f;

    CUtensorMap tma, tmb;
    int rc;
    if ((rc = make_fmap_map(&tma, fmap1, B, C, H, W1, kBlockM)) != 0) return rc;
    if ((rc = make_fmap_map(&tmb, fmap2, B, C, H, W2, p.block_n)) != 0) return rc;
    // TMA stores need 16-byte row strides on both levels; otherwise the epilogue falls back to ordinary stores
    // (measured 373 us vs 354 us for the ordinary stores at 540p x 8: each warp has a single staging box, so the
    // bulk store serialises with the next chunk; opt in with TCS_FUSED_TMA_STORE=1)
    { const char* e = getenv("TCS_FUSED_TMA_STORE"); p.tma_store = (kEpiStageBytes >= 6144) && (W2 % 8 == 0) && e != nullptr && atoi(e) != 0 && (num_levels < 2 || lv[1] != nullptr); }
    CUtensorMap tml0 = tma, tml1 = tma;
    if (p.tma_store) {
        if ((rc = make_level_map(&tml0, lv[0], B * H, W1, W2, 32, CU_TENSOR_MAP_SWIZZLE_128B)) != 0) return rc;
        if (num_levels > 1 && (rc = make_level_map(&tml1, lv[1], B * H, W1, W2 >> 1, 16, CU_TENSOR_MAP_SWIZZLE_64B)) != 0) return rc;
    }

    TCS_ONCE_PER_DEVICE(
        TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_build_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_build_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    );
    const int units = p.num_rows * p.num_m;
    const int ctas = kCtasPerSm * num_sms();
    const int grid = units < ctas ? units : ctas;
    p.tile_mode = kCtasPerSm == 2 ? 1 : 0;
    { const char* e = getenv("TCS_FUSED_TILE_MODE"); if (e != nullptr) p.tile_mode = atoi(e) != 0; }
    p.stagger_ns = 0;
    if (kCtasPerSm == 2 && grid > num_sms()) { const char* e = getenv("TCS_FUSED_STAGGER_NS"); p.stagger_ns = e != nullptr ? atoi(e) : 5000; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (fp16) corr_build_fused_kernel<true><<<grid, kThreads, kSmemBytes, s>>>(tma, tmb, tml0, tml1, p);
    else corr_build_fused_kernel<false><<<grid, kThreads, kSmemBytes, s>>>(tma, tmb, tml0, tml1, p);
    TCS_CHECK_LAUNCH("tcs_corr_build_fused");
    return 0;
}
