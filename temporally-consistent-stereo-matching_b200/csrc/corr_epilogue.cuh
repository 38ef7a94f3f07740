// Epilogue shared by the correlation-build kernels: one finished 128 x N accumulator tile in TMEM ->
// pyramid levels 0..3 in global memory.
// ref: core/corr.py:15-23 (level l+1 = avg_pool2d(level l, [1,2]) == (even + odd) * 0.5, floor on odd widths).
// A null pointer for level 1 or level 3 means "do not store it": the row-aligned lookup kernels re-pool the odd levels from
// the even ones (corr_lookup.cu), so for them a third of the pyramid's bytes need never be written.
#pragma once

#include "tcs_common.cuh"
#include "sm100_ptx.cuh"

namespace tcs {

struct EpilogueArgs {
    float* lvl[TCS_MAX_LEVELS];
    int W1, W2, num_levels;
    float scale;
};

__device__ __forceinline__ void store4(float* __restrict__ row, int col, int limit, bool vec_ok, const float4& v) {
    if (vec_ok && col + 3 < limit) {
        // write-once data: cache-streaming so the levels do not push the (re-read) feature maps out of L2
        asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(row + col), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    } else {
        if (col < limit) row[col] = v.x;
        if (col + 1 < limit) row[col + 1] = v.y;
        if (col + 2 < limit) row[col + 2] = v.z;
        if (col + 3 < limit) row[col + 3] = v.w;
    }
}

// One epilogue warp's share of a tile.  The warp owns TMEM lanes [32*quarter, +32) (= tile rows) and the
// 32-column chunks parity, parity + kChunkStride, ... (kChunkStride warps per quarter split the columns).  After tcgen05.ld each
// thread holds one w1 row, so the pooling cascade along w2 is intra-thread.  Levels 0 and 1 go through a
// swizzled shared-memory transpose so that each store instruction covers whole 128-byte row segments;
// levels 2 and 3 (32 B / 16 B per row and chunk) are written by the owning thread.
//   taddr       TMEM address of (lane quarter, first column of the accumulator)
//   n0, n_end   level-0 column range of the tile (n0 a multiple of 32)
//   row0        first w1 row of this warp's quarter;  rbase = (b*H + h) * W1
//   stage0/1    per-warp staging (32-bit shared addresses): [32 rows][8 float4] and [32 rows][4 float4]; they
//               may alias when kAliasedStage (then an extra warp barrier separates the two uses)
//   bar_tempty  mbarrier to arrive on (all 32 lanes) once this warp has read its last chunk from TMEM
template <bool kAliasedStage, int kChunkStride = 2>
__device__ __forceinline__ void epilogue_tile(const EpilogueArgs& p, uint32_t taddr, int n0, int n_end, int row0,
                                              size_t rbase, int parity, int lane, uint32_t stage0, uint32_t stage1,
                                              uint32_t bar_tempty) {
    const int W1 = p.W1, W2 = p.W2;
    const int W2_1 = W2 >> 1, W2_2 = W2 >> 2, W2_3 = W2 >> 3;
    const bool vec0 = (W2 & 3) == 0, vec1 = (W2_1 & 3) == 0, vec2 = (W2_2 & 3) == 0, vec3 = (W2_3 & 3) == 0;
    const float scale = p.scale;
    const int n_chunks = (n_end - n0 + 31) >> 5;
    const int ch_last = parity + ((n_chunks - 1 - parity) / kChunkStride) * kChunkStride;   // this warp's last chunk (if any)
    if (parity >= n_chunks) {                                      // nothing to read: release the accumulator at once
        ptx::tc_fence_before_sync();
        ptx::mbar_arrive(bar_tempty);
    }
    for (int ch = parity; ch < n_chunks; ch += kChunkStride) {
        float v[32];
        ptx::tmem_ld_32x32(taddr + ch * 32, v);
        if (ch == ch_last) {
            // all TMEM reads of this accumulator by this warp are done: hand it back to the MMA warp
            ptx::tc_fence_before_sync();
            ptx::mbar_arrive(bar_tempty);
        }
        const int cg = n0 + ch * 32;  // first level-0 column of the chunk
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= scale;
        // ---- level 0: transpose through smem so that each store instruction covers 4 full 128 B rows
#pragma unroll
        for (int s = 0; s < 8; ++s)
            sts_v4_f32(stage0 + 16u * (lane * 8 + (s ^ (lane & 7))), make_float4(v[4 * s], v[4 * s + 1], v[4 * s + 2], v[4 * s + 3]));
        // ---- level 1 (needed by every deeper level too)
        float l1[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) l1[j] = (v[2 * j] + v[2 * j + 1]) * 0.5f;
        const bool want1 = p.num_levels > 1 && p.lvl[1] != nullptr;
        if (!kAliasedStage && want1) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
                sts_v4_f32(stage1 + 16u * (lane * 4 + (s ^ ((lane >> 1) & 3))), make_float4(l1[4 * s], l1[4 * s + 1], l1[4 * s + 2], l1[4 * s + 3]));
        }
        __syncwarp();
        {
            const int s = lane & 7;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + (lane >> 3);
                const int row = row0 + rr;
                if (row < W1) {
                    const float4 val = lds_v4_f32(stage0 + 16u * (rr * 8 + (s ^ (rr & 7))));
                    store4(p.lvl[0] + (rbase + row) * W2, cg + 4 * s, n_end, vec0, val);
                }
            }
        }
        if (kAliasedStage && want1) {
            __syncwarp();   // level-0 reads of the shared staging area are done; reuse it for level 1
#pragma unroll
            for (int s = 0; s < 4; ++s)
                sts_v4_f32(stage1 + 16u * (lane * 4 + (s ^ ((lane >> 1) & 3))), make_float4(l1[4 * s], l1[4 * s + 1], l1[4 * s + 2], l1[4 * s + 3]));
            __syncwarp();
        }
        if (want1) {
            const int s = lane & 3;
            const int lim = min(n_end >> 1, W2_1);
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int rr = it * 8 + (lane >> 2);
                const int row = row0 + rr;
                if (row < W1) {
                    const float4 val = lds_v4_f32(stage1 + 16u * (rr * 4 + (s ^ ((rr >> 1) & 3))));
                    store4(p.lvl[1] + (rbase + row) * W2_1, (cg >> 1) + 4 * s, lim, vec1, val);
                }
            }
        }
        __syncwarp();
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

}  // namespace tcs
