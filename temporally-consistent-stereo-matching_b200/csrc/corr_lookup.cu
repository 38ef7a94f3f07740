// Per-GRU-iteration consumers of the correlation pyramid:
//   tcs_corr_lookup       fused all-level linear-interpolated lookup        (ref: core/corr.py:33-52)
//   tcs_corr_lookup_alt   same result with no materialised volume           (new; contract = corr.py:33-52)
//   tcs_corr_argmax       first-frame winner-take-all + uniqueness test     (ref: core/corr.py:67-79)
//   tcs_corr_cost_volume  masked [b,w2,h,w1] copy for the training loss     (ref: core/corr.py:25-31,64-65)
//
// Sampling arithmetic follows bilinear_sampler -> F.grid_sample (core/utils/utils.py:82-97) to the
// rounding: x = k + coords/2^l ; xg = 2x/(W-1) - 1 ; ix = ((xg+1)/2)*(W-1) ; x0 = floor(ix) ;
// out = v[x0]*(x0+1-ix) + v[x0+1]*(ix-x0), out-of-range taps contributing 0 (zeros padding).
#include "tcs_common.cuh"
#include "sm100_ptx.cuh"
#include <cstdlib>

namespace tcs {

struct LevelPtrs {
    const float* p[TCS_MAX_LEVELS];
};
// Kernel parameters live in constant memory; indexing the array with a run-time value would make the compiler
// copy it to local memory first (STL/LDL on the critical path).  Select with compares instead.
__device__ __forceinline__ const float* level_ptr(const LevelPtrs& lv, int l) {
    return l == 0 ? lv.p[0] : l == 1 ? lv.p[1] : l == 2 ? lv.p[2] : lv.p[3];
}

// grid_sample's normalise / un-normalise round trip, step by step in fp32 (no contraction).
__device__ __forceinline__ float sample_pos(float xk, float wm1) {
    const float xg = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, xk), wm1), 1.0f);       // utils.py:86
    return __fmul_rn(__fmul_rn(__fadd_rn(xg, 1.0f), 0.5f), wm1);                 // ATen unnormalize, align_corners
}

// grid_sample position of tap xk on a level of width wm1 + 1: same roundings as sample_pos();
// 2*x and *0.5 are exact, so they are folded into their neighbours (hwm1 = 0.5 * wm1).
__device__ __forceinline__ float sample_pos_fast(float xk, float wm1, float rc, float hwm1) {
    const float xg = __fmaf_rn(2.0f, div_by_const(xk, wm1, rc), -1.0f);
    return __fmul_rn(__fadd_rn(xg, 1.0f), hwm1);
}

// ---- fused lookup, radius 4: one thread per (pixel, level pair) -----------------------------------------
// Bytes, not instructions, bound this kernel: a 40-byte tap window costs whole 128-byte lines, and at the
// coarse levels neighbouring rows are so short that a call drags the entire level through DRAM.  So only
// EVEN levels are read: the thread fetches one span of 24 floats of level 2g — in-row indices
// [2(fu-5), 2(fu+6)+1], fu = floor(coords / 2^(2g+1)) — which contains both the 12-float window of level 2g
// and, as pairs, the 12-entry window of level 2g+1, whose values are recomputed on the fly as
// (a + b) * 0.5: the very expression the build epilogue (and avg_pool2d) used, hence bit-identical.
// The span is fetched as 16-byte-aligned quads (at most 7 LDG.128, none outside the row), parked in the
// thread's own column of a transposed shared-memory tile (conflict-free, private, so no barrier) and the
// 2 x 9 taps are interpolated from there.  Lanes are consecutive pixels: every store is a full 128-byte
// line of one tap plane.
constexpr int kLookThreads = 128;
constexpr int kLookQuads = 7;

// One tap, branch-free: position -> (x0, weights); taps that fall outside the level get weight 0 on both
// sides and a clamped (always valid) shared-memory index, so the loads need no predicate.
struct TapPos {
    float w_lo, w_hi;   // already zeroed for entries outside [0, W)
    int x0;
};

__device__ __forceinline__ TapPos tap_position(float xk, float wm1, float rc, float hwm1, int W) {
    TapPos t;
    const float ix = sample_pos_fast(xk, wm1, rc, hwm1);
    const float x0f = floorf(ix);
    // NaN / far-away positions: the float compares fail, the weights vanish and x0 is forced to 0
    const bool in0 = (x0f >= 0.0f) && (x0f <= wm1);                 // x0 inside the level
    const bool in1 = (x0f >= -1.0f) && (x0f < wm1);                 // x0 + 1 inside the level
    t.x0 = (in0 || in1) ? (int)x0f : 0;
    t.w_hi = in1 ? __fsub_rn(ix, x0f) : 0.0f;
    t.w_lo = in0 ? __fsub_rn(__fadd_rn(x0f, 1.0f), ix) : 0.0f;
    (void)W;
    return t;
}

// The same when the span registers already hold exact zeros wherever the level ends (rows that start and end on
// 16-byte boundaries: a quad is then entirely inside or entirely outside its row, and span_load() leaves the
// outside ones at zero).  Zero padding then needs no predicate at all: 0 * w is the reference's "skip this corner".
// The caller keeps coords finite (sane_coord), so w is finite.
__device__ __forceinline__ TapPos tap_position_padded(float xk, float wm1, float rc, float hwm1) {
    TapPos t;
    const float ix = sample_pos_fast(xk, wm1, rc, hwm1);
    const float x0f = floorf(ix);
    t.x0 = (int)x0f;                                                // saturating; only its clamped offset is used
    t.w_hi = __fsub_rn(ix, x0f);
    t.w_lo = __fsub_rn(__fadd_rn(x0f, 1.0f), ix);
    return t;
}

// NaN / infinite / absurd coordinates -> a finite one far outside every level (all taps are then zero padding,
// which is also what the reference's grid_sample returns for them).
__device__ __forceinline__ float sane_coord(float c) { return (fabsf(c) <= 1.0e9f) ? c : 1.0e9f; }

// The span of one (pixel, level pair): where it starts in the row and its 16-byte quads.
struct Span {
    float4 q[kLookQuads];
    int win_first;     // in-row index (level 2g) of q[0].x
    int span_first;    // in-row index of the first float that may be needed (>= win_first, < win_first + 4)
    float cb;          // coords / 2^(2g)
};

// Issue the loads of a span (no use of the data: the caller overlaps them with other work).
__device__ __forceinline__ void span_load(Span& sp, const float* __restrict__ base, long long p, long long npix, float c0,
                                          int lb, int Wb, bool upper) {
    const int Wu = Wb >> 1;
    const long long row_start = p * Wb;
    const long long readable = ((npix * Wb + 3) >> 2) << 2;     // caller pads each level to 16 B
    sp.cb = c0 * (1.0f / (float)(1 << lb));                     // coords / 2**lb (exact)
    int span_first, span_last;                                  // in-row indices of level lb that may be needed
    if (upper) {
        const int fu = (int)fminf(fmaxf(floorf(sp.cb * 0.5f), -16.0f), (float)(Wu + 16));
        span_first = 2 * (fu - 5);
        span_last = 2 * (fu + 6) + 1;
    } else {
        const int fb = (int)fminf(fmaxf(floorf(sp.cb), -16.0f), (float)(Wb + 16));
        span_first = fb - 5;
        span_last = fb + 6;
    }
    const long long a_abs = ((row_start + span_first) >> 2) << 2;   // floor to a multiple of 4 floats (16 B)
    sp.win_first = (int)(a_abs - row_start);
    sp.span_first = span_first;
    const int need_lo = max(span_first, 0), need_hi = min(span_last, Wb - 1);
#pragma unroll
    for (int k = 0; k < kLookQuads; ++k) {
        const int q_lo = sp.win_first + 4 * k;                  // in-row index of this quad's first float
        const long long idx = a_abs + 4 * k;
        sp.q[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q_lo + 3 >= need_lo && q_lo <= need_hi && idx >= 0 && idx + 3 < readable)
            sp.q[k] = ldg_stream64_f4(reinterpret_cast<const float4*>(base + idx));
    }
}

// Park the span in the thread's own shared-memory column and interpolate the 9 (+9) taps from there.  kKeep: the
// taps stay in registers (keep[0..17]) for a fused consumer instead of going to global memory.
template <bool kKeep>
__device__ __forceinline__ void span_taps(const Span& sp, float (*win)[kLookThreads], int tid, float* __restrict__ out,
                                          long long out_px, int HW, int num_levels, int b, int lb, int Wb, bool upper,
                                          float* keep) {
    const int Wu = Wb >> 1;
#pragma unroll
    for (int k = 0; k < kLookQuads; ++k) {
        win[4 * k + 0][tid] = sp.q[k].x;
        win[4 * k + 1][tid] = sp.q[k].y;
        win[4 * k + 2][tid] = sp.q[k].z;
        win[4 * k + 3][tid] = sp.q[k].w;
    }
    constexpr int kWin = 4 * kLookQuads;
    const float* col = &win[0][tid];                            // entry i of this thread's span: col[i * kLookThreads]
    const int win_first = sp.win_first;
    const float cb = sp.cb, cu = sp.cb * 0.5f;
    {   // ---- level lb: taps straight from the span
        const float wm1 = (float)(Wb - 1), rc = __frcp_rn(wm1), hwm1 = __fmul_rn(0.5f, wm1);
        float* o = out + (((long long)b * num_levels + lb) * 9) * HW + out_px;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const TapPos tp = tap_position(__fadd_rn((float)(t - 4), cb), wm1, rc, hwm1, Wb);   // corr.py:43
            const int i0 = min(max(tp.x0 - win_first, 0), kWin - 2);
            const float v0 = col[i0 * kLookThreads], v1 = col[(i0 + 1) * kLookThreads];
            // a zero weight stands for "outside the level" (zeros padding): the product must be 0 even if the
            // clamped slot holds a non-finite value
            const float a0 = (tp.w_lo != 0.0f) ? __fmul_rn(v0, tp.w_lo) : 0.0f;
            const float r = (tp.w_hi != 0.0f) ? fmaf(v1, tp.w_hi, a0) : a0;
            if (kKeep) keep[t] = r; else stg_stream_f1(o, r);
            o += HW;
        }
    }
    if (kKeep && !upper) {
#pragma unroll
        for (int t = 0; t < 9; ++t) keep[9 + t] = 0.0f;
    }
    if (upper) {   // ---- level lb + 1: entries re-pooled from pairs of the span
        const float wm1 = (float)(Wu - 1), rc = __frcp_rn(wm1), hwm1 = __fmul_rn(0.5f, wm1);
        float* o = out + (((long long)b * num_levels + lb + 1) * 9) * HW + out_px;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const TapPos tp = tap_position(__fadd_rn((float)(t - 4), cu), wm1, rc, hwm1, Wu);
            const int i0 = min(max(2 * tp.x0 - win_first, 0), kWin - 4);   // span index of the pair behind entry x0
            const float v0 = __fmul_rn(__fadd_rn(col[i0 * kLookThreads], col[(i0 + 1) * kLookThreads]), 0.5f);
            const float v1 = __fmul_rn(__fadd_rn(col[(i0 + 2) * kLookThreads], col[(i0 + 3) * kLookThreads]), 0.5f);
            const float a0 = (tp.w_lo != 0.0f) ? __fmul_rn(v0, tp.w_lo) : 0.0f;
            const float r = (tp.w_hi != 0.0f) ? fmaf(v1, tp.w_hi, a0) : a0;
            if (kKeep) keep[9 + t] = r; else stg_stream_f1(o, r);
            o += HW;
        }
    }
}

// The same taps with the span held in REGISTERS (level pair with both levels present).  Shared memory costs this
// kernel L1 capacity, and L1 capacity is what bounds its loads in flight; so the run-time indexing is done with
// selects instead.  After re-aligning the quads to the span start (offset 0..3), the left neighbour of base-level
// tap t sits at span position t + 6 + s with s in {-1, 0, 1, 2} (floor(coords) is 2*fu or 2*fu + 1, and
// grid_sample's round trip moves a tap by at most one), and that of upper-level tap t at pooled position
// t + 1 + e with e in {-1, 0, 1}: a 4-way and a 3-way select per tap.
template <bool kKeep, bool kPadded>
__device__ __forceinline__ void span_taps_R(const float* R, int span_first, float cb_, float* __restrict__ out, long long out_px,
                                            int HW, int num_levels, int b, int lb, int Wb, float* keep);

// kPadded: the rows of this level start on 16-byte boundaries (Wb % 4 == 0), see tap_position_padded().
template <bool kKeep, bool kPadded>
__device__ __forceinline__ void span_taps_reg(const Span& sp, float* __restrict__ out, long long out_px, int HW,
                                              int num_levels, int b, int lb, int Wb, float* keep) {
    const float* f = reinterpret_cast<const float*>(sp.q);      // 28 floats, statically indexed below
    const int off = sp.span_first - sp.win_first;               // 0..3
    float t1[27], R[24];
#pragma unroll
    for (int j = 0; j < 27; ++j) t1[j] = (off & 1) ? f[j + 1] : f[j];
#pragma unroll
    for (int j = 0; j < 24; ++j) R[j] = (off & 2) ? t1[j + 2] : t1[j];
    span_taps_R<kKeep, kPadded>(R, sp.span_first, sp.cb, out, out_px, HW, num_levels, b, lb, Wb, keep);
}

// The nine taps of a REGULAR pixel (see span_taps_R): tap t interpolates V[t] and V[t + 1] at sample_pos_fast(t - 4 + c).  Taps
// are processed two at a time with packed fp32 (TCS_LOOKUP_X2, default on): the same round-to-nearest operations in the same
// order per lane, so the result is bit-identical to the scalar form; what changes is the issue-slot count (this kernel is
// co-limited by instruction issue and load latency: 49 % issue-active at 36 % occupancy, profiles/r02_kernel_lookup.md).
#ifndef TCS_LOOKUP_X2
#define TCS_LOOKUP_X2 1
#endif
template <bool kKeep>
__device__ __forceinline__ void regular_taps(const float* V, float c, float wm1, float rc, float hwm1, float* __restrict__ o, int HW,
                                             float* keep) {
#if TCS_LOOKUP_X2
    const f32x2 c2 = pk2(c), rc2 = pk2(rc), nw2 = pk2(-wm1), two2 = pk2(2.0f), m1 = pk2(-1.0f), one2 = pk2(1.0f), h2 = pk2(hwm1);
#pragma unroll
    for (int t = 0; t < 8; t += 2) {
        const f32x2 xk = add2(pk2((float)(t - 4), (float)(t - 3)), c2);
        const f32x2 q0 = mul2(xk, rc2);                                   // div_by_const(), lane-wise
        const f32x2 q = fma2(fma2(q0, nw2, xk), rc2, q0);
        const f32x2 ix = mul2(add2(fma2(two2, q, m1), one2), h2);         // sample_pos_fast()
        const f32x2 x0 = pk2(floorf(lo2(ix)), floorf(hi2(ix)));
        const f32x2 w_hi = sub2(ix, x0), w_lo = sub2(add2(x0, one2), ix);
        const f32x2 r = fma2(pk2(V[t + 1], V[t + 2]), w_hi, mul2(pk2(V[t], V[t + 1]), w_lo));
        if (kKeep) { keep[t] = lo2(r); keep[t + 1] = hi2(r); }
        else { stg_stream_f1(o, lo2(r)); stg_stream_f1(o + HW, hi2(r)); }
        o += 2 * (long long)HW;
    }
    {
        const float ix = sample_pos_fast(__fadd_rn(4.0f, c), wm1, rc, hwm1);
        const float x0f = floorf(ix);
        const float r = fmaf(V[9], __fsub_rn(ix, x0f), __fmul_rn(V[8], __fsub_rn(__fadd_rn(x0f, 1.0f), ix)));
        if (kKeep) keep[8] = r; else stg_stream_f1(o, r);
    }
#else
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float ix = sample_pos_fast(__fadd_rn((float)(t - 4), c), wm1, rc, hwm1);
        const float x0f = floorf(ix);
        const float r = fmaf(V[t + 1], __fsub_rn(ix, x0f), __fmul_rn(V[t], __fsub_rn(__fadd_rn(x0f, 1.0f), ix)));
        if (kKeep) keep[t] = r; else stg_stream_f1(o, r);
        o += HW;
    }
#endif
}

// R[0..23]: the span itself (entry j = in-row index span_first + j of level lb, zero outside the row when kPadded).
template <bool kKeep, bool kPadded>
__device__ __forceinline__ void span_taps_R(const float* R, int span_first, float cb_, float* __restrict__ out, long long out_px,
                                            int HW, int num_levels, int b, int lb, int Wb, float* keep) {
    const int Wu = Wb >> 1;
    const float cb = cb_, cu = cb_ * 0.5f;
    const int fu = (span_first >> 1) + 5;                       // span_first = 2 * (fu - 5)
    // Regular pixels: grid_sample's round trip moves a tap by less than 2e-4 px (levels up to 1024 wide), so when the
    // fractional part of the level coordinate keeps 2^-9 away from 0 and 1 every tap's floor is floor(coordinate) +
    // (t - 4) and the neighbours sit at ONE statically known offset: no per-tap index arithmetic or selects.  Taken
    // only when the whole warp is regular (exact-integer coordinates, e.g. a zero-flow first iteration, are not).
    constexpr float kRegLo = 1.0f / 512.0f, kRegHi = 1.0f - 1.0f / 512.0f;
    const unsigned active = __activemask();
    {   // ---- level lb
        const float wm1 = (float)(Wb - 1), rc = __frcp_rn(wm1), hwm1 = __fmul_rn(0.5f, wm1);
        float* o = out + (((long long)b * num_levels + lb) * 9) * HW + out_px;
        const float fbf = floorf(cb), frac = cb - fbf;
        const int cs = (int)fbf - span_first - 9;               // the select index c every tap would get: 1 or 2
        const bool regular = kPadded && Wb <= 1024 && frac >= kRegLo && frac <= kRegHi && (cs == 1 || cs == 2);
        if (kPadded && __all_sync(active, regular)) {
            float Q[10];
#pragma unroll
            for (int i = 0; i < 10; ++i) Q[i] = (cs == 2) ? R[i + 7] : R[i + 6];
            regular_taps<kKeep>(Q, cb, wm1, rc, hwm1, o, HW, keep);
        } else {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float xk = __fadd_rn((float)(t - 4), cb);                    // corr.py:43
            const TapPos tp = kPadded ? tap_position_padded(xk, wm1, rc, hwm1) : tap_position(xk, wm1, rc, hwm1, Wb);
            const int c = min(max(tp.x0 - span_first - (t + 5), 0), 3);     // s + 1
            const float lo01 = (c & 1) ? R[t + 6] : R[t + 5], lo23 = (c & 1) ? R[t + 8] : R[t + 7];
            const float hi01 = (c & 1) ? R[t + 7] : R[t + 6], hi23 = (c & 1) ? R[t + 9] : R[t + 8];
            const float v0 = (c & 2) ? lo23 : lo01, v1 = (c & 2) ? hi23 : hi01;
            float r;
            if (kPadded) {
                r = fmaf(v1, tp.w_hi, __fmul_rn(v0, tp.w_lo));
            } else {
                const float a0 = (tp.w_lo != 0.0f) ? __fmul_rn(v0, tp.w_lo) : 0.0f;
                r = (tp.w_hi != 0.0f) ? fmaf(v1, tp.w_hi, a0) : a0;
            }
            if (kKeep) keep[t] = r; else stg_stream_f1(o, r);
            o += HW;
        }
        }
    }
    {   // ---- level lb + 1: entries re-pooled from pairs of the span
        float P[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) P[j] = __fmul_rn(__fadd_rn(R[2 * j], R[2 * j + 1]), 0.5f);
        const float wm1 = (float)(Wu - 1), rc = __frcp_rn(wm1), hwm1 = __fmul_rn(0.5f, wm1);
        float* o = out + (((long long)b * num_levels + lb + 1) * 9) * HW + out_px;
        const float fuf = floorf(cu), frac = cu - fuf;
        const bool regular = kPadded && Wu <= 1024 && frac >= kRegLo && frac <= kRegHi && (int)fuf == fu;   // fu unclamped
        if (kPadded && __all_sync(active, regular)) {
            regular_taps<kKeep>(P + 1, cu, wm1, rc, hwm1, o, HW, kKeep ? keep + 9 : keep);   // the select index is 1 for every tap
        } else {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const float xk = __fadd_rn((float)(t - 4), cu);
            const TapPos tp = kPadded ? tap_position_padded(xk, wm1, rc, hwm1) : tap_position(xk, wm1, rc, hwm1, Wu);
            const int c = min(max(tp.x0 - (fu - 5) - t, 0), 2);                  // e + 1
            const float v0 = (c == 0) ? P[t] : (c == 1) ? P[t + 1] : P[t + 2];
            const float v1 = (c == 0) ? P[t + 1] : (c == 1) ? P[t + 2] : P[t + 3];
            float r;
            if (kPadded) {
                r = fmaf(v1, tp.w_hi, __fmul_rn(v0, tp.w_lo));
            } else {
                const float a0 = (tp.w_lo != 0.0f) ? __fmul_rn(v0, tp.w_lo) : 0.0f;
                r = (tp.w_hi != 0.0f) ? fmaf(v1, tp.w_hi, a0) : a0;
            }
            if (kKeep) keep[9 + t] = r; else stg_stream_f1(o, r);
            o += HW;
        }
        }
    }
}

// Level 0 of a row-aligned pyramid read as four 32-byte loads instead of seven 16-byte ones: what bounds this kernel
// is the number of load requests the SM keeps in flight, not their bytes (23.0 -> 20.4 us).  Needs W2 % 8 == 0 and a
// 32-byte aligned base: rows then start on 32-byte boundaries and an oct is entirely inside or outside its row.
// (The same for level 2, whose 240-byte rows would need half-oct fix-ups, is no faster than quads: 23.1 us.)
// TCS_LOOKUP_L2HINT (experiment): 1 = the window loads carry an L2 evict_last policy and the tap
// stores an evict_first one (successive GRU iterations read nearly the same windows: ~76 MB of granules, which would fit L2
// if the 36 MB of output planes per call did not push them out); 2 = only the loads are hinted.
#ifndef TCS_LOOKUP_L2HINT
#define TCS_LOOKUP_L2HINT 0
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void ldg_oct(float* r, const float* p) {
#if TCS_LOOKUP_L2HINT
    const uint64_t pol = l2_policy_evict_last();
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::64B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "l"(p), "l"(pol));
#else
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]) : "l"(p));
#endif
}

struct SpanOct {
    float o[32];
    int off;           // span_first - (in-row index of o[0]): 0, 2, 4 or 6
    int span_first;
    float cb;
};

__device__ __forceinline__ void span_load_oct(SpanOct& sp, const float* __restrict__ base, long long p, float c0, int lb, int Wb) {
    const int Wu = Wb >> 1;
    sp.cb = c0 * (1.0f / (float)(1 << lb));
    const int fu = (int)fminf(fmaxf(floorf(sp.cb * 0.5f), -16.0f), (float)(Wu + 16));
    const int span_first = 2 * (fu - 5);
    sp.span_first = span_first;
    const int win_first = span_first & ~7;                           // row starts are multiples of 8 floats
    sp.off = span_first - win_first;
    const float* row = base + p * Wb;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int q_lo = win_first + 8 * k;
        float* r = sp.o + 8 * k;
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = 0.0f;
        if (q_lo >= 0 && q_lo < Wb && q_lo <= span_first + 23) ldg_oct(r, row + q_lo);
    }
}

template <bool kKeep>
__device__ __forceinline__ void span_taps_oct(const SpanOct& sp, float* __restrict__ out, long long out_px, int HW, int b, int lb,
                                              int Wb, float* keep) {
    float t1[30], R[24];
#pragma unroll
    for (int j = 0; j < 30; ++j) t1[j] = (sp.off & 2) ? sp.o[j + 2] : sp.o[j];
#pragma unroll
    for (int j = 0; j < 24; ++j) R[j] = (sp.off & 4) ? t1[j + 4] : t1[j];
    span_taps_R<kKeep, true>(R, sp.span_first, sp.cb, out, out_px, HW, 4, b, lb, Wb, keep);
}

// One thread per (pixel, level pair), blockIdx.y = pair: 50 registers instead of 80, so 40 warps per SM hide the
// latency of the scattered loads instead of 24 (20.4 -> 19.5 us; the coordinate is simply read twice).
// Block size 64 / 96 / 128 are equivalent, 256 is 3 % slower.  With the regular path and the packed taps the kernel would take
// 72 registers; capped at 64 (8 CTAs per SM: 0.667 vs 0.676 ms per 32 calls; 9 -> 56 registers 0.675; 10 -> 48 registers and
// 32 bytes spilled 0.717; tools/sweep_lookup.sh).
#ifndef TCS_LOOKUP_MINBLOCKS
#define TCS_LOOKUP_MINBLOCKS 8
#endif
#ifndef TCS_LOOKUP_PDL
#define TCS_LOOKUP_PDL 1
#endif

__global__ void __launch_bounds__(kLookThreads, TCS_LOOKUP_MINBLOCKS)
corr_lookup_r4x4o_kernel(const LevelPtrs lv, const float* __restrict__ coords, long long coords_bstride,
                         float* __restrict__ out, int HW, int W2, int W2p) {   // W2p: row pitch of level 0 (>= W2, zeros beyond W2)
    const int b = blockIdx.z;
#if TCS_LOOKUP_PDL
    // Programmatic dependent launch: this grid's CTAs may be scheduled while the previous kernel of the stream drains; nothing is
    // read before that kernel has completed and flushed (wait), and the NEXT kernel may start launching at once.
    pdl_wait_then_release();
#endif
    const int hw = blockIdx.x * blockDim.x + threadIdx.x;           // launched with kLookThreads, or 32 for a small frame
    if (hw >= HW) return;
    const long long npix = (long long)gridDim.z * HW;
    const long long p = (long long)b * HW + hw;
    const float c0 = sane_coord(__ldg(coords + b * coords_bstride + hw));
    if (blockIdx.y == 0) {
        SpanOct s0;
        span_load_oct(s0, lv.p[0], p, c0, 0, W2p);                      // addressing: the pitch; sampling arithmetic: the width
        span_taps_oct<false>(s0, out, hw, HW, b, 0, W2, nullptr);
    } else {
        Span s1;
        span_load(s1, lv.p[2], p, npix, c0, 2, W2p >> 2, true);
        span_taps_reg<false, true>(s1, out, hw, HW, 4, b, 2, W2 >> 2, nullptr);
    }
}

// The standard configuration (4 levels): no shared memory at all.  kPadded: W2 % 16 == 0, i.e. the rows of levels
// 0 and 2 start on 16-byte boundaries and the taps need no bounds predicates.
template <bool kPadded>
__global__ void __launch_bounds__(kLookThreads)
corr_lookup_r4x4_kernel(const LevelPtrs lv, const float* __restrict__ coords, long long coords_bstride,
                        float* __restrict__ out, int HW, int W2, int W2p) {
    const int b = blockIdx.z;
    const int hw = blockIdx.x * blockDim.x + threadIdx.x;
    if (hw >= HW) return;
    const long long npix = (long long)gridDim.z * HW;
    const long long p = (long long)b * HW + hw;
    float c0 = __ldg(coords + b * coords_bstride + hw);
    if (kPadded) c0 = sane_coord(c0);
    Span sp;                                   // blockIdx.y = level pair (see corr_lookup_r4x4o_kernel)
    if (blockIdx.y == 0) {
        span_load(sp, lv.p[0], p, npix, c0, 0, W2p, true);
        span_taps_reg<false, kPadded>(sp, out, hw, HW, 4, b, 0, W2, nullptr);
    } else {
        span_load(sp, lv.p[2], p, npix, c0, 2, W2p >> 2, true);
        span_taps_reg<false, kPadded>(sp, out, hw, HW, 4, b, 2, W2 >> 2, nullptr);
    }
}

// One thread per pixel, both level pairs: the loads of pair 1 (levels 2,3) are issued before the taps of pair 0
// (levels 0,1) are evaluated, so their latency hides behind ~600 instructions of interpolation.
__global__ void __launch_bounds__(kLookThreads)
corr_lookup_r4_kernel(const LevelPtrs lv, const float* __restrict__ coords, long long coords_bstride,
                      float* __restrict__ out, int HW, int W2, int num_levels) {
    __shared__ float win[4 * kLookQuads][kLookThreads];
    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int hw = blockIdx.x * kLookThreads + tid;
    if (hw >= HW) return;
    const long long npix = (long long)gridDim.z * HW;
    const long long p = (long long)b * HW + hw;
    const float c0 = __ldg(coords + b * coords_bstride + hw);
    const bool two = num_levels > 2;
    Span s0, s1;
    span_load(s0, lv.p[0], p, npix, c0, 0, W2, num_levels > 1);
    if (two) span_load(s1, lv.p[2], p, npix, c0, 2, W2 >> 2, num_levels > 3);
    span_taps<false>(s0, win, tid, out, hw, HW, num_levels, b, 0, W2, num_levels > 1, nullptr);
    if (two) span_taps<false>(s1, win, tid, out, hw, HW, num_levels, b, 2, W2 >> 2, num_levels > 3, nullptr);
}

// ---- lookup fused with the motion encoder's first layer ------------------------------------------------------
// ref: core/update.py:97,104 (BasicMotionEncoder: cor = relu(convc1(corr)), convc1 = Conv2d(36, 64, 1)) applied to
// the result of core/corr.py:33-52.  The 36 taps never go to global memory: they stay in the thread's registers and
// the per-pixel 36 -> Cout product runs against weights broadcast from shared memory (SURVEY.md section 8f, rank 1:
// saves the 144 B/pixel store and its re-read every GRU iteration, and one launch).
constexpr int kEncTaps = 36;
constexpr int kEncMaxOut = 128;

template <int kMode>   // 0: any width; 1: W2 % 16 == 0 (no bounds predicates); 2: + level 0 32-byte aligned (32-byte loads)
__global__ void __launch_bounds__(kLookThreads)
corr_lookup_encode_kernel(const LevelPtrs lv, const float* __restrict__ coords, long long coords_bstride,
                          const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ out,
                          int HW, int W2, int W2p, int Cout, int relu) {
    __shared__ __align__(16) float s_w[kEncMaxOut * kEncTaps];
    __shared__ float s_b[kEncMaxOut];
    const int tid = threadIdx.x;
    for (int i = tid; i < Cout * kEncTaps; i += kLookThreads) s_w[i] = __ldg(weight + i);
    for (int i = tid; i < Cout; i += kLookThreads) s_b[i] = (bias != nullptr) ? __ldg(bias + i) : 0.0f;
    __syncthreads();
    const int b = blockIdx.z;
    const int hw = blockIdx.x * kLookThreads + tid;
    if (hw >= HW) return;
    const long long npix = (long long)gridDim.z * HW;
    const long long p = (long long)b * HW + hw;
    constexpr bool kPadded = kMode != 0;
    float c0 = __ldg(coords + b * coords_bstride + hw);
    if (kPadded) c0 = sane_coord(c0);
    float tp[kEncTaps];
    Span s1;
    if (kMode == 2) {
        SpanOct s0;
        span_load_oct(s0, lv.p[0], p, c0, 0, W2p);
        span_load(s1, lv.p[2], p, npix, c0, 2, W2p >> 2, true);
        span_taps_oct<true>(s0, nullptr, 0, HW, b, 0, W2, tp);
    } else {
        Span s0;
        span_load(s0, lv.p[0], p, npix, c0, 0, W2p, true);
        span_load(s1, lv.p[2], p, npix, c0, 2, W2p >> 2, true);
        span_taps_reg<true, kPadded>(s0, nullptr, 0, HW, 4, b, 0, W2, tp);
    }
    span_taps_reg<true, kPadded>(s1, nullptr, 0, HW, 4, b, 2, W2 >> 2, tp + 18);
    float* o = out + (long long)b * Cout * HW + hw;
    for (int oc = 0; oc < Cout; oc += 4) {
        float acc[4] = {s_b[oc], s_b[oc + 1], s_b[oc + 2], s_b[oc + 3]};
#pragma unroll
        for (int k = 0; k < kEncTaps; k += 4) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 w = *reinterpret_cast<const float4*>(s_w + (oc + j) * kEncTaps + k);   // broadcast
                acc[j] = fmaf(w.x, tp[k], acc[j]);
                acc[j] = fmaf(w.y, tp[k + 1], acc[j]);
                acc[j] = fmaf(w.z, tp[k + 2], acc[j]);
                acc[j] = fmaf(w.w, tp[k + 3], acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) stg_stream_f1(o + (long long)(oc + j) * HW, relu ? fmaxf(acc[j], 0.0f) : acc[j]);
    }
}

// ---- the same on the tensor cores (Cout = 64, row-aligned levels) ------------------------------------------------------------------
// The 36 -> 64 product of 128 pixels is a small GEMM.  A tile = 128 pixels, a CTA = 256 threads: TWO threads per pixel, one per
// level pair, exactly as in the plain lookup (the pair's 18 taps in one thread keep it at the lookup's register count and hence
// at its number of resident warps, which is what the scattered window loads live on: with all 36 taps in one thread the kernel
// needed 95 registers and ran at 46 us against the lookup's 21, profiles/r02_encode_tc.md).  Each thread splits its first 16 taps
// into fp16 hi + lo (x 2^8, as the build does) and writes them as two whole 16-byte chunks of its pixel's row of the A operand
// (K-major SWIZZLE_64B, [128 x 32]: K index k < 16 = tap k of levels 0-1, 16 <= k < 32 = tap k - 16 of levels 2-3); its last two
// taps go to a small shared table.  The B operand is the weight matrix in the same K order, split and swizzled ONCE by
// encode_pack_weights_kernel (scaled by a power of two that puts its largest entry in [2^9, 2^10)).  One thread issues
// hi.hi + hi.lo + lo.hi as 6 tcgen05.mma 128 x 64 x 16 into 64 TMEM columns; then all eight warps read the accumulator back
// (warp w: lanes 32 (w % 4).., columns 32 (w / 4)..: thread = its pixel again, half of the outputs), un-scale, add the bias and
// the four left-over taps' share (4 fp32 FMAs per output, weights broadcast from shared memory), apply ReLU and store 32
// coalesced planes each.  CTAs are persistent (TMEM, barrier and weights set up once).
namespace enc_tc {
constexpr int kN = 64;                      // output channels
constexpr int kRows = 128;                  // pixels per tile = UMMA M
constexpr int kThreads = 2 * kRows;         // two threads per pixel
constexpr int kK = 32;                      // taps on the tensor cores: 16 of each level pair
constexpr int kTail = kEncTaps - kK;        // 4 taps on the CUDA cores: taps 16, 17 of each level pair
constexpr int kATile = kRows * 64;          // 8 KB (32 fp16 = 64 B per row)
constexpr int kBTile = kN * 64;             // 4 KB
constexpr int kPackWeights = 2 * kBTile;    // B_hi, B_lo
constexpr int kPackTail = kN * kTail * 4;   // fp32 weights of the four left-over taps, [64][4]
constexpr int kPackBytes = kPackWeights + kPackTail + kN * 4 + 16;   // + bias[64] + {2^-(8+s), pad}
constexpr int kTailBytes = 2 * kRows * kTail * 4;                    // left-over taps [tile parity][128][4] fp32
constexpr int kSmemBytes = 1024 + 2 * kATile + kPackBytes + 16 + kTailBytes;
constexpr int kTmemCols = 64;

// byte offset of fp16 element (row r, k < 32) inside a K-major SWIZZLE_64B operand block
__host__ __device__ constexpr uint32_t sw64_offset(int r, int k) {
    return (uint32_t)((r >> 3) * 512 + (r & 7) * 64 + (((k >> 3) ^ ((r >> 1) & 3)) << 4) + (k & 7) * 2);
}
// lookup output channel (level * 9 + tap) behind GEMM K index k / behind left-over tap q
__host__ __device__ constexpr int tap_of_k(int k) { return k < 16 ? k : 18 + (k - 16); }
__host__ __device__ constexpr int tap_of_tail(int q) { return q < 2 ? 16 + q : 34 + (q - 2); }
}  // namespace enc_tc

// packed: [B_hi | B_lo | tail weights | bias | 2^-(8+s)]: see enc_tc.  One block.
__global__ void __launch_bounds__(256)
encode_pack_weights_kernel(const float* __restrict__ weight, const float* __restrict__ bias, unsigned char* __restrict__ packed) {
    using namespace enc_tc;
    __shared__ float red[256];
    float m = 0.0f;
    for (int i = threadIdx.x; i < kN * kEncTaps; i += 256) m = fmaxf(m, fabsf(__ldg(weight + i)));
    red[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    m = red[0];
    const int s = (m > 0.0f && isfinite(m)) ? 9 - ilogbf(m) : 0;       // max |w| * 2^s in [2^9, 2^10)
    const float scale = ldexpf(1.0f, s);
    for (int e = threadIdx.x; e < kN * kK; e += 256) {
        const int n = e / kK, k = e % kK;
        const float v = __ldg(weight + n * kEncTaps + tap_of_k(k)) * scale;
        const __half h = __float2half_rn(v);
        const __half l = __float2half_rn(v - __half2float(h));
        const uint32_t off = sw64_offset(n, k);
        *reinterpret_cast<__half*>(packed + off) = h;
        *reinterpret_cast<__half*>(packed + kBTile + off) = l;
    }
    float* tail = reinterpret_cast<float*>(packed + kPackWeights);
    for (int i = threadIdx.x; i < kN * kTail; i += 256) tail[i] = __ldg(weight + (i / kTail) * kEncTaps + tap_of_tail(i % kTail));
    float* bs = tail + kN * kTail;
    for (int i = threadIdx.x; i < kN; i += 256) bs[i] = bias != nullptr ? __ldg(bias + i) : 0.0f;
    if (threadIdx.x == 0) { bs[kN] = ldexpf(1.0f, -(8 + s)); bs[kN + 1] = 0.0f; bs[kN + 2] = 0.0f; bs[kN + 3] = 0.0f; }
}

#ifndef TCS_ENCODE_TC_MINBLOCKS
#define TCS_ENCODE_TC_MINBLOCKS 4              // 256 threads x 64 registers
#endif
template <int kMode>   // 1: W2 pitch % 16 == 0; 2: + level 0 32-byte aligned
__global__ void __launch_bounds__(enc_tc::kThreads, TCS_ENCODE_TC_MINBLOCKS)
corr_lookup_encode_tc_kernel(const LevelPtrs lv, const float* __restrict__ coords, long long coords_bstride,
                             const unsigned char* __restrict__ packed, float* __restrict__ out, int HW, int W2, int W2p, int relu,
                             int tiles_per_sample, int num_tiles) {
    using namespace enc_tc;
    extern __shared__ uint8_t enc_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(enc_smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    const uint32_t a_hi = smem_u32(smem), a_lo = a_hi + kATile;
    uint8_t* pk = smem + 2 * kATile;                                   // the packed weights, copied as they are
    const uint32_t b_hi = smem_u32(pk), b_lo = b_hi + kBTile;
    const float* s_tail = reinterpret_cast<const float*>(pk + kPackWeights);
    const float* s_bias = s_tail + kN * kTail;
    uint64_t* bar_ptr = reinterpret_cast<uint64_t*>(pk + kPackBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_ptr + 1);
    float* left = reinterpret_cast<float*>(pk + kPackBytes + 16);      // [2][128][4]
    const uint32_t bar = smem_u32(bar_ptr);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int px = tid & (kRows - 1), pair = tid >> 7;                 // warps 0-3: levels 0-1, warps 4-7: levels 2-3

    if (warp == 0) {
        if (tid == 0) { ptx::mbar_init(bar, 1); ptx::fence_barrier_init(); }
        __syncwarp();
        ptx::tmem_alloc(smem_u32(tmem_slot), kTmemCols);
        ptx::tmem_relinquish();
    }
    for (int i = tid; i < kPackBytes / 16; i += kThreads)
        reinterpret_cast<uint4*>(pk)[i] = __ldg(reinterpret_cast<const uint4*>(packed) + i);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    ptx::tc_fence_before_sync();
    __syncthreads();
    ptx::tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const float inv = s_bias[kN];
    const long long npix = (long long)(num_tiles / tiles_per_sample) * HW;
    uint32_t phase = 0;

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        // ---- the lookup: this thread's 18 taps (a pixel past the end repeats the last one and is not stored)
        const int b = tile / tiles_per_sample;
        const int hw_raw = (tile - b * tiles_per_sample) * kRows + px;
        const int hw = min(hw_raw, HW - 1);
        const long long p = (long long)b * HW + hw;
        const float c0 = sane_coord(__ldg(coords + b * coords_bstride + hw));
        float tp[18];
        if (pair == 0) {                                               // warp-uniform
            if (kMode == 2) {
                SpanOct s0;
                span_load_oct(s0, lv.p[0], p, c0, 0, W2p);
                span_taps_oct<true>(s0, nullptr, 0, HW, b, 0, W2, tp);
            } else {
                Span s0;
                span_load(s0, lv.p[0], p, npix, c0, 0, W2p, true);
                span_taps_reg<true, true>(s0, nullptr, 0, HW, 4, b, 0, W2, tp);
            }
        } else {
            Span s1;
            span_load(s1, lv.p[2], p, npix, c0, 2, W2p >> 2, true);
            span_taps_reg<true, true>(s1, nullptr, 0, HW, 4, b, 2, W2 >> 2, tp);
        }
        // ---- A operand: row = pixel, this thread's two 16-byte chunks (K = 16 pair .. 16 pair + 15); left-over taps to the table
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float x0 = tp[8 * j + 2 * q] * 256.0f, x1 = tp[8 * j + 2 * q + 1] * 256.0f;
                const __half2 h = __floats2half2_rn(x0, x1);
                const float2 back = __half22float2(h);
                const __half2 l = __floats2half2_rn(x0 - back.x, x1 - back.y);
                hi[q] = *reinterpret_cast<const uint32_t*>(&h);
                lo[q] = *reinterpret_cast<const uint32_t*>(&l);
            }
            const uint32_t off = sw64_offset(px, 16 * pair + 8 * j);
            sts_v4_u32(a_hi + off, hi[0], hi[1], hi[2], hi[3]);
            sts_v4_u32(a_lo + off, lo[0], lo[1], lo[2], lo[3]);
        }
        float* mine = left + ((phase * kRows + px) * kTail + 2 * pair);
        mine[0] = tp[16];
        mine[1] = tp[17];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
        ptx::tc_fence_before_sync();
        __syncthreads();                                               // (also: every warp has read the previous tile's accumulator)
        if (tid == 0) {
            ptx::tc_fence_after_sync();
            constexpr uint32_t idesc = ptx::make_idesc_f16(0u, kRows, kN);
            uint32_t acc = 0;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {                     // hi.hi, hi.lo, lo.hi: the build's order
                const uint64_t da = ptx::make_kmajor_sw64_desc(pass == 2 ? a_lo : a_hi);
                const uint64_t db = ptx::make_kmajor_sw64_desc(pass == 1 ? b_lo : b_hi);
#pragma unroll
                for (int ks = 0; ks < kK / 16; ++ks) {
                    ptx::umma_f16(tmem_base, da + 2 * ks, db + 2 * ks, idesc, acc);
                    acc = 1;
                }
            }
            ptx::umma_commit(bar);
        }
        ptx::mbar_wait(bar, phase);
        ptx::tc_fence_after_sync();
        // ---- epilogue: this thread's pixel, outputs 32 pair .. 32 pair + 31 (the warp's TMEM lane quarter is warp % 4 = px / 32)
        const float4 t4 = *reinterpret_cast<const float4*>(left + (phase * kRows + px) * kTail);
        phase ^= 1;
        float* o = out + ((long long)b * kN + 32 * pair) * HW + hw;
        float v[32];
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 32 * pair, v);
        if (hw_raw < HW) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int oc = 32 * pair + i;
                const float4 wt = *reinterpret_cast<const float4*>(s_tail + oc * kTail);         // broadcast
                float r = fmaf(v[i], inv, s_bias[oc]);
                r = fmaf(wt.x, t4.x, r);
                r = fmaf(wt.y, t4.y, r);
                r = fmaf(wt.z, t4.z, r);
                r = fmaf(wt.w, t4.w, r);
                stg_stream_f1(o + (long long)i * HW, relu ? fmaxf(r, 0.0f) : r);
            }
        }
        ptx::tc_fence_before_sync();                                   // the accumulator reads above precede the next tile's MMAs
    }
    ptx::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after_sync();
        ptx::tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- generic lookup (any radius <= 8): one thread per (pixel, level), scalar loads ---------------------
__global__ void __launch_bounds__(256)
corr_lookup_generic_kernel(const LevelPtrs lv, const float* __restrict__ coords, long long coords_bstride,
                           float* __restrict__ out, int HW, int W2, int num_levels, int radius, long long npix) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const int l = blockIdx.y;
    const int Wl = W2 >> l;
    const float wm1 = (float)(Wl - 1);
    const long long b = p / HW, hw = p - b * HW;
    const float cl = __ldg(coords + b * coords_bstride + hw) * (1.0f / (float)(1 << l));
    const float* __restrict__ row = level_ptr(lv, l) + p * Wl;
    const int taps = 2 * radius + 1;
    float* o = out + ((b * num_levels + l) * taps) * (long long)HW + hw;
    for (int t = 0; t < taps; ++t) {
        const float xk = __fadd_rn((float)(t - radius), cl);
        const float ix = sample_pos(xk, wm1);
        float r = 0.0f;
        if (ix > -1.0f && ix < (float)Wl) {
            const float x0f = floorf(ix);
            const int x0 = (int)x0f;
            const float w_hi = __fsub_rn(ix, x0f);
            const float w_lo = __fsub_rn(__fadd_rn(x0f, 1.0f), ix);
            const float v0 = (x0 >= 0) ? __ldg(row + x0) : 0.0f;
            const float v1 = (x0 + 1 < Wl) ? __ldg(row + x0 + 1) : 0.0f;
            r = fmaf(v1, w_hi, __fmul_rn(v0, w_lo));
        }
        o[(long long)t * HW] = r;
    }
}

// ---- backward of the lookup w.r.t. the volume (training) ---------------------------------
// ref: autograd of core/corr.py:33-52 (grid_sample's gradient w.r.t. its INPUT: every output tap sends g * weight to its
// two in-range neighbours; coords are detached, tc_stereo.py:176) folded through core/corr.py:21-23 (avg_pool2d's
// gradient: a level-l entry hands 2^-l of its gradient to each of its 2^l level-0 columns).  Output: d(volume)
// [B,H,W1,W2], dense, written exactly once per element (no atomics, deterministic: torch's own grid_sampler backward
// scatters with atomicAdd).  One warp per pixel: lane l < num_levels gathers its level's taps into that level's window
// gradient (shared memory, private to the lane), then all lanes write the pixel's row of d(volume), coalesced.
constexpr int kBwdWin = 2 * TCS_MAX_RADIUS + 4;     // window entries per level: 2r + 1 taps, their right neighbours, +-1 of round trip

__global__ void __launch_bounds__(256)
corr_lookup_backward_kernel(const float* __restrict__ gout, const float* __restrict__ coords, long long coords_bstride,
                            float* __restrict__ dvol, int HW, int W2, int num_levels, int radius, long long npix) {
    __shared__ float s_g[8][TCS_MAX_LEVELS][kBwdWin];
    __shared__ int s_first[8][TCS_MAX_LEVELS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long p = (long long)blockIdx.x * 8 + warp;
    if (p >= npix) return;                                              // warp-uniform
    const long long b = p / HW, hw = p - b * HW;
    const float c0 = sane_coord(__ldg(coords + b * coords_bstride + hw));
    const int taps = 2 * radius + 1;
    if (lane < num_levels) {
        const int l = lane, Wl = W2 >> l;
        const float wm1 = (float)(Wl - 1), rc = __frcp_rn(wm1), hwm1 = __fmul_rn(0.5f, wm1);
        const float cl = c0 * (1.0f / (float)(1 << l));
        const int first = (int)fminf(fmaxf(floorf(cl), -64.0f), (float)(Wl + 64)) - radius - 1;
        float* g = s_g[warp][l];
        for (int k = 0; k < kBwdWin; ++k) g[k] = 0.0f;
        const float* go = gout + ((b * num_levels + l) * taps) * (long long)HW + hw;
        for (int t = 0; t < taps; ++t) {
            const TapPos tp = tap_position(__fadd_rn((float)(t - radius), cl), wm1, rc, hwm1, Wl);
            const float gv = __ldg(go + (long long)t * HW);
            const int k = tp.x0 - first;
            if (tp.w_lo != 0.0f && k >= 0 && k < kBwdWin) g[k] = fmaf(gv, tp.w_lo, g[k]);
            if (tp.w_hi != 0.0f && k + 1 >= 0 && k + 1 < kBwdWin) g[k + 1] = fmaf(gv, tp.w_hi, g[k + 1]);
        }
        s_first[warp][l] = first;
    }
    __syncwarp();
    float* row = dvol + p * W2;
    for (int j = lane; j < W2; j += 32) {
        float v = 0.0f;
        float scale = 1.0f;
        for (int l = 0; l < num_levels; ++l) {
            const int e = j >> l;                                       // the level-l entry this column was pooled into
            const int k = e - s_first[warp][l];
            if (e < (W2 >> l) && k >= 0 && k < kBwdWin) v = fmaf(s_g[warp][l][k], scale, v);
            scale *= 0.5f;
        }
        row[j] = v;
    }
}

// ---- alternate path: dot products only at the taps -------------------------------------------------------
// One warp per pixel.  Lanes split the C channels (float4 x C/128 per lane), the left feature vector
// stays in registers, each candidate column of the (pooled) right features is a coalesced C*4-byte
// read (L1/L2 resident: neighbouring pixels share almost all columns), and the per-column partial
// sums are combined with a halving butterfly (16 values in 16 shuffles instead of 80).
template <int kC4>  // float4 per lane = C / 128
__global__ void __launch_bounds__(256)
corr_lookup_alt_kernel(const float* __restrict__ a, const LevelPtrs bl, const float* __restrict__ coords,
                       long long coords_bstride, float* __restrict__ out, int HW, int W1, int W2,
                       int num_levels, int radius, long long npix) {
    const int lane = threadIdx.x & 31;
    const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= npix) return;  // warp-uniform
    constexpr int C = kC4 * 128;
    const long long b = p / HW, hw = p - b * HW;
    const long long bh = p / W1;  // row index b*H + h
    const float c0 = __ldg(coords + b * coords_bstride + hw);
    float4 av[kC4];
#pragma unroll
    for (int i = 0; i < kC4; ++i) av[i] = __ldg(reinterpret_cast<const float4*>(a + p * C) + lane + 32 * i);
    const int taps = 2 * radius + 1;
    const int ncols = taps + 3;  // columns [fc-r-1, fc+r+2]

    for (int l = 0; l < num_levels; ++l) {
        const int Wl = W2 >> l;
        const float wm1 = (float)(Wl - 1);
        const float cl = c0 * (1.0f / (float)(1 << l));
        const float fcf = fminf(fmaxf(floorf(cl), -32.0f), (float)(Wl + 32));
        const int col0 = (int)fcf - radius - 1;
        const float* __restrict__ brow = level_ptr(bl, l) + bh * (long long)Wl * C;
        // dots[g][i]: partial sums of columns col0 + 16*g + i  (ncols <= 20 -> up to two groups of 16)
        for (int g = 0; g * 16 < ncols; ++g) {
            float part[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int col = col0 + g * 16 + i;
                float s = 0.0f;
                if (g * 16 + i < ncols && col >= 0 && col < Wl) {
                    const float4* bp = reinterpret_cast<const float4*>(brow + (long long)col * C);
#pragma unroll
                    for (int q = 0; q < kC4; ++q) {
                        const float4 bv = __ldg(bp + lane + 32 * q);
                        s = fmaf(av[q].x, bv.x, s); s = fmaf(av[q].y, bv.y, s);
                        s = fmaf(av[q].z, bv.z, s); s = fmaf(av[q].w, bv.w, s);
                    }
                }
                part[i] = s;
            }
            // halving butterfly: after it, lane L holds the full sum of column index (L >> 1) & 15 ... see below
#pragma unroll
            for (int i = 0; i < 8; ++i) {   // xor 16: lanes < 16 keep columns 0..7, lanes >= 16 keep 8..15
                const bool up = (lane & 16) != 0;
                const float send = up ? part[i] : part[i + 8];
                const float keep = up ? part[i + 8] : part[i];
                part[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {   // xor 8
                const bool up = (lane & 8) != 0;
                const float send = up ? part[i] : part[i + 4];
                const float keep = up ? part[i + 4] : part[i];
                part[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {   // xor 4
                const bool up = (lane & 4) != 0;
                const float send = up ? part[i] : part[i + 2];
                const float keep = up ? part[i + 2] : part[i];
                part[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            {                               // xor 2
                const bool up = (lane & 2) != 0;
                const float send = up ? part[0] : part[1];
                const float keep = up ? part[1] : part[0];
                part[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
            part[0] += __shfl_xor_sync(0xffffffffu, part[0], 1);
            // lane L now owns column i(L) = 8*b4 + 4*b3 + 2*b2 + b1 (bits of L); column i lives in lane 2*rev... :
            // owner lane of column i: bit4=i>>3&1, bit3=i>>2&1, bit2=i>>1&1, bit1=i&1  ->  lane = 2*i (bit0 free)
            const float mine = part[0];
            // every lane t < taps computes tap t of this level (two groups are merged through shuffles below)
            for (int t0 = 0; t0 < taps; t0 += 32) {
                const int t = t0 + lane;
                float r = 0.0f;
                int i0 = -1000;
                float w_lo = 0.0f, w_hi = 0.0f;
                bool inb = false;
                int x0 = 0;
                if (t < taps) {
                    const float xk = __fadd_rn((float)(t - radius), cl);
                    const float ix = sample_pos(xk, wm1);
                    if (ix > -1.0f && ix < (float)Wl) {
                        const float x0f = floorf(ix);
                        x0 = (int)x0f;
                        w_hi = __fsub_rn(ix, x0f);
                        w_lo = __fsub_rn(__fadd_rn(x0f, 1.0f), ix);
                        i0 = x0 - col0 - g * 16;
                        inb = true;
                    }
                }
                // fetch column sums i0 and i0+1 of this group (if they are in this group)
                const int s0 = ((unsigned)i0 < 16u) ? 2 * i0 : 0;
                const int s1 = ((unsigned)(i0 + 1) < 16u) ? 2 * (i0 + 1) : 0;
                const float v0 = __shfl_sync(0xffffffffu, mine, s0);
                const float v1 = __shfl_sync(0xffffffffu, mine, s1);
                if (inb) {
                    // contributions from this group only; the two groups partition the columns, so adding
                    // the per-group contributions reproduces v0*w_lo + v1*w_hi exactly when g covers both.
                    const float c_lo = ((unsigned)i0 < 16u && x0 >= 0) ? v0 : 0.0f;
                    const float c_hi = ((unsigned)(i0 + 1) < 16u && x0 + 1 < Wl) ? v1 : 0.0f;
                    r = fmaf(c_hi, w_hi, __fmul_rn(c_lo, w_lo));
                }
                if (t < taps) {
                    float* o = out + ((b * num_levels + l) * taps + t) * (long long)HW + hw;
                    if (g == 0) *o = r; else *o += r;
                }
            }
        }
    }
}

// ---- argmax_disp ---------------------------------------------------------------------------------------
// One warp per (b,h,w1) row of level 0.  Pass 1: max over the masked row (w2 > w1 -> 0), first index on
// ties (torch.max semantics).  Pass 2: max with w2 in {idx-1, idx, idx+1} zeroed.  Values stay in
// registers between the passes (W2 <= 1024); the per-lane count is a template parameter so that a 240-wide row
// costs 8 loads and 8 registers, not 32 predicated ones.
constexpr int kArgMaxPerLane = 32;   // upper bound: rows of up to 1024 columns

template <int kPer>
__global__ void __launch_bounds__(256)
corr_argmax_kernel(const float* __restrict__ lvl0, float* __restrict__ sparse_disp, float* __restrict__ main_cost,
                   float* __restrict__ mask_out, int W1, int W2, float thres, long long nrows) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= nrows) return;
    const int w1 = (int)(row % W1);
    const float* __restrict__ r = lvl0 + row * W2;
    float vals[kPer];
    float best = -INFINITY;
    int best_i = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int w2 = lane + 32 * k;
        float v = -INFINITY;
        if (w2 < W2) {
            v = __ldg(r + w2);
            // corr.py:27-31: cost_volume * mask with mask = 0 where w1 < w2 (so -0.0 for negative entries)
            if (w1 < w2) v = __fmul_rn(v, 0.0f);
            if (v > best) { best = v; best_i = w2; }   // ascending w2 per lane: first index wins ties
        }
        vals[k] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    float sub = -INFINITY;
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
        const int w2 = lane + 32 * k;
        if (w2 < W2) {
            float v = vals[k];
            if (w2 >= best_i - 1 && w2 <= best_i + 1) v = 0.0f;   // corr.py:71
            sub = fmaxf(sub, v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sub = fmaxf(sub, __shfl_xor_sync(0xffffffffu, sub, o));
    if (lane == 0) {
        const float m = (__fsub_rn(best, sub) > thres) ? 1.0f : 0.0f;
        sparse_disp[row] = __fmul_rn((float)(w1 - best_i), m);   // int * float mask: keeps the reference's -0.0
        main_cost[row] = __fmul_rn(best, m);
        mask_out[row] = m;
    }
}

// ---- get_cost_volume: [b,h,w1,w2] -> [b,w2,h,w1], zeroed where w1 < w2 -----------------------------------
__global__ void __launch_bounds__(256)
corr_cost_volume_kernel(const float* __restrict__ lvl0, float* __restrict__ out, int H, int W1, int W2) {
    __shared__ float tile[32][33];
    const int bh = blockIdx.z;
    const int b = bh / H, h = bh - b * H;
    const int w1_0 = blockIdx.y * 32, w2_0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows per pass
    for (int i = ty; i < 32; i += 8) {
        const int w1 = w1_0 + i, w2 = w2_0 + tx;
        float v = 0.0f;
        if (w1 < W1 && w2 < W2) {
            v = __ldg(lvl0 + ((long long)bh * W1 + w1) * W2 + w2);
            if (w1 < w2) v = __fmul_rn(v, 0.0f);
        }
        tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int w2 = w2_0 + i, w1 = w1_0 + tx;
        if (w1 < W1 && w2 < W2) out[(((long long)b * W2 + w2) * H + h) * W1 + w1] = tile[tx][i];
    }
}

}  // namespace tcs

// ================================================ C ABI ==================================================

// odd_optional: the caller's kernels read only levels 0 and 2 when num_levels == 4 and radius == 4 (they re-pool 1 and 3), so
// those two pointers may be NULL (a pyramid built without them, see tcs_corr_build).
static int check_lookup_args(const char* fn, const float* const* lv, const float* coords, const float* out,
                             int B, int H, int W1, int W2, int num_levels, int radius, bool odd_optional = false) {
    using namespace tcs;
    TCS_REQUIRE(coords != nullptr && out != nullptr, TCS_E_BADARG, "%s: null coords/out", fn);
    TCS_REQUIRE(num_levels >= 1 && num_levels <= TCS_MAX_LEVELS, TCS_E_SHAPE, "%s: num_levels=%d not in [1,4]", fn, num_levels);
    TCS_REQUIRE(radius >= 0 && radius <= TCS_MAX_RADIUS, TCS_E_SHAPE, "%s: radius=%d not in [0,8]", fn, radius);
    TCS_REQUIRE(B > 0 && H > 0 && W1 > 0 && (W2 >> (num_levels - 1)) >= 2, TCS_E_SHAPE, "%s: bad sizes (coarsest level needs width >= 2)", fn);
    const bool skip_odd = odd_optional && num_levels == 4 && radius == 4;
    for (int l = 0; l < num_levels; ++l)
        TCS_REQUIRE((lv[l] != nullptr || (skip_odd && (l & 1))) && aligned16(lv[l]), TCS_E_ALIGN,
                    "%s: level %d pointer null or not 16-byte aligned", fn, l);
    return 0;
}

extern "C" int tcs_corr_lookup(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                               const float* coords, long long coords_bstride, float* out,
                               int B, int H, int W1, int W2, int num_levels, int radius, int W2_pitch, void* stream) {
    using namespace tcs;
    const float* lv[4] = {lvl0, lvl1, lvl2, lvl3};
    int rc = check_lookup_args("tcs_corr_lookup", lv, coords, out, B, H, W1, W2, num_levels, radius, true);
    if (rc != 0) return rc;
    const int W2p = W2_pitch > 0 ? W2_pitch : W2;
    TCS_REQUIRE(W2p == W2 || (num_levels == 4 && radius == 4 && W2p > W2 && W2p % 16 == 0 && W2 % 8 == 0), TCS_E_SHAPE,
                "tcs_corr_lookup: a row pitch (%d) other than W2 (%d) needs 4 levels, radius 4, W2 %% 8 == 0 and a pitch that is a multiple of 16", W2p, W2);
    LevelPtrs lp;
    for (int l = 0; l < 4; ++l) lp.p[l] = lv[l];
    const long long npix = (long long)B * H * W1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (radius == 4) {
        TCS_REQUIRE(B <= 65535, TCS_E_SHAPE, "tcs_corr_lookup: B must be <= 65535");
        dim3 grid((unsigned)ceil_div(H * W1, kLookThreads), 1, B);
        TCS_ONCE_PER_DEVICE(   // leave most of the unified array to L1: the loads stream through it (27 vs 68 us)
            const int carve = carveout_percent("TCS_CARVE_LOOKUP", 25);
            if (carve >= 0) TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_lookup_r4_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        );
        // W2 % 16 == 0: the rows of levels 0 and 2 start on 16-byte boundaries (no bounds predicates); with a 32-byte
        // aligned level 0 its span comes as 32-byte loads
        // 4 levels: one thread per (pixel, level pair).  A single frame (batch 1: 2 x 255 CTAs of 128 threads at 540p) leaves
        // most of the 148 SMs' warp slots empty and the call is one latency chain long; 32-thread CTAs spread the same threads
        // over four times as many CTAs (TCS_LOOKUP_SMALL_CTA=0 keeps 128).
        static const int small_cta = carveout_percent("TCS_LOOKUP_SMALL_CTA", 1);
        const long long ctas128 = (long long)grid.x * 2 * B;
        const int threads = (small_cta > 0 && ctas128 < 8LL * num_sms()) ? 32 : kLookThreads;
        const dim3 grid2((unsigned)ceil_div(H * W1, threads), 2, B);
        if (num_levels == 4 && W2p % 16 == 0 && (reinterpret_cast<uintptr_t>(lvl0) & 31) == 0) {
#if TCS_LOOKUP_PDL
            TCS_CHECK_CUDA(launch_pdl(corr_lookup_r4x4o_kernel, grid2, dim3(threads), 0, s, lp, coords, coords_bstride, out, H * W1, W2, W2p));
#else
            corr_lookup_r4x4o_kernel<<<grid2, threads, 0, s>>>(lp, coords, coords_bstride, out, H * W1, W2, W2p);
#endif
        }
        else if (num_levels == 4 && W2p % 16 == 0)
            corr_lookup_r4x4_kernel<true><<<grid2, threads, 0, s>>>(lp, coords, coords_bstride, out, H * W1, W2, W2p);
        else if (num_levels == 4)     // any width: span in registers, no shared memory (+2 % in the step)
            corr_lookup_r4x4_kernel<false><<<grid2, threads, 0, s>>>(lp, coords, coords_bstride, out, H * W1, W2, W2p);
        else
            corr_lookup_r4_kernel<<<grid, kLookThreads, 0, s>>>(lp, coords, coords_bstride, out, H * W1, W2, num_levels);
    } else {
        dim3 grid((unsigned)ceil_div_ll(npix, 256), num_levels);
        corr_lookup_generic_kernel<<<grid, 256, 0, s>>>(lp, coords, coords_bstride, out, H * W1, W2, num_levels, radius, npix);
    }
    TCS_CHECK_LAUNCH("tcs_corr_lookup");
    return 0;
}

extern "C" int tcs_corr_lookup_encode(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                                      const float* coords, long long coords_bstride, const float* weight,
                                      const float* bias, float* out, int B, int H, int W1, int W2, int num_levels,
                                      int radius, int Cout, int relu, int W2_pitch, void* stream) {
    using namespace tcs;
    const float* lv[4] = {lvl0, lvl1, lvl2, lvl3};
    int rc = check_lookup_args("tcs_corr_lookup_encode", lv, coords, out, B, H, W1, W2, num_levels, radius, true);
    if (rc != 0) return rc;
    const int W2p = W2_pitch > 0 ? W2_pitch : W2;
    TCS_REQUIRE(W2p == W2 || (W2p > W2 && W2p % 16 == 0 && W2 % 8 == 0), TCS_E_SHAPE,
                "tcs_corr_lookup_encode: a row pitch (%d) other than W2 (%d) needs W2 %% 8 == 0 and a pitch that is a multiple of 16", W2p, W2);
    TCS_REQUIRE(num_levels == 4 && radius == 4, TCS_E_SHAPE, "tcs_corr_lookup_encode: implemented for num_levels=4, radius=4 (36 taps)");
    TCS_REQUIRE(weight != nullptr, TCS_E_BADARG, "tcs_corr_lookup_encode: null weight");
    TCS_REQUIRE(Cout > 0 && Cout <= kEncMaxOut && Cout % 4 == 0, TCS_E_SHAPE, "tcs_corr_lookup_encode: Cout=%d must be a multiple of 4, <= %d", Cout, kEncMaxOut);
    TCS_REQUIRE(B <= 65535, TCS_E_SHAPE, "tcs_corr_lookup_encode: B must be <= 65535");
    LevelPtrs lp;
    for (int l = 0; l < 4; ++l) lp.p[l] = lv[l];
    dim3 grid((unsigned)ceil_div(H * W1, kLookThreads), 1, B);
    TCS_ONCE_PER_DEVICE(
        const int carve = carveout_percent("TCS_CARVE_LOOKUP_ENC", 25);
        if (carve >= 0) {
            TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_lookup_encode_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_lookup_encode_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
            TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_lookup_encode_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        }
    );
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (W2p % 16 == 0 && (reinterpret_cast<uintptr_t>(lvl0) & 31) == 0)
        corr_lookup_encode_kernel<2><<<grid, kLookThreads, 0, s>>>(lp, coords, coords_bstride, weight, bias, out, H * W1, W2, W2p, Cout, relu);
    else if (W2p % 16 == 0)
        corr_lookup_encode_kernel<1><<<grid, kLookThreads, 0, s>>>(lp, coords, coords_bstride, weight, bias, out, H * W1, W2, W2p, Cout, relu);
    else
        corr_lookup_encode_kernel<0><<<grid, kLookThreads, 0, s>>>(lp, coords, coords_bstride, weight, bias, out, H * W1, W2, W2p, Cout, relu);
    TCS_CHECK_LAUNCH("tcs_corr_lookup_encode");
    return 0;
}

extern "C" int tcs_corr_encode_packed_bytes(void) { return tcs::enc_tc::kPackBytes; }

extern "C" int tcs_corr_encode_pack_weights(const float* weight, const float* bias, void* packed, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(weight != nullptr && packed != nullptr && aligned16(packed), TCS_E_BADARG, "tcs_corr_encode_pack_weights: null or unaligned pointer");
    encode_pack_weights_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(weight, bias, static_cast<unsigned char*>(packed));
    TCS_CHECK_LAUNCH("tcs_corr_encode_pack_weights");
    return 0;
}

extern "C" int tcs_corr_lookup_encode_tc(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                                         const float* coords, long long coords_bstride, const void* packed, float* out,
                                         int B, int H, int W1, int W2, int relu, int W2_pitch, void* stream) {
    using namespace tcs;
    const float* lv[4] = {lvl0, lvl1, lvl2, lvl3};
    int rc = check_lookup_args("tcs_corr_lookup_encode_tc", lv, coords, out, B, H, W1, W2, 4, 4, true);
    if (rc != 0) return rc;
    TCS_REQUIRE(packed != nullptr && aligned16(packed), TCS_E_BADARG, "tcs_corr_lookup_encode_tc: packed weights null or unaligned");
    const int W2p = W2_pitch > 0 ? W2_pitch : W2;
    TCS_REQUIRE(W2p % 16 == 0 && (W2p == W2 || (W2p > W2 && W2 % 8 == 0)), TCS_E_SHAPE,
                "tcs_corr_lookup_encode_tc: needs a row pitch (%d) that is a multiple of 16 (W2=%d); use tcs_corr_lookup_encode", W2p, W2);
    TCS_REQUIRE(B <= 65535, TCS_E_SHAPE, "tcs_corr_lookup_encode_tc: B must be <= 65535");
    LevelPtrs lp;
    for (int l = 0; l < 4; ++l) lp.p[l] = lv[l];
    TCS_ONCE_PER_DEVICE(
        TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_lookup_encode_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, enc_tc::kSmemBytes));
        TCS_CHECK_CUDA(cudaFuncSetAttribute(corr_lookup_encode_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, enc_tc::kSmemBytes));
    );
    const int tiles_per_sample = ceil_div(H * W1, enc_tc::kRows);
    const long long num_tiles = (long long)tiles_per_sample * B;
    TCS_REQUIRE(num_tiles < 0x7fffffffLL, TCS_E_SHAPE, "tcs_corr_lookup_encode_tc: too many pixels");
    int ctas_per_sm = TCS_ENCODE_TC_MINBLOCKS;                         // 64 registers x 256 threads; 29 KB of shared memory each
    { const char* e = getenv("TCS_ENCODE_TC_CTAS"); if (e != nullptr && atoi(e) > 0) ctas_per_sm = atoi(e); }
    const long long resident = (long long)ctas_per_sm * num_sms();
    const unsigned grid = (unsigned)(num_tiles < resident ? num_tiles : resident);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned char* pk = static_cast<const unsigned char*>(packed);
    if ((reinterpret_cast<uintptr_t>(lvl0) & 31) == 0)
        corr_lookup_encode_tc_kernel<2><<<grid, enc_tc::kThreads, enc_tc::kSmemBytes, s>>>(lp, coords, coords_bstride, pk, out, H * W1, W2, W2p, relu,
                                                                                        tiles_per_sample, (int)num_tiles);
    else
        corr_lookup_encode_tc_kernel<1><<<grid, enc_tc::kThreads, enc_tc::kSmemBytes, s>>>(lp, coords, coords_bstride, pk, out, H * W1, W2, W2p, relu,
                                                                                        tiles_per_sample, (int)num_tiles);
    TCS_CHECK_LAUNCH("tcs_corr_lookup_encode_tc");
    return 0;
}

extern "C" int tcs_corr_lookup_alt(const float* a_n32, const float* b0, const float* b1, const float* b2, const float* b3,
                                   const float* coords, long long coords_bstride, float* out,
                                   int B, int H, int W1, int W2, int C, int num_levels, int radius, void* stream) {
    using namespace tcs;
    const float* lv[4] = {b0, b1, b2, b3};
    int rc = check_lookup_args("tcs_corr_lookup_alt", lv, coords, out, B, H, W1, W2, num_levels, radius);
    if (rc != 0) return rc;
    TCS_REQUIRE(a_n32 != nullptr && aligned16(a_n32), TCS_E_ALIGN, "tcs_corr_lookup_alt: a_n32 null or unaligned");
    TCS_REQUIRE(C == 128 || C == 256 || C == 384 || C == 512, TCS_E_SHAPE, "tcs_corr_lookup_alt: C=%d must be 128, 256, 384 or 512", C);
    LevelPtrs lp;
    for (int l = 0; l < 4; ++l) lp.p[l] = lv[l];
    const long long npix = (long long)B * H * W1;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const unsigned grid = (unsigned)ceil_div_ll(npix, 8);
    switch (C / 128) {
        case 1: corr_lookup_alt_kernel<1><<<grid, 256, 0, s>>>(a_n32, lp, coords, coords_bstride, out, H * W1, W1, W2, num_levels, radius, npix); break;
        case 2: corr_lookup_alt_kernel<2><<<grid, 256, 0, s>>>(a_n32, lp, coords, coords_bstride, out, H * W1, W1, W2, num_levels, radius, npix); break;
        case 3: corr_lookup_alt_kernel<3><<<grid, 256, 0, s>>>(a_n32, lp, coords, coords_bstride, out, H * W1, W1, W2, num_levels, radius, npix); break;
        default: corr_lookup_alt_kernel<4><<<grid, 256, 0, s>>>(a_n32, lp, coords, coords_bstride, out, H * W1, W1, W2, num_levels, radius, npix); break;
    }
    TCS_CHECK_LAUNCH("tcs_corr_lookup_alt");
    return 0;
}

extern "C" int tcs_corr_argmax(const float* lvl0, float* sparse_disp, float* main_cost, float* mask,
                               int B, int H, int W1, int W2, float thres, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(lvl0 != nullptr && sparse_disp != nullptr && main_cost != nullptr && mask != nullptr, TCS_E_BADARG, "tcs_corr_argmax: null pointer");
    TCS_REQUIRE(B > 0 && H > 0 && W1 > 0 && W2 > 0, TCS_E_BADARG, "tcs_corr_argmax: non-positive size");
    TCS_REQUIRE(W2 <= 32 * kArgMaxPerLane, TCS_E_SHAPE, "tcs_corr_argmax: W2=%d exceeds %d", W2, 32 * kArgMaxPerLane);
    const long long nrows = (long long)B * H * W1;
    const unsigned grid = (unsigned)ceil_div_ll(nrows, 8);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (W2 <= 256) corr_argmax_kernel<8><<<grid, 256, 0, s>>>(lvl0, sparse_disp, main_cost, mask, W1, W2, thres, nrows);
    else if (W2 <= 512) corr_argmax_kernel<16><<<grid, 256, 0, s>>>(lvl0, sparse_disp, main_cost, mask, W1, W2, thres, nrows);
    else corr_argmax_kernel<32><<<grid, 256, 0, s>>>(lvl0, sparse_disp, main_cost, mask, W1, W2, thres, nrows);
    TCS_CHECK_LAUNCH("tcs_corr_argmax");
    return 0;
}

extern "C" int tcs_corr_cost_volume(const float* lvl0, float* out, int B, int H, int W1, int W2, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(lvl0 != nullptr && out != nullptr, TCS_E_BADARG, "tcs_corr_cost_volume: null pointer");
    TCS_REQUIRE(B > 0 && H > 0 && W1 > 0 && W2 > 0 && (long long)B * H <= 65535, TCS_E_SHAPE, "tcs_corr_cost_volume: bad sizes (B*H <= 65535)");
    dim3 grid(ceil_div(W2, 32), ceil_div(W1, 32), B * H);
    corr_cost_volume_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(lvl0, out, H, W1, W2);
    TCS_CHECK_LAUNCH("tcs_corr_cost_volume");
    return 0;
}

extern "C" int tcs_corr_lookup_backward(const float* grad_out, const float* coords, long long coords_bstride, float* grad_volume,
                                        int B, int H, int W1, int W2, int num_levels, int radius, void* stream) {
    using namespace tcs;
    TCS_REQUIRE(grad_out && coords && grad_volume, TCS_E_BADARG, "tcs_corr_lookup_backward: null pointer");
    TCS_REQUIRE(num_levels >= 1 && num_levels <= TCS_MAX_LEVELS && radius >= 0 && radius <= TCS_MAX_RADIUS, TCS_E_SHAPE,
                "tcs_corr_lookup_backward: num_levels in [1,%d], radius in [0,%d]", TCS_MAX_LEVELS, TCS_MAX_RADIUS);
    TCS_REQUIRE(B > 0 && H > 0 && W1 > 0 && (W2 >> (num_levels - 1)) >= 2, TCS_E_SHAPE, "tcs_corr_lookup_backward: bad sizes");
    const long long npix = (long long)B * H * W1;
    corr_lookup_backward_kernel<<<(unsigned)ceil_div_ll(npix, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        grad_out, coords, coords_bstride, grad_volume, H * W1, W2, num_levels, radius, npix);
    TCS_CHECK_LAUNCH("tcs_corr_lookup_backward");
    return 0;
}
