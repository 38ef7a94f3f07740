/*
 * tcs_b200.h — C-ABI of the B200-native cost-volume hot path of TC-Stereo.
 *
 * One shared library (libtcs_b200.so, hand-written sm_100a CUDA) exports the entry points below.
 * Every entry point
 *   - takes raw DEVICE pointers, plain ints/floats and the CUDA stream to launch on (as void*:
 *     a cudaStream_t), never a torch type;
 *   - allocates nothing and frees nothing: the caller owns every buffer, scratch included;
 *   - never synchronises the host with the device;
 *   - returns 0 on success, a negative TCS_E_* code for a rejected argument, or a positive
 *     cudaError_t value when a launch failed.  tcs_last_error() gives the text (thread-local).
 *
 * "ref:" comments cite the reference interface each function replaces, relative to the upstream
 * repository root (jiaxiZeng/Temporally-Consistent-Stereo-Matching).
 *
 * Layouts (all row-major, innermost last):
 *   fmap        [B, C, H, W]        fp32   the reference's NCHW feature maps
 *   operands    [B, H, W, C]        16-bit the L2-normalised features, channels last (K-major GEMM operands)
 *   level l     [B, H, W1, W2 >> l] fp32   correlation pyramid level l  (ref: corr_pyramid[l] viewed 4-D)
 *   lookup out  [B, L*(2r+1), H, W1] fp32  (ref: CorrBlock1D.__call__ result)
 */
#ifndef TCS_B200_H
#define TCS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define TCS_ABI_VERSION 10

/* argument errors (negative); CUDA launch errors are returned as positive cudaError_t values */
#define TCS_E_BADARG   (-1)   /* null pointer / non-positive size / unsupported combination */
#define TCS_E_SHAPE    (-2)   /* shape outside what the kernel supports (see each function)  */
#define TCS_E_ALIGN    (-3)   /* pointer not aligned as required                              */
#define TCS_E_DRIVER   (-4)   /* cuTensorMapEncodeTiled unavailable or failed                 */

/* precision of the tensor-core correlation build */
#define TCS_PREC_BF16    0    /* one bf16 pass, fp32 accumulate (north_star's fast mode)      */
#define TCS_PREC_BF16X3  1    /* hi/lo split: hi*hi + hi*lo + lo*hi, fp32-level accuracy      */
#define TCS_PREC_FP16    2    /* one fp16 pass on 2^8-scaled unit vectors                     */
#define TCS_PREC_FP16X3  3    /* fp16 hi/lo split                                             */

#define TCS_WARP_PER_SAMPLE_MEAN 1
#define TCS_WARP_DETERMINISTIC   2

#define TCS_MAX_LEVELS   4
#define TCS_MAX_RADIUS   8

int         tcs_abi_version(void);
const char* tcs_last_error(void);

/* ---- (1) correlation build ------------------------------------------------------------------ */

/* L2-normalise over channels and re-lay out NCHW fp32 -> channels-last 16-bit GEMM operands.
 * ref: core/corr.py:58-59 (F.normalize(fmap, dim=1), eps 1e-12).
 *   fmap  [B,C,H,W] fp32 (in)
 *   hi    [B,H,W,C] bf16/fp16 (out)            the rounded normalised feature (x 2^8 for fp16)
 *   lo    [B,H,W,C] bf16/fp16 (out, nullable)  the rounding residual, for the *X3 modes
 *   n32   [B,H,W,C] fp32 (out, nullable)       the fp32 normalised feature (alternate path operand)
 * prec selects bf16 vs fp16 rounding.  Requires C % 64 == 0, C <= 512. */
int tcs_corr_prepass(const float* fmap, void* hi, void* lo, float* n32,
                     int B, int C, int H, int W, int prec, void* stream);

/* The same 16-bit operands K-BLOCK-MAJOR, [B,H,C/64,W,64]: every 64-channel block of an image row is one contiguous
 * run of 128-byte pixel rows, so the TMA boxes of tcs_corr_lookup_alt_tc are whole contiguous lines.
 * ref: core/corr.py:58-59.  Same requirements as tcs_corr_prepass. */
int tcs_corr_prepass_kblocked(const float* fmap, void* hi, void* lo,
                              int B, int C, int H, int W, int prec, void* stream);

/* All-pairs 1-D correlation of one frame, all pyramid levels in one pass (tcgen05 + TMEM + TMA).
 * ref: core/corr.py:54-62 (CorrBlock1D.corr: einsum 'aijk,aijh->ajkh') and core/corr.py:15-23
 * (the avg_pool2d([1,2]) pyramid).  Level l+1 = 0.5*(even + odd column) of level l, floor on odd
 * widths, exactly as avg_pool2d does.
 *   a_hi,a_lo  [B,H,W1,C] operands of the left image  (lo nullable unless prec is *X3)
 *   b_hi,b_lo  [B,H,W2,C] operands of the right image
 *   lvl[l]     [B,H,W1,W2>>l] fp32 (out), l < num_levels <= 4; every level pointer 16-byte aligned.  lvl1 and / or lvl3 may be
 *              NULL: that level is then not stored (it is still the input of the next one).  tcs_corr_lookup at radius 4 with
 *              4 levels, tcs_corr_lookup_encode[_tc], tcs_corr_argmax and tcs_corr_cost_volume read levels 0 and 2 only — they
 *              re-pool the odd levels as (a + b) * 0.5, the expression used here — so a third of the pyramid need not exist
 *   W2_pitch   0 (dense rows) or the row pitch of level 0 in floats (level l: W2_pitch >> l), a multiple of 16 above W2
 *              with W2 % 8 == 0: the columns past W2 of every level are written as exact zeros, which puts widths like the
 *              KITTI shape's 312 on tcs_corr_lookup's predicate-free path (rows that start on 16-byte boundaries)
 * Requires C % 64 == 0, W1 >= 1, W2 >= 8. */
int tcs_corr_build(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo,
                   float* lvl0, float* lvl1, float* lvl2, float* lvl3,
                   int B, int H, int W1, int W2, int C, int num_levels, int prec, int W2_pitch, void* stream);

/* The same result in ONE kernel straight from the fp32 NCHW feature maps: normalisation, the 16-bit hi/lo
 * split and the K-major operand layout happen on chip, so the operands never make a round trip through HBM
 * (halves the build's traffic).  ref: core/corr.py:54-62 + :15-23.
 *   fmap1 [B,C,H,W1], fmap2 [B,C,H,W2] fp32;  lvl[l] as for tcs_corr_build.
 * Requires C % 32 == 0, 8 <= W2 <= 240, W1 <= 256, W1 and W2 multiples of 4 and 16-byte aligned maps
 * (TCS_E_SHAPE otherwise: use the two-step path). */
int tcs_corr_build_fused(const float* fmap1, const float* fmap2,
                         float* lvl0, float* lvl1, float* lvl2, float* lvl3,
                         int B, int H, int W1, int W2, int C, int num_levels, int prec, void* stream);

/* Exact-fp32 CUDA-core build (no tensor cores): the strict-parity mode and the in-library check of
 * the tensor-core path.  Same outputs as tcs_corr_build; operands are the fp32 normalised
 * channels-last features written by tcs_corr_prepass (n32). */
int tcs_corr_build_fp32(const float* a_n32, const float* b_n32,
                        float* lvl0, float* lvl1, float* lvl2, float* lvl3,
                        int B, int H, int W1, int W2, int C, int num_levels, void* stream);

/* ---- (2) fused pyramid lookup ----------------------------------------------------------------- */

/* ref: core/corr.py:33-52 (CorrBlock1D.__call__) + core/utils/utils.py:82-97 (bilinear_sampler ->
 * F.grid_sample, align_corners=True, zeros padding), including grid_sample's normalise /
 * un-normalise fp32 round trip.
 *   lvl[l]  as written by tcs_corr_build; each level must be readable up to the next 16-byte
 *           boundary past its end (the Python wrapper pads).  With num_levels == 4 and radius == 4 only levels 0 and 2 are
 *           read (1 and 3 are re-pooled from them on the fly, bit for bit): lvl1 and lvl3 may then be NULL — here, in
 *           tcs_corr_lookup_encode and in tcs_corr_lookup_encode_tc
 *   coords  fp32, x coordinate of pixel (b,h,w1) at  coords[b*coords_bstride + h*W1 + w1]
 *           (coords_bstride lets the caller pass channel 0 of a [B,2,H,W1] tensor)
 *   out     [B, num_levels*(2r+1), H, W1] fp32
 *   W2_pitch  0, or the row pitch the levels were built with (see tcs_corr_build; 4 levels and radius 4 only).  The
 *           sampling arithmetic (normalisation by W2 - 1, zero padding from W2 on) always uses W2. */
int tcs_corr_lookup(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                    const float* coords, long long coords_bstride, float* out,
                    int B, int H, int W1, int W2, int num_levels, int radius, int W2_pitch, void* stream);

/* Lookup fused with the motion encoder's first layer: out = [relu](W . taps + bias), the 36 taps never leave
 * the registers.  ref: core/update.py:97,104 (BasicMotionEncoder.convc1, a 1x1 Conv2d(36, 64), then F.relu)
 * applied to core/corr.py:33-52.  (the first "next" row after the path itself.)
 *   weight [Cout, 36] fp32 (the conv weight viewed 2-D), bias [Cout] (nullable), out [B, Cout, H, W1] fp32.
 * Requires num_levels == 4, radius == 4, Cout % 4 == 0, Cout <= 128. */
int tcs_corr_lookup_encode(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                           const float* coords, long long coords_bstride, const float* weight, const float* bias,
                           float* out, int B, int H, int W1, int W2, int num_levels, int radius, int Cout, int relu,
                           int W2_pitch, void* stream);

/* The same on the tensor cores, for the model's own shape (Cout = 64) and row-aligned levels (row pitch % 16 == 0): the 36 -> 64
 * product of 128 pixels is one 128 x 64 x 48 tcgen05 GEMM in fp16 hi/lo split (hi.hi + hi.lo + lo.hi, fp32 accumulation in TMEM),
 * the taps written as the A operand by the threads that sampled them.  `packed` (tcs_corr_encode_packed_bytes() bytes, 16-byte
 * aligned) is written once per weight by tcs_corr_encode_pack_weights(weight [64,36] fp32, bias [64] or NULL): the weights scaled
 * by a power of two, split and laid out as the K-major SWIZZLE_64B B operand, then the bias and the un-scaling factor.
 * ref: core/update.py:97,104 on core/corr.py:33-52, as tcs_corr_lookup_encode. */
int tcs_corr_encode_packed_bytes(void);
int tcs_corr_encode_pack_weights(const float* weight, const float* bias, void* packed, void* stream);
int tcs_corr_lookup_encode_tc(const float* lvl0, const float* lvl1, const float* lvl2, const float* lvl3,
                              const float* coords, long long coords_bstride, const void* packed, float* out,
                              int B, int H, int W1, int W2, int relu, int W2_pitch, void* stream);

/* ---- (3) alternate (on-the-fly) lookup ---------------------------------------------------------- */

/* Average-pool the fp32 normalised right features along W (channels last): out[b,h,j,:] =
 * 0.5*(in[b,h,2j,:] + in[b,h,2j+1,:]).  Pooling the features == pooling the volume (linearity). */
int tcs_fmap_pool_w(const float* in, float* out, int B, int H, int W, int C, void* stream);

/* Same result as tcs_corr_lookup without a materialised volume: dot products only at the taps,
 * warp-level reductions.  (New capability; the reference has no such path — its contract is
 * "equals CorrBlock1D.__call__", ref: core/corr.py:33-52.)
 *   a_n32     [B,H,W1,C]      fp32 normalised left features
 *   b_n32[l]  [B,H,W2>>l,C]   fp32 normalised right features pooled l times */
int tcs_corr_lookup_alt(const float* a_n32,
                        const float* b0_n32, const float* b1_n32, const float* b2_n32, const float* b3_n32,
                        const float* coords, long long coords_bstride, float* out,
                        int B, int H, int W1, int W2, int C, int num_levels, int radius, void* stream);

/* The same on the tensor cores (tcgen05 + TMEM + TMA), 4 levels, radius 4: per 128-pixel tile the band of the row's
 * correlation block that the tile's coordinates can touch is built into TMEM from the 16-bit operands of
 * tcs_corr_prepass and the 36 taps are sampled in the epilogue; nothing of the volume is written.  Bit-identical to
 * tcs_corr_build (same precision) + tcs_corr_lookup.  Contract: ref core/corr.py:33-52.
 *   a_hi,a_lo [B,H,C/64,W1,64], b_hi,b_lo [B,H,C/64,W2,64] K-block-major operands of tcs_corr_prepass_kblocked
 *   (lo nullable unless prec is *X3); out [B,36,H,W1].
 * Requires C % 64 == 0, W2 >= 16. */
int tcs_corr_lookup_alt_tc(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo,
                           const float* coords, long long coords_bstride, float* out,
                           int B, int H, int W1, int W2, int C, int prec, void* stream);

/* ---- first-frame initialisation ------------------------------------------------------------------ */

/* ref: core/corr.py:67-79 (CorrBlock1D.argmax_disp) on the masked volume of core/corr.py:25-31.
 *   lvl0 [B,H,W1,W2] -> sparse_disp, main_cost, mask : each [B,1,H,W1] fp32.  thres: 0.3 in the reference. */
int tcs_corr_argmax(const float* lvl0, float* sparse_disp, float* main_cost, float* mask,
                    int B, int H, int W1, int W2, float thres, void* stream);

/* ref: core/corr.py:25-31,64-65 (get_cost_volume): out[b,w2,h,w1] = lvl0[b,h,w1,w2] * (w2 <= w1). */
int tcs_corr_cost_volume(const float* lvl0, float* out, int B, int H, int W1, int W2, void* stream);

/* ---- (4) temporal step -------------------------------------------------------------------------- */

/* Scratch sizes (bytes) the caller must provide to tcs_warp_forward. */
long long tcs_warp_scratch_bytes(int B, int C, int H, int W);

/* Forward-warp the previous frame's disparity and features into the current view.
 * ref: core/utils/geo_utils.py:158-198 (warp) with helpers :7-57,:135-145, core/utils/utils.py:100-103,
 * core/utils/splatting/softsplat.py:232-274 (softsplat 'soft-clipeps') and :284-335 (softsplat_out).
 *   disp      [B,1,H,W]  previous disparity (>= 0)
 *   fmap      [B,C,H,W]  previous left features
 *   rel_T     [B,4,4]    previous->current camera transform;  K, K_inv [B,3,3];  baseline [B]
 *   cur_fmap  [B,C,H,W]  current left features (nullable: then cost is not computed)
 *   out_disp  [B,1,H,W], out_fmap [B,C,H,W] (nullable: the model reads only the cost of the warped features,
 *             core/tc_stereo.py:139, so a caller that wants just the cost saves the 1 KB/pixel store),
 *             out_mask [B,1,H,W], out_cost [B,1,H,W] (nullable)
 *   out_cost = sum_c normalize(cur_fmap)*normalize(out_fmap) * out_mask   (ref: core/tc_stereo.py:139-140)
 *   flags     TCS_WARP_PER_SAMPLE_MEAN: each sample's own mean disparity for the soft-splat metric instead of the
 *             reference's batch-global mean (geo_utils.py:193) — for batching independent sequences.
 *             TCS_WARP_DETERMINISTIC: collect the splat from the target's side through per-target contributor lists
 *             sorted by source pixel (fixed summation order, no accumulator, no floating-point atomics; any flow).
 *   cur_t_out [B*H*W][C] (nullable; cost-only call, i.e. out_fmap null and out_cost given): cur_fmap as pixel-major rows
 *             in the library's private channel order, written while its tile is in shared memory for the cost anyway.
 *   fmap_t    (nullable; TCS_WARP_DETERMINISTIC only) the cur_t_out of the call in which `fmap` was cur_fmap, i.e. the
 *             previous frame's call (core/tc_stereo.py carries fmap1 to the next frame as last_fmap1): the list
 *             formulation then skips its transposition of fmap.  The caller vouches that fmap is unchanged since.
 *   scratch   tcs_warp_scratch_bytes() bytes, 16-byte aligned. */
int tcs_warp_forward(const float* disp, const float* fmap, const float* rel_T, const float* K,
                     const float* K_inv, const float* baseline, const float* cur_fmap,
                     float* out_disp, float* out_fmap, float* out_mask, float* out_cost,
                     const float* fmap_t, float* cur_t_out,
                     void* scratch, int B, int C, int H, int W, int flags, void* stream);

/* ref: core/utils/geo_utils.py:148-155 (cal_relative_transformation): out = T2 * inv(T1), batched 4x4 world2cam
 * poses, computed in fp64 and rounded once (the reference: fp32 LU with a host sync on `info`, fp32 matmul).
 *   T1, T2, out  [B,4,4] fp32 row-major. */
int tcs_relative_pose(const float* T1, const float* T2, float* out, int B, void* stream);

/* ref: core/utils/geo_utils.py:201-236 (get_backward_grid).  disp [B,1,H,W] -> grid [B,2,H,W] (x,y). */
int tcs_backward_grid(const float* disp, const float* rel_T, const float* K, const float* K_inv,
                      const float* baseline, float* grid, int B, int H, int W, void* stream);

/* ref: core/utils/utils.py:82-97 (bilinear_sampler: grid_sample bilinear, zeros, align_corners=True,
 * pixel coordinates).  img [B,C,Hi,Wi]; grid_xy [B,2,Ho,Wo] planar (x plane then y plane);
 * out [B,C,Ho,Wo]. */
int tcs_bilinear_sample(const float* img, const float* grid_xy, float* out,
                        int B, int C, int Hi, int Wi, int Ho, int Wo, void* stream);

/* ref: core/tc_stereo.py:163: grid <- 0.5 * F.interpolate(grid, scale_factor=0.5, 'bilinear',
 * align_corners=True).  in [B,2,H,W] -> out [B,2,H/2,W/2]. */
int tcs_grid_halve(const float* in, float* out, int B, int H, int W, void* stream);

/* ref: core/tc_stereo.py:159-163, the whole loop in one launch: the three hidden-state maps net[l] [B,C_l,H>>l,W>>l] are
 * sampled (tcs_bilinear_sample's arithmetic) at the backward grid [B,2,H,W] halved l times (tcs_grid_halve's arithmetic,
 * evaluated on the fly), so the results are bit-identical to the chain of three samples and two halvings.
 *   out[l] [B,C_l,H>>l,W>>l].  Requires H, W >= 4. */
int tcs_warp_hidden_states(const float* net0, const float* net1, const float* net2, const float* grid,
                           float* out0, float* out1, float* out2, int B, int C0, int C1, int C2, int H, int W,
                           void* stream);

/* ---- (5) "next" row: the per-GRU-iteration 3x3 stencils on the disparity ---------- */

/* ref: core/utils/geo_utils.py:115-132 (disp2disp_gradient_xy).  disp [N,1,H,W] -> grads [N,2,H,W] (forward
 * differences in x and y on the replicate-padded map) and edge_mask [N,1,H,W] bytes (|gx| < 5 && |gy| < 5; nullable). */
int tcs_disp_gradient_xy(const float* disp, float* grads, unsigned char* edge_mask, int N, int H, int W, void* stream);

/* ref: core/utils/geo_utils.py:73-101 (disp2disp_grad_candidates).  disp [N,1,H,W] -> out [N,2,8*levels,H,W]:
 * -n_xy / n_z of the normals of the triangles (centre, ring[k], ring[k+2]) on the zero-padded map, rings at distance
 * 1..levels concatenated before the pairing.  levels 1..4. */
int tcs_disp_grad_candidates(const float* disp, float* out, int N, int H, int W, int levels, void* stream);

/* ref: core/update.py:259-289 (DispRefine.propagate_disparity).  grad [N,2,H,W], disp [N,1,H,W] -> prop [N,9,H,W]
 * (the neighbour's disparity extrapolated along its gradient to the centre) and matrix [N,18,H,W] (|grad_c - grad_nb|,
 * channel = component * 9 + neighbour). */
int tcs_disp_propagate(const float* grad, const float* disp, float* prop, float* matrix, int N, int H, int W, void* stream);

/* ref: core/tc_stereo.py:75-88 (TCStereo.upsample_flow).  flow [N,D,H,W], mask
 * [N,9*factor^2,H,W] -> out [N,D,factor*H,factor*W]: softmax over the 9 logits of each sub-pixel, convex combination of
 * the 3x3 zero-padded neighbours of (factor *) flow.  factor 2, 4 or 8; scale != 0 multiplies flow by factor. */
int tcs_convex_upsample(const float* flow, const float* mask, float* out, int N, int D, int H, int W,
                        int factor, int scale, void* stream);

/* ---- (7) "next" row: backward of the lookup, for training ------------------------------------ */

/* ref: the autograd of core/corr.py:33-52 (grid_sample's gradient w.r.t. its input; coords are detached, tc_stereo.py:176)
 * folded through core/corr.py:21-23 (avg_pool2d's gradient): d(lookup output) -> d(volume), every element written once.
 *   grad_out [B, num_levels*(2r+1), H, W1] fp32, coords as for tcs_corr_lookup, grad_volume [B,H,W1,W2] fp32 (out, dense). */
int tcs_corr_lookup_backward(const float* grad_out, const float* coords, long long coords_bstride, float* grad_volume,
                             int B, int H, int W1, int W2, int num_levels, int radius, void* stream);

/* ref: train_stereo.py:150-172 (init_loss: the per-pixel terms of the cost-volume initialisation loss), on level 0 of the pyramid
 * instead of the masked transposed cost volume of core/corr.py:25-31 (cv[b,w2,h,w1] = level0[b,h,w1,w2] * [w2 <= w1]):
 *   phi[b,h,w1]       = frac * rho(df + 1) + (1 - frac) * rho(df), df = floor(index_gt), rho(i) = cv[clip(i, 0, W2-1)]   (:151-158)
 *   cost_nm[b,j,h,w1] = j-th largest entry along w2 of cv with [index_gt - 1.5, index_gt + 1.5) and every column of a pixel
 *                       whose mask is 0 filled with 0 (:166-171), j < k <= 8
 *   idx_nm[b,j,h,w1]  = its w2, or -1 when it is a filled / masked zero (no gradient reaches the volume through it)
 * level0 [B,H,W1,W2_pitch] fp32 (W2_pitch 0 = dense), index_gt [B,H,W1] fp32 already clipped to [0, W2-1] (:164), mask [B,H,W1]
 * bytes (:162-163).  W2 <= 512.  Ties go to the lowest w2. */
int tcs_init_loss_forward(const float* level0, int W2_pitch, const float* index_gt, const unsigned char* mask,
                          float* phi, float* cost_nm, int* idx_nm, int B, int H, int W1, int W2, int k, void* stream);

/* The autograd of the above (torch.gather / masked_fill / topk backward in the reference): grad_level0 [B,H,W1,W2] fp32, dense,
 * zero-filled here, then (1 - frac) g_phi and frac g_phi at the two gathered entries and grad_cost_nm at idx_nm >= 0. */
int tcs_init_loss_backward(const float* grad_phi, const float* grad_cost_nm, const float* index_gt, const int* idx_nm,
                           float* grad_level0, int B, int H, int W1, int W2, int k, void* stream);

/* ---- (6) "next" row: input stems of the disparity completion network ------------------ */

/* ref: core/update.py:312-323,375-378 (DisparityCompletor: conv_disp_stem, conv_cost_stem, conv_mask_stem, the cat and
 * conv_disp_fuse): a per-pixel 3 -> 64 perceptron, eight 1x1 convolutions in the reference, one kernel here.
 *   disp, cost, mask [N,1,H,W] fp32 (the stems' inputs, i.e. disp/10 and mask-0.5 as update.py:373-374 prepares them);
 *   weights: tcs_completor_stems_weight_floats() floats, 16-byte aligned, packed in this order, matrices TRANSPOSED to
 *   [input][output]: w1d[64] b1d[64] W2d[64][64] b2d[64] | w1c[32] b1c[32] W2c[32][32] b2c[32] | w1m[32] b1m[32] W2m[32][32]
 *   b2m[32] | W3[128][128] b3[128] | W4[128][64] b4[64];   out [N,64,H,W] fp32. */
int tcs_completor_stems_weight_floats(void);
int tcs_completor_stems(const float* disp, const float* cost, const float* mask, const float* weights, float* out,
                        int N, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TCS_B200_H */
