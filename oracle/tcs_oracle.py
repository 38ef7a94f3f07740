"""CPU oracle of TC-Stereo's cost-volume hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy (fp32, every operation single-rounded; FMA only where stated) restatement of what the reference computes on this path, written
from the reference's source and pinned against outputs of the reference itself (tests/golden/*.npz, made
by tests/golden/make_golden.py which imports /root/reference).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the product
(temporally-consistent-stereo-matching_b200/) never does and has no CPU path.

Pinning status: the reference ships no tests or golden vectors, so the pin is
"outputs of the reference's own PyTorch code run in the build container on seeded inputs".  One piece of
the reference cannot run anywhere without cupy — the soft-splat CUDA kernel (softsplat.py:284-335) — so
for `softsplat_scatter` the golden generator substitutes this file's restatement of that kernel inside the
reference's own `softsplat`/`warp` code; the scatter arithmetic itself is therefore pinned only by reading
the kernel source (it is 20 lines of index arithmetic), everything around it by execution.

Each function cites the reference file:line it follows (paths relative to the upstream repository root).
"""
import numpy as np

F32 = np.float32


def _f(x):
    return np.asarray(x, dtype=F32)


# ------------------------------------------------------------------------------------------------------
# (1) correlation build + pyramid
# ------------------------------------------------------------------------------------------------------

def normalize_features(fmap):
    """F.normalize(fmap, dim=1): x / max(||x||_2, 1e-12).  ref: core/corr.py:58-59."""
    fmap = _f(fmap)
    nrm = np.sqrt(np.sum(fmap.astype(np.float64) ** 2, axis=1, keepdims=True)).astype(F32)
    return fmap / np.maximum(nrm, F32(1e-12))


def corr_volume(fmap1, fmap2, dtype=np.float32):
    """einsum('aijk,aijh->ajkh') of the normalised features -> [B,H,W1,W2].  ref: core/corr.py:54-62.
    dtype=np.float64 gives the exact-arithmetic yardstick used to rank fp32 / bf16x3 / bf16 errors."""
    n1 = normalize_features(fmap1).astype(dtype)
    n2 = normalize_features(fmap2).astype(dtype)
    a = np.ascontiguousarray(n1.transpose(0, 2, 3, 1))       # [B,H,W1,C]
    b = np.ascontiguousarray(n2.transpose(0, 2, 1, 3))       # [B,H,C,W2]
    return np.matmul(a, b)


def corr_pyramid(volume, num_levels=4):
    """avg_pool2d(corr, [1,2], stride=[1,2]) cascade along w2, floor on odd widths.  ref: core/corr.py:18-23
    (the reference appends num_levels extra entries; only the first num_levels are ever read, :39-40)."""
    levels = [_f(volume)]
    for _ in range(num_levels - 1):
        v = levels[-1]
        w = (v.shape[-1] // 2) * 2
        levels.append((v[..., 0:w:2] + v[..., 1:w:2]) * F32(0.5))
    return levels


def masked_cost_volume(volume):
    """[B,H,W1,W2] -> [B,W2,H,W1] multiplied by the mask [w2 <= w1].  ref: core/corr.py:25-31."""
    v = _f(volume).transpose(0, 3, 1, 2)
    W2, W1 = v.shape[1], v.shape[3]
    mask = (np.arange(W1)[None, None, None, :] >= np.arange(W2)[None, :, None, None]).astype(F32)
    return v * mask


# ------------------------------------------------------------------------------------------------------
# (2) pyramid lookup
# ------------------------------------------------------------------------------------------------------

def _unnormalized_x(x, size):
    """bilinear_sampler's 2*x/(W-1)-1 (core/utils/utils.py:86) followed by grid_sample's
    align_corners=True un-normalisation ((g+1)/2)*(W-1), each step rounded to fp32."""
    wm1 = F32(size - 1)
    g = (F32(2.0) * x) / wm1 - F32(1.0)
    return ((g + F32(1.0)) * F32(0.5)) * wm1


def corr_lookup(levels, coords, radius=4):
    """CorrBlock1D.__call__: linear interpolation at coords/2^l + k, zeros padding.  ref: core/corr.py:33-52,
    core/utils/utils.py:82-97.   levels[l] [B,H,W1,W2_l], coords [B,>=1,H,W1] -> [B, L*(2r+1), H, W1]."""
    coords = _f(coords)[:, 0]                                  # corr.py:35
    B, H, W1 = coords.shape
    outs = []
    taps = np.arange(-radius, radius + 1, dtype=F32)           # linspace(-r, r, 2r+1)
    for l, lv in enumerate(levels):
        lv = _f(lv)
        Wl = lv.shape[-1]
        x = taps[None, None, None, :] + (coords / F32(2 ** l))[..., None]     # corr.py:43
        ix = _unnormalized_x(x.astype(F32), Wl)
        x0f = np.floor(ix)
        w1 = ix - x0f
        w0 = (x0f + F32(1.0)) - ix
        with np.errstate(invalid="ignore"):
            x0 = np.clip(x0f, -2, Wl + 1).astype(np.int64)
        x1 = x0 + 1
        in0 = (x0 >= 0) & (x0 < Wl)
        in1 = (x1 >= 0) & (x1 < Wl)
        v0 = np.take_along_axis(lv, np.clip(x0, 0, Wl - 1), axis=-1) * in0
        v1 = np.take_along_axis(lv, np.clip(x1, 0, Wl - 1), axis=-1) * in1
        outs.append((v0.astype(F32) * w0 + v1.astype(F32) * w1).astype(F32))
    out = np.concatenate(outs, axis=-1)                        # [B,H,W1,L*(2r+1)]
    return np.ascontiguousarray(out.transpose(0, 3, 1, 2))


def corr_lookup_encoded(levels, coords, weight, bias=None, relu=True, radius=4):
    """relu(conv1x1(lookup)): the first layer of the motion encoder on the lookup result.
    ref: core/update.py:97,104 (BasicMotionEncoder.convc1 = Conv2d(36, 64, 1); cor = F.relu(convc1(corr)))."""
    taps = corr_lookup(levels, coords, radius).astype(np.float64)          # [B,36,H,W]
    w = np.asarray(weight, np.float64).reshape(weight.shape[0], -1)
    out = np.einsum("ok,bkhw->bohw", w, taps)
    if bias is not None:
        out = out + np.asarray(bias, np.float64)[None, :, None, None]
    if relu:
        out = np.maximum(out, 0.0)
    return out.astype(F32)


def corr_lookup_alternate(fmap1, fmap2, coords, num_levels=4, radius=4):
    """The on-the-fly path's contract: same result as corr_lookup without a volume, using
    pool(volume) == <n1, pool(n2)> (pooling happens after normalisation).  Not in the reference."""
    n1 = normalize_features(fmap1)
    n2 = normalize_features(fmap2)
    levels = []
    for _ in range(num_levels):
        a = np.ascontiguousarray(n1.transpose(0, 2, 3, 1))
        b = np.ascontiguousarray(n2.transpose(0, 2, 1, 3))
        levels.append(np.matmul(a, b))
        w = (n2.shape[-1] // 2) * 2
        n2 = (n2[..., 0:w:2] + n2[..., 1:w:2]) * F32(0.5)
    return corr_lookup(levels, coords, radius)


# ------------------------------------------------------------------------------------------------------
# first-frame initialisation
# ------------------------------------------------------------------------------------------------------

def argmax_disp(volume, thres=0.3):
    """Winner-take-all with a uniqueness test on the masked volume.  ref: core/corr.py:67-79.
    Ties resolve to the lowest index (torch.max).  -> sparse_disp, main_cost, mask, each [B,1,H,W1]."""
    cv = masked_cost_volume(volume)                            # [B,W2,H,W1]
    B, W2, H, W1 = cv.shape
    idx = np.argmax(cv, axis=1)[:, None]                       # first maximum, like torch.max
    main = np.take_along_axis(cv, idx, axis=1)
    k = np.arange(W2, dtype=np.float64)[None, :, None, None]
    near = (k >= idx - 1.5) & (k < idx + 1.5)                  # corr.py:71
    sub = np.where(near, F32(0.0), cv).max(axis=1, keepdims=True)
    mask = ((main - sub) > F32(thres)).astype(F32)
    disp = (np.arange(W1)[None, None, None, :] - idx).astype(F32)
    return disp * mask, main * mask, mask


# ------------------------------------------------------------------------------------------------------
# (4) temporal step: geometry
# ------------------------------------------------------------------------------------------------------

def _fma(a, b, c):
    """fp32 fused multiply-add: the product is exact in fp64, so one fp64 add and one rounding to fp32
    reproduce fmaf (up to double-rounding ties, probability ~2^-29 per operation)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F32)


def _dot3(m, x, y, z):
    """One row of a 3-vector product as the FMA chain a GEMM micro-kernel runs over k = 0, 1, 2:
    acc = m0*x; acc = fma(m1, y, acc); acc = fma(m2, z, acc).  Bit-identical to torch.matmul on the CPU
    for these 3x3 / 4x4 products (checked while generating tests/golden)."""
    acc = (m[0] * x).astype(F32)
    acc = _fma(m[1], y, acc)
    return _fma(m[2], z, acc)


def _finite_or_m1(v):
    return np.where(np.isnan(v) | np.isinf(v), F32(-1.0), v).astype(F32)


def _project(disp_clipped, rel_T, K, K_inv, baseline):
    """disp -> depth -> 3-D point -> rigid transform.  ref: geo_utils.py:7-16 (disp2depth), :32-42
    (pixel2point), :135-145 (relative_transform).  Returns bf [B] and the transformed point [3][B,H,W]."""
    B, _, H, W = disp_clipped.shape
    K, K_inv, rel_T = _f(K), _f(K_inv), _f(rel_T)
    bf = (_f(baseline).reshape(B) * K[:, 0, 0]).astype(F32)
    depth = bf[:, None, None] / disp_clipped[:, 0]
    xs = np.broadcast_to(np.arange(W, dtype=F32)[None, None, :], (B, H, W))
    ys = np.broadcast_to(np.arange(H, dtype=F32)[None, :, None], (B, H, W))
    one = F32(1.0)
    q = [depth * _dot3([K_inv[:, i, j][:, None, None] for j in range(3)], xs, ys, one) for i in range(3)]
    P = [(_dot3([rel_T[:, i, j][:, None, None] for j in range(3)], q[0], q[1], q[2]) + rel_T[:, i, 3][:, None, None]).astype(F32)
         for i in range(3)]
    return bf, [p.astype(F32) for p in P]


def _reproject(P, K):
    """(K P)_{0,1} / z with NaN/Inf -> -1.  ref: geo_utils.py:45-57 (point2pixel)."""
    K = _f(K)
    with np.errstate(divide="ignore", invalid="ignore"):
        u = _dot3([K[:, 0, j][:, None, None] for j in range(3)], P[0], P[1], P[2]) / P[2]
        v = _dot3([K[:, 1, j][:, None, None] for j in range(3)], P[0], P[1], P[2]) / P[2]
    return _finite_or_m1(u), _finite_or_m1(v)


def warp_geometry(disp, rel_T, K, K_inv, baseline):
    """First half of warp(): where every previous-frame pixel lands.  ref: geo_utils.py:170-192.
    -> disp' [B,H,W], target x, target y (x + flow, as the splat kernel forms them), valid (bool)."""
    disp = _f(disp)
    B, _, H, W = disp.shape
    bf, P = _project(np.maximum(disp, F32(0.001)), rel_T, K, K_inv, baseline)
    with np.errstate(divide="ignore", invalid="ignore"):
        d1 = _finite_or_m1(bf[:, None, None] / P[2])          # depth2disp, geo_utils.py:19-29
    valid = (d1 > 0) & (d1 < F32(W))                          # geo_utils.py:185
    u, v = _reproject(P, K)
    xs = np.arange(W, dtype=F32)[None, None, :]
    ys = np.arange(H, dtype=F32)[None, :, None]
    tx = xs + (u - xs)                                        # flow = coords' - coords0; target = x + flow
    ty = ys + (v - ys)
    return d1, tx.astype(F32), ty.astype(F32), valid


def softsplat_scatter(ten_in, target_x, target_y):
    """The forward soft-splat kernel: every source pixel adds in*w into the four integer neighbours of its
    target, bounds-checked, non-finite targets skipped.  ref: softsplat.py:284-335 (softsplat_out).
    ten_in [B,C,H,W]; target_x/y [B,H,W] = x + flow_x, y + flow_y.  -> [B,C,H,W]."""
    ten_in = _f(ten_in)
    B, C, H, W = ten_in.shape
    out = np.zeros((B, H * W, C), dtype=F32)
    src = ten_in.transpose(0, 2, 3, 1).reshape(B, H * W, C)
    fx = _f(target_x).reshape(B, H * W)
    fy = _f(target_y).reshape(B, H * W)
    finite = np.isfinite(fx) & np.isfinite(fy)
    fxs = np.where(finite, fx, F32(0))
    fys = np.where(finite, fy, F32(0))
    nwx = np.floor(fxs)
    nwy = np.floor(fys)
    sex, sey = nwx + F32(1), nwy + F32(1)
    corners = [(nwx, nwy, (sex - fxs) * (sey - fys)),          # north-west   softsplat.py:314
               (sex, nwy, (fxs - nwx) * (sey - fys)),          # north-east   :315
               (nwx, sey, (sex - fxs) * (fys - nwy)),          # south-west   :316
               (sex, sey, (fxs - nwx) * (fys - nwy))]          # south-east   :317
    for b in range(B):
        for cx, cy, wgt in corners:
            ok = finite[b] & (cx[b] >= 0) & (cx[b] < W) & (cy[b] >= 0) & (cy[b] < H)
            tgt = (cy[b][ok].astype(np.int64) * W + cx[b][ok].astype(np.int64))
            np.add.at(out[b], tgt, src[b][ok] * wgt[b][ok][:, None].astype(F32))
    return np.ascontiguousarray(out.reshape(B, H, W, C).transpose(0, 3, 1, 2))


def warp(disp, fmap, rel_T, K, K_inv, baseline, per_sample_mean=False):
    """Forward-warp disparity and features into the current view.  ref: geo_utils.py:158-198 (warp),
    softsplat.py:232-274 (softsplat, 'soft-clipeps').  -> disp' [B,1,H,W], fmap' [B,C,H,W], mask [B,1,H,W]."""
    fmap = _f(fmap)
    d1, tx, ty, valid = warp_geometry(disp, rel_T, K, K_inv, baseline)
    if per_sample_mean:
        mean = d1.astype(np.float64).mean(axis=(1, 2), keepdims=True).astype(F32)
    else:
        mean = F32(d1.astype(np.float64).mean())               # geo_utils.py:193 (batch-global mean)
    metric = np.clip(d1 - mean, F32(-50), F32(50))
    e = np.exp(metric.astype(np.float64)).astype(F32)[:, None]
    vm = valid.astype(F32)[:, None]
    feats = np.concatenate([d1[:, None], fmap], axis=1)        # geo_utils.py:195
    ten_in = np.concatenate([(feats * vm) * e, e * vm], axis=1)   # softsplat.py:236,250
    out = softsplat_scatter(ten_in, tx, ty)
    norm = out[:, -1:]
    mask = (norm != 0).astype(F32)                             # softsplat.py:258
    out = out[:, :-1] / np.maximum(norm, F32(1e-7))            # softsplat.py:267-271
    return out[:, :1], out[:, 1:], mask


def matching_cost(fmap1, warped_fmap1, mask):
    """sum_c normalize(fmap1) * normalize(warped) * mask.  ref: core/tc_stereo.py:139-140."""
    c = np.sum(normalize_features(fmap1) * normalize_features(warped_fmap1), axis=1, keepdims=True, dtype=np.float64)
    return c.astype(F32) * _f(mask)


def backward_grid(disp, rel_T, K, K_inv, baseline):
    """Where each current pixel was in the previous frame.  ref: geo_utils.py:201-236.  -> [B,2,H,W]."""
    disp = np.maximum(_f(disp), F32(0.01))                     # geo_utils.py:217
    _, P = _project(np.maximum(disp, F32(0.001)), rel_T, K, K_inv, baseline)
    u, v = _reproject(P, K)
    ok = P[2] > 0                                              # geo_utils.py:229
    return np.stack([np.where(ok, u, F32(-1)), np.where(ok, v, F32(-1))], axis=1).astype(F32)


def bilinear_sample(img, grid_xy):
    """grid_sample(bilinear, zeros, align_corners=True) on pixel coordinates.  ref: core/utils/utils.py:82-97.
    img [B,C,Hi,Wi], grid_xy [B,2,Ho,Wo] -> [B,C,Ho,Wo]."""
    img = _f(img)
    g = _f(grid_xy)
    B, C, Hi, Wi = img.shape
    ix = _unnormalized_x(g[:, 0], Wi)
    if Hi > 1:
        iy = _unnormalized_x(g[:, 1], Hi)
    else:
        iy = ((g[:, 1] + F32(1.0)) * F32(0.5)) * F32(0.0)      # utils.py:87: y is not normalised when H == 1
    x0f, y0f = np.floor(ix), np.floor(iy)
    x1f, y1f = x0f + F32(1), y0f + F32(1)
    wts = [((x1f - ix) * (y1f - iy), x0f, y0f), ((ix - x0f) * (y1f - iy), x1f, y0f),
           ((x1f - ix) * (iy - y0f), x0f, y1f), ((ix - x0f) * (iy - y0f), x1f, y1f)]
    flat = img.reshape(B, C, Hi * Wi)
    out = np.zeros((B, C) + ix.shape[1:], dtype=F32)
    for wgt, xf, yf in wts:
        with np.errstate(invalid="ignore"):
            ok = (xf >= 0) & (xf <= Wi - 1) & (yf >= 0) & (yf <= Hi - 1)
        xi = np.where(ok, xf, 0).astype(np.int64)
        yi = np.where(ok, yf, 0).astype(np.int64)
        lin = (yi * Wi + xi).reshape(B, 1, -1)
        v = np.take_along_axis(flat, np.broadcast_to(lin, (B, C, lin.shape[-1])), axis=2).reshape(out.shape)
        out = out + np.where(ok[:, None], v * wgt[:, None].astype(F32), F32(0))
    return out.astype(F32)


def grid_halve(grid_xy):
    """0.5 * F.interpolate(grid, scale_factor=0.5, mode='bilinear', align_corners=True).
    ref: core/tc_stereo.py:163.  [B,2,H,W] -> [B,2,H//2,W//2]."""
    g = _f(grid_xy)
    B, C, H, W = g.shape
    Ho, Wo = H // 2, W // 2
    sh = F32(H - 1) / F32(Ho - 1) if Ho > 1 else F32(0)
    sw = F32(W - 1) / F32(Wo - 1) if Wo > 1 else F32(0)
    ys = (sh * np.arange(Ho, dtype=F32)).astype(F32)
    xs = (sw * np.arange(Wo, dtype=F32)).astype(F32)
    y0, x0 = ys.astype(np.int64), xs.astype(np.int64)
    y1 = y0 + (y0 < H - 1)
    x1 = x0 + (x0 < W - 1)
    ly1 = (ys - y0.astype(F32))[:, None]
    lx1 = (xs - x0.astype(F32))[None, :]
    ly0, lx0 = F32(1) - ly1, F32(1) - lx1
    top = lx0 * g[:, :, y0][:, :, :, x0] + lx1 * g[:, :, y0][:, :, :, x1]
    bot = lx0 * g[:, :, y1][:, :, :, x0] + lx1 * g[:, :, y1][:, :, :, x1]
    return (F32(0.5) * (ly0 * top + ly1 * bot)).astype(F32)


def warp_hidden_states(net_list, grid_xy):
    """ref: core/tc_stereo.py:159-163."""
    out = []
    g = grid_xy
    for i, net in enumerate(net_list):
        out.append(bilinear_sample(net, g))
        if i + 1 < len(net_list):
            g = grid_halve(g)
    return out


# ---- "next" row: the per-GRU-iteration 3x3 stencils ------------------------------------

RING_VU = [(0, 0), (0, 1), (0, 2), (1, 2), (2, 2), (2, 1), (2, 0), (1, 0)]     # geo_utils.py:83


def disp_gradient_xy(disp):
    """ref: core/utils/geo_utils.py:115-132 (disp2disp_gradient_xy).  disp [N,1,H,W] -> (grads [N,2,H,W], edge_mask
    [N,1,H,W] bool): forward differences on the replicate-padded map; the mask keeps |gx| < 5 and |gy| < 5."""
    d = _f(disp)
    pad = np.pad(d, ((0, 0), (0, 0), (1, 1), (1, 1)), mode="edge")
    c = pad[:, :, 1:-1, 1:-1]
    gx = pad[:, :, 1:-1, 2:] - c                                    # kernel (v,u) = (1,2) minus the centre
    gy = pad[:, :, 2:, 1:-1] - c                                    # kernel (2,1) minus the centre
    grads = np.concatenate([gx, gy], axis=1).astype(F32)
    return grads, (np.abs(gx) < 5) & (np.abs(gy) < 5)


def disp_grad_candidates(disp, level=1):
    """ref: core/utils/geo_utils.py:73-101 (disp2disp_grad_candidates).  disp [N,1,H,W] -> [N,2,8*level,H,W]."""
    d = _f(disp)
    N, _, H, W = d.shape
    vecs = []                                                        # per candidate: (dx, dy, ddisp) as [N,H,W] arrays
    for i in range(level):
        r = i + 1
        pad = np.pad(d[:, 0], ((0, 0), (r, r), (r, r)))             # zeros (F.pad default)
        c = pad[:, r:r + H, r:r + W]
        for v, u in RING_VU:                                         # dilation r: neighbour at ((v-1) r, (u-1) r)
            dy, dx = (v - 1) * r, (u - 1) * r
            nb = pad[:, r + dy:r + dy + H, r + dx:r + dx + W]
            vecs.append((np.full((N, H, W), dx, F32), np.full((N, H, W), dy, F32), (nb - c).astype(F32)))
    K = len(vecs)
    out = np.empty((N, 2, K, H, W), F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        for k in range(K):
            ax, ay, ad = vecs[k]
            bx, by, bd = vecs[(k + 2) % K]                           # torch.roll(grads, -2, dims=2)
            c0 = (ay * bd - ad * by).astype(F32)                     # torch.cross along (x, y, disp)
            c1 = (ad * bx - ax * bd).astype(F32)
            c2 = (ax * by - ay * bx).astype(F32)
            out[:, 0, k] = (-c0) / c2
            out[:, 1, k] = (-c1) / c2
    return out


def disp_propagate(disparity_grad, disparity_map):
    """ref: core/update.py:259-289 (DispRefine.propagate_disparity).  grad [N,2,H,W], disp [N,1,H,W] ->
    (propagated [N,9,H,W], matrix [N,18,H,W])."""
    g = _f(disparity_grad)
    d = _f(disparity_map)
    N, _, H, W = g.shape
    gp = np.pad(g, ((0, 0), (0, 0), (1, 1), (1, 1)))                # zeros
    dp = np.pad(d, ((0, 0), (0, 0), (1, 1), (1, 1)), mode="edge")   # replicate
    prop = np.empty((N, 9, H, W), F32)
    matrix = np.empty((N, 18, H, W), F32)
    for k in range(9):
        v, u = divmod(k, 3)                                          # update.py:221: row-major 3x3
        m = dp[:, 0, v:v + H, u:u + W]
        gx = gp[:, 0, v:v + H, u:u + W]
        gy = gp[:, 1, v:v + H, u:u + W]
        cx, cy = F32(1 - u), F32(1 - v)                              # centre minus neighbour coordinates
        prop[:, k] = ((m + gx * cx).astype(F32) + gy * cy).astype(F32)
        matrix[:, k] = np.abs(g[:, 0] - gx)
        matrix[:, 9 + k] = np.abs(g[:, 1] - gy)
    return prop, matrix


def convex_upsample(flow, mask, factor=4, scale=True):
    """ref: core/tc_stereo.py:75-88 (TCStereo.upsample_flow).  flow [N,D,H,W], mask
    [N,9*factor^2,H,W] -> [N,D,factor*H,factor*W]."""
    fl = _f(flow)
    N, D, H, W = fl.shape
    m = _f(mask).reshape(N, 1, 9, factor, factor, H, W)
    m = m - m.max(axis=2, keepdims=True)
    e = np.exp(m).astype(F32)
    w = (e / e.sum(axis=2, keepdims=True, dtype=F32)).astype(F32)                     # softmax over the 9 neighbours
    src = (F32(factor) * fl if scale else fl).astype(F32)
    pad = np.pad(src, ((0, 0), (0, 0), (1, 1), (1, 1)))                               # F.unfold(.., [3,3], padding=1)
    up = np.stack([pad[:, :, v:v + H, u:u + W] for v in range(3) for u in range(3)], axis=2)   # [N,D,9,H,W]
    out = np.zeros((N, D, factor, factor, H, W), F32)
    for k in range(9):
        out = (out + w[:, :, k] * up[:, :, k][:, :, None, None]).astype(F32)
    return out.transpose(0, 1, 4, 2, 5, 3).reshape(N, D, factor * H, factor * W)     # permute(0,1,4,2,5,3)


# ------------------------------------------------------------------------------------------------------
# (7) training: the cost-volume initialisation loss
# ------------------------------------------------------------------------------------------------------

def _interp_nearest_down(x, scale):
    """F.interpolate(x, scale_factor=scale, mode='nearest') for 1/scale an integer: output (i, j) takes input
    (floor(i / scale), floor(j / scale)); output size floor(in * scale)."""
    step = int(round(1.0 / scale))
    Ho, Wo = int(np.floor(x.shape[-2] * scale)), int(np.floor(x.shape[-1] * scale))
    return x[..., :Ho * step:step, :Wo * step:step]


def _interp_bilinear_ac(x, scale):
    """F.interpolate(x, scale_factor=scale, mode='bilinear', align_corners=True): source = dst * (in - 1) / (out - 1)."""
    B, C, H, W = x.shape
    Ho, Wo = int(np.floor(H * scale)), int(np.floor(W * scale))
    ys = np.arange(Ho, dtype=F32) * (F32(H - 1) / F32(max(Ho - 1, 1)))
    xs = np.arange(Wo, dtype=F32) * (F32(W - 1) / F32(max(Wo - 1, 1)))
    y0 = np.minimum(np.floor(ys).astype(np.int64), H - 1)
    x0 = np.minimum(np.floor(xs).astype(np.int64), W - 1)
    y1, x1 = np.minimum(y0 + 1, H - 1), np.minimum(x0 + 1, W - 1)
    wy, wx = (ys - y0.astype(F32)).reshape(1, 1, Ho, 1), (xs - x0.astype(F32)).reshape(1, 1, 1, Wo)
    top = x[:, :, y0][:, :, :, x0] * (F32(1) - wx) + x[:, :, y0][:, :, :, x1] * wx
    bot = x[:, :, y1][:, :, :, x0] * (F32(1) - wx) + x[:, :, y1][:, :, :, x1] * wx
    return _f(top * (F32(1) - wy) + bot * wy)


def init_loss(cost_volume, flow_gt, valid, max_flow=700, k=1, scale=0.25, threshold=0.1, valid_interp=None):
    """ref: train_stereo.py:138-182.  cost_volume [B,D,H,W] (corr.py:25-31), flow_gt [B,1,Hf,Wf], valid [B,1,Hf,Wf].
    Returns a dict: the loss and its terms, the per-pixel phi_gt / cost_nm / mask, and d loss / d cost_volume (the gradient
    torch.gather, masked_fill and topk pass back; ties in topk: lowest index).
    valid_interp: torch's own result of line 143 (the bilinear interpolation of `valid`).  The reference then tests it with
    `== 1`, so the LAST BIT of a float interpolation of ones decides whether a pixel counts, and that bit differs between this
    restatement, torch's CPU kernel and torch's CUDA kernel (0.99999994 vs 1.0 at 2 of 480 pixels of the golden case); the pin
    against the reference's output therefore takes torch's interpolation as given."""
    cv = _f(cost_volume)
    B, D, H, W = cv.shape
    flow = _f(F32(scale) * _interp_nearest_down(_f(flow_gt), scale))                                   # :141
    val = _f(valid_interp) if valid_interp is not None else _interp_bilinear_ac(_f(valid), scale)      # :143
    mag = np.sqrt(np.sum(flow * flow, axis=1, keepdims=True))                                          # :145
    val = (val == 1) & (mag < F32(max_flow * scale))                                                   # :148
    index_gt = _f(np.arange(W, dtype=F32).reshape(1, 1, 1, W) - (-flow))                               # :160-161
    mask = (index_gt >= 0) & (index_gt <= D - 1) & val                                                 # :162-163
    index_gt = np.clip(index_gt, 0, D - 1).astype(F32)                                                 # :164
    bb, hh, ww = np.meshgrid(np.arange(B), np.arange(H), np.arange(W), indexing="ij")

    def rho(d):                                                                                        # :150-152
        d = np.clip(d, 0, D - 1)
        return cv[bb, d[:, 0], hh, ww][:, None]

    df = np.floor(index_gt).astype(np.int64)                                                           # :155
    frac = _f(index_gt - df.astype(F32))
    phi = _f(_f(frac * rho(df + 1)) + _f(_f(F32(1) - frac) * rho(df)))                                 # :158
    n_mask = int(mask.sum())
    gt_loss = F32(1) - F32(phi[mask].astype(np.float64).mean()) if n_mask else F32(np.nan)             # :166
    rng = np.arange(D, dtype=F32).reshape(1, D, 1, 1)
    low, high = _f(index_gt - F32(1.5)), _f(index_gt + F32(1.5))                                       # :169-170
    filled = ((rng >= low) & (rng < high)) | ~mask                                                     # :171
    cv_nm = np.where(filled, F32(0), cv)
    order = np.argsort(-cv_nm, axis=1, kind="stable")[:, :k]                                           # :173 topk, descending
    cost_nm = np.take_along_axis(cv_nm, order, axis=1)
    hinge = _f(cost_nm + F32(threshold) - phi)                                                         # :174
    active = (hinge > 0) & np.repeat(mask, k, axis=1)
    nm_loss = F32(np.clip(hinge, 0, None)[np.repeat(mask, k, axis=1)].astype(np.float64).mean()) if n_mask else F32(np.nan)
    # ---- gradient w.r.t. cost_volume
    g = np.zeros_like(cv)
    if n_mask:
        gphi = np.where(mask, F32(-1.0 / n_mask), F32(0))                                              # d gt_loss / d phi (phi is detached in nm_loss)
        i0, i1 = np.clip(df, 0, D - 1), np.clip(df + 1, 0, D - 1)
        np.add.at(g, (bb, i0[:, 0], hh, ww), (_f(F32(1) - frac) * gphi)[:, 0])
        np.add.at(g, (bb, i1[:, 0], hh, ww), (frac * gphi)[:, 0])
        gnm = np.where(active, F32(1.0 / (n_mask * k)), F32(0))
        for j in range(k):
            through = ~np.take_along_axis(filled, order[:, j:j + 1], axis=1)                           # masked_fill passes no gradient
            np.add.at(g, (bb, order[:, j], hh, ww), (gnm[:, j:j + 1] * through)[:, 0])
    return {"loss": F32(gt_loss + nm_loss), "gt_loss": gt_loss, "nm_loss": nm_loss,
            "forward_mask_rate": F32(((cost_nm[:, :1] + F32(0.3) - phi) > 0).astype(F32).mean()),       # :180
            "phi": phi, "cost_nm": cost_nm, "mask": mask, "index_gt": index_gt, "grad_cost_volume": g}
