"""Multi-threaded CPU port of the hot path on torch's CPU operators — TEST / BASELINE INFRASTRUCTURE ONLY.

Purpose: the `cpu_baseline` leg and `--impl reference` arm of bench.py.  The reference's implementation of
this path IS a sequence of torch library calls (F.normalize, einsum, avg_pool2d, grid_sample, max, matmul;
core/corr.py, core/utils/geo_utils.py, core/utils/utils.py) plus one cupy kernel that cannot run on a CPU;
the reference itself cannot travel to the GPU box, so this file restates that call sequence on the same
torch operators (kind = "port"), with the splat as index_add_ (ref: softsplat.py:284-335).  It uses all
the host threads torch is given.  It is checked against the numpy oracle and the golden vectors in
tests/test_oracle_golden.py; the product never imports it.
"""
import torch
import torch.nn.functional as F


def build_block(fmap1, fmap2, num_levels=4):
    """ref: core/corr.py:8-31,54-62 — normalise, all-pairs product, pooled pyramid, masked transposed volume."""
    n1 = F.normalize(fmap1, dim=1)
    n2 = F.normalize(fmap2, dim=1)
    vol = torch.einsum('bchi,bchj->bhij', n1, n2).contiguous()           # [B,H,W1,W2]
    B, H, W1, W2 = vol.shape
    level = vol.clone().reshape(B * H * W1, 1, 1, W2)
    pyramid = [level]
    for _ in range(num_levels):                                          # the reference pools num_levels times
        level = F.avg_pool2d(level, [1, 2], stride=[1, 2])
        pyramid.append(level)
    keep = torch.arange(W1).view(1, 1, 1, W1) >= torch.arange(W2).view(1, W2, 1, 1)
    cost_volume = vol.permute(0, 3, 1, 2).contiguous() * keep.to(vol.dtype)
    return pyramid, cost_volume


def lookup(pyramid, coords, num_levels=4, radius=4):
    """ref: core/corr.py:33-52 + core/utils/utils.py:82-97."""
    B, _, H, W1 = coords.shape
    x = coords[:, :1].permute(0, 2, 3, 1).reshape(B * H * W1, 1, 1, 1)
    taps = torch.linspace(-radius, radius, 2 * radius + 1).view(2 * radius + 1, 1)
    outs = []
    for l in range(num_levels):
        lv = pyramid[l]
        Wl = lv.shape[-1]
        xl = taps + x / 2 ** l
        grid = torch.cat([2 * xl / (Wl - 1) - 1, torch.zeros_like(xl)], dim=-1)
        outs.append(F.grid_sample(lv, grid, align_corners=True).view(B, H, W1, -1))
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def argmax_disp(cost_volume, thres=0.3):
    """ref: core/corr.py:67-79."""
    B, W2, H, W1 = cost_volume.shape
    main, idx = cost_volume.max(dim=1, keepdim=True)
    k = torch.arange(W2).view(1, W2, 1, 1)
    near = (k >= idx - 1.5) & (k < idx + 1.5)
    sub = torch.where(near, torch.zeros_like(cost_volume), cost_volume).max(dim=1, keepdim=True)[0]
    mask = (main - sub > thres).float()
    disp = (torch.arange(W1).view(1, 1, 1, W1) - idx) * mask
    return disp, main * mask, mask


def _pixel_grid(B, H, W):
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    return torch.stack([xs, ys], 0)[None].repeat(B, 1, 1, 1)


def _lift_and_move(disp, rel_T, K, K_inv, baseline):
    """disp -> depth -> point -> transformed point (geo_utils.py:7-16,32-42,135-145)."""
    B, _, H, W = disp.shape
    bf = (baseline.view(-1) * K[:, 0, 0]).view(B, 1, 1, 1)
    depth = bf / torch.clip(disp, min=0.001)
    pix = torch.cat([_pixel_grid(B, H, W), torch.ones(B, 1, H, W)], 1).view(B, 3, -1)
    P = depth.view(B, 1, -1) * torch.matmul(K_inv, pix)
    P = torch.matmul(rel_T, torch.cat([P, torch.ones(B, 1, H * W)], 1))[:, :3]
    return bf, P


def _to_pixels(P, K, H, W):
    uv = torch.matmul(K, P) / P[:, 2:3]
    uv = torch.where(torch.isnan(uv) | torch.isinf(uv), -torch.ones_like(uv), uv)
    return uv[:, :2].reshape(-1, 2, H, W)


def splat(ten_in, flow):
    """ref: softsplat.py:284-335 as index_add_ (one pass per corner)."""
    B, C, H, W = ten_in.shape
    tgt = _pixel_grid(B, H, W) + flow
    fx, fy = tgt[:, 0].reshape(B, -1), tgt[:, 1].reshape(B, -1)
    ok = torch.isfinite(fx) & torch.isfinite(fy)
    fx, fy = torch.where(ok, fx, torch.zeros_like(fx)), torch.where(ok, fy, torch.zeros_like(fy))
    x0, y0 = fx.floor(), fy.floor()
    x1, y1 = x0 + 1, y0 + 1
    src = ten_in.reshape(B, C, H * W)
    out = torch.zeros(B, C, H * W)
    for cx, cy, wgt in ((x0, y0, (x1 - fx) * (y1 - fy)), (x1, y0, (fx - x0) * (y1 - fy)),
                        (x0, y1, (x1 - fx) * (fy - y0)), (x1, y1, (fx - x0) * (fy - y0))):
        inb = ok & (cx >= 0) & (cx < W) & (cy >= 0) & (cy < H)
        lin = (cy.clamp(0, H - 1) * W + cx.clamp(0, W - 1)).long()
        w = wgt * inb
        for b in range(B):
            out[b].index_add_(1, lin[b], src[b] * w[b])
    return out.view(B, C, H, W)


def warp(disp, fmap, rel_T, K, K_inv, baseline):
    """ref: geo_utils.py:158-198 + softsplat.py:232-274 ('soft-clipeps')."""
    B, _, H, W = disp.shape
    bf, P = _lift_and_move(disp, rel_T, K, K_inv, baseline)
    z = P[:, 2:3].reshape(B, 1, H, W)
    d1 = bf / z
    d1 = torch.where(torch.isnan(d1) | torch.isinf(d1), -torch.ones_like(d1), d1)
    valid = ((d1 > 0) & (d1 < W)).float()
    flow = _to_pixels(P, K, H, W) - _pixel_grid(B, H, W)
    e = (d1 - d1.mean()).clamp(-50, 50).exp()
    feats = torch.cat([d1, fmap], 1) * valid
    out = splat(torch.cat([feats * e, e * valid], 1), flow)
    norm = out[:, -1:]
    mask = (norm != 0).float()
    out = out[:, :-1] / norm.clip(1e-7, None)
    return out[:, :1], out[:, 1:], mask


def matching_cost(fmap1, warped, mask):
    """ref: core/tc_stereo.py:139-140."""
    return torch.sum(F.normalize(fmap1, dim=1) * F.normalize(warped, dim=1), dim=1, keepdim=True) * mask


def backward_grid(disp, rel_T, K, K_inv, baseline):
    """ref: geo_utils.py:201-236."""
    B, _, H, W = disp.shape
    _, P = _lift_and_move(torch.clip(disp, 0.01), rel_T, K, K_inv, baseline)
    uv = _to_pixels(P, K, H, W)
    return torch.where(P[:, 2:3].reshape(B, 1, H, W) > 0, uv, -torch.ones_like(uv))


def warp_hidden(net_list, grid):
    """ref: core/tc_stereo.py:159-163 (bilinear_sampler = normalise + grid_sample, utils.py:82-97)."""
    out = []
    for net in net_list:
        H, W = net.shape[-2:]
        g = grid.permute(0, 2, 3, 1)
        gx = 2 * g[..., :1] / (W - 1) - 1
        gy = 2 * g[..., 1:] / (H - 1) - 1 if H > 1 else g[..., 1:]
        out.append(F.grid_sample(net, torch.cat([gx, gy], -1), align_corners=True))
        grid = 0.5 * F.interpolate(grid, scale_factor=0.5, mode="bilinear", align_corners=True)
    return out


def frame(fmap1, fmap2, coords_seq, state=None, rel_T=None, rel_T_inv=None, K=None, K_inv=None, baseline=None,
          num_levels=4, radius=4):
    """One frame of the hot path on the CPU, same order as tcs_b200.hot_path_frame."""
    pyramid, cost_volume = build_block(fmap1, fmap2, num_levels)
    warped_net = None
    if state is None:
        sparse_disp, cost, mask = argmax_disp(cost_volume)
    else:
        last_disp, last_fmap1, last_nets = state
        sparse_disp, wf, mask = warp(last_disp, last_fmap1, rel_T, K, K_inv, baseline)
        cost = matching_cost(fmap1, wf, mask)
        if last_nets is not None:
            warped_net = warp_hidden(last_nets, backward_grid(sparse_disp, rel_T_inv, K, K_inv, baseline))
    out = None
    for it in range(coords_seq.shape[0]):
        out = lookup(pyramid, coords_seq[it], num_levels, radius)
    return {"corr": out, "sparse_disp": sparse_disp, "cost": cost, "mask": mask, "warped_net": warped_net}
