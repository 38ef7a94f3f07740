"""Temporal step on top of libtcs_b200.so — the reference's geometry functions, same names and meaning.

Mirrors core/utils/geo_utils.py (warp :158-198, get_backward_grid :201-236, cal_relative_transformation
:148-155) and core/utils/utils.py (bilinear_sampler :82-97) of the reference, plus the hidden-state warp
loop of core/tc_stereo.py:159-163.  CUDA tensors only, inference only, no fallback.
"""
import os

import torch

from . import _lib


_HIDDEN_ONE_LAUNCH = os.environ.get("TCS_B200_HIDDEN_ONE_LAUNCH", "1") != "0"   # development aid: 0 chains the single ops


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _no_grad_path(name, t):
    """The kernels have no backward (inference only).  In the reference, gradients flow through
    these calls (corr -> fnet, propagate_disparity -> disp_grad, upsample_flow -> up_mask), so a training run on
    the drop-in would silently train with them cut: refuse instead."""
    if t.requires_grad and torch.is_grad_enabled():
        raise RuntimeError("%s requires grad, but libtcs_b200 is inference only (no backward kernels): run under "
                           "torch.no_grad() / model.eval() with test_mode=True, or uninstall() the drop-in for training" % name)


def _f32c(name, t, shape=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA tensor (libtcs_b200 has no CPU path)" % name)
    _no_grad_path(name, t)
    if t.dtype != torch.float32:
        t = t.float()
    t = t.contiguous()
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError("%s must have shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
    return t


def _warp_scratch(B, C, H, W, device):
    """Scratch of tcs_warp_forward, taken from torch's caching allocator on every call: stream-ordered (two streams
    never share a block), owned by the graph's pool when the call is captured into a CUDA graph, and free once warm."""
    n = _lib.warp_scratch_bytes(B, C, H, W)
    if n <= 0:
        raise ValueError("bad warp shape B=%d C=%d H=%d W=%d" % (B, C, H, W))
    return torch.empty(n, dtype=torch.uint8, device=device)


def _camera_args(relative_T, K, K_inv, baseline, B):
    relative_T = _f32c("relative_T", relative_T, (B, 4, 4))
    K = _f32c("K", K, (B, 3, 3))
    K_inv = _f32c("K_inv", K_inv, (B, 3, 3))
    baseline = _f32c("baseline", baseline).reshape(-1)
    if baseline.numel() != B:
        raise ValueError("baseline must have %d entries, got %d" % (B, baseline.numel()))
    return relative_T, K, K_inv, baseline


class WarpCarry:
    """The current frame's features, already transposed for the next frame's warp.

    The cost-only call of warp_with_cost has cur_fmap in shared memory anyway and writes it out as pixel-major rows
    (`rows`, [B*H*W, C] in the library's private channel order).  TC-Stereo hands fmap1 to the next frame as
    last_fmap1 (core/tc_stereo.py:137, evaluate_stereo.py:192-197); when that next call receives this object as
    `carry_in` and its `fmap` is the memory recorded here, unchanged, the list formulation skips its
    transposition pass.  The key is (storage pointer, shape, strides, version counter), not object identity: the model
    hands back `fmap1.detach()` (tc_stereo.py:227,242), a new Python object over the same storage and version counter.
    The tensor is kept referenced so that its storage cannot be freed and reused in between."""

    def __init__(self):
        self.tensor = None
        self.version = -1
        self.rows = None

    def record(self, fmap):
        self.tensor = fmap
        self.version = fmap._version

    def reserve(self, like):
        """Allocate the row buffer for feature maps shaped like `like` [B,C,H,W] now (e.g. outside a timed loop)."""
        B, C, H, W = like.shape
        if self.rows is None or tuple(self.rows.shape) != (B * H * W, C) or self.rows.device != like.device:
            self.rows = torch.empty((B * H * W, C), dtype=torch.float32, device=like.device)
            self.tensor = None
        return self

    def matches(self, fmap):
        t = self.tensor
        return (t is not None and self.rows is not None and fmap.dtype == torch.float32 and fmap.is_contiguous()
                and fmap.device == t.device and fmap.data_ptr() == t.data_ptr() and fmap.shape == t.shape
                and fmap.stride() == t.stride() and fmap._version == self.version and t._version == self.version)


def warp_with_cost(disp, fmap, relative_T, K, K_inv, baseline, cur_fmap=None, per_sample_mean=False, want_fmap=True,
                   deterministic=False, carry_in=None, carry_out=None):
    """warp() plus the matching cost of core/tc_stereo.py:139-140 fused into the normalise kernel.

    -> (disp', fmap', mask, cost)  with cost None when cur_fmap is None.  want_fmap=False skips materialising
    fmap' (TCStereo.forward only ever reads its cost, tc_stereo.py:139-140) and returns None in its place.
    deterministic=True collects the splat from the target's side through sorted contributor lists (bitwise
    repeatable, no accumulator, as fast as the default atomic scatter).
    carry_out (a WarpCarry; cost-only calls) receives cur_fmap transposed for the next frame; carry_in (the
    WarpCarry filled when `fmap` was the current frame) selects the list formulation and saves its transposition."""
    disp = _f32c("disp", disp)
    fmap = _f32c("fmap", fmap)
    if disp.dim() != 4 or disp.shape[1] != 1:
        raise ValueError("disp must be [B,1,H,W], got %s" % (tuple(disp.shape),))
    B, _, H, W = disp.shape
    if fmap.dim() != 4 or fmap.shape[0] != B or fmap.shape[2:] != disp.shape[2:]:
        raise ValueError("fmap must be [B,C,H,W] matching disp, got %s" % (tuple(fmap.shape),))
    C = fmap.shape[1]
    relative_T, K, K_inv, baseline = _camera_args(relative_T, K, K_inv, baseline, B)
    if cur_fmap is not None:
        cur_fmap = _f32c("cur_fmap", cur_fmap, fmap.shape)
    dev = disp.device
    out_disp = torch.empty_like(disp)
    if not want_fmap and cur_fmap is None:
        raise ValueError("want_fmap=False needs cur_fmap (otherwise nothing of the warped features is returned)")
    out_fmap = torch.empty_like(fmap) if want_fmap else None
    out_mask = torch.empty_like(disp)
    out_cost = torch.empty_like(disp) if cur_fmap is not None else None
    fmap_t = None
    if carry_in is not None and carry_in.matches(fmap) and tuple(carry_in.rows.shape) == (B * H * W, C):
        fmap_t = carry_in.rows
        deterministic = True                                  # the list formulation is the one that reads rows
    cur_t = None
    if carry_out is not None:
        if want_fmap or cur_fmap is None:
            raise ValueError("carry_out is filled by the cost-only call (want_fmap=False with cur_fmap)")
        carry_out.reserve(fmap)
        if fmap_t is not None and carry_out.rows.data_ptr() == fmap_t.data_ptr():
            raise ValueError("carry_in and carry_out must be different WarpCarry objects")
        cur_t = carry_out.rows
        carry_out.tensor = None                               # the rows are about to be overwritten: valid again only on success
    with torch.cuda.device(dev):
        scratch = _warp_scratch(B, C, H, W, dev)
        _lib.call("tcs_warp_forward", disp.data_ptr(), fmap.data_ptr(), relative_T.data_ptr(), K.data_ptr(),
                  K_inv.data_ptr(), baseline.data_ptr(), cur_fmap.data_ptr() if cur_fmap is not None else None,
                  out_disp.data_ptr(), out_fmap.data_ptr() if out_fmap is not None else None, out_mask.data_ptr(),
                  out_cost.data_ptr() if out_cost is not None else None,
                  fmap_t.data_ptr() if fmap_t is not None else None, cur_t.data_ptr() if cur_t is not None else None,
                  scratch.data_ptr(),
                  B, C, H, W, (1 if per_sample_mean else 0) | (2 if deterministic else 0), _stream())
    if carry_out is not None:
        carry_out.record(cur_fmap)                            # the fp32 contiguous tensor the kernel read
    return out_disp, out_fmap, out_mask, out_cost


def warp(disp, fmap, relative_T, K, K_inv, baseline):
    """ref: geo_utils.py:158-198.  -> (current_disp [B,1,H,W], current_fmap [B,C,H,W], warped_mask [B,1,H,W])."""
    d, f, m, _ = warp_with_cost(disp, fmap, relative_T, K, K_inv, baseline)
    return d, f, m


def get_backward_grid(disp, relative_T, K, K_inv, baseline):
    """ref: geo_utils.py:201-236.  disp [B,1,H,W] -> previous-frame pixel coordinates [B,2,H,W] (x, y)."""
    disp = _f32c("disp", disp)
    if disp.dim() != 4 or disp.shape[1] != 1:
        raise ValueError("disp must be [B,1,H,W], got %s" % (tuple(disp.shape),))
    B, _, H, W = disp.shape
    relative_T, K, K_inv, baseline = _camera_args(relative_T, K, K_inv, baseline, B)
    grid = torch.empty((B, 2, H, W), dtype=torch.float32, device=disp.device)
    with torch.cuda.device(disp.device):
        _lib.call("tcs_backward_grid", disp.data_ptr(), relative_T.data_ptr(), K.data_ptr(), K_inv.data_ptr(),
                  baseline.data_ptr(), grid.data_ptr(), B, H, W, _stream())
    return grid


def sample_planar(img, grid_xy):
    """img [B,C,Hi,Wi] sampled at pixel coordinates grid_xy [B,2,Ho,Wo] (x plane, y plane)."""
    img = _f32c("img", img)
    grid_xy = _f32c("grid_xy", grid_xy)
    B, C, Hi, Wi = img.shape
    if grid_xy.dim() != 4 or grid_xy.shape[0] != B or grid_xy.shape[1] != 2:
        raise ValueError("grid_xy must be [B,2,Ho,Wo], got %s" % (tuple(grid_xy.shape),))
    Ho, Wo = grid_xy.shape[2:]
    out = torch.empty((B, C, Ho, Wo), dtype=torch.float32, device=img.device)
    with torch.cuda.device(img.device):
        _lib.call("tcs_bilinear_sample", img.data_ptr(), grid_xy.data_ptr(), out.data_ptr(), B, C, Hi, Wi, Ho, Wo, _stream())
    return out


def bilinear_sampler(img, coords, mode='bilinear', mask=False, align_corners=True):
    """ref: utils.py:82-97.  coords [B,Ho,Wo,2] in pixels (x, y).  Only the reference's own use
    (bilinear, align_corners=True) is implemented."""
    if mode != 'bilinear' or not align_corners:
        raise NotImplementedError("libtcs_b200 implements bilinear_sampler(mode='bilinear', align_corners=True) only")
    if coords.dim() != 4 or coords.shape[-1] != 2:
        raise ValueError("coords must be [B,Ho,Wo,2], got %s" % (tuple(coords.shape),))
    out = sample_planar(img, coords.permute(0, 3, 1, 2))   # a no-copy view when coords came from a planar grid
    if mask:
        H, W = img.shape[-2:]
        xg = 2 * coords[..., :1] / (W - 1) - 1
        yg = 2 * coords[..., 1:] / (H - 1) - 1 if H > 1 else coords[..., 1:]
        return out, ((xg > -1) & (yg > -1) & (xg < 1) & (yg < 1)).float()
    return out


def halve_grid(grid_xy):
    """ref: tc_stereo.py:163.  0.5 * F.interpolate(grid, scale_factor=0.5, 'bilinear', align_corners=True)."""
    grid_xy = _f32c("grid_xy", grid_xy)
    B, two, H, W = grid_xy.shape
    if two != 2:
        raise ValueError("grid_xy must be [B,2,H,W]")
    out = torch.empty((B, 2, H // 2, W // 2), dtype=torch.float32, device=grid_xy.device)
    with torch.cuda.device(grid_xy.device):
        _lib.call("tcs_grid_halve", grid_xy.data_ptr(), out.data_ptr(), B, H, W, _stream())
    return out


def warp_hidden_states(net_list, backward_grid):
    """ref: tc_stereo.py:159-163.  Sample each hidden-state level with the (progressively halved) grid.  The model's
    three-level case (each level half the previous one's size) is one launch; anything else chains the single ops."""
    if len(net_list) == 3 and backward_grid.dim() == 4 and backward_grid.shape[1] == 2 and _HIDDEN_ONE_LAUNCH:
        B, _, H, W = backward_grid.shape
        dims = [(H, W), (H // 2, W // 2), (H // 2 // 2, W // 2 // 2)]
        if H >= 4 and W >= 4 and all(n.dim() == 4 and n.shape[0] == B and tuple(n.shape[2:]) == d for n, d in zip(net_list, dims)):
            nets = [_f32c("net_list[%d]" % i, n) for i, n in enumerate(net_list)]
            grid = _f32c("backward_grid", backward_grid)
            outs = [torch.empty_like(n) for n in nets]
            with torch.cuda.device(grid.device):
                _lib.call("tcs_warp_hidden_states", nets[0].data_ptr(), nets[1].data_ptr(), nets[2].data_ptr(), grid.data_ptr(),
                          outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), B, nets[0].shape[1], nets[1].shape[1],
                          nets[2].shape[1], H, W, _stream())
            return outs
    out = []
    grid = backward_grid
    for i, net in enumerate(net_list):
        out.append(sample_planar(net, grid))
        if i + 1 < len(net_list):
            grid = halve_grid(grid)
    return out


def cal_relative_transformation(T1, T2):
    """ref: geo_utils.py:148-155.  T2 @ inv(T1) for world2cam poses [B,4,4] (or [4,4]), one launch, no host sync
    (torch.linalg.inv reads its LU `info` back on the host: two syncs per temporal frame in tc_stereo.py:127,159)."""
    single = T1.dim() == 2
    a = _f32c("T1", T1.reshape(-1, 4, 4))
    b = _f32c("T2", T2.reshape(-1, 4, 4))
    if a.shape != b.shape or tuple(a.shape[1:]) != (4, 4):
        raise ValueError("T1 and T2 must both be [B,4,4], got %s and %s" % (tuple(T1.shape), tuple(T2.shape)))
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _lib.call("tcs_relative_pose", a.data_ptr(), b.data_ptr(), out.data_ptr(), a.shape[0], _stream())
    return out[0] if single else out


# ---- "next" row: the per-GRU-iteration 3x3 stencils on the disparity ------------------

def _disp_map(name, t):
    t = _f32c(name, t)
    if t.dim() != 4 or t.shape[1] != 1:
        raise ValueError("%s must be [N,1,H,W], got %s" % (name, tuple(t.shape)))
    return t


def disp2disp_gradient_xy(disp):
    """ref: geo_utils.py:115-132.  disp [N,1,H,W] -> (grads [N,2,H,W], edge_mask [N,1,H,W] bool), one kernel."""
    disp = _disp_map("disp", disp)
    N, _, H, W = disp.shape
    grads = torch.empty((N, 2, H, W), dtype=torch.float32, device=disp.device)
    edge = torch.empty((N, 1, H, W), dtype=torch.bool, device=disp.device)
    with torch.cuda.device(disp.device):
        _lib.call("tcs_disp_gradient_xy", disp.data_ptr(), grads.data_ptr(), edge.data_ptr(), N, H, W, _stream())
    return grads, edge


def disp2disp_grad_candidates(disp, level=1):
    """ref: geo_utils.py:73-101.  disp [N,1,H,W] -> [N,2,8*level,H,W], one kernel (bit-identical for level <= 2, the
    model's setting; level 3 and 4 multiply by 3 and agree to an ulp)."""
    disp = _disp_map("disp", disp)
    if not 1 <= int(level) <= 4:
        raise ValueError("level must be 1..4, got %r" % (level,))
    N, _, H, W = disp.shape
    out = torch.empty((N, 2, 8 * int(level), H, W), dtype=torch.float32, device=disp.device)
    with torch.cuda.device(disp.device):
        _lib.call("tcs_disp_grad_candidates", disp.data_ptr(), out.data_ptr(), N, H, W, int(level), _stream())
    return out


def propagate_disparity(disparity_grad, disparity_map):
    """ref: update.py:259-289 (DispRefine.propagate_disparity).  grad [N,2,H,W], disp [N,1,H,W] ->
    (propagated [N,9,H,W], matrix [N,18,H,W]), one kernel."""
    disparity_map = _disp_map("disparity_map", disparity_map)
    N, _, H, W = disparity_map.shape
    disparity_grad = _f32c("disparity_grad", disparity_grad, (N, 2, H, W))
    prop = torch.empty((N, 9, H, W), dtype=torch.float32, device=disparity_map.device)
    matrix = torch.empty((N, 18, H, W), dtype=torch.float32, device=disparity_map.device)
    with torch.cuda.device(disparity_map.device):
        _lib.call("tcs_disp_propagate", disparity_grad.data_ptr(), disparity_map.data_ptr(), prop.data_ptr(), matrix.data_ptr(),
                  N, H, W, _stream())
    return prop, matrix


def convex_upsample(flow, mask, factor=4, scale=True):
    """ref: tc_stereo.py:75-88 (TCStereo.upsample_flow).  flow [N,D,H,W], mask
    [N,9*factor^2,H,W] -> [N,D,factor*H,factor*W]: softmax + unfold + weighted sum + re-layout in one kernel."""
    flow = _f32c("flow", flow)
    if flow.dim() != 4:
        raise ValueError("flow must be [N,D,H,W], got %s" % (tuple(flow.shape),))
    N, D, H, W = flow.shape
    factor = int(factor)
    if factor not in (2, 4, 8):
        raise ValueError("factor must be 2, 4 or 8, got %r" % (factor,))
    mask = _f32c("mask", mask, (N, 9 * factor * factor, H, W))
    out = torch.empty((N, D, factor * H, factor * W), dtype=torch.float32, device=flow.device)
    with torch.cuda.device(flow.device):
        _lib.call("tcs_convex_upsample", flow.data_ptr(), mask.data_ptr(), out.data_ptr(), N, D, H, W, factor,
                  1 if scale else 0, _stream())
    return out
