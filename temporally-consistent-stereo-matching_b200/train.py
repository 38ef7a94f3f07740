"""Differentiable correlation block: the volume path's backward, for training on the drop-in.

In the reference, gradients reach the feature encoder through the correlation block only (tc_stereo.py:116,177,238 ->
corr.py): lookup (grid_sample) -> pyramid (avg_pool2d) -> volume (einsum) -> F.normalize -> fmap1 / fmap2; the warp's
outputs, the per-iteration coordinates and the temporal state are all detached (geo_utils.py:198, tc_stereo.py:176,221-227).
`DifferentiableCorrBlock1D` keeps the forward on the kernels (build, lookup, argmax) and adds that backward:

    lookup backward   tcs_corr_lookup_backward (one kernel: taps' gradients -> d volume, pooling folded in, no atomics)
    volume backward   d n1 = d vol . n2,  d n2 = d vol^T . n1   (two batched GEMMs: torch.bmm, i.e. cuBLAS - a library GEMM)
    normalise         d f = (d n - n <n, d n>) / max(||f||, eps)                                   (torch elementwise)

`install(core.tc_stereo, training=True)` binds it (and nothing that would cut a gradient: no stencils, no fused encoder).
get_cost_volume() is differentiable too (the training loss reads it, train_stereo.py:385): it returns a deferred volume that
`init_loss` below evaluates on level 0 of the pyramid with one kernel each way (tcs_init_loss_forward / _backward) and that
any other consumer - the reference's own init_loss included - sees as the [B,W2,H,W1] tensor of corr.py:25-31.
What is NOT here: the splat's backward (softsplat.py:357-528) - dead code in this model, since warp() detaches its outputs.
"""
import torch
import torch.nn.functional as F

from . import _lib
from .corr import CorrBlock1D, _coords_plane, _stream
from .lazy import LazyTensorOps


class _Build(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fmap1, fmap2, block):
        ctx.save_for_backward(fmap1, fmap2)
        return block._levels[0].detach()

    @staticmethod
    def backward(ctx, dvol):
        f1, f2 = ctx.saved_tensors
        f1, f2 = f1.float(), f2.float()
        dvol = dvol.contiguous()
        norm1 = f1.norm(dim=1, keepdim=True).clamp_min(1e-12)
        norm2 = f2.norm(dim=1, keepdim=True).clamp_min(1e-12)
        n1, n2 = f1 / norm1, f2 / norm2
        dn1 = torch.einsum("bhij,bchj->bchi", dvol, n2)           # corr.py:60, transposed
        dn2 = torch.einsum("bhij,bchi->bchj", dvol, n1)
        df1 = (dn1 - n1 * (n1 * dn1).sum(dim=1, keepdim=True)) / norm1      # F.normalize's backward (norm above eps)
        df2 = (dn2 - n2 * (n2 * dn2).sum(dim=1, keepdim=True)) / norm2
        return df1, df2, None


class _Lookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vol, coords, block):
        ctx.block = block
        ctx.save_for_backward(coords)
        return CorrBlock1D.__call__(block, coords)

    @staticmethod
    def backward(ctx, gout):
        (coords,) = ctx.saved_tensors
        b = ctx.block
        gout = gout.float().contiguous()
        c, cptr, cstride = _coords_plane(coords, b.B, b.H, b.W1)
        dvol = torch.empty((b.B, b.H, b.W1, b.W2), dtype=torch.float32, device=gout.device)
        with torch.cuda.device(gout.device):
            _lib.call("tcs_corr_lookup_backward", gout.data_ptr(), cptr, cstride, dvol.data_ptr(),
                      b.B, b.H, b.W1, b.W2, b.num_levels, b.radius, _stream())
        return dvol, None, None


class DifferentiableCorrBlock1D(CorrBlock1D):
    """ref: core/corr.py:7-79 with gradients to fmap1 / fmap2 (pyramid mode)."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, thres=0.2, precision=None, mode=None):
        if mode not in (None, "pyramid"):
            raise ValueError("the differentiable block materialises the pyramid (mode='pyramid')")
        with torch.no_grad():
            super().__init__(fmap1.detach(), fmap2.detach(), num_levels, radius, thres, precision=precision, mode="pyramid")
        self._vol = _Build.apply(fmap1, fmap2, self) if (fmap1.requires_grad or fmap2.requires_grad) else self._levels[0]

    def __call__(self, coords):
        if not self._vol.requires_grad:
            return super().__call__(coords)
        return _Lookup.apply(self._vol, coords.detach(), self)

    def get_cost_volume(self):
        """ref: corr.py:25-31: [B,W2,H,W1], zero where w2 > w1; differentiable w.r.t. the volume.  Deferred: see LazyCostVolume."""
        return LazyCostVolume(self)

    def _cost_volume_tensor(self):
        if not self._vol.requires_grad:
            return super().get_cost_volume()
        w1 = torch.arange(self.W1, device=self.device).view(1, 1, 1, self.W1)
        w2 = torch.arange(self.W2, device=self.device).view(1, self.W2, 1, 1)
        vol = self._vol[..., :self.W2] if self._vol.shape[-1] != self.W2 else self._vol
        return vol.permute(0, 3, 1, 2) * (w2 <= w1).to(vol.dtype)


class LazyCostVolume(LazyTensorOps):
    """get_cost_volume() not yet evaluated: `init_loss` reads level 0 of the block's pyramid instead; every other use (torch
    functions, indexing, attributes) materialises the masked transposed tensor of corr.py:25-31, gradients included."""

    def __init__(self, block):
        self.block, self._value = block, None

    def materialize(self):
        if self._value is None:
            self._value = self.block._cost_volume_tensor()
        return self._value

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        unwrap = lambda a: a.materialize() if isinstance(a, LazyCostVolume) else a
        return func(*tuple(unwrap(a) for a in args), **{k: unwrap(v) for k, v in (kwargs or {}).items()})

    @property
    def shape(self):
        b = self.block
        return torch.Size((b.B, b.W2, b.H, b.W1))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]


class _InitLossTerms(torch.autograd.Function):
    """phi(index_gt) and the k largest non-matching costs per pixel, from level 0 (tcs_init_loss_forward / _backward)."""

    @staticmethod
    def forward(ctx, vol, index_gt, mask, block, k):
        B, H, W1, W2 = block.B, block.H, block.W1, block.W2
        dev = index_gt.device
        index_gt = index_gt.reshape(B, H, W1).float().contiguous()
        mask8 = mask.reshape(B, H, W1).to(torch.uint8).contiguous()
        phi = torch.empty((B, 1, H, W1), dtype=torch.float32, device=dev)
        cost_nm = torch.empty((B, k, H, W1), dtype=torch.float32, device=dev)
        idx_nm = torch.empty((B, k, H, W1), dtype=torch.int32, device=dev)
        lv0 = block._levels[0]
        with torch.cuda.device(dev):
            _lib.call("tcs_init_loss_forward", lv0.data_ptr(), block._pitch_arg(), index_gt.data_ptr(), mask8.data_ptr(),
                      phi.data_ptr(), cost_nm.data_ptr(), idx_nm.data_ptr(), B, H, W1, W2, k, _stream())
        ctx.save_for_backward(index_gt, idx_nm)
        ctx.dims = (B, H, W1, W2, k)
        ctx.mark_non_differentiable(idx_nm)
        return phi, cost_nm, idx_nm

    @staticmethod
    def backward(ctx, g_phi, g_nm, _):
        index_gt, idx_nm = ctx.saved_tensors
        B, H, W1, W2, k = ctx.dims
        g_phi = g_phi.float().contiguous()
        g_nm = g_nm.float().contiguous()
        dvol = torch.empty((B, H, W1, W2), dtype=torch.float32, device=g_phi.device)
        with torch.cuda.device(g_phi.device):
            _lib.call("tcs_init_loss_backward", g_phi.data_ptr(), g_nm.data_ptr(), index_gt.data_ptr(), idx_nm.data_ptr(),
                      dvol.data_ptr(), B, H, W1, W2, k, _stream())
        return dvol, None, None, None, None


def init_loss(cost_volume, flow_gt, valid, max_flow=700, k=1, scale=0.25, threshold=0.1):
    """ref: train_stereo.py:138-182, same arguments and return value (loss, metrics).  `cost_volume` is what the model returned
    under 'cost_volume' with the drop-in installed for training (a LazyCostVolume) or the DifferentiableCorrBlock1D itself; the
    volume-sized work (two gathers, the range mask, the top-k along w2 and their backward) is one kernel each way on level 0."""
    block = cost_volume.block if isinstance(cost_volume, LazyCostVolume) else cost_volume
    if not isinstance(block, DifferentiableCorrBlock1D):
        raise TypeError("init_loss needs the deferred cost volume of tcs_b200.install(core.tc_stereo, training=True) "
                        "(or the block itself); for a plain tensor call the reference's own init_loss")
    if flow_gt.shape[1] != 1:
        raise ValueError("flow_gt must have one channel (the horizontal flow), got %d" % flow_gt.shape[1])
    D, W = block.W2, block.W1
    flow_gt = scale * F.interpolate(flow_gt, scale_factor=scale, mode="nearest")                                  # :141
    valid = F.interpolate(valid.float(), scale_factor=scale, mode="bilinear", align_corners=True)                 # :143
    mag = torch.sum(flow_gt ** 2, dim=1, keepdim=True).sqrt()
    valid = (valid == 1) & (mag < max_flow * scale)                                                               # :148
    index_gt = torch.arange(W, device=flow_gt.device).view(1, 1, 1, -1) - (-flow_gt)                              # :160-161
    if tuple(index_gt.shape) != (block.B, 1, block.H, W):
        raise ValueError("flow_gt scaled by %g is %s, the volume is for %s" % (scale, tuple(index_gt.shape), (block.B, 1, block.H, W)))
    mask = (index_gt >= 0) & (index_gt <= D - 1) & valid                                                          # :162-163
    index_gt = torch.clip(index_gt, 0, D - 1)                                                                     # :164
    phi_gt, cost_nm, _ = _InitLossTerms.apply(block._vol, index_gt, mask, block, k)
    gt_loss = 1 - phi_gt[mask].mean()                                                                             # :166
    nm_loss = torch.clip(cost_nm + threshold - phi_gt.detach(), min=0)[mask.repeat(1, k, 1, 1)].mean()            # :172-173
    loss = gt_loss + nm_loss
    metrics = {
        "init_loss": loss.item(),
        "init_gt_loss": gt_loss.item(),
        "init_nm_loss": nm_loss.item(),
        "forward_mask_rate": ((cost_nm[:, :1] + 0.3 - phi_gt) > 0).float().mean().item(),
    }
    return loss, metrics
