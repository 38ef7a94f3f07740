"""Differentiable correlation block: the volume path's backward, for training on the drop-in.

In the reference, gradients reach the feature encoder through the correlation block only (tc_stereo.py:116,177,238 ->
corr.py): lookup (grid_sample) -> pyramid (avg_pool2d) -> volume (einsum) -> F.normalize -> fmap1 / fmap2; the warp's
outputs, the per-iteration coordinates and the temporal state are all detached (geo_utils.py:198, tc_stereo.py:176,221-227).
`DifferentiableCorrBlock1D` keeps the forward on the kernels (build, lookup, argmax) and adds that backward:

    lookup backward   tcs_corr_lookup_backward (one kernel: taps' gradients -> d volume, pooling folded in, no atomics)
    volume backward   d n1 = d vol . n2,  d n2 = d vol^T . n1   (two batched GEMMs: torch.bmm, i.e. cuBLAS - a library GEMM)
    normalise         d f = (d n - n <n, d n>) / max(||f||, eps)                                   (torch elementwise)

`install(core.tc_stereo, training=True)` binds it (and nothing that would cut a gradient: no stencils, no fused encoder).
get_cost_volume() is differentiable too (the training loss reads it, train_stereo.py:385).  What is NOT here: the splat's
backward (softsplat.py:357-528) - dead code in this model, since warp() detaches its outputs.
"""
import torch
import torch.nn.functional as F

from . import _lib
from .corr import CorrBlock1D, _coords_plane, _stream


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
        """ref: corr.py:25-31: [B,W2,H,W1], zero where w2 > w1; differentiable w.r.t. the volume."""
        if not self._vol.requires_grad:
            return super().get_cost_volume()
        w1 = torch.arange(self.W1, device=self.device).view(1, 1, 1, self.W1)
        w2 = torch.arange(self.W2, device=self.device).view(1, self.W2, 1, 1)
        vol = self._vol[..., :self.W2] if self._vol.shape[-1] != self.W2 else self._vol
        return vol.permute(0, 3, 1, 2) * (w2 <= w1).to(vol.dtype)
