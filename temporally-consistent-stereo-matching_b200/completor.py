"""Input stems of the reference's DisparityCompletor as one kernel (SURVEY.md section 8f rank 3's other half).

ref: core/update.py:312-323 (the four nn.Sequential stems) and :375-378 (their use in DisparityCompletor.forward):

    disp_f4 = self.conv_disp_stem(disp); cost_f4 = self.conv_cost_stem(cost); mask_f4 = self.conv_mask_stem(mask)
    x4_disp = self.conv_disp_fuse(torch.cat((disp_f4, cost_f4, mask_f4), dim=1))

`fuse_completor_stems(model.disp_completor)` replaces the `forward` of those four module INSTANCES: the three stems return a
marker that remembers their input, `torch.cat` of exactly those three markers (dim=1) returns a marker of the three inputs, and
conv_disp_fuse on that marker runs tcs_completor_stems.  DisparityCompletor.forward itself is untouched; anything else done with
a marker materialises it with the module's original layers.  fp32, inference only.
"""
import torch

from . import _lib
from .lazy import LazyTensorOps

_STEMS = ("conv_disp_stem", "conv_cost_stem", "conv_mask_stem")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def pack_stem_weights(completor):
    """The 16 parameters of the four stems in the kernel's layout (matrices transposed to [input][output])."""
    def two(seq):
        a, b = seq[0], seq[2]
        return [a.weight.reshape(-1), a.bias, b.weight.reshape(b.out_channels, b.in_channels).t().reshape(-1), b.bias]
    parts = two(completor.conv_disp_stem) + two(completor.conv_cost_stem) + two(completor.conv_mask_stem)
    f0, f2 = completor.conv_disp_fuse[0], completor.conv_disp_fuse[2]
    parts += [f0.weight.reshape(f0.out_channels, f0.in_channels).t().reshape(-1), f0.bias,
              f2.weight.reshape(f2.out_channels, f2.in_channels).t().reshape(-1), f2.bias]
    packed = torch.cat([p.detach().float().reshape(-1) for p in parts]).contiguous()
    n = _lib.load().tcs_completor_stems_weight_floats()
    if packed.numel() != n:
        raise ValueError("the stems do not have the reference's widths (1->64->64, 1->32->32, 1->32->32, 128->128->64): "
                         "%d parameters, the kernel packs %d" % (packed.numel(), n))
    return packed


def completor_stems(disp, cost, mask, packed):
    """-> x4_disp [N,64,H,W] from the stems' inputs [N,1,H,W] (ref: update.py:375-378)."""
    from .geo import _f32c
    disp = _f32c("disp", disp)
    if disp.dim() != 4 or disp.shape[1] != 1:
        raise ValueError("disp must be [N,1,H,W], got %s" % (tuple(disp.shape),))
    cost = _f32c("cost", cost, disp.shape)
    mask = _f32c("mask", mask, disp.shape)
    N, _, H, W = disp.shape
    out = torch.empty((N, 64, H, W), dtype=torch.float32, device=disp.device)
    with torch.cuda.device(disp.device):
        _lib.call("tcs_completor_stems", disp.data_ptr(), cost.data_ptr(), mask.data_ptr(), packed.data_ptr(), out.data_ptr(),
                  N, H, W, _stream())
    return out


class _StemOut(LazyTensorOps):
    """Output of one patched stem (`which` in 0..2), or of the cat of all three (`which` = 3, value = the three inputs)."""

    def __init__(self, owner, which, value):
        self._owner, self._which, self._in, self._value = owner, which, value, None

    @property
    def shape(self):
        x = self._in[0] if self._which == 3 else self._in
        return torch.Size((x.shape[0], (64, 32, 32, 128)[self._which], x.shape[2], x.shape[3]))

    def materialize(self):
        if self._value is None:
            o = self._owner
            if self._which == 3:
                self._value = torch.cat([o.original[n](x) for n, x in zip(_STEMS, self._in)], dim=1)
            else:
                self._value = o.original[_STEMS[self._which]](self._in)
        return self._value

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is torch.cat and len(args) >= 1 and isinstance(args[0], (list, tuple)) and len(args[0]) == 3:
            a = args[0]
            dim = kwargs.get("dim", args[1] if len(args) > 1 else 0)
            if dim == 1 and all(isinstance(x, _StemOut) and x._which == i and x._value is None and x._owner is a[0]._owner
                                for i, x in enumerate(a)):
                return _StemOut(a[0]._owner, 3, tuple(x._in for x in a))
        unwrap = lambda v: v.materialize() if isinstance(v, _StemOut) else ([unwrap(x) for x in v] if isinstance(v, (list, tuple)) else v)
        return func(*[unwrap(v) for v in args], **{k: unwrap(v) for k, v in kwargs.items()})


class FusedStems:
    def __init__(self, completor):
        self.completor = completor
        self.original = {}
        self._packed, self._key = None, None
        self.fused_calls = 0

    def packed(self):
        c = self.completor
        params = [p for n in _STEMS + ("conv_disp_fuse",) for p in getattr(c, n).parameters()]
        key = tuple((p.data_ptr(), p._version) for p in params)
        if key != self._key:
            self._packed, self._key = pack_stem_weights(c), key
        return self._packed


def fuse_completor_stems(completor):
    """completor: the model's DisparityCompletor instance (model.disp_completor).  Returns the FusedStems handle."""
    h = FusedStems(completor)
    h.packed()                                                   # validates the widths now
    for i, n in enumerate(_STEMS):
        seq = getattr(completor, n)
        h.original[n] = type(seq).forward.__get__(seq)

        def stem_forward(x, _i=i, _n=n):
            if isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and not torch.is_autocast_enabled() \
                    and not (x.requires_grad and torch.is_grad_enabled()):
                return _StemOut(h, _i, x)
            return h.original[_n](x)
        seq.forward = stem_forward
    fuse = completor.conv_disp_fuse
    h.original["conv_disp_fuse"] = type(fuse).forward.__get__(fuse)

    def fuse_forward(x):
        if isinstance(x, _StemOut) and x._which == 3 and x._value is None:
            h.fused_calls += 1
            return completor_stems(x._in[0], x._in[1], x._in[2], h.packed())
        return h.original["conv_disp_fuse"](x.materialize() if isinstance(x, _StemOut) else x)
    fuse.forward = fuse_forward
    return h


def unfuse_completor_stems(completor):
    for n in _STEMS + ("conv_disp_fuse",):
        m = getattr(completor, n)
        if "forward" in m.__dict__:
            del m.forward
