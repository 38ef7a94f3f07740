"""Rebinds the hot-path names inside the reference's model module.

core/tc_stereo.py imports CorrBlock1D, warp, get_backward_grid, cal_relative_transformation and
bilinear_sampler by name (tc_stereo.py:6-8), so the drop-in is an assignment into that module's namespace;
TCStereo.forward (tc_stereo.py:114-116,127,137,142,159-163,177) then runs unmodified on libtcs_b200.
"""
import torch
import torch.nn.functional as F

from .lazy import LazyTensorOps

_saved = {}
_NAMES = ("CorrBlock1D", "warp", "get_backward_grid", "bilinear_sampler", "cal_relative_transformation")


class _FrameContext:
    """What one TCStereo.forward leaves behind for the next call of the patched `warp` (fuse_cost=True): the current
    frame's fmap1 (seen by the CorrBlock1D constructor, tc_stereo.py:116, one statement before the warp, :137) and two
    alternating WarpCarry buffers (each frame's fmap1 transposed by its own cost kernel for the next frame's warp)."""

    def __init__(self):
        self.cur_fmap1 = None
        self.carries = None
        self.fused_calls = 0
        self.carried_calls = 0


_ctx = _FrameContext()


class LazyWarpedFmap(LazyTensorOps):
    """warp()'s second output under install(..., fuse_cost=True): the warped features, not materialised.

    TCStereo.forward uses them for exactly one expression (tc_stereo.py:139-140),

        cost = torch.sum(F.normalize(fmap1, dim=1) * F.normalize(warped_fmap1, dim=1), dim=1, keepdim=True) * sparse_mask

    which the cost kernel of tcs_warp_forward already evaluated against the current frame's fmap1 while it
    normalised the splat (the 256-channel warped map is never stored).  F.normalize(self, dim=1) -> a marker;
    normalised_fmap1 * marker -> a marker that remembers the other factor's shape; torch.sum(marker, dim=1,
    keepdim=True) -> the fused cost (already times the mask, and the mask is 0/1, so the reference's second
    multiplication changes nothing).  Any other use of the object materialises the warped features with the
    ordinary warp kernels and continues as a tensor."""

    _NORMALIZED, _PRODUCT = 1, 2

    def __init__(self, cost, shape, rerun, stage=0):
        self._cost, self._shape, self._rerun, self._stage, self._value = cost, torch.Size(shape), rerun, stage, None

    @property
    def shape(self):
        return self._shape

    def materialize(self):
        if self._stage != 0:
            raise RuntimeError("the fused matching cost was used outside the expression of tc_stereo.py:139-140; "
                               "install the drop-in without fuse_cost=True for this model")
        if self._value is None:
            self._value = self._rerun()
        return self._value

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        me = next((a for a in args if isinstance(a, LazyWarpedFmap)), None)
        name = getattr(func, "__name__", "")
        if me is not None and me._value is None:
            if me._stage == 0 and func is F.normalize and args[0] is me and \
                    (kwargs.get("dim", args[2] if len(args) > 2 else 1) == 1) and kwargs.get("p", 2.0) in (2, 2.0):
                return LazyWarpedFmap(me._cost, me._shape, None, cls._NORMALIZED)
            if me._stage == cls._NORMALIZED and name in ("mul", "__mul__", "__rmul__") and len(args) == 2:
                other = args[1] if args[0] is me else args[0]
                if isinstance(other, torch.Tensor) and other.shape == me._shape:
                    return LazyWarpedFmap(me._cost, me._shape, None, cls._PRODUCT)
            if me._stage == cls._PRODUCT and name == "sum" and args[0] is me and \
                    (kwargs.get("dim", args[1] if len(args) > 1 else None) in (1, [1], (1,))) and \
                    kwargs.get("keepdim", args[2] if len(args) > 2 else False):
                return me._cost
        unwrap = lambda a: a.materialize() if isinstance(a, LazyWarpedFmap) else a
        return func(*[unwrap(a) for a in args], **{k: unwrap(v) for k, v in kwargs.items()})


def _fused_warp(disp, fmap, relative_T, K, K_inv, baseline):
    """geo.warp for TCStereo.forward with the matching cost fused (ref: geo_utils.py:158-198 + tc_stereo.py:139-140)."""
    from . import geo

    cur = _ctx.cur_fmap1
    if cur is None or not isinstance(fmap, torch.Tensor) or cur.shape != fmap.shape or cur.device != fmap.device:
        return geo.warp(disp, fmap, relative_T, K, K_inv, baseline)
    if _ctx.carries is None:
        _ctx.carries = [geo.WarpCarry(), geo.WarpCarry()]
    fm = fmap if (fmap.dtype == torch.float32 and fmap.is_contiguous()) else None
    cin = next((c for c in _ctx.carries if fm is not None and c.matches(fm)), None)
    cout = next(c for c in _ctx.carries if c is not cin)
    d, _, m, cost = geo.warp_with_cost(disp, fmap, relative_T, K, K_inv, baseline, cur_fmap=cur, per_sample_mean=False,
                                       want_fmap=False, deterministic=True, carry_in=cin, carry_out=cout)
    _ctx.fused_calls += 1
    _ctx.carried_calls += cin is not None
    _ctx.cur_fmap1 = None

    def rerun():
        return geo.warp(disp, fmap, relative_T, K, K_inv, baseline)[1]

    return d, LazyWarpedFmap(cost, fmap.shape, rerun), m


def strip_asserts(*modules):
    """What `python -O` does, for the given (reference) modules only and at run time: every function and method defined
    in them gets its code object re-compiled from the module's own source with the assert statements left out.

    The reference guards its tensors with `assert not torch.isnan(x).any()` (geo_utils.py:14-15,26,52-53,211-235,
    corr.py:77-78, tc_stereo.py:160, update.py:27-391: ten or more per GRU iteration); each one is a device -> host
    synchronisation, which is what bounds the model's GPU time once the hot path is a handful of kernels (SURVEY.md
    section 3.1).  Nothing else of the modules changes (same source, same line numbers, same globals).  Undone by
    restore_asserts().  Returns the number of functions re-compiled."""
    import inspect
    import types

    count = 0
    for mod in modules:
        src = inspect.getsource(mod)
        top = compile(src, getattr(mod, "__file__", "<reference>"), "exec", optimize=1)
        stripped = {}

        def collect(code):
            for c in code.co_consts:
                if isinstance(c, types.CodeType):
                    stripped[(c.co_name, c.co_firstlineno)] = c
                    collect(c)
        collect(top)

        def functions(ns):
            for v in list(vars(ns).values()):
                f = v.__func__ if isinstance(v, (staticmethod, classmethod)) else v
                if isinstance(f, types.FunctionType) and f.__module__ == mod.__name__:
                    yield f
                elif isinstance(v, type) and v.__module__ == mod.__name__ and ns is mod:
                    yield from functions(v)
        for f in functions(mod):
            new = stripped.get((f.__code__.co_name, f.__code__.co_firstlineno))
            if new is not None and new.co_freevars == f.__code__.co_freevars and new is not f.__code__:
                _saved.setdefault(("code", id(f)), (f, f.__code__))
                f.__code__ = new
                count += 1
    return count


def restore_asserts():
    for key in [k for k in _saved if k[0] == "code"]:
        f, code = _saved.pop(key)
        f.__code__ = code


def install(tc_stereo_module, precision=None, mode=None, fuse_motion_encoder=None, stencils=None, fuse_cost=False, training=False):
    """tc_stereo_module: the imported `core.tc_stereo`.  Returns the dict of names that were replaced.

    fuse_cost: `warp` also evaluates the matching cost of tc_stereo.py:139-140 in its normalise kernel (against the
    fmap1 the CorrBlock1D constructor saw one statement earlier), never stores the 256-channel warped features, runs
    the deterministic list formulation, and hands each frame's transposed fmap1 to the next frame's call (WarpCarry).
    The second return value is then a LazyWarpedFmap (see there) instead of a tensor.

    stencils: the imported `core.update`.  When given, the three per-iteration 3x3 stencils also run as single kernels
    (SURVEY.md section 8f rank 2): `disp2disp_gradient_xy` (tc_stereo.py:192), `disp2disp_grad_candidates`
    (update.py:202) and `DispRefine.propagate_disparity` (update.py:294), and so does the convex upsampling
    `TCStereo.upsample_flow` (tc_stereo.py:75-88, rank 3); fp32, inference only.

    fuse_motion_encoder: the imported `core.update`.  When given, corr_fn(coords) returns a deferred lookup and
    BasicMotionEncoder.forward (update.py:103-112) evaluates `relu(convc1(corr))` with the fused lookup + 1x1
    kernel; fp32 only — under autocast the reference's conv runs in fp16."""
    from . import corr, geo

    if training:
        # gradients flow through the correlation block only (train.py); nothing that would cut one may be installed
        if fuse_motion_encoder is not None or stencils is not None or fuse_cost or mode not in (None, "pyramid"):
            raise ValueError("training=True installs the differentiable correlation block only: no fuse_motion_encoder, stencils, "
                             "fuse_cost or mode='alternate' (their kernels have no backward)")
        from .train import DifferentiableCorrBlock1D

        class _Trainable(DifferentiableCorrBlock1D):
            def __init__(self, fmap1, fmap2, num_levels=4, radius=4, thres=0.2):
                super().__init__(fmap1, fmap2, num_levels, radius, thres, precision=precision)
        _Trainable.__name__ = "CorrBlock1D"
        new = {"CorrBlock1D": _Trainable, "warp": geo.warp, "get_backward_grid": geo.get_backward_grid,
               "bilinear_sampler": geo.bilinear_sampler, "cal_relative_transformation": geo.cal_relative_transformation}
        for name in _NAMES:
            if not hasattr(tc_stereo_module, name):
                raise AttributeError("%s has no attribute %r; is it the reference's core.tc_stereo?" % (tc_stereo_module, name))
            _saved.setdefault((id(tc_stereo_module), name), getattr(tc_stereo_module, name))
            setattr(tc_stereo_module, name, new[name])
        return new

    block = corr.CorrBlock1D
    if precision is not None or mode is not None or fuse_motion_encoder is not None or fuse_cost:
        lazy = fuse_motion_encoder is not None

        class _Configured(corr.CorrBlock1D):
            def __init__(self, fmap1, fmap2, num_levels=4, radius=4, thres=0.2):
                super().__init__(fmap1, fmap2, num_levels, radius, thres, precision=precision, mode=mode)
                if fuse_cost:
                    _ctx.cur_fmap1 = self.fmap1          # the fp32 contiguous tensor the build read

            def __call__(self, coords):
                if lazy and self.mode == "pyramid" and self.num_levels == 4 and self.radius == 4:
                    return self.lazy(coords)
                return super().__call__(coords)
        _Configured.__name__ = "CorrBlock1D"
        block = _Configured
    if fuse_motion_encoder is not None:
        _patch_motion_encoder(fuse_motion_encoder)
    if stencils is not None:
        _patch_stencils(tc_stereo_module, stencils)
    new = {"CorrBlock1D": block, "warp": _fused_warp if fuse_cost else geo.warp, "get_backward_grid": geo.get_backward_grid,
           "bilinear_sampler": geo.bilinear_sampler, "cal_relative_transformation": geo.cal_relative_transformation}
    for name in _NAMES:
        if not hasattr(tc_stereo_module, name):
            raise AttributeError("%s has no attribute %r; is it the reference's core.tc_stereo?" % (tc_stereo_module, name))
        _saved.setdefault((id(tc_stereo_module), name), getattr(tc_stereo_module, name))
        setattr(tc_stereo_module, name, new[name])
    return new


def _patch_motion_encoder(update_module):
    import torch
    import torch.nn.functional as F
    from .corr import LazyLookup

    enc = update_module.BasicMotionEncoder
    _saved.setdefault((id(update_module), "BasicMotionEncoder.forward"), enc.forward)

    def forward(self, flow, corr):                      # update.py:103-112 with the first layer fused
        if isinstance(corr, LazyLookup):
            cor = corr.encode(self.convc1)              # relu(convc1(lookup)) in one kernel
        else:
            cor = F.relu(self.convc1(corr))
        cor = F.relu(self.convc2(cor))
        flo = F.relu(self.convf1(flow))
        flo = F.relu(self.convf2(flo))
        out = F.relu(self.conv(torch.cat([cor, flo], dim=1)))
        return torch.cat([out, flow], dim=1)

    enc.forward = forward


def _patch_stencils(tc_stereo_module, update_module):
    from . import geo

    _saved.setdefault((id(tc_stereo_module), "disp2disp_gradient_xy"), tc_stereo_module.disp2disp_gradient_xy)
    tc_stereo_module.disp2disp_gradient_xy = geo.disp2disp_gradient_xy
    _saved.setdefault((id(update_module), "disp2disp_grad_candidates"), update_module.disp2disp_grad_candidates)
    update_module.disp2disp_grad_candidates = geo.disp2disp_grad_candidates
    ref = update_module.DispRefine
    _saved.setdefault((id(update_module), "DispRefine.propagate_disparity"), ref.propagate_disparity)

    def propagate_disparity(self, disparity_grad, disparity_map):       # update.py:259-289
        return geo.propagate_disparity(disparity_grad, disparity_map)

    ref.propagate_disparity = propagate_disparity
    model = getattr(tc_stereo_module, "TCStereo", None)              # rank 3: the convex upsampling of the last iteration
    if model is not None:
        _saved.setdefault((id(tc_stereo_module), "TCStereo.upsample_flow"), model.upsample_flow)

        def upsample_flow(self, flow, mask, scale=True):                # tc_stereo.py:75-88
            return geo.convex_upsample(flow, mask, 2 ** self.args.n_downsample, scale)

        model.upsample_flow = upsample_flow


def uninstall(tc_stereo_module, update_module=None):
    if update_module is not None:
        old = _saved.pop((id(update_module), "BasicMotionEncoder.forward"), None)
        if old is not None:
            update_module.BasicMotionEncoder.forward = old
        old = _saved.pop((id(update_module), "disp2disp_grad_candidates"), None)
        if old is not None:
            update_module.disp2disp_grad_candidates = old
        old = _saved.pop((id(update_module), "DispRefine.propagate_disparity"), None)
        if old is not None:
            update_module.DispRefine.propagate_disparity = old
    old = _saved.pop((id(tc_stereo_module), "disp2disp_gradient_xy"), None)
    if old is not None:
        tc_stereo_module.disp2disp_gradient_xy = old
    old = _saved.pop((id(tc_stereo_module), "TCStereo.upsample_flow"), None)
    if old is not None:
        tc_stereo_module.TCStereo.upsample_flow = old
    for name in _NAMES:
        old = _saved.pop((id(tc_stereo_module), name), None)
        if old is not None:
            setattr(tc_stereo_module, name, old)
    _ctx.cur_fmap1, _ctx.carries = None, None
