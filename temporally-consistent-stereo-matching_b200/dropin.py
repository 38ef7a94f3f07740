"""Rebinds the hot-path names inside the reference's model module.

core/tc_stereo.py imports CorrBlock1D, warp, get_backward_grid, cal_relative_transformation and
bilinear_sampler by name (tc_stereo.py:6-8), so the drop-in is an assignment into that module's namespace;
TCStereo.forward (tc_stereo.py:114-116,137,142,159-163,177) then runs unmodified on libtcs_b200.
"""
_saved = {}
_NAMES = ("CorrBlock1D", "warp", "get_backward_grid", "bilinear_sampler")


def install(tc_stereo_module, precision=None, mode=None, fuse_motion_encoder=None, stencils=None):
    """tc_stereo_module: the imported `core.tc_stereo`.  Returns the dict of names that were replaced.

    stencils: the imported `core.update`.  When given, the three per-iteration 3x3 stencils also run as single kernels
    (SURVEY.md section 8f rank 2): `disp2disp_gradient_xy` (tc_stereo.py:192), `disp2disp_grad_candidates`
    (update.py:202) and `DispRefine.propagate_disparity` (update.py:294), and so does the convex upsampling
    `TCStereo.upsample_flow` (tc_stereo.py:75-88, rank 3); fp32, inference only.

    fuse_motion_encoder: the imported `core.update`.  When given, corr_fn(coords) returns a deferred lookup and
    BasicMotionEncoder.forward (update.py:103-112) evaluates `relu(convc1(corr))` with the fused lookup + 1x1
    kernel; fp32 only — under autocast the reference's conv runs in fp16."""
    from . import corr, geo

    block = corr.CorrBlock1D
    if precision is not None or mode is not None or fuse_motion_encoder is not None:
        lazy = fuse_motion_encoder is not None

        class _Configured(corr.CorrBlock1D):
            def __init__(self, fmap1, fmap2, num_levels=4, radius=4, thres=0.2):
                super().__init__(fmap1, fmap2, num_levels, radius, thres, precision=precision, mode=mode)

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
    new = {"CorrBlock1D": block, "warp": geo.warp, "get_backward_grid": geo.get_backward_grid,
           "bilinear_sampler": geo.bilinear_sampler}
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
