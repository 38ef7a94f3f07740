"""Rebinds the hot-path names inside the reference's model module.

core/tc_stereo.py imports CorrBlock1D, warp, get_backward_grid, cal_relative_transformation and
bilinear_sampler by name (tc_stereo.py:6-8), so the drop-in is an assignment into that module's namespace;
TCStereo.forward (tc_stereo.py:114-116,137,142,159-163,177) then runs unmodified on libtcs_b200.
"""
_saved = {}
_NAMES = ("CorrBlock1D", "warp", "get_backward_grid", "bilinear_sampler")


def install(tc_stereo_module, precision=None, mode=None):
    """tc_stereo_module: the imported `core.tc_stereo`.  Returns the dict of names that were replaced."""
    from . import corr, geo

    block = corr.CorrBlock1D
    if precision is not None or mode is not None:
        class _Configured(corr.CorrBlock1D):
            def __init__(self, fmap1, fmap2, num_levels=4, radius=4, thres=0.2):
                super().__init__(fmap1, fmap2, num_levels, radius, thres, precision=precision, mode=mode)
        _Configured.__name__ = "CorrBlock1D"
        block = _Configured
    new = {"CorrBlock1D": block, "warp": geo.warp, "get_backward_grid": geo.get_backward_grid,
           "bilinear_sampler": geo.bilinear_sampler}
    for name in _NAMES:
        if not hasattr(tc_stereo_module, name):
            raise AttributeError("%s has no attribute %r; is it the reference's core.tc_stereo?" % (tc_stereo_module, name))
        _saved.setdefault((id(tc_stereo_module), name), getattr(tc_stereo_module, name))
        setattr(tc_stereo_module, name, new[name])
    return new


def uninstall(tc_stereo_module):
    for name in _NAMES:
        old = _saved.pop((id(tc_stereo_module), name), None)
        if old is not None:
            setattr(tc_stereo_module, name, old)
