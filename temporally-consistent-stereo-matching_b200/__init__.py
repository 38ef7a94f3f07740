"""B200-native cost-volume hot path of TC-Stereo behind the reference's own interfaces.

The directory name carries hyphens (it is the project's name), so import it through the `tcs_b200` shim at
the repository root, or with importlib.import_module("temporally-consistent-stereo-matching_b200").

    from tcs_b200 import CorrBlock1D, warp, get_backward_grid, bilinear_sampler, install

Everything computes in libtcs_b200.so (hand-written sm_100a CUDA, C-ABI in include/tcs_b200.h).  Importing
this package loads that library and fails if it has not been built: there is no CPU or PyTorch fallback.
"""
from . import _lib

_lib.load()

from .corr import CorrBlock1D, build_pyramid, normalized_operands  # noqa: E402
from .geo import (bilinear_sampler, cal_relative_transformation, get_backward_grid, halve_grid,  # noqa: E402
                  sample_planar, warp, warp_hidden_states, warp_with_cost, WarpCarry,
                  disp2disp_gradient_xy, disp2disp_grad_candidates, propagate_disparity, convex_upsample)
from .dropin import install, uninstall, strip_asserts, restore_asserts  # noqa: E402
from .graphed import graph_modules, ungraph_modules  # noqa: E402
from .completor import completor_stems, fuse_completor_stems, unfuse_completor_stems, pack_stem_weights  # noqa: E402
from .train import DifferentiableCorrBlock1D, LazyCostVolume, init_loss  # noqa: E402
from .sequence import HotPathRunner, hot_path_frame, shard_sequences  # noqa: E402

__all__ = ["CorrBlock1D", "DifferentiableCorrBlock1D", "LazyCostVolume", "init_loss", "build_pyramid", "normalized_operands", "warp", "warp_with_cost", "WarpCarry", "get_backward_grid",
           "disp2disp_gradient_xy", "disp2disp_grad_candidates", "propagate_disparity", "convex_upsample",
           "bilinear_sampler", "sample_planar", "halve_grid", "warp_hidden_states", "cal_relative_transformation",
           "install", "uninstall", "strip_asserts", "restore_asserts", "graph_modules", "ungraph_modules", "completor_stems", "fuse_completor_stems", "unfuse_completor_stems", "pack_stem_weights", "HotPathRunner", "hot_path_frame", "shard_sequences"]
