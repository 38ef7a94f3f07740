"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.py; or
/root/reference in the build container) so that its own TCStereo.forward, CorrBlock1D, warp, ... can run beside the
kernels: on the GPU box with the reference's own soft-splat CUDA kernel (oracle/cupy_shim.py), on the CPU with the
oracle's restatement of the scatter patched over softsplat_func.apply (softsplat.py's CPU branch is assert(False)).

Only tests/, bench.py's reference legs (`--impl reference`, `--impl reference-gpu`) and __graft_entry__.smoke()
import this.  The product package never does.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF_DIRS = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")
_loaded = None


def reference_root():
    for d in _REF_DIRS:
        if os.path.exists(os.path.join(d, "core", "tc_stereo.py")):
            return d
    return None


def load():
    """-> namespace(tc_stereo, update, corr, utils, geo, softsplat, root).  Raises if the reference is not installed."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = reference_root()
    if root is None:
        raise ImportError("the reference is not installed: run `python baseline/install_ref.py` in the build container "
                          "(baseline/_ref/ travels to the GPU box)")
    from oracle import cupy_shim
    cupy_shim.install()
    if root not in sys.path:
        sys.path.insert(0, root)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # SyntaxWarning / FutureWarning of the 2023 sources under torch 2.11
        import core.corr as rcorr
        import core.tc_stereo as rtc
        import core.update as rupdate
        import core.utils.geo_utils as rgeo
        import core.utils.utils as rutils
        import core.utils.splatting.softsplat as rsplat
    _loaded = types.SimpleNamespace(tc_stereo=rtc, update=rupdate, corr=rcorr, utils=rutils, geo=rgeo, softsplat=rsplat,
                                    root=root, _gpu_apply=rsplat.softsplat_func.apply)
    return _loaded


def load_init_loss(root=None):
    """The reference's own init_loss (train_stereo.py:138-182) as a function.  train_stereo.py cannot be imported (wandb,
    skimage, pykitti are absent), so the function definition is cut out of its source with `ast` and compiled on its own; it
    needs only torch and torch.nn.functional."""
    import ast
    import torch
    import torch.nn.functional as F
    root = root or reference_root()
    path = os.path.join(root, "train_stereo.py") if root else None
    if not path or not os.path.exists(path):
        raise ImportError("train_stereo.py of the reference is not installed (python baseline/install_ref.py)")
    src = open(path).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "init_loss"]
    if len(fn) != 1:
        raise ImportError("init_loss not found in %s" % path)
    ns = {"torch": torch, "F": F}
    exec(compile(ast.Module(body=fn, type_ignores=[]), path, "exec"), ns)
    return ns["init_loss"]


def use_cpu_splat(ref, threaded=False):
    """CPU runs: replace the cupy kernel launch by a restatement of softsplat.py:284-335 — the numpy oracle's
    (sequential, the parity checker) or, threaded=True, torch_port.splat (index_add_ on all host threads: the CPU
    baseline's timing leg)."""
    from oracle import tcs_oracle as orc
    from oracle import torch_port

    def splat_apply(ten_in, ten_flow):
        if ten_in.is_cuda:
            return ref._gpu_apply(ten_in, ten_flow)
        if threaded:
            return torch_port.splat(ten_in, ten_flow)
        B, C, H, W = ten_in.shape
        xs = torch.arange(W, dtype=torch.float32).view(1, 1, W)
        ys = torch.arange(H, dtype=torch.float32).view(1, H, 1)
        tx = (xs + ten_flow[:, 0]).numpy()
        ty = (ys + ten_flow[:, 1]).numpy()
        return torch.from_numpy(orc.softsplat_scatter(ten_in.numpy(), tx, ty))

    ref.softsplat.softsplat_func.apply = staticmethod(splat_apply)


def model_args(**over):
    """The flag set every shipped script uses (*_evaluate.sh)."""
    a = dict(hidden_dims=[128] * 3, shared_backbone=True, corr_levels=4, corr_radius=4, n_downsample=2,
             context_norm="none", slow_fast_gru=False, n_gru_layers=3, mixed_precision=False, init_thres=0.5,
             temporal=True)
    a.update(over)
    return types.SimpleNamespace(**a)


def make_model(device="cpu", seed=1234, **over):
    ref = load()
    torch.manual_seed(seed)                       # train_stereo.py:295
    model = ref.tc_stereo.TCStereo(model_args(**over)).eval().to(device)
    for p in model.parameters():
        p.requires_grad_(False)
    return model


def synthetic_sequence(frames, height, width, seed=1234, device="cpu", batch=1):
    """Images U(0,255), TartanAir intrinsics scaled to the image (evaluate_stereo.py:138-142), baseline 0.25 and
    world2cam poses of a camera advancing 0.05 m and yawing 0.2 degrees per frame."""
    g = torch.Generator().manual_seed(seed)
    imgs = [(torch.rand(batch, 3, height, width, generator=g) * 255, torch.rand(batch, 3, height, width, generator=g) * 255)
            for _ in range(frames)]
    K = torch.tensor([[0.5 * width, 0, 0.5 * width], [0, 0.5 * width, 0.5 * height], [0, 0, 1.0]], dtype=torch.float32)
    poses = []
    for t in range(frames):
        yaw = np.deg2rad(0.2 * t)
        c, s = float(np.cos(yaw)), float(np.sin(yaw))
        cam2world = torch.tensor([[c, 0, s, 0.002 * t], [0, 1, 0, 0], [-s, 0, c, 0.05 * t], [0, 0, 0, 1.0]], dtype=torch.float64)
        poses.append(torch.linalg.inv(cam2world).float())
    rep = lambda m: m[None].repeat(batch, 1, 1).contiguous().to(device)
    return ([(a.to(device), b.to(device)) for a, b in imgs], rep(K), [rep(p) for p in poses],
            torch.full((batch, 1), 0.25, device=device))


def run_sequence(model, imgs, K, poses, baseline, iters, on_frame=None):
    """The temporal loop of evaluate_stereo.py:170-197 (state carried as flow_q / net_list / fmap1 / previous_T).
    Images are already padded to a multiple of 32 by the caller's choice of size.  -> list of forward outputs."""
    outs, state = [], None
    with torch.no_grad():
        for t, ((im1, im2), T) in enumerate(zip(imgs, poses)):
            params = None
            if state is not None:
                params = {"K": K, "T": T, "previous_T": state["T"], "last_disp": state["flow_q"],
                          "last_net_list": state["net_list"], "fmap1": state["fmap1"], "baseline": baseline}
            out = model(im1, im2, iters=iters, test_mode=True, params=params)
            state = {"T": T, "flow_q": out["flow_q"], "net_list": out["net_list"], "fmap1": out["fmap1"]}
            outs.append(out)
            if on_frame is not None:
                on_frame(t, out)
    return outs


def hot_path_frame(ref, fmap1, fmap2, coords_seq, state, rel_T, rel_T_inv, K, K_inv, baseline, num_levels=4, radius=4,
                   timers=None):
    """The hot path's share of one temporal frame written as the REFERENCE's own calls, in the order and with the inline
    expressions of TCStereo.forward (tc_stereo.py:116,137-140,159-163,177).  The learned blocks between them are not
    run (their outputs are inputs here, exactly as in tcs_b200.hot_path_frame).  Works on CPU (after use_cpu_splat) and
    on the GPU (reference's own splat kernel).  timers: optional callable(label) invoked at the phase boundaries."""
    import torch.nn.functional as F
    tick = timers or (lambda label: None)
    tick("start")
    blk = ref.corr.CorrBlock1D(fmap1, fmap2, radius=radius, num_levels=num_levels)
    tick("build")
    last_disp, last_fmap1, last_nets = state
    wdisp, wfmap, wmask = ref.geo.warp(last_disp, last_fmap1, rel_T, K, K_inv, baseline)
    cost = torch.sum(F.normalize(fmap1, dim=1) * F.normalize(wfmap, dim=1), dim=1, keepdim=True)
    cost = cost * wmask
    grid = ref.geo.get_backward_grid(wdisp.clamp_min(0), rel_T_inv, K, K_inv, baseline)
    warped = []
    for net in last_nets:
        warped.append(ref.utils.bilinear_sampler(net.float(), grid.permute(0, 2, 3, 1)))
        grid = 0.5 * F.interpolate(grid, scale_factor=0.5, mode="bilinear", align_corners=True)
    tick("warp")
    out = None
    for it in range(coords_seq.shape[0]):
        out = blk(coords_seq[it])
    tick("lookups")
    return {"corr": out, "sparse_disp": wdisp, "cost": cost, "mask": wmask, "warped_net": warped}
