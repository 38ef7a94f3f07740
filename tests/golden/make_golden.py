"""Generates tests/golden/*.npz by running the REFERENCE's own code (imported from /root/reference) on
seeded inputs.  Runs only in the build container (the reference does not travel to the GPU box); the
.npz files it writes are committed and are what the tests read.

    python tests/golden/make_golden.py [/root/reference]

What is reference-executed: core/corr.py (CorrBlock1D build, pyramid, __call__, argmax_disp,
get_cost_volume), core/utils/utils.py (bilinear_sampler), core/utils/geo_utils.py (warp,
get_backward_grid and their helpers), the hidden-state warp loop and the matching cost copied as call
sequences from core/tc_stereo.py:139-140,159-163 (they are inline code in TCStereo.forward, not functions).
What cannot be: softsplat_func.forward is a cupy-JIT CUDA kernel (cupy is not installable here and the CPU
branch is assert(False)); it is replaced by oracle.tcs_oracle.softsplat_scatter, so the scatter itself is
pinned by source reading only — everything around it in softsplat()/warp() is the reference's code.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
sys.path.insert(0, ROOT)

from oracle import tcs_oracle as orc  # noqa: E402


def import_reference():
    cp = types.ModuleType("cupy")
    cp.int32 = int
    cp.float32 = float
    cp.memoize = lambda for_each_device=False: (lambda f: f)
    cp.cuda = types.SimpleNamespace()
    sys.modules["cupy"] = cp
    sys.path.insert(0, REF)
    import core.corr as rcorr
    import core.utils.utils as rutils
    import core.utils.geo_utils as rgeo
    import core.utils.splatting.softsplat as rsplat

    def splat_apply(ten_in, ten_flow):
        B, C, H, W = ten_in.shape
        xs = torch.arange(W, dtype=torch.float32).view(1, 1, W)
        ys = torch.arange(H, dtype=torch.float32).view(1, H, 1)
        tx = (xs + ten_flow[:, 0]).numpy()
        ty = (ys + ten_flow[:, 1]).numpy()
        return torch.from_numpy(orc.softsplat_scatter(ten_in.numpy(), tx, ty))

    rsplat.softsplat_func.apply = staticmethod(splat_apply)
    return rcorr, rutils, rgeo


def camera(B, H, W, rng, baseline=0.25, tz=None):
    """Feature-resolution pinhole camera + a small forward/yaw motion with per-sample jitter.  tz: a fixed translation
    along the optical axis instead (metres): large positive values put near points behind the current camera (the
    forward warp's invalid branch), negative ones put them behind the previous camera (get_backward_grid's -1 branch)."""
    K = np.zeros((B, 3, 3), np.float64)
    K[:, 0, 0] = K[:, 1, 1] = 0.5 * W
    K[:, 0, 2], K[:, 1, 2], K[:, 2, 2] = 0.5 * W - 0.5, 0.5 * H - 0.5, 1.0
    Kinv = np.linalg.inv(K)
    T = np.tile(np.eye(4), (B, 1, 1))
    for b in range(B):
        yaw = np.deg2rad(0.6 + 0.3 * b)
        c, s = np.cos(yaw), np.sin(yaw)
        cam2world = np.array([[c, 0, s, 0.02 * (b + 1)], [0, 1, 0, 0.01], [-s, 0, c, (0.12 + 0.05 * b) if tz is None else tz], [0, 0, 0, 1.0]])
        T[b] = np.linalg.inv(cam2world)
    base = np.full((B, 1), baseline)
    f32 = lambda a: a.astype(np.float32)
    return f32(K), f32(Kinv), f32(T), f32(np.linalg.inv(T)), f32(base)


def corr_case(rcorr, name, B, C, H, W, seed, correlated_shift):
    g = torch.Generator().manual_seed(seed)
    fmap1 = torch.randn(B, C, H, W, generator=g)
    if correlated_shift:   # a real disparity signal so that argmax_disp's confidence mask is not vacuous
        fmap2 = torch.roll(fmap1, -correlated_shift, dims=3) + 0.3 * torch.randn(B, C, H, W, generator=g)
    else:
        fmap2 = torch.randn(B, C, H, W, generator=g)
    blk = rcorr.CorrBlock1D(fmap1, fmap2, num_levels=4, radius=4)
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    coords = xs - torch.rand(B, 1, H, W, generator=g) * (W / 4)
    flat = coords.view(-1)
    n = flat.numel()
    pick = torch.randperm(n, generator=g)
    flat[pick[: n // 50]] = -7.5                      # far left: every tap out of range
    flat[pick[n // 50: n // 25]] = W + 9.25           # far right
    flat[pick[n // 25: n // 10]] = flat[pick[n // 25: n // 10]].round()   # exact integers
    coords2 = torch.cat([coords, torch.zeros_like(coords)], dim=1)        # the model passes 1 channel; accept 2
    lookup = blk(coords2)
    sparse_disp, main_cost, mask = blk.argmax_disp()
    out = {
        "fmap1": fmap1.numpy(), "fmap2": fmap2.numpy(), "coords": coords2.numpy(),
        "lookup": lookup.numpy(), "cost_volume": blk.get_cost_volume().numpy(),
        "sparse_disp": sparse_disp.numpy(), "main_cost": main_cost.numpy(), "mask": mask.numpy(),
    }
    for l in range(4):
        out["level%d" % l] = blk.corr_pyramid[l].view(B, H, W, -1).numpy()
    np.savez(os.path.join(HERE, name + ".npz"), **out)
    print(name, "mask density %.3f" % mask.mean().item(), {k: v.shape for k, v in out.items() if k.startswith("level")})


def warp_case(rutils, rgeo, name, B, C, H, W, seed, tz=None):
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    K, Kinv, T, Tinv, base = camera(B, H, W, rng, tz=tz)
    disp = 0.5 + torch.rand(B, 1, H, W, generator=g) * (W / 4)
    disp.view(-1)[:: 37] = 0.0                        # exact zeros hit the clip(disp, 1e-3) branch
    fmap = torch.randn(B, C, H, W, generator=g)
    cur_fmap = torch.randn(B, C, H, W, generator=g)
    tK, tKinv, tT, tTinv, tb = map(torch.from_numpy, (K, Kinv, T, Tinv, base))
    wdisp, wfmap, wmask = rgeo.warp(disp, fmap, tT, tK, tKinv, tb)
    cost = torch.sum(F.normalize(cur_fmap, dim=1) * F.normalize(wfmap, dim=1), dim=1, keepdim=True) * wmask   # tc_stereo.py:139-140
    disp_init = (wdisp * wmask).clamp_min(0)
    if tz is not None:   # the completed disparity is a network output, not the warped one: independent values reach z <= 0
        disp_init = 0.5 + torch.rand(B, 1, H, W, generator=g) * (W / 4)
    grid = rgeo.get_backward_grid(disp_init, tTinv, tK, tKinv, tb)
    nets = [torch.tanh(torch.randn(B, 8, H >> i, W >> i, generator=g)) for i in range(3)]
    warped, gg, grids = [], grid, [grid]
    for net in nets:                                  # tc_stereo.py:161-163
        warped.append(rutils.bilinear_sampler(net.float(), gg.permute(0, 2, 3, 1)))
        gg = 0.5 * F.interpolate(gg, scale_factor=0.5, mode="bilinear", align_corners=True)
        grids.append(gg)
    out = {"disp": disp.numpy(), "fmap": fmap.numpy(), "cur_fmap": cur_fmap.numpy(), "K": K, "K_inv": Kinv,
           "rel_T": T, "rel_T_inv": Tinv, "baseline": base,
           "warped_disp": wdisp.numpy(), "warped_fmap": wfmap.numpy(), "warped_mask": wmask.numpy(), "cost": cost.numpy(),
           "disp_init": disp_init.numpy(), "backward_grid": grid.numpy()}
    for i in range(3):
        out["net%d" % i] = nets[i].numpy()
        out["warped_net%d" % i] = warped[i].numpy()
        out["grid%d" % i] = grids[i].numpy()
    np.savez(os.path.join(HERE, name + ".npz"), **out)
    print(name, "splat mask density %.3f" % wmask.mean().item(), "grid -1 fraction %.3f" % (grid == -1).float().mean().item())


def stencil_case(rgeo, name, N, H, W, seed):
    """The per-iteration 3x3 stencils: geo_utils.py:73-101, :115-132, update.py:259-289."""
    import core.update as rupdate

    class Args:
        n_downsample = 2

    g = torch.Generator().manual_seed(seed)
    disp = torch.rand(N, 1, H, W, generator=g) * 24
    disp[0, 0, 2:5, 3:9] = 0.0                                      # flow_q is clipped at 0: exact zeros are common
    disp[-1, 0, :, W // 2:] += 11.0                                 # a depth edge: gradients beyond the |g| < 5 mask
    grad = torch.randn(N, 2, H, W, generator=g)
    grads, edge = rgeo.disp2disp_gradient_xy(disp)
    cands1 = rgeo.disp2disp_grad_candidates(disp, level=1)
    cands2 = rgeo.disp2disp_grad_candidates(disp, level=2)         # the level the model uses (update.py:202)
    prop, matrix = rupdate.DispRefine(Args()).propagate_disparity(grad, disp)
    # convex upsampling (tc_stereo.py:75-88) reads only self.args: call it unbound on a stand-in
    import core.tc_stereo as rtc
    up_mask = torch.randn(N, 9 * 16, H, W, generator=g) * 3
    up = rtc.TCStereo.upsample_flow(types.SimpleNamespace(args=Args()), -disp, up_mask)
    out = {"disp": disp.numpy(), "grad": grad.numpy(), "grads": grads.numpy(), "edge_mask": edge.numpy(),
           "cands1": cands1.numpy(), "cands2": cands2.numpy(), "prop": prop.numpy(), "matrix": matrix.numpy(),
           "up_mask": up_mask.numpy(), "up": up.numpy()}
    np.savez(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()}, "edge mask density %.3f" % edge.float().mean().item())


def main():
    torch.manual_seed(1234)
    rcorr, rutils, rgeo = import_reference()
    with torch.no_grad():
        stencil_case(rgeo, "stencils_small", N=2, H=11, W=19, seed=77)
        corr_case(rcorr, "corr_small", B=2, C=128, H=3, W=40, seed=1234, correlated_shift=5)
        corr_case(rcorr, "corr_oddwidth", B=1, C=128, H=2, W=78, seed=4321, correlated_shift=0)
        warp_case(rutils, rgeo, "warp_small", B=2, C=128, H=12, W=16, seed=1234)
        # round 2: the branches the first set left vacuous: an odd width WITH a disparity signal (argmax mask
        # density > 0, floor-pooled levels), near points behind the current camera (invalid sources in the forward warp) and
        # behind the previous one (get_backward_grid's where(valid, uv, -1))
        corr_case(rcorr, "corr_oddshift", B=1, C=128, H=3, W=77, seed=99, correlated_shift=4)
        warp_case(rutils, rgeo, "warp_forward_jump", B=2, C=128, H=12, W=16, seed=77, tz=1.2)
        warp_case(rutils, rgeo, "warp_backward_jump", B=2, C=128, H=12, W=16, seed=78, tz=-1.2)


if __name__ == "__main__":
    main()
