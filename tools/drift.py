"""End-to-end drift of the build precisions through the UNMODIFIED reference model (north_star's
"EPE drift after 32 iterations" gate).  Three stages, because the reference model cannot travel to the GPU box
and the kernels cannot run in the build container:

  1. python tools/drift.py inputs      (build container, CPU)  random-init TCStereo -> fmaps -> drift_work/fmaps.npz
  2. python tools/drift.py gpu         (GPU box, via gpurun)   libtcs_b200 pyramids per precision -> gpurun_out/drift_<prec>.npz
  3. python tools/drift.py eval        (build container, CPU)  reference forward with each injected pyramid vs the
                                                               untouched reference -> profiles/r01_drift.md

The injected object is the reference's own CorrBlock1D with corr_pyramid / cost_volume replaced by the GPU-built
levels; the lookups, GRUs and everything else stay the reference's CPU code, so the number isolates the build.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
WORK = os.path.join(ROOT, "drift_work")
OUT = os.path.join(ROOT, "gpurun_out")
REF = "/root/reference"
PRECISIONS = ["fp32", "fp16x3", "bf16x3", "fp16", "bf16"]
HEIGHT, WIDTH, ITERS = 256, 320, 32


def reference_model():
    cp = types.ModuleType("cupy")
    cp.int32, cp.float32 = int, float
    cp.memoize = lambda for_each_device=False: (lambda f: f)
    cp.cuda = types.SimpleNamespace()
    sys.modules["cupy"] = cp
    sys.path.insert(0, REF)
    import core.tc_stereo as tcs
    args = types.SimpleNamespace(hidden_dims=[128] * 3, shared_backbone=True, corr_levels=4, corr_radius=4, n_downsample=2,
                                 context_norm="none", slow_fast_gru=False, n_gru_layers=3, mixed_precision=False,
                                 init_thres=0.5, temporal=True)
    torch.manual_seed(1234)
    model = tcs.TCStereo(args).eval()
    g = torch.Generator().manual_seed(1234)
    img1 = torch.rand(1, 3, HEIGHT, WIDTH, generator=g) * 255
    img2 = torch.rand(1, 3, HEIGHT, WIDTH, generator=g) * 255
    return tcs, model, img1, img2


def stage_inputs():
    tcs, model, img1, img2 = reference_model()
    captured = {}
    orig = tcs.CorrBlock1D

    class Spy(orig):
        def __init__(self, fmap1, fmap2, **kw):
            captured["fmap1"], captured["fmap2"] = fmap1.detach().numpy().copy(), fmap2.detach().numpy().copy()
            super().__init__(fmap1, fmap2, **kw)

    tcs.CorrBlock1D = Spy
    with torch.no_grad():
        model(img1, img2, iters=1, test_mode=True)
    tcs.CorrBlock1D = orig
    os.makedirs(WORK, exist_ok=True)
    np.savez(os.path.join(WORK, "fmaps.npz"), **captured)
    print("wrote", os.path.join(WORK, "fmaps.npz"), captured["fmap1"].shape)


def stage_gpu():
    import tcs_b200
    z = np.load(os.path.join(WORK, "fmaps.npz"))
    f1, f2 = torch.from_numpy(z["fmap1"]).cuda(), torch.from_numpy(z["fmap2"]).cuda()
    os.makedirs(OUT, exist_ok=True)
    for prec in PRECISIONS:
        _, levels = tcs_b200.build_pyramid(f1, f2, 4, prec)
        np.savez(os.path.join(OUT, "drift_%s.npz" % prec), **{"level%d" % l: lv.cpu().numpy() for l, lv in enumerate(levels)})
        print(prec, "done")


def stage_eval():
    tcs, model, img1, img2 = reference_model()
    orig = tcs.CorrBlock1D

    def run(block_cls):
        tcs.CorrBlock1D = block_cls
        with torch.no_grad():
            out = model(img1, img2, iters=ITERS, test_mode=True)
        tcs.CorrBlock1D = orig
        return out["flow"].numpy(), out["flow_q"].numpy()

    def injected(levels, noise=0.0):
        class Injected(orig):
            def __init__(self, fmap1, fmap2, **kw):
                super().__init__(fmap1, fmap2, **kw)
                if levels is not None:
                    B, H, W1 = fmap1.shape[0], fmap1.shape[2], fmap1.shape[3]
                    for l in range(4):
                        self.corr_pyramid[l] = torch.from_numpy(levels["level%d" % l]).reshape(B * H * W1, 1, 1, -1)
                if noise:
                    g = torch.Generator().manual_seed(7)
                    self.corr_pyramid = [c + noise * torch.randn(c.shape, generator=g) for c in self.corr_pyramid]
        return Injected

    ref_flow, ref_q = run(orig)
    rows = []
    floor_flow, floor_q = run(injected(None, noise=1e-7))
    rows.append(("fp32 re-ordering noise floor (reference volume + N(0,1e-7))", np.abs(floor_flow - ref_flow).mean(), np.abs(floor_q - ref_q).mean(), None))
    for prec in PRECISIONS:
        path = os.path.join(OUT, "drift_%s.npz" % prec)
        if not os.path.exists(path):
            continue
        lv = dict(np.load(path))
        tcs.CorrBlock1D = orig
        with torch.no_grad():
            own = orig(torch.from_numpy(np.load(os.path.join(WORK, "fmaps.npz"))["fmap1"]), torch.from_numpy(np.load(os.path.join(WORK, "fmaps.npz"))["fmap2"]))
        vol_err = float(np.abs(lv["level0"].reshape(-1) - own.corr_pyramid[0].numpy().reshape(-1)).max())
        flow, q = run(injected(lv))
        rows.append(("libtcs_b200 build, precision %s" % prec, np.abs(flow - ref_flow).mean(), np.abs(q - ref_q).mean(), vol_err))
    lines = ["# End-to-end disparity drift after %d GRU iterations (reference TCStereo.forward on CPU, random-init weights, %dx%d noise images)" % (ITERS, HEIGHT, WIDTH), "",
             "Mean |d disparity| against the untouched fp32 reference; `flow` is full resolution, `flow_q` 1/4 resolution. The",
             "random-init network amplifies any perturbation of the volume (SURVEY.md section 0), so the first row is the floor", "any correct fp32 implementation sits on.", "",
             "| corr volume | mean abs d flow (px, full res) | mean abs d flow_q (px, 1/4 res) | max abs d level 0 vs reference |", "|---|---|---|---|"]
    for name, a, b, v in rows:
        lines.append("| %s | %.3e | %.3e | %s |" % (name, a, b, "-" if v is None else "%.2e" % v))
    lines += ["", "mean |flow| of the reference output: %.2f px (full res)" % np.abs(ref_flow).mean(), ""]
    text = "\n".join(lines)
    open(os.path.join(ROOT, "profiles", "r01_drift.md"), "w").write(text)
    print(text)


if __name__ == "__main__":
    {"inputs": stage_inputs, "gpu": stage_gpu, "eval": stage_eval}[sys.argv[1]]()
