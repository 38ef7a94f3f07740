"""Per-kernel timings of libtcs_b200 at the bench shape (CUDA events, inputs larger than L2).  Development aid.

    python tools/kbench.py [--B 8] [--hw 136 240] [--gran 32|64|128] [--only lookup,build,...]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tcs_b200  # noqa: E402
from tcs_b200 import sequence  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--hw", type=int, nargs=2, default=[136, 240])
ap.add_argument("--gran", type=int, default=0)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--only", default="")
ap.add_argument("--smooth", action="store_true", help="piecewise-smooth disparity instead of i.i.d.")
args = ap.parse_args()
B, (H, W), C = args.B, args.hw, 256
dev = torch.device("cuda")
if args.gran:
    try:
        from cuda.bindings import runtime as cudart
    except ImportError:
        from cuda import cudart
    print("set L2 fetch granularity", args.gran, cudart.cudaDeviceSetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity, args.gran),
          cudart.cudaDeviceGetLimit(cudart.cudaLimit.cudaLimitMaxL2FetchGranularity))
only = set(args.only.split(",")) if args.only else None
HBM = 6538.0


def timeit(name, fn, nbytes, reps=args.reps, flops=0.0):
    """Median-free but launch-overhead-free: `reps` calls captured in one CUDA graph, replayed 3 times."""
    if only and name.split("[")[0] not in only:
        return
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(reps):
            fn()
    graph.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    med = sorted(ts)[1] * 1e3
    extra = "  %.1f TFLOP/s" % (flops / med / 1e6) if flops else ""
    print("%-28s %9.1f us   alg %8.1f MB  %7.0f GB/s  %5.1f%% of %g%s" % (name, med, nbytes / 1e6, nbytes / med / 1e3, 100 * nbytes / med / 1e3 / HBM, HBM, extra))
    del graph


g = torch.Generator().manual_seed(0)
f1 = torch.randn(B, C, H, W, generator=g).to(dev)
f2 = torch.randn(B, C, H, W, generator=g).to(dev)
npix = B * H * W
xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
if args.smooth:
    base = torch.nn.functional.interpolate(torch.rand(B, 1, H // 8, W // 8, generator=g) * (W / 16), size=(H, W), mode="bilinear")
    disp = 0.5 + base + 0.05 * torch.randn(B, 1, H, W, generator=g)
else:
    disp = 0.5 + torch.rand(B, 1, H, W, generator=g) * (W / 16)
coords = (xs - disp).to(dev)

for prec in ("fp16", "fp16x3"):
    nb = npix * C * (4 + (4 if "x3" in prec else 2))
    timeit("prepass[%s]" % prec, lambda: tcs_b200.normalized_operands(f1, prec), nb)
timeit("prepass[n32]", lambda: tcs_b200.normalized_operands(f1, want_hi=False, want_n32=True), npix * C * 8)
for prec in ("fp16", "fp16x3", "bf16x3"):
    a_hi, a_lo, _ = tcs_b200.normalized_operands(f1, prec)
    b_hi, b_lo, _ = tcs_b200.normalized_operands(f2, prec)
    flat, levels = tcs_b200.corr.alloc_pyramid(B, H, W, W, 4, dev)
    ptrs = [lv.data_ptr() for lv in levels]
    from tcs_b200 import _lib
    st = lambda: torch.cuda.current_stream().cuda_stream
    nb = npix * C * 2 * (4 if "x3" in prec else 2) + npix * W * 4 * 1.875
    timeit("build[%s]" % prec, lambda: _lib.call("tcs_corr_build", a_hi.data_ptr(), a_lo.data_ptr() if a_lo is not None else None,
                                                  b_hi.data_ptr(), b_lo.data_ptr() if b_lo is not None else None, *ptrs, B, H, W, W, C, 4,
                                                  _lib.PRECISIONS[prec], st()), nb, flops=2.0 * npix * W * C * (3 if "x3" in prec else 1))
    del a_hi, a_lo, b_hi, b_lo, flat, levels
for prec in ("fp16", "fp16x3"):
    flat, levels = tcs_b200.corr.alloc_pyramid(B, H, W, W, 4, dev)
    ptrs = [lv.data_ptr() for lv in levels]
    nb = npix * C * 8 + npix * W * 4 * 1.875
    timeit("fused[%s]" % prec, lambda: _lib.call("tcs_corr_build_fused", f1.data_ptr(), f2.data_ptr(), *ptrs, B, H, W, W, C, 4,
                                                   _lib.PRECISIONS[prec], st()), nb, flops=2.0 * npix * W * C * (3 if "x3" in prec else 1))
    del flat, levels
blk = tcs_b200.CorrBlock1D(f1, f2, precision="fp16x3")
timeit("lookup", lambda: blk(coords), 308 * npix)
wenc = torch.randn(64, 36, device=dev) * 0.3
benc = torch.randn(64, device=dev) * 0.1
timeit("lookup+conv1x1[fused]", lambda: blk.lookup_encoded(coords, wenc, benc), (4 + 160 + 256) * npix)
timeit("lookup+conv1x1[torch]", lambda: torch.relu(torch.nn.functional.conv2d(blk(coords), wenc[:, :, None, None], benc)), (4 + 160 + 256) * npix)
timeit("argmax", lambda: blk.argmax_disp(), npix * W * 4)
alt = tcs_b200.CorrBlock1D(f1, f2, mode="alternate")
timeit("lookup_alt", lambda: alt(coords), npix * (1024 + 144 + 4), reps=5)
del alt, blk
K, K_inv = sequence.synthetic_intrinsics(B, 4 * H, 4 * W, dev)
fwd, inv = sequence.relative_pose(torch.stack([sequence.synthetic_pose(0, s) for s in range(B)]), torch.stack([sequence.synthetic_pose(1, s) for s in range(B)]))
fwd, inv = fwd.to(dev), inv.to(dev)
base = torch.full((B, 1), 0.25, device=dev)
dd = disp.to(dev)
timeit("warp+cost", lambda: tcs_b200.warp_with_cost(dd, f1, fwd, K, K_inv, base, cur_fmap=f2, per_sample_mean=True), npix * (4 + 1024 * 3 + 12))
timeit("warp+cost[no fmap]", lambda: tcs_b200.warp_with_cost(dd, f1, fwd, K, K_inv, base, cur_fmap=f2, per_sample_mean=True, want_fmap=False),
       npix * (4 + 1024 * 2 + 12))
grid = tcs_b200.get_backward_grid(dd, inv, K, K_inv, base)
timeit("backward_grid", lambda: tcs_b200.get_backward_grid(dd, inv, K, K_inv, base), npix * 12)
for i in range(3):
    net = torch.tanh(torch.randn(B, 128, H >> i, W >> i, generator=g)).to(dev)
    timeit("sample[l%d]" % i, lambda: tcs_b200.sample_planar(net, grid), net.numel() * 8 + grid.numel() * 4)
    grid = tcs_b200.halve_grid(grid)

# ---- "next" row, rank 2: the per-iteration stencils; beside them the reference's own op sequence (grouped conv2d on
# padded copies) written out with torch ops on the GPU
import torch.nn.functional as F  # noqa: E402

ggrad = torch.randn(B, 2, H, W, generator=g).to(dev)


def ref_gradient_xy(disp):                                           # geo_utils.py:115-132
    pad = F.pad(disp, (1, 1, 1, 1), mode="replicate")
    k = torch.zeros(2, 1, 3, 3, device=disp.device)
    k[:, :, 1, 1] = -1
    k[0, :, 1, 2] = 1
    k[1, :, 2, 1] = 1
    gr = F.conv2d(pad.repeat(1, 2, 1, 1), k, groups=2)
    return gr, (gr[:, :1].abs() < 5) & (gr[:, 1:].abs() < 5)


def ref_grad_candidates(disp, level=2):                              # geo_utils.py:73-101
    N = disp.shape[0]
    k = torch.zeros(8, 1, 3, 3, device=disp.device)
    k[:, :, 1, 1] = -1
    for i, (v, u) in enumerate([(0, 0), (0, 1), (0, 2), (1, 2), (2, 2), (2, 1), (2, 0), (1, 0)]):
        k[i, :, v, u] = 1
    cands = []
    for i in range(level):
        p = 1 + i
        dp = F.pad(disp, (p, p, p, p))
        Hp, Wp = H + 2 * p, W + 2 * p
        ys, xs = torch.meshgrid(torch.arange(Hp, device=disp.device), torch.arange(Wp, device=disp.device), indexing="ij")
        coord = torch.stack([xs, ys], 0).float()[None].repeat(N, 1, 1, 1)
        cd = torch.cat((coord, dp), 1).reshape(-1, 1, Hp, Wp).repeat(1, 8, 1, 1)
        cands.append(F.conv2d(cd, k, groups=8, dilation=p).reshape(N, 3, 8, H, W))
    gr = torch.cat(cands, 2)
    c = torch.cross(gr, torch.roll(gr, -2, 2), dim=1)
    return -c[:, :2] / c[:, 2:]


timeit("gradient_xy", lambda: tcs_b200.disp2disp_gradient_xy(dd), npix * 13)
timeit("gradient_xy[torch ops]", lambda: ref_gradient_xy(dd), npix * 13)
timeit("grad_candidates[l2]", lambda: tcs_b200.disp2disp_grad_candidates(dd, 2), npix * 132)
timeit("grad_candidates[l2, torch ops]", lambda: ref_grad_candidates(dd, 2), npix * 132)
timeit("propagate", lambda: tcs_b200.propagate_disparity(ggrad, dd), npix * 120)


def ref_upsample(flow, mask, factor=4):                              # tc_stereo.py:75-88
    N, D, Hh, Ww = flow.shape
    m = mask.view(N, 1, 9, factor, factor, Hh, Ww)
    m = torch.softmax(m - torch.max(m, dim=2, keepdim=True)[0], dim=2)
    up = F.unfold(factor * flow, [3, 3], padding=1).view(N, D, 9, 1, 1, Hh, Ww)
    up = torch.sum(m * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(N, D, factor * Hh, factor * Ww)


umask = torch.randn(B, 144, H, W, generator=g).to(dev)
timeit("convex_upsample", lambda: tcs_b200.convex_upsample(dd, umask, 4, True), npix * (4 + 144 * 4 + 64))
timeit("convex_upsample[torch ops]", lambda: ref_upsample(dd, umask, 4), npix * (4 + 144 * 4 + 64))
