import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import tcs_b200 as tcs
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_parity import make_coords
for (B, H, W1, W2) in [(1, 136, 240, 240), (1, 7, 480, 480), (1, 5, 78, 78), (1, 6, 200, 72), (1, 3, 130, 300), (1, 272, 480, 480)]:
    for prec in ("fp16x3", "bf16"):
        g = torch.Generator().manual_seed(W1 + W2)
        f1 = torch.randn(B, 256, H, W1, generator=g).cuda(); f2 = torch.randn(B, 256, H, W2, generator=g).cuda()
        coords = make_coords(B, H, W1, 11).cuda()
        alt = tcs.CorrBlock1D(f1, f2, mode="alternate", precision=prec)
        _, levels = tcs.build_pyramid(f1, f2, 4, prec, fused=False)
        pyr = tcs.CorrBlock1D.from_levels(levels)
        a, p = alt(coords), pyr(coords)
        bad = (a != p)
        n = int(bad.sum())
        msg = ""
        if n:
            idx = bad.nonzero()
            planes = torch.bincount(idx[:, 1], minlength=36).view(4, 9).sum(1).tolist()
            d = (a - p).abs()
            first = idx[0].tolist()
            c = coords[first[0], 0, first[2], first[3]].item()
            msg = "per-level %s max|d| %.3g first %s coord %.4f a %.6g p %.6g; w1 of bad min/max %d/%d" % (planes, d.max().item(), first, c, a[tuple(first)].item(), p[tuple(first)].item(), int(idx[:, 3].min()), int(idx[:, 3].max()))
        print((B, H, W1, W2), prec, "mismatches", n, "of", a.numel(), msg, flush=True)
