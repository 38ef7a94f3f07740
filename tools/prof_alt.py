"""A few calls of the tensor-core alternate lookup at BASELINE config 5's shape, for ncu / timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tcs_b200 as tcs
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
B, H, W = 2, 272, 480
g = torch.Generator().manual_seed(1)
f1 = torch.randn(B, 256, H, W, generator=g).cuda()
f2 = torch.randn(B, 256, H, W, generator=g).cuda()
xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
coords = (xs - (0.5 + torch.rand(B, 1, H, W, generator=g) * (W / 16.0))).cuda()
smooth = (xs - 12.3).expand(B, 1, H, W).contiguous().cuda()
blk = tcs.CorrBlock1D(f1, f2, mode="alternate", precision=prec)
for c, name in ((coords, "iid"), (smooth, "smooth")):
    for _ in range(3):
        blk(c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        blk(c)
    e1.record()
    torch.cuda.synchronize()
    print(prec, name, "us per call", 100 * e0.elapsed_time(e1))
