"""Times the fused lookup + 1x1 at 8 x 136x240: tensor-core form, CUDA-core form, and the plain lookup."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tcs_b200 as tcs

def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n

for (B, H, W) in [(8, 136, 240), (1, 136, 240)]:
    g = torch.Generator().manual_seed(0)
    f1, f2 = torch.randn(B, 256, H, W, generator=g).cuda(), torch.randn(B, 256, H, W, generator=g).cuda()
    blk = tcs.CorrBlock1D(f1, f2)
    coords = (torch.arange(W).view(1, 1, 1, W) - torch.rand(B, 1, H, W, generator=g) * 15).cuda()
    w, b = (torch.randn(64, 36, generator=g) * 0.3).cuda(), torch.randn(64, generator=g).cuda()
    conv = torch.nn.Conv2d(36, 64, 1).cuda()
    os.environ["TCS_B200_ENCODE_TC"] = "1"
    us_tc = timed(lambda: blk.lookup_encoded(coords, w, b))
    os.environ["TCS_B200_ENCODE_TC"] = "0"
    us_cc = timed(lambda: blk.lookup_encoded(coords, w, b))
    del os.environ["TCS_B200_ENCODE_TC"]                      # default: by size (corr._ENCODE_TC_MIN_PIXELS)
    us_lk = timed(lambda: blk(coords))
    with torch.no_grad():
        us_ref = timed(lambda: torch.relu(conv(blk(coords))))
    print((B, H, W), "tensor-core %.1f us | CUDA-core %.1f us | plain lookup %.1f us | lookup + cuDNN 1x1 + relu %.1f us" % (us_tc, us_cc, us_lk, us_ref))
