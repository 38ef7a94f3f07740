import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tcs_b200 as tcs
for (B, H, W) in [(8, 96, 312), (8, 136, 240), (2, 272, 480)]:
    f = torch.randn(B, 256, H, W, device="cuda")
    for kb in (False, True):
        for _ in range(3): tcs.normalized_operands(f, "fp16x3", kblocked=kb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): tcs.normalized_operands(f, "fp16x3", kblocked=kb)
        e1.record(); torch.cuda.synchronize()
        us = 50 * e0.elapsed_time(e1)
        print((B, H, W), "kblocked" if kb else "pixel-major", "%.1f us" % us, "%.0f GB/s" % (f.numel() * 8 / us / 1e3))
