"""Times the cost-volume initialisation loss (forward + backward to the volume) at a training crop: the reference's own
init_loss (train_stereo.py:138-182, torch ops on the materialised [B,W2,H,W1] cost volume) against tcs_b200.init_loss."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tcs_b200 as tcs
from oracle import ref_model

ref_init_loss = ref_model.load_init_loss()
for (B, H, W, k) in [(8, 80, 180, 3), (8, 96, 312, 3)]:
    g = torch.Generator().manual_seed(1)
    f1 = torch.randn(B, 256, H, W, generator=g).cuda().requires_grad_(True)
    f2 = torch.randn(B, 256, H, W, generator=g).cuda().requires_grad_(True)
    flow = (-4.0 * torch.rand(B, 1, 4 * H, 4 * W, generator=g) * W / 5).cuda()
    valid = torch.ones(B, 1, 4 * H, 4 * W).cuda()
    blk = tcs.DifferentiableCorrBlock1D(f1, f2)

    def step(fused):
        cv = blk.get_cost_volume()
        loss, _ = (tcs.init_loss if fused else ref_init_loss)(cv if fused else cv.materialize(), flow, valid, k=k, scale=0.25, threshold=0.5)
        (dvol,) = torch.autograd.grad(loss, blk._vol)
        return dvol

    for fused in (False, True):
        for _ in range(3): step(fused)
        torch.cuda.synchronize(); torch.cuda.reset_peak_memory_stats(); base = torch.cuda.memory_allocated()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): step(fused)
        e1.record(); torch.cuda.synchronize()
        print((B, H, W, k), "kernels" if fused else "reference torch ops", "%.3f ms per loss fwd+bwd (incl. 4 .item() syncs)" % (e0.elapsed_time(e1) / 10),
              "peak extra memory %.0f MB" % ((torch.cuda.max_memory_allocated() - base) / 1e6))
