"""Generates tests/golden/init_loss_small.npz by running the REFERENCE's own init_loss (train_stereo.py:138-182, cut out of the
source with ast because the module itself needs wandb) on a cost volume made by the reference's own CorrBlock1D.

    python tests/golden/make_golden_init_loss.py [/root/reference]

Runs only in the build container; the .npz is committed and is what tests/test_oracle_golden.py reads."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_model  # noqa: E402
import make_golden  # noqa: E402


def main():
    rcorr, _, _ = make_golden.import_reference()
    init_loss = ref_model.load_init_loss(REF)
    g = torch.Generator().manual_seed(2024)
    B, C, H, W, k = 2, 64, 6, 40, 3
    disp = torch.rand(B, 1, H, W, generator=g) * 12.0                       # quarter-resolution disparity of the scene
    f2 = torch.randn(B, C, H, W, generator=g)
    xs = torch.arange(W).view(1, 1, 1, W) - disp.round().long()             # fmap1[w1] ~ fmap2[w1 - disp]
    f1 = torch.gather(f2, 3, xs.clamp(0, W - 1).expand(B, C, H, W)) + 0.4 * torch.randn(B, C, H, W, generator=g)
    cv = rcorr.CorrBlock1D(f1, f2).get_cost_volume().detach().clone().requires_grad_(True)
    # full-resolution ground truth: flow = -4 * disparity, piecewise constant over 4x4 blocks plus sub-pixel jitter; some
    # out-of-range (index_gt < 0), some huge (mag >= max_flow * scale) and an invalid band
    flow = -4.0 * disp.repeat_interleave(4, 2).repeat_interleave(4, 3) + 0.8 * torch.rand(B, 1, 4 * H, 4 * W, generator=g)
    flow[0, 0, :4, :8] = -4.0 * 45.0
    flow[1, 0, 8:12, 100:108] = -4.0 * 800.0
    valid = torch.ones(B, 1, 4 * H, 4 * W)
    valid[:, :, :, 60:75] = 0.0
    valid[0, 0, 13, :] = 0.0
    loss, metrics = init_loss(cv, flow, valid, k=k, scale=0.25, threshold=0.5)
    loss.backward()
    np.savez(os.path.join(HERE, "init_loss_small.npz"), fmap1=f1.numpy(), fmap2=f2.numpy(), cost_volume=cv.detach().numpy(),
             flow_gt=flow.numpy(), valid=valid.numpy(), k=np.int32(k), threshold=np.float32(0.5), loss=np.float32(loss.item()),
             gt_loss=np.float32(metrics["init_gt_loss"]), nm_loss=np.float32(metrics["init_nm_loss"]),
             forward_mask_rate=np.float32(metrics["forward_mask_rate"]), grad_cost_volume=cv.grad.numpy(),
             valid_interp=torch.nn.functional.interpolate(valid, scale_factor=0.25, mode="bilinear", align_corners=True).numpy())
    print("init_loss_small:", loss.item(), metrics)


if __name__ == "__main__":
    main()
