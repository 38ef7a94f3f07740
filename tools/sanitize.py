"""Small shapes through the kernels added in round 2, for `compute-sanitizer --tool memcheck|racecheck` (one tool per gpurun call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tcs_b200 as tcs
g = torch.Generator().manual_seed(0)
for (B, H, W1, W2) in [(1, 3, 200, 72), (1, 2, 130, 300), (2, 2, 64, 64)]:
    f1 = torch.randn(B, 128, H, W1, generator=g).cuda(); f2 = torch.randn(B, 128, H, W2, generator=g).cuda()
    xs = torch.arange(W1, dtype=torch.float32).view(1, 1, 1, W1)
    coords = (xs - torch.rand(B, 1, H, W1, generator=g) * (W1 / 4)).cuda()
    coords.view(-1)[3] = float("nan"); coords.view(-1)[7] = -500.0
    for prec in ("fp16x3", "bf16"):
        out = tcs.CorrBlock1D(f1, f2, mode="alternate", precision=prec)(coords)
        assert torch.isfinite(out).all()
grid = torch.rand(2, 2, 12, 20, generator=g).cuda() * 24 - 2
nets = [torch.randn(2, 20, 12 >> l, 20 >> l, generator=g).cuda() for l in range(3)]
outs = tcs.warp_hidden_states(nets, grid)
T = torch.eye(4).repeat(3, 1, 1).cuda(); T[:, :3, 3] = 1.0
print(tcs.cal_relative_transformation(T, T)[0, 0, 0].item(), [o.shape for o in outs])
torch.cuda.synchronize()
print("sanitize script ok")
