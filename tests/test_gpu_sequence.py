"""BASELINE config 2: a TartanAir-shape 480x640 sequence of 50 frames with synthetic poses, temporal warp of the
disparity / features / hidden states enabled, run through HotPathRunner; a few frames are checked against the
oracle with the runner's own carried state as input, and the whole run must stay finite."""
import numpy as np
import pytest
import torch

from conftest import assert_close, assert_exact
from oracle import tcs_oracle as orc

pytestmark = pytest.mark.gpu


def test_fifty_frame_sequence_480x640():
    import tcs_b200
    from tcs_b200 import sequence
    dev = torch.device("cuda")
    H, W = sequence.feature_shape(480, 640)
    assert (H, W) == (120, 160)
    B, C, iters, frames = 1, 256, 32, 50                # BASELINE config 2: 32 GRU iterations, 128-channel hidden states
    check = {0, 1, 24, 49}
    g = torch.Generator().manual_seed(1234)
    runner = tcs_b200.HotPathRunner()
    K, K_inv = sequence.synthetic_intrinsics(B, 4 * H, 4 * W, dev)
    baseline = torch.full((B, 1), 0.25, device=dev)
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    base_f = torch.randn(B, C, H, W + frames, generator=g)          # a scene that slides by one feature pixel per frame
    prev_T = None
    for t in range(frames):
        f1 = (base_f[..., t:t + W] + 0.05 * torch.randn(B, C, H, W, generator=g)).contiguous()
        f2 = torch.roll(f1, -4, dims=3) + 0.3 * torch.randn(B, C, H, W, generator=g)
        coords = xs - (2.0 + 6.0 * torch.rand(iters, B, 1, H, W, generator=g))
        nets = [torch.tanh(torch.randn(B, 128, H >> i, W >> i, generator=g)) for i in range(3)]
        T = torch.stack([sequence.synthetic_pose(t, s) for s in range(B)])
        kw = {}
        if prev_T is not None:
            fwd, inv = sequence.relative_pose(prev_T, T)
            kw = dict(rel_T=fwd.to(dev), rel_T_inv=inv.to(dev), K=K, K_inv=K_inv, baseline=baseline)
        if t in check and prev_T is not None:
            st = (runner.last_disp.cpu().numpy(), runner.last_fmap1.cpu().numpy(), [n.cpu().numpy() for n in runner.last_net_list])
        out = runner.frame(f1.to(dev), f2.to(dev), coords.to(dev), net_list=[n.to(dev) for n in nets], **kw)
        for k in ("corr", "sparse_disp", "cost", "mask"):
            assert torch.isfinite(out[k]).all(), "frame %d: %s not finite" % (t, k)
        if t in check:
            lv = [x.cpu().numpy() for x in out["corr_fn"]._levels]
            assert_close(lv[0], orc.corr_volume(f1.numpy(), f2.numpy(), np.float64), rtol=1e-5, atol=1e-6, what="frame %d volume" % t)
            assert_close(out["corr"].cpu().numpy(), orc.corr_lookup(lv, coords[-1].numpy(), 4), what="frame %d lookup" % t)
            if prev_T is None:
                rd, rc, rm = orc.argmax_disp(lv[0])
                assert_exact(out["mask"].cpu().numpy(), rm, what="frame 0 argmax mask")
                assert_exact(out["sparse_disp"].cpu().numpy(), rd, what="frame 0 sparse disp")
                assert rm.mean() > 0.5
            else:
                Kn, Kin, bn = K.cpu().numpy(), K_inv.cpu().numpy(), baseline.cpu().numpy()
                rd, rf, rm = orc.warp(st[0], st[1], fwd.numpy(), Kn, Kin, bn, per_sample_mean=True)
                assert_exact(out["mask"].cpu().numpy(), rm, what="frame %d splat mask" % t)
                assert_close(out["sparse_disp"].cpu().numpy(), rd, rtol=1e-5, atol=2e-6, what="frame %d warped disparity" % t)
                grid = orc.backward_grid(out["sparse_disp"].cpu().numpy(), inv.numpy(), Kn, Kin, bn)
                for a, r in zip(out["warped_net"], orc.warp_hidden_states(st[2], grid)):
                    assert_close(a.cpu().numpy(), r, rtol=1e-5, atol=2e-6, what="frame %d hidden state" % t)
        prev_T = T


def test_runner_is_bitwise_repeatable():
    """A HotPathRunner warps through sorted contributor lists from its first temporal frame on (the carried
    transposition from the third): two runs over the same frames give identical bits, which the reference's atomic
    scatter does not."""
    import tcs_b200
    from tcs_b200 import sequence
    dev = torch.device("cuda")
    B, C, H, W, iters, frames = 2, 128, 40, 96, 2, 5
    K, K_inv = sequence.synthetic_intrinsics(B, 4 * H, 4 * W, dev)
    baseline = torch.full((B, 1), 0.25, device=dev)
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)

    def run():
        g = torch.Generator().manual_seed(99)
        runner = tcs_b200.HotPathRunner()
        outs, prev_T = [], None
        for t in range(frames):
            f1 = torch.randn(B, C, H, W, generator=g).to(dev)
            f2 = torch.randn(B, C, H, W, generator=g).to(dev)
            coords = (xs - 8.0 * torch.rand(iters, B, 1, H, W, generator=g)).to(dev)      # i.i.d. disparities: long, irregular lists
            nets = [torch.randn(B, 8, H >> i, W >> i, generator=g).to(dev) for i in range(3)]
            T = torch.stack([sequence.synthetic_pose(t, s) for s in range(B)])
            kw = {}
            if prev_T is not None:
                fwd, inv = sequence.relative_pose(prev_T, T)
                kw = dict(rel_T=fwd.to(dev), rel_T_inv=inv.to(dev), K=K, K_inv=K_inv, baseline=baseline)
            out = runner.frame(f1, f2, coords, net_list=nets, **kw)
            outs.append([out[k].clone() for k in ("corr", "sparse_disp", "cost", "mask")] + [n.clone() for n in (out["warped_net"] or [])])
            prev_T = T
        return outs

    a, b = run(), run()
    for fa, fb in zip(a, b):
        for x, y in zip(fa, fb):
            assert torch.equal(x, y)
