"""The oracle against outputs of the reference itself (tests/golden/*.npz, see make_golden.py).

Integer-valued and mask outputs must match exactly; floats within |d| <= 1e-6 + 1e-5 |ref|."""
import numpy as np
import pytest

from conftest import assert_close, assert_exact, load_golden
from oracle import tcs_oracle as orc

CORR_CASES = ["corr_small", "corr_oddwidth", "corr_oddshift"]
WARP_CASES = ["warp_small", "warp_forward_jump", "warp_backward_jump"]


@pytest.mark.parametrize("case", CORR_CASES)
def test_volume_and_pyramid(case):
    g = load_golden(case)
    vol = orc.corr_volume(g["fmap1"], g["fmap2"])
    levels = orc.corr_pyramid(vol, 4)
    for l in range(4):
        assert_close(levels[l], g["level%d" % l], what="%s level %d" % (case, l))


@pytest.mark.parametrize("case", CORR_CASES)
def test_lookup_on_reference_pyramid(case):
    g = load_golden(case)
    out = orc.corr_lookup([g["level%d" % l] for l in range(4)], g["coords"], radius=4)
    assert_close(out, g["lookup"], what=case + " lookup")


@pytest.mark.parametrize("case", CORR_CASES)
def test_lookup_alternate_identity(case):
    g = load_golden(case)
    out = orc.corr_lookup_alternate(g["fmap1"], g["fmap2"], g["coords"], 4, 4)
    assert_close(out, g["lookup"], rtol=1e-5, atol=2e-6, what=case + " alternate lookup")


@pytest.mark.parametrize("case", CORR_CASES)
def test_argmax_and_cost_volume(case):
    g = load_golden(case)
    cv = orc.masked_cost_volume(g["level0"])
    assert_exact(cv, g["cost_volume"], what=case + " cost volume")
    sparse_disp, main_cost, mask = orc.argmax_disp(g["level0"])
    assert_exact(mask, g["mask"], what=case + " argmax mask")
    assert_exact(sparse_disp, g["sparse_disp"], what=case + " sparse_disp")
    assert_exact(main_cost, g["main_cost"], what=case + " main_cost")


@pytest.mark.parametrize("case", WARP_CASES)
def test_warp_forward(case):
    g = load_golden(case)
    d, f, m = orc.warp(g["disp"], g["fmap"], g["rel_T"], g["K"], g["K_inv"], g["baseline"])
    assert_exact(m, g["warped_mask"], what="splat mask")
    assert_close(d, g["warped_disp"], rtol=1e-5, atol=1e-5, what="warped disparity")
    assert_close(f, g["warped_fmap"], rtol=1e-5, atol=1e-5, what="warped features")
    cost = orc.matching_cost(g["cur_fmap"], g["warped_fmap"], g["warped_mask"])
    assert_close(cost, g["cost"], what="matching cost")


@pytest.mark.parametrize("case", WARP_CASES)
def test_backward_grid_and_hidden_state_warp(case):
    g = load_golden(case)
    grid = orc.backward_grid(g["disp_init"], g["rel_T_inv"], g["K"], g["K_inv"], g["baseline"])
    assert_exact(grid == -1, g["backward_grid"] == -1, what="behind-the-camera entries (where(valid, uv, -1))")
    if case == "warp_backward_jump":
        assert 0.3 < (g["backward_grid"] == -1).mean() < 0.9          # the branch is not vacuous in this case
    assert_close(grid, g["backward_grid"], rtol=1e-5, atol=1e-4, what="backward grid")
    gg = g["grid0"]
    for i in range(3):
        assert_close(orc.bilinear_sample(g["net%d" % i], gg), g["warped_net%d" % i], rtol=1e-5, atol=2e-6,
                     what="hidden state level %d" % i)
        if i < 2:
            nxt = orc.grid_halve(gg)
            assert_close(nxt, g["grid%d" % (i + 1)], rtol=1e-5, atol=1e-5, what="halved grid %d" % i)
            gg = g["grid%d" % (i + 1)]
    outs = orc.warp_hidden_states([g["net%d" % i] for i in range(3)], g["grid0"])
    for i in range(3):
        assert_close(outs[i], g["warped_net%d" % i], rtol=1e-4, atol=1e-4, what="chained hidden state level %d" % i)


def test_pooling_floor_on_odd_width():
    v = np.arange(2 * 7, dtype=np.float32).reshape(1, 1, 2, 7)
    lv = orc.corr_pyramid(v, 3)
    assert lv[1].shape[-1] == 3 and lv[2].shape[-1] == 1
    np.testing.assert_array_equal(lv[1][0, 0, 0], [0.5, 2.5, 4.5])


def test_lookup_far_out_of_range_is_zero():
    lv = [np.ones((1, 1, 4, 16 >> l), np.float32) for l in range(4)]
    coords = np.full((1, 1, 1, 4), -100.0, np.float32)
    assert not orc.corr_lookup(lv, coords, 4).any()
    coords[:] = np.nan
    assert not np.isnan(orc.corr_lookup(lv, coords, 4)).all() or True  # NaN coords must not crash


def test_torch_port_matches_golden():
    """The multi-threaded torch CPU port used as bench.py's CPU baseline computes the same things."""
    import torch
    from oracle import torch_port as tp
    t = torch.from_numpy
    for case in CORR_CASES:
        g = load_golden(case)
        pyr, cv = tp.build_block(t(g["fmap1"]), t(g["fmap2"]))
        B, H, W = g["level0"].shape[:3]
        for l in range(4):
            assert_close(pyr[l].view(B, H, W, -1).numpy(), g["level%d" % l], what="port level %d" % l)
        assert_close(tp.lookup(pyr, t(g["coords"])).numpy(), g["lookup"], what="port lookup")
        d, c, m = tp.argmax_disp(cv)
        assert_exact(m.numpy(), g["mask"], what="port argmax mask")
        assert_exact(d.numpy(), g["sparse_disp"], what="port sparse_disp")
    g = load_golden("warp_small")
    d, f, m = tp.warp(t(g["disp"]), t(g["fmap"]), t(g["rel_T"]), t(g["K"]), t(g["K_inv"]), t(g["baseline"]))
    assert_exact(m.numpy(), g["warped_mask"], what="port splat mask")
    assert_close(f.numpy(), g["warped_fmap"], rtol=1e-5, atol=1e-5, what="port warped features")
    assert_close(tp.matching_cost(t(g["cur_fmap"]), f, m).numpy(), g["cost"], rtol=1e-5, atol=2e-6, what="port cost")
    grid = tp.backward_grid(t(g["disp_init"]), t(g["rel_T_inv"]), t(g["K"]), t(g["K_inv"]), t(g["baseline"]))
    assert_close(grid.numpy(), g["backward_grid"], rtol=1e-5, atol=1e-4, what="port backward grid")
    outs = tp.warp_hidden([t(g["net%d" % i]) for i in range(3)], t(g["backward_grid"]))
    for i in range(3):
        assert_close(outs[i].numpy(), g["warped_net%d" % i], rtol=1e-5, atol=2e-6, what="port hidden %d" % i)


def test_stencil_oracles_match_reference_bit_for_bit():
    """geo_utils.py:73-101, :115-132 and update.py:259-289 restated in numpy: every product has an exact small-integer
    factor, so the restatement reproduces the reference's outputs exactly (level 2 is what the model uses)."""
    g = load_golden("stencils_small")
    grads, edge = orc.disp_gradient_xy(g["disp"])
    assert_exact(grads, g["grads"], what="disp2disp_gradient_xy")
    assert np.array_equal(edge, g["edge_mask"]) and 0 < g["edge_mask"].mean() < 1
    assert_exact(orc.disp_grad_candidates(g["disp"], 1), g["cands1"], what="disp2disp_grad_candidates level 1")
    assert_exact(orc.disp_grad_candidates(g["disp"], 2), g["cands2"], what="disp2disp_grad_candidates level 2")
    prop, matrix = orc.disp_propagate(g["grad"], g["disp"])
    assert_exact(prop, g["prop"], what="propagate_disparity")
    assert_exact(matrix, g["matrix"], what="propagate_disparity matrix")
    up = orc.convex_upsample(-g["disp"], g["up_mask"], 4, True)                  # tc_stereo.py:75-88 (exp: not bit-exact)
    assert_close(up, g["up"], rtol=1e-5, atol=1e-5, what="upsample_flow")


def test_init_loss_oracle_matches_the_reference_function():
    """train_stereo.py:138-182 executed by the reference itself (make_golden_init_loss.py) against the restatement: the loss
    terms, the metric, and d loss / d cost_volume on w2 <= w1 — outside that triangle the volume is exactly 0 (corr.py:28-31),
    torch.topk breaks the ties among those zeros in an unspecified order, and no gradient reaches the features from there."""
    g = load_golden("init_loss_small")
    o = orc.init_loss(g["cost_volume"], g["flow_gt"], g["valid"], k=int(g["k"]), scale=0.25, threshold=float(g["threshold"]),
                      valid_interp=g["valid_interp"])
    for name in ("loss", "gt_loss", "nm_loss"):
        assert abs(float(o[name]) - float(g[name])) <= 1e-6, name
    assert abs(float(o["forward_mask_rate"]) - float(g["forward_mask_rate"])) <= 1e-7
    assert 0.3 < o["mask"].mean() < 0.9, "the case must have masked AND unmasked pixels"
    B, D, H, W = g["cost_volume"].shape
    tri = np.arange(D).reshape(1, D, 1, 1) <= np.arange(W).reshape(1, 1, 1, W)
    assert_close(o["grad_cost_volume"] * tri, g["grad_cost_volume"] * tri, rtol=1e-6, atol=1e-9, what="d loss / d cost_volume")
    assert ((g["grad_cost_volume"] * tri) != 0).sum() > 500
    # without torch's interpolation of `valid` the restatement differs only where `valid == 1` hangs on a last bit
    own = orc.init_loss(g["cost_volume"], g["flow_gt"], g["valid"], k=int(g["k"]), scale=0.25, threshold=float(g["threshold"]))
    assert (own["mask"] != o["mask"]).sum() <= 4
