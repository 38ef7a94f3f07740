"""CPU-side checks of the harness that runs the UNMODIFIED reference beside the kernels (oracle/ref_model.py,
oracle/cupy_shim.py, baseline/install_ref.py) and of the drop-in's deferred results.  No GPU needed: NVRTC compiles for
sm_100 without a device."""
import filecmp
import os
import sys

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_model  # noqa: E402

needs_ref = pytest.mark.skipif(ref_model.reference_root() is None, reason="reference not installed (baseline/install_ref.py)")


@needs_ref
def test_installed_reference_is_unmodified():
    src = "/root/reference"
    if not os.path.isdir(src):
        pytest.skip("no reference checkout on this machine")
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import install_ref
    dest = install_ref.install(src, quiet=True)
    for rel in install_ref.FILES:
        assert filecmp.cmp(os.path.join(src, rel), os.path.join(dest, rel), shallow=False), rel


@needs_ref
def test_reference_splat_kernel_compiles_through_the_cupy_shim():
    """The reference's own preprocessing (softsplat.py:27-216) + NVRTC on its own kernel string, for sm_100."""
    from oracle import cupy_shim
    ref = ref_model.load()
    ss = ref.softsplat
    ss.objCudacache.setdefault("device", "NVIDIA B200")          # cuda_kernel() asks torch for the device name otherwise
    text = open(os.path.join(ref.root, "core", "utils", "splatting", "softsplat.py")).read()
    kernel = text.split("cuda_kernel('softsplat_out', '''")[1].split("''', {")[0]
    t = torch.zeros(2, 258, 12, 16)
    key = ss.cuda_kernel("softsplat_out", kernel, {"tenIn": t, "tenFlow": torch.zeros(2, 2, 12, 16), "tenOut": torch.zeros_like(t)})
    src = ss.objCudacache[key]["strKernel"]
    assert "atomicAdd" in src and "SIZE_" not in src and "VALUE_" not in src and "OFFSET_" not in src
    cubin = cupy_shim.compile_source(src, ["-I /usr/local/cuda", "-I /usr/local/cuda/include"], arch="sm_100")
    assert len(cubin) > 1000 and b"softsplat_out" in cubin


@needs_ref
def test_reference_model_runs_a_temporal_sequence_on_cpu():
    ref = ref_model.load()
    ref_model.use_cpu_splat(ref)
    model = ref_model.make_model()
    imgs, K, poses, base = ref_model.synthetic_sequence(2, 64, 96)
    outs = ref_model.run_sequence(model, imgs, K, poses, base, iters=2)
    assert outs[1]["flow"].shape == (1, 1, 64, 96) and torch.isfinite(outs[1]["flow"]).all()
    # and the same frame through the reference's hot-path call sequence alone
    f = outs[0]["fmap1"]
    xs = torch.arange(24, dtype=torch.float32).view(1, 1, 1, 24)
    coords = (xs - 2.0).expand(1, 1, 16, 24)[None].contiguous()
    Ks = K * torch.tensor([0.25, 0.25, 1]).view(1, 3, 1)
    rel = ref.geo.cal_relative_transformation(poses[0], poses[1])
    out = ref_model.hot_path_frame(ref, f, f, coords, (-outs[0]["flow_q"], f, outs[0]["net_list"]), rel, torch.linalg.inv(rel), Ks,
                                   torch.linalg.inv(Ks), base)
    assert out["corr"].shape == (1, 36, 16, 24) and len(out["warped_net"]) == 3


def test_fused_cost_marker_resolves_only_the_models_expression():
    """dropin.LazyWarpedFmap (install(..., fuse_cost=True)): tc_stereo.py:139-140 resolves to the fused cost without
    materialising; anything else sees the materialised tensor; a half-used marker refuses."""
    from tcs_b200 import dropin
    cost = torch.full((2, 1, 3, 4), 7.0)
    f = torch.randn(2, 8, 3, 4)
    calls = []

    def mk():
        return dropin.LazyWarpedFmap(cost, (2, 8, 3, 4), lambda: calls.append(1) or torch.ones(2, 8, 3, 4))

    assert torch.sum(F.normalize(f, dim=1) * F.normalize(mk(), dim=1), dim=1, keepdim=True) is cost
    assert torch.sum(F.normalize(mk(), dim=1) * F.normalize(f, dim=1), dim=1, keepdim=True) is cost
    assert (F.normalize(mk(), dim=1) * F.normalize(f, dim=1)).sum(dim=1, keepdim=True) is cost
    assert not calls
    lz = mk()
    assert lz.shape == (2, 8, 3, 4) and not calls
    assert torch.equal(lz + 1, torch.full((2, 8, 3, 4), 2.0)) and len(calls) == 1
    assert float(lz.mean()) == 1.0 and len(calls) == 1                 # materialised once
    assert torch.equal(F.normalize(lz, dim=1), F.normalize(torch.ones(2, 8, 3, 4), dim=1))   # now an ordinary tensor
    with pytest.raises(RuntimeError, match="139-140"):
        torch.sum(F.normalize(mk(), dim=1))                             # not the model's expression
    with pytest.raises(RuntimeError, match="139-140"):
        F.normalize(mk(), dim=1) * torch.ones(2, 8, 3, 5)               # another shape


def test_kernels_refuse_tensors_that_require_grad():
    from tcs_b200 import corr, geo
    t = torch.zeros(1, 8, 2, 16, requires_grad=True)
    for check in (lambda: corr._check_fmap("fmap1", t), lambda: geo._f32c("disp", t)):
        with pytest.raises((RuntimeError, TypeError)):                  # TypeError: CPU tensor is refused first
            check()


@needs_ref
def test_strip_asserts_is_python_O_for_the_reference_modules_only():
    """tcs_b200.strip_asserts re-compiles the reference's functions without their assert statements (each one a
    device -> host sync on the GPU) and restore_asserts puts the original code objects back; results are unchanged."""
    import dis
    import tcs_b200
    ref = ref_model.load()
    ref_model.use_cpu_splat(ref)
    has_assert = lambda f: any(i.opname == "LOAD_ASSERTION_ERROR" for i in dis.get_instructions(f))
    targets = [ref.geo.get_backward_grid, ref.update.BasicMultiUpdateBlock.forward, ref.corr.CorrBlock1D.argmax_disp, ref.geo.disp2depth]
    assert all(has_assert(f) for f in targets)
    model = ref_model.make_model()
    imgs, K, poses, base = ref_model.synthetic_sequence(2, 64, 96)
    want = ref_model.run_sequence(model, imgs, K, poses, base, iters=2)
    try:
        n = tcs_b200.strip_asserts(ref.geo, ref.update, ref.corr, ref.tc_stereo)
        assert n >= 40 and not any(has_assert(f) for f in targets)
        assert has_assert(ref.softsplat.softsplat)                     # a module that was not named keeps its asserts
        got = ref_model.run_sequence(model, imgs, K, poses, base, iters=2)
    finally:
        tcs_b200.restore_asserts()
    assert all(has_assert(f) for f in targets)
    assert all(torch.equal(a["flow"], b["flow"]) and torch.equal(a["flow_q"], b["flow_q"]) for a, b in zip(got, want))


def test_graph_argument_flattening_round_trips():
    from tcs_b200 import graphed
    a, b, c = torch.zeros(2), torch.ones(3), torch.full((1,), 2.0)
    args = ([a, b], [[c, None], [a]], None, True, 3, "x")
    tensors = []
    spec = graphed._flatten((args, {"flag": False, "t": b}), tensors)
    assert [t is u for t, u in zip(tensors, (a, b, c, a, b))] == [True] * 5
    hash(spec)                                                           # the spec is a dict key
    (ra, rk) = graphed._rebuild(spec, tensors)
    assert ra[0][0] is a and ra[1][0][1] is None and ra[3] is True and ra[4] == 3 and rk["t"] is b and rk["flag"] is False
    assert isinstance(ra[0], list) and isinstance(ra, tuple)
    spec2 = graphed._flatten((([a, b], [[c, None], [a]], None, False, 3, "x"), {"flag": False, "t": b}), [])
    assert spec2 != spec                                                 # a flag is part of the signature
    out = graphed._fresh_containers([a, (b, [c])])
    assert out[0] is a and out[1][1][0] is c and out is not None


def test_level_pitch_and_pitched_allocation():
    from tcs_b200 import corr
    assert corr.level_pitch(312) == 320 and corr.level_pitch(240) == 240 and corr.level_pitch(480) == 480
    assert corr.level_pitch(78) == 78                                   # W2 % 8 != 0: a pooled entry would mix real and padding columns
    assert corr.level_pitch(312, num_levels=3) == 312 and corr.level_pitch(312, radius=3) == 312
    flat, levels = corr.alloc_pyramid(2, 3, 312, 312, 4, "cpu", pitch=320, zero=True)
    assert [tuple(l.shape) for l in levels] == [(2, 3, 312, 312 >> i) for i in range(4)]
    assert [l.stride(2) for l in levels] == [320 >> i for i in range(4)]
    assert all(l.data_ptr() % 128 == flat.data_ptr() % 128 for l in levels) and float(flat.abs().sum()) == 0.0
    lv = levels[1]
    assert lv.view(2 * 3 * 312, 1, 1, 156).shape == (1872, 1, 1, 156)   # the reference's corr_pyramid view still works on pitched rows


@needs_ref
def test_completor_stem_weights_are_packed_in_the_kernels_layout():
    """pack_stem_weights against the reference's own DisparityCompletor layers on the CPU: a numpy perceptron over the packed
    buffer (matrices [input][output]) must reproduce update.py:375-378."""
    import numpy as np
    from tcs_b200 import completor as cm
    ref = ref_model.load()
    torch.manual_seed(5)
    comp = ref.update.DisparityCompletor()
    for p in comp.parameters():
        torch.nn.init.normal_(p, std=0.3)
    pk = cm.pack_stem_weights(comp).numpy()
    assert pk.size == 31296
    d, c, m = 0.7, -0.2, 0.5
    x = torch.tensor
    with torch.no_grad():
        one = lambda v: x([[[[v]]]], dtype=torch.float32)
        want = comp.conv_disp_fuse(torch.cat((comp.conv_disp_stem(one(d)), comp.conv_cost_stem(one(c)), comp.conv_mask_stem(one(m))), dim=1)).reshape(-1).numpy()
    o = 0
    feats = []
    for n, v in ((64, d), (32, c), (32, m)):
        w1, b1 = pk[o:o + n], pk[o + n:o + 2 * n]
        W2 = pk[o + 2 * n:o + 2 * n + n * n].reshape(n, n)
        b2 = pk[o + 2 * n + n * n:o + 3 * n + n * n]
        feats.append(np.maximum(w1 * v + b1, 0) @ W2 + b2)
        o += 3 * n + n * n
    cat = np.concatenate(feats)
    W3 = pk[o:o + 128 * 128].reshape(128, 128); b3 = pk[o + 128 * 128:o + 128 * 128 + 128]; o += 128 * 128 + 128
    W4 = pk[o:o + 128 * 64].reshape(128, 64); b4 = pk[o + 128 * 64:o + 128 * 64 + 64]
    got = np.maximum(cat @ W3 + b3, 0) @ W4 + b4
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)
