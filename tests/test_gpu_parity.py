"""Parity of the CUDA path (through the C-ABI of libtcs_b200.so) with the oracle and the golden vectors.

Gates: integer-valued / mask outputs bit-exact; fp32 floats |d| <= 1e-6 + 1e-5 |ref|;
plain bf16 build |d corr| <= 2^-8; bf16x3 / fp16x3 build at fp32 level.  Outputs whose conditioning is
set by pixel coordinates (splat, backward grid, hidden-state gather) carry the looser bound stated inline.
"""
import numpy as np
import pytest
import torch

from conftest import assert_close, assert_exact, load_golden
from oracle import tcs_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tcs():
    import tcs_b200
    return tcs_b200


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def make_fmaps(B, C, H, W, seed, shift=0):
    g = torch.Generator().manual_seed(seed)
    f1 = torch.randn(B, C, H, W, generator=g)
    f2 = torch.roll(f1, -shift, dims=3) + 0.3 * torch.randn(B, C, H, W, generator=g) if shift else torch.randn(B, C, H, W, generator=g)
    return f1, f2


def make_coords(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    c = xs - torch.rand(B, 1, H, W, generator=g) * (W / 4)
    flat = c.view(-1)
    n = flat.numel()
    pick = torch.randperm(n, generator=g)
    flat[pick[: n // 100]] = -9.0
    flat[pick[n // 100: n // 50]] = W + 7.5
    flat[pick[n // 50: n // 10]] = flat[pick[n // 50: n // 10]].round()
    return c


# ---------------------------------------------------------------------------------------------------------
# (1) build
# ---------------------------------------------------------------------------------------------------------

def test_prepass_normalise(tcs):
    f1, _ = make_fmaps(2, 256, 5, 37, 7)
    hi, lo, n32 = tcs.normalized_operands(f1.cuda(), "bf16x3", want_n32=True)
    ref = orc.normalize_features(f1.numpy()).transpose(0, 2, 3, 1)
    assert_close(host(n32), ref, rtol=1e-6, atol=1e-7, what="n32")
    split = host(hi.float()) + host(lo.float())
    assert np.abs(split - host(n32)).max() <= 2.0 ** -16 * np.abs(host(n32)).max(), "hi+lo does not carry 16 mantissa bits"
    hi16, lo16, _ = tcs.normalized_operands(f1.cuda(), "fp16x3")
    split16 = (host(hi16.float()) + host(lo16.float())) / 256.0
    assert np.abs(split16 - host(n32)).max() <= 2.0 ** -20, "fp16 hi+lo"


@pytest.mark.parametrize("C", [64, 128, 192, 256, 320, 512])
@pytest.mark.parametrize("W", [1, 31, 33, 100])
def test_prepass_channel_counts_and_ragged_widths(tcs, C, W):
    """Every channel-group split of the pre-pass (1, 2 or 4 groups per thread, idle warps at C = 64) and widths that leave
    partial 32-pixel tiles; the K-block-major operands are the pixel-major ones re-laid out, bit for bit."""
    f1, _ = make_fmaps(2, C, 3, W, 100 + C + W)
    f1[0, :, 1, 0] = 0.0                                           # an all-zero pixel: x / max(0, 1e-12) = 0 (corr.py:58)
    hi, lo, n32 = tcs.normalized_operands(f1.cuda(), "fp16x3", want_n32=True)
    ref = orc.normalize_features(f1.numpy()).transpose(0, 2, 3, 1)
    assert_close(host(n32), ref, rtol=1e-6, atol=1e-7, what="n32 C=%d W=%d" % (C, W))
    assert np.all(host(n32)[0, 1, 0] == 0.0)
    split = (host(hi.float()) + host(lo.float())) / 256.0
    assert np.abs(split - host(n32)).max() <= 2.0 ** -20
    hk, lk, _ = tcs.normalized_operands(f1.cuda(), "fp16x3", kblocked=True)
    for a, k in ((hi, hk), (lo, lk)):
        relaid = a.view(2, 3, W, C // 64, 64).permute(0, 1, 3, 2, 4).contiguous()
        assert torch.equal(relaid, k), "K-block-major operands differ from the pixel-major ones"
    hb, lb, _ = tcs.normalized_operands(f1.cuda(), "bf16x3")
    assert np.abs(host(hb.float()) + host(lb.float()) - host(n32)).max() <= 2.0 ** -16


@pytest.mark.parametrize("case", ["corr_small", "corr_oddwidth", "corr_oddshift"])
@pytest.mark.parametrize("precision,rtol,atol", [("fp32", 1e-5, 1e-6), ("fp16x3", 1e-5, 1e-6), ("bf16x3", 1e-5, 8e-6),
                                                 ("bf16", 0.0, 2.0 ** -8), ("fp16", 0.0, 2.0 ** -10)])
def test_build_golden(tcs, case, precision, rtol, atol):
    g = load_golden(case)
    _, levels = tcs.build_pyramid(cuda(g["fmap1"]), cuda(g["fmap2"]), 4, precision)
    for l in range(4):
        assert_close(host(levels[l]), g["level%d" % l], rtol=rtol, atol=atol, what="%s %s level %d" % (case, precision, l))


SHAPES = [(1, 136, 240), (2, 120, 160), (1, 96, 312), (1, 17, 100), (1, 3, 250),
          (1, 48, 480)]   # BASELINE config 5 rows (1088x1920 -> 272x480; two N tiles, four M tiles): the fp64 oracle takes ~10 s per 48 rows


@pytest.mark.parametrize("B,H,W", SHAPES)
@pytest.mark.parametrize("precision,rtol,atol", [("fp32", 1e-5, 1e-6), ("fp16x3", 1e-5, 1e-6), ("bf16x3", 1e-5, 8e-6), ("bf16", 0.0, 2.0 ** -8)])
def test_build_full_size(tcs, B, H, W, precision, rtol, atol):
    f1, f2 = make_fmaps(B, 256, H, W, 1234 + H, shift=7)
    _, levels = tcs.build_pyramid(f1.cuda(), f2.cuda(), 4, precision)
    ref64 = orc.corr_volume(f1.numpy(), f2.numpy(), np.float64)
    lv = [host(x) for x in levels]
    assert_close(lv[0], ref64, rtol=rtol, atol=atol, what="%s level 0 %dx%d" % (precision, H, W))
    # pooling is exact arithmetic on the kernel's own level below: (a + b) * 0.5 in fp32
    pooled = orc.corr_pyramid(lv[0], 4)
    for l in range(1, 4):
        assert lv[l].shape[-1] == W >> l
        assert_exact(lv[l], pooled[l], what="%s level %d is not the exact pool of level %d" % (precision, l, l - 1))


@pytest.mark.parametrize("B,H,W1,W2", [(1, 136, 240, 240), (2, 120, 160, 160), (1, 9, 200, 72), (1, 5, 256, 240), (3, 2, 20, 40), (1, 150, 128, 100)])
@pytest.mark.parametrize("precision", ["fp16x3", "bf16x3", "fp16", "bf16"])
def test_build_fused_matches_two_step(tcs, B, H, W1, W2, precision):
    """The single fused kernel (normalise + split + UMMA + pyramid) against the pre-pass + build pair: same
    operands up to the last ulp of x/||x||, so the volumes agree far inside either precision's error."""
    g = torch.Generator().manual_seed(W1 * 7 + W2)
    f1 = torch.randn(B, 256, H, W1, generator=g).cuda()
    f2 = torch.randn(B, 256, H, W2, generator=g).cuda()
    _, fused = tcs.build_pyramid(f1, f2, 4, precision, fused=True)
    _, two = tcs.build_pyramid(f1, f2, 4, precision, fused=False)
    tol = {"fp16x3": 2e-6, "bf16x3": 2e-6, "fp16": 1e-4, "bf16": 1e-3}[precision]   # an operand ulp can flip a 16-bit rounding
    for l in range(4):
        assert_close(host(fused[l]), host(two[l]), rtol=0.0, atol=tol, what="%s level %d" % (precision, l))
    pooled = orc.corr_pyramid(host(fused[0]), 4)
    for l in range(1, 4):
        assert_exact(host(fused[l]), pooled[l], what="fused level %d is not the exact pool" % l)
    if precision == "fp16x3":
        assert_close(host(fused[0]), orc.corr_volume(f1.cpu().numpy(), f2.cpu().numpy(), np.float64), rtol=1e-5, atol=1e-6, what="fused vs fp64")


def test_build_fused_rejects_wide_rows(tcs):
    f = torch.randn(1, 64, 2, 312).cuda()
    with pytest.raises(RuntimeError):
        tcs.build_pyramid(f, f, 4, "fp16x3", fused=True)
    tcs.build_pyramid(f, f, 4, "fp16x3")            # default falls back to the two-step path


def test_build_properties_540p(tcs):
    """Size-independent properties at the headline shape: self-correlation has a unit diagonal, symmetric
    volume, values in [-1, 1], level means preserved."""
    f1, _ = make_fmaps(1, 256, 136, 240, 99)
    _, levels = tcs.build_pyramid(f1.cuda(), f1.cuda(), 4, "fp16x3")
    v = levels[0]
    diag = torch.diagonal(v, dim1=2, dim2=3)
    # the tensor core accumulates its fp32 sums with truncation: 48 UMMA steps leave up to ~3e-6 on a sum of 1
    assert (diag - 1).abs().max().item() < 5e-6
    assert (v - v.transpose(2, 3)).abs().max().item() < 2e-6
    assert v.abs().max().item() <= 1 + 2e-6
    for l in range(1, 4):
        assert abs(levels[l].double().mean().item() - v.double().mean().item()) < 1e-7


def test_build_two_n_tiles_and_m_tail(tcs):
    """W2 > 256 takes two UMMA N tiles; W1 not a multiple of 128 exercises the masked M tail."""
    g = torch.Generator().manual_seed(5)
    f1 = torch.randn(1, 128, 4, 200, generator=g)
    f2 = torch.randn(1, 128, 4, 480, generator=g)
    _, levels = tcs.build_pyramid(f1.cuda(), f2.cuda(), 4, "fp16x3")
    ref = orc.corr_volume(f1.numpy(), f2.numpy(), np.float64)
    assert_close(host(levels[0]), ref, rtol=1e-5, atol=1e-6, what="W1=200 W2=480")
    pooled = orc.corr_pyramid(host(levels[0]), 4)
    for l in range(1, 4):
        assert_exact(host(levels[l]), pooled[l], what="level %d" % l)


# ---------------------------------------------------------------------------------------------------------
# (2) lookup, (3) alternate
# ---------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("case", ["corr_small", "corr_oddwidth", "corr_oddshift"])
def test_lookup_golden(tcs, case):
    g = load_golden(case)
    blk = tcs.CorrBlock1D.from_levels([cuda(g["level%d" % l]) for l in range(4)], radius=4)
    out = blk(cuda(g["coords"]))
    assert out.shape == g["lookup"].shape and out.is_contiguous() and out.dtype == torch.float32
    assert_close(host(out), g["lookup"], what=case + " lookup vs reference")


@pytest.mark.parametrize("B,H,W", [(1, 136, 240), (3, 120, 160), (1, 96, 312), (1, 5, 67), (1, 272, 480)])
@pytest.mark.parametrize("radius", [4, 3])
def test_lookup_full_size(tcs, B, H, W, radius):
    f1, f2 = make_fmaps(B, 128, H, W, 11 + W)
    blk = tcs.CorrBlock1D(f1.cuda(), f2.cuda(), num_levels=4, radius=radius, precision="fp32")
    coords = make_coords(B, H, W, 3)
    out = blk(coords.cuda())
    ref = orc.corr_lookup([host(x) for x in blk._levels], coords.numpy(), radius)
    assert_close(host(out), ref, what="lookup %dx%dx%d r=%d" % (B, H, W, radius))


def test_lookup_integer_coords_return_volume_entries(tcs):
    f1, f2 = make_fmaps(1, 128, 8, 64, 21)
    blk = tcs.CorrBlock1D(f1.cuda(), f2.cuda(), precision="fp32")
    xs = torch.arange(64, dtype=torch.float32).view(1, 1, 1, 64).expand(1, 1, 8, 64).contiguous()
    out = host(blk(xs.cuda()))                       # tap k of level 0 == volume[w1, w1 + k]
    vol = host(blk._levels[0])
    for k in range(-4, 5):
        w1 = np.arange(max(0, -k), min(64, 64 - k))
        # not bit-exact: grid_sample's normalise / un-normalise round trip moves integer x by an ulp
        assert_close(out[0, k + 4][:, w1], vol[0][:, w1, w1 + k], rtol=1e-5, atol=2e-6, what="tap %d" % k)


def test_lookup_level0_alignment_variants_agree(tcs):
    """W2 % 16 == 0 picks the predicate-free kernels: 32-byte loads when level 0 is 32-byte aligned, 16-byte loads
    otherwise.  Both must reproduce the oracle (and hence each other) bit for bit on the same volume."""
    B, H, W = 2, 24, 160
    f1, f2 = make_fmaps(B, 128, H, W, 5)
    blk = tcs.CorrBlock1D(f1.cuda(), f2.cuda(), num_levels=4, radius=4, precision="fp32")
    assert blk._levels[0].data_ptr() % 32 == 0
    coords = make_coords(B, H, W, 9).cuda()
    w = torch.randn(64, 36, device="cuda") * 0.2
    out_oct = blk(coords)
    enc_oct = blk.lookup_encoded(coords, w)
    n0 = blk._levels[0].numel()
    shifted = torch.empty(n0 + 12, dtype=torch.float32, device="cuda")
    lvl0 = shifted[4:4 + n0].view_as(blk._levels[0])             # 16-byte aligned, not 32
    assert lvl0.data_ptr() % 32 == 16
    lvl0.copy_(blk._levels[0])
    blk._levels[0] = lvl0
    out_quad = blk(coords)
    enc_quad = blk.lookup_encoded(coords, w)
    assert torch.equal(out_oct, out_quad)
    assert torch.equal(enc_oct, enc_quad)
    ref = orc.corr_lookup([host(x) for x in blk._levels], host(coords), 4)
    assert_close(host(out_oct), ref, what="lookup, both alignments")


def test_lookup_regular_fast_path_is_bit_identical(tcs):
    """The aligned 4-level kernel takes a select-free path when a whole warp's coordinates keep away from integers.
    The 3-level call runs the general shared-memory kernel on the same levels: levels 0..2 must agree bit for bit,
    for coordinates that are regular, integer (never regular) and a fraction of 1e-3 from an integer (borderline)."""
    B, H, W = 2, 16, 240
    f1, f2 = make_fmaps(B, 128, H, W, 31)
    blk = tcs.CorrBlock1D(f1.cuda(), f2.cuda(), num_levels=4, radius=4, precision="fp32")
    g = torch.Generator().manual_seed(4)
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    base = xs - torch.rand(B, 1, H, W, generator=g) * (W / 8)
    variants = {"regular": base,
                "integer": base.round(),
                "borderline": base.round() + (torch.rand(B, 1, H, W, generator=g) - 0.5) * 4e-3,
                "mixed": torch.where(torch.rand(B, 1, H, W, generator=g) < 0.02, base.round(), base),
                "outside": base - 200.0}
    lv = [x for x in blk._levels]
    three = tcs.CorrBlock1D.from_levels([x.clone() for x in lv[:3]], radius=4)
    for name, c in variants.items():
        c = c.cuda()
        a = blk(c)
        b3 = three(c)
        assert torch.equal(a[:, :27], b3), name
        ref = orc.corr_lookup([host(x) for x in lv], host(c), 4)
        assert_close(host(a), ref, what="lookup (%s coordinates)" % name)


def test_lookup_coords_view_and_nan(tcs):
    g = load_golden("corr_small")
    blk = tcs.CorrBlock1D.from_levels([cuda(g["level%d" % l]) for l in range(4)])
    two = cuda(g["coords"])                          # [B,2,H,W]: channel 0 is read in place
    one = two[:, :1].contiguous()
    assert torch.equal(blk(two), blk(one))
    bad = one.clone()
    bad[0, 0, 0, :5] = float("nan")
    bad[0, 0, 1, :5] = float("inf")
    out = blk(bad)                                   # must not fault; other pixels unchanged
    assert torch.equal(out[1], blk(one)[1])


# ---------------------------------------------------------------------------------------------------------
# "next" row: lookup fused with BasicMotionEncoder.convc1 + relu
# ---------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("case", ["corr_small", "corr_oddwidth", "corr_oddshift"])
def test_lookup_encoded_golden(tcs, case):
    """Against the reference's own lookup output pushed through torch's conv2d + relu (what update.py:104 does)."""
    g = load_golden(case)
    gen = torch.Generator().manual_seed(3)
    conv = torch.nn.Conv2d(36, 64, 1)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(64, 36, 1, 1, generator=gen) * 0.3)
        conv.bias.copy_(torch.randn(64, generator=gen) * 0.1)
        ref = torch.relu(conv(torch.from_numpy(g["lookup"]))).numpy()
    blk = tcs.CorrBlock1D.from_levels([cuda(g["level%d" % l]) for l in range(4)])
    out = blk.lookup_encoded(cuda(g["coords"]), conv.weight.cuda(), conv.bias.cuda())
    assert_close(host(out), ref, rtol=1e-5, atol=2e-6, what=case + " fused lookup + convc1 + relu vs reference ops")
    lazy = blk.lazy(cuda(g["coords"]))
    assert torch.equal(lazy.encode(conv.cuda()), out)
    assert not torch.isnan(lazy).any()                       # any other torch use sees the plain lookup
    assert_close(host(lazy.materialize()), g["lookup"], what="lazy lookup materialised")


@pytest.mark.parametrize("wscale", [0.3, 3.0e3, 1.0e-4, 0.0])
@pytest.mark.parametrize("B,H,W", [(1, 136, 240), (2, 5, 248), (1, 3, 48)])
def test_lookup_encoded_tensor_core_against_cuda_core_and_oracle(tcs, monkeypatch, B, H, W, wscale):
    """tcs_corr_lookup_encode_tc (Cout = 64, row pitch % 16 == 0: one tcgen05 GEMM per 128 pixels, fp16 hi/lo split of taps and
    weights) against the CUDA-core kernel and the oracle, for weights of any magnitude (the pack kernel scales by a power of two)
    and a ragged last tile (H * W not a multiple of 128)."""
    f1, f2 = make_fmaps(B, 128, H, W, 21 + W)
    blk = tcs.CorrBlock1D(f1.cuda(), f2.cuda())                  # (the two-step build pitches 248 -> 256)
    assert blk.W2p % 16 == 0
    coords = make_coords(B, H, W, 6).cuda()
    gen = torch.Generator().manual_seed(10)
    w = (torch.randn(64, 36, generator=gen) * wscale).cuda()
    bias = (torch.randn(64, generator=gen) * max(wscale, 1e-3)).cuda()
    monkeypatch.setenv("TCS_B200_ENCODE_TC", "1")             # small calls take the CUDA-core kernel by default
    tc = blk.lookup_encoded(coords, w, bias)
    monkeypatch.setenv("TCS_B200_ENCODE_TC", "0")
    cc = blk.lookup_encoded(coords, w, bias)
    monkeypatch.setenv("TCS_B200_ENCODE_TC", "1")
    assert not torch.equal(tc, cc) or wscale == 0.0, "the two forms round differently: identical bits mean the switch did nothing"
    ref = orc.corr_lookup_encoded([host(x) for x in blk._levels], host(coords), host(w), host(bias), True)
    scale = max(wscale, 1e-30)
    assert_close(host(cc), ref, rtol=1e-5, atol=2e-6 * max(scale / 0.3, 1e-3), what="CUDA-core fused lookup + 1x1")
    assert_close(host(tc), ref, rtol=1e-5, atol=2e-6 * max(scale / 0.3, 1e-3), what="tensor-core fused lookup + 1x1")
    if wscale == 0.0:
        assert torch.equal(tc, torch.relu(bias).view(1, 64, 1, 1).expand_as(tc))
    # no ReLU, no bias
    assert_close(host(blk.lookup_encoded(coords, w, None, relu=False)),
                 orc.corr_lookup_encoded([host(x) for x in blk._levels], host(coords), host(w), None, False),
                 rtol=1e-5, atol=2e-6 * max(scale / 0.3, 1e-3), what="tensor-core, no bias, no relu")
    # the packed weights follow an in-place update of the parameter
    w.mul_(2.0)
    assert_close(host(blk.lookup_encoded(coords, w, None, relu=False)),
                 2.0 * orc.corr_lookup_encoded([host(x) for x in blk._levels], host(coords), host(w) / 2.0, None, False),
                 rtol=1e-5, atol=4e-6 * max(scale / 0.3, 1e-3), what="after an in-place weight update")


@pytest.mark.parametrize("B,H,W,cout,relu", [(1, 136, 240, 64, True), (2, 30, 160, 64, False), (1, 9, 67, 8, True)])
def test_lookup_encoded_full_size(tcs, B, H, W, cout, relu):
    f1, f2 = make_fmaps(B, 128, H, W, 13 + W)
    blk = tcs.CorrBlock1D(f1.cuda(), f2.cuda(), precision="fp32")
    coords = make_coords(B, H, W, 5)
    gen = torch.Generator().manual_seed(9)
    w = torch.randn(cout, 36, generator=gen) * 0.3
    bias = torch.randn(cout, generator=gen) * 0.1 if relu else None
    out = blk.lookup_encoded(coords.cuda(), w.cuda(), bias.cuda() if bias is not None else None, relu=relu)
    ref = orc.corr_lookup_encoded([host(x) for x in blk._levels], coords.numpy(), w.numpy(), None if bias is None else bias.numpy(), relu)
    assert_close(host(out), ref, rtol=1e-5, atol=2e-6, what="fused lookup + 1x1")
    torch.backends.cudnn.allow_tf32 = False          # cuDNN convolutions default to TF32 (1e-3): compare in true fp32
    assert_close(host(out), host(torch.relu(torch.nn.functional.conv2d(blk(coords.cuda()), w.cuda()[:, :, None, None], bias.cuda())) if relu
                                 else torch.nn.functional.conv2d(blk(coords.cuda()), w.cuda()[:, :, None, None])),
                 rtol=1e-4, atol=1e-5, what="fused vs lookup + torch conv2d on the GPU")


@pytest.mark.parametrize("case", ["corr_small", "corr_oddwidth", "corr_oddshift"])
def test_alternate_golden(tcs, case):
    g = load_golden(case)
    blk = tcs.CorrBlock1D(cuda(g["fmap1"]), cuda(g["fmap2"]), mode="alternate")
    out = blk(cuda(g["coords"]))
    assert_close(host(out), g["lookup"], rtol=1e-5, atol=2e-6, what=case + " alternate lookup vs reference")


@pytest.mark.parametrize("B,H,W", [(1, 136, 240), (1, 96, 312)])
def test_alternate_equals_pyramid(tcs, B, H, W):
    f1, f2 = make_fmaps(B, 256, H, W, 5)
    coords = make_coords(B, H, W, 9).cuda()
    a = tcs.CorrBlock1D(f1.cuda(), f2.cuda(), mode="alternate")(coords)
    p = tcs.CorrBlock1D(f1.cuda(), f2.cuda(), mode="pyramid", precision="fp32")(coords)
    assert_close(host(a), host(p), rtol=1e-5, atol=2e-6, what="alternate vs pyramid")


@pytest.mark.parametrize("B,H,W1,W2", [(1, 136, 240, 240), (2, 9, 312, 312), (1, 7, 480, 480), (1, 5, 78, 78), (1, 6, 200, 72), (1, 3, 130, 300)])
@pytest.mark.parametrize("precision", ["fp16x3", "bf16"])
def test_alternate_tensor_core_is_bit_identical_to_the_pyramid_path(tcs, B, H, W1, W2, precision):
    """tcs_corr_lookup_alt_tc builds each tile's band from the operands of tcs_corr_build with the same MMAs and pools /
    samples with the expressions of its epilogue and of tcs_corr_lookup.  Single-pass precisions agree with the pyramid
    path BIT FOR BIT; the hi/lo split ones issue their three passes per 32-channel block where tcs_corr_build issues
    them per 64-channel block, i.e. the same products summed in another order inside the fp32 accumulator: <= 3e-7 on
    values in [-1, 1].  Coordinates spread over a quarter of the width (bands wider than 256 columns: several chunks
    per tile), out-of-range, exact-integer and non-finite ones, odd level widths, W1 != W2 and partial M tiles."""
    g = torch.Generator().manual_seed(W1 + W2)
    f1 = torch.randn(B, 256, H, W1, generator=g).cuda()
    f2 = torch.randn(B, 256, H, W2, generator=g).cuda()
    coords = make_coords(B, H, W1, 11)
    coords.view(-1)[5] = float("nan")
    coords.view(-1)[17] = float("inf")
    coords = coords.cuda()
    alt = tcs.CorrBlock1D(f1, f2, mode="alternate", precision=precision)
    assert alt._alt_tc and alt._levels is None
    _, levels = tcs.build_pyramid(f1, f2, 4, precision, fused=False)
    pyr = tcs.CorrBlock1D.from_levels(levels)
    def same(x, y, what):
        if precision.endswith("x3"):
            assert_close(host(x), host(y), rtol=0.0, atol=3e-7, what=what)
        else:
            assert_exact(host(x), host(y), what=what)

    same(alt(coords), pyr(coords), "tensor-core alternate vs pyramid (%s)" % precision)
    smooth = (torch.arange(W1, dtype=torch.float32).view(1, 1, 1, W1) - 3.25).expand(B, 1, H, W1).contiguous().cuda()
    same(alt(smooth), pyr(smooth), "constant disparity: one chunk per tile")
    edge = (torch.arange(W1, dtype=torch.float32).view(1, 1, 1, W1) - 8.0 + 1e-4).expand(B, 1, H, W1).contiguous().cuda()
    same(alt(edge), pyr(edge), "coords/8 within 2^-9 of an integer: the wider level-3 window")
    far = torch.full((B, 1, H, W1), -1000.0).cuda()
    assert float(alt(far).abs().max()) == 0.0                      # no tile touches the row: zero chunks, zero output


def test_alternate_tensor_core_full_size_1080p(tcs):
    """BASELINE config 5's shape (1088x1920 -> 272x480): against the pyramid path bit for bit on the whole frame, and
    against the oracle's fp64 volume on a band of rows (the oracle takes ~1 s per 10 rows at this width)."""
    B, H, W = 1, 272, 480
    f1, f2 = make_fmaps(B, 256, H, W, 31, shift=11)
    coords = make_coords(B, H, W, 13)
    alt = tcs.CorrBlock1D(f1.cuda(), f2.cuda(), mode="alternate")
    pyr = tcs.CorrBlock1D(f1.cuda(), f2.cuda(), mode="pyramid")           # two-step build at this width
    a = alt(coords.cuda())
    assert_close(host(a), host(pyr(coords.cuda())), rtol=0.0, atol=3e-7, what="1080p alternate vs pyramid (fp16x3: summation order)")
    assert_exact(host(tcs.CorrBlock1D(f1.cuda(), f2.cuda(), mode="alternate", precision="fp16")(coords.cuda())),
                 host(tcs.CorrBlock1D(f1.cuda(), f2.cuda(), mode="pyramid", precision="fp16")(coords.cuda())), what="1080p alternate vs pyramid, fp16")
    rows = slice(100, 124)
    vol = orc.corr_volume(f1[:, :, rows].numpy(), f2[:, :, rows].numpy(), np.float64)
    want = orc.corr_lookup(orc.corr_pyramid(vol.astype(np.float32), 4), coords[:, :, rows].numpy(), 4)
    assert_close(host(a)[:, :, rows], want, rtol=1e-5, atol=2e-6, what="1080p alternate vs oracle")


def test_alternate_cuda_core_fallback_still_matches(tcs):
    """Radius != 4 (or TCS_B200_ALT_TC=0) keeps the dot-per-tap kernel."""
    f1, f2 = make_fmaps(1, 128, 6, 64, 3)
    coords = make_coords(1, 6, 64, 4).cuda()
    alt = tcs.CorrBlock1D(f1.cuda(), f2.cuda(), mode="alternate", radius=3)
    assert not alt._alt_tc
    pyr = tcs.CorrBlock1D(f1.cuda(), f2.cuda(), mode="pyramid", precision="fp32", radius=3)
    assert_close(host(alt(coords)), host(pyr(coords)), rtol=1e-5, atol=2e-6, what="CUDA-core alternate, radius 3")


# ---------------------------------------------------------------------------------------------------------
# first frame: argmax_disp, cost volume
# ---------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("case", ["corr_small", "corr_oddwidth", "corr_oddshift"])
def test_argmax_and_cost_volume_golden(tcs, case):
    g = load_golden(case)
    blk = tcs.CorrBlock1D.from_levels([cuda(g["level%d" % l]) for l in range(4)])
    sparse_disp, main_cost, mask = blk.argmax_disp()
    assert_exact(host(mask), g["mask"], what="argmax mask")
    assert_exact(host(sparse_disp), g["sparse_disp"], what="sparse_disp")
    assert_exact(host(main_cost), g["main_cost"], what="main_cost")
    assert_exact(host(blk.get_cost_volume()), g["cost_volume"], what="cost volume")


@pytest.mark.parametrize("B,H,W,shift", [(1, 136, 240, 9), (2, 60, 160, 3), (1, 20, 312, 0)])
def test_argmax_full_size(tcs, B, H, W, shift):
    f1, f2 = make_fmaps(B, 128, H, W, 77, shift=shift)
    blk = tcs.CorrBlock1D(f1.cuda(), f2.cuda(), precision="fp32")
    outs = blk.argmax_disp()
    ref = orc.argmax_disp(host(blk._levels[0]))
    for name, a, r in zip(("sparse_disp", "main_cost", "mask"), outs, ref):
        assert_exact(host(a), r, what="%s %dx%d" % (name, H, W))
    if shift:
        assert host(outs[2]).mean() > 0.5 and np.median(host(outs[0])[host(outs[2]) > 0]) == shift
    assert_exact(host(blk.get_cost_volume()), orc.masked_cost_volume(host(blk._levels[0])), what="cost volume")


# ---------------------------------------------------------------------------------------------------------
# (4) temporal step
# ---------------------------------------------------------------------------------------------------------

WARP_CASES = ["warp_small", "warp_forward_jump", "warp_backward_jump"]


@pytest.mark.parametrize("case", WARP_CASES)
def test_warp_golden(tcs, case):
    g = load_golden(case)
    args = [cuda(g[k]) for k in ("disp", "fmap", "rel_T", "K", "K_inv", "baseline")]
    d, f, m, c = tcs.warp_with_cost(*args, cur_fmap=cuda(g["cur_fmap"]))
    assert_exact(host(m), g["warped_mask"], what="splat mask vs reference")
    # splat outputs are ratios of sums weighted by (x - floor x): conditioned by the target coordinates' ulp
    assert_close(host(d), g["warped_disp"], rtol=1e-5, atol=1e-5, what="warped disparity vs reference")
    assert_close(host(f), g["warped_fmap"], rtol=1e-5, atol=1e-5, what="warped features vs reference")
    assert_close(host(c), g["cost"], rtol=1e-5, atol=2e-6, what="matching cost vs reference")
    d2, f2, m2 = tcs.warp(*args)
    assert torch.equal(m2, m)
    d3, f3, m3, c3 = tcs.warp_with_cost(*args, cur_fmap=cuda(g["cur_fmap"]), want_fmap=False)   # cost without materialising fmap'
    assert f3 is None and torch.equal(m3, m)
    assert_close(host(c3), g["cost"], rtol=1e-5, atol=2e-6, what="matching cost (no fmap output) vs reference")


def camera(B, H, W, seed):
    rng = np.random.default_rng(seed)
    K = np.zeros((B, 3, 3))
    K[:, 0, 0] = K[:, 1, 1] = 0.5 * W
    K[:, 0, 2], K[:, 1, 2], K[:, 2, 2] = 0.5 * W - 0.5, 0.5 * H - 0.5, 1.0
    T = np.tile(np.eye(4), (B, 1, 1))
    for b in range(B):
        yaw = np.deg2rad(rng.uniform(-0.6, 0.6))
        c, s = np.cos(yaw), np.sin(yaw)
        cam2world = np.array([[c, 0, s, rng.uniform(-0.03, 0.03)], [0, 1, 0, rng.uniform(-0.01, 0.01)],
                              [-s, 0, c, rng.uniform(0.02, 0.15)], [0, 0, 0, 1.0]])
        T[b] = np.linalg.inv(cam2world)
    f32 = lambda a: a.astype(np.float32)
    return f32(K), f32(np.linalg.inv(K)), f32(T), f32(np.linalg.inv(T)), f32(np.full((B, 1), 0.25))


@pytest.mark.parametrize("deterministic", [False, True])
@pytest.mark.parametrize("B,H,W,per_sample", [(1, 120, 160, False), (2, 136, 240, False), (2, 30, 75, True)])
def test_warp_full_size(tcs, B, H, W, per_sample, deterministic):
    K, Kinv, T, Tinv, base = camera(B, H, W, 3)
    g = torch.Generator().manual_seed(17)
    disp = 0.5 + torch.rand(B, 1, H, W, generator=g) * (W / 16)
    disp.view(-1)[::53] = 0.0
    fmap = torch.randn(B, 256, H, W, generator=g)
    cur = torch.randn(B, 256, H, W, generator=g)
    d, f, m, c = tcs.warp_with_cost(disp.cuda(), fmap.cuda(), cuda(T), cuda(K), cuda(Kinv), cuda(base),
                                    cur_fmap=cur.cuda(), per_sample_mean=per_sample, deterministic=deterministic)
    rd, rf, rm = orc.warp(disp.numpy(), fmap.numpy(), T, K, Kinv, base, per_sample_mean=per_sample)
    assert_exact(host(m), rm, what="splat mask")
    assert 0.5 < rm.mean() < 1.0
    # same geometry bit for bit (oracle == kernel), so what is left is the summation order of <= ~10 contributions and
    # expf's last ulp: north_star's fp32 gate, with the absolute part covering cancellation in sums of O(1) features
    assert_close(host(d), rd, rtol=1e-5, atol=2e-6, what="warped disparity")
    assert_close(host(f), rf, rtol=1e-5, atol=2e-6, what="warped features")
    rc = orc.matching_cost(cur.numpy(), host(f), host(m))
    assert_close(host(c), rc, rtol=1e-5, atol=2e-6, what="matching cost")
    if not deterministic:                      # the cost-only kernel (no warped-feature output) gives the same cost
        d2, f2, m2, c2 = tcs.warp_with_cost(disp.cuda(), fmap.cuda(), cuda(T), cuda(K), cuda(Kinv), cuda(base),
                                            cur_fmap=cur.cuda(), per_sample_mean=per_sample, want_fmap=False)
        assert f2 is None and torch.equal(m2, m)
        assert_close(host(d2), rd, rtol=1e-5, atol=2e-6, what="warped disparity (cost-only path)")
        assert_close(host(c2), rc, rtol=1e-5, atol=2e-6, what="matching cost (cost-only path)")


@pytest.mark.parametrize("kind", ["large_flow", "irregular_flow"])
def test_warp_lists_take_any_flow(tcs, kind):
    """Deterministic mode (sorted per-target contributor lists): large flows and flows that jump from pixel to pixel, where
    a target collects from many, far-apart sources, give the same splat as the oracle."""
    B, H, W = 1, 40, 96
    K, Kinv, T, _, base = camera(B, H, W, 5)
    g = torch.Generator().manual_seed(23)
    if kind == "large_flow":
        yaw = np.deg2rad(25.0)                       # ~ 0.5 * W * tan(25 deg) = 22 px of flow and more
        c, s_ = np.cos(yaw), np.sin(yaw)
        T[0] = np.linalg.inv(np.array([[c, 0, s_, 0.0], [0, 1, 0, 0], [-s_, 0, c, 0.3], [0, 0, 0, 1.0]])).astype(np.float32)
        disp = 1.0 + torch.rand(B, 1, H, W, generator=g) * 3
    else:
        T[0] = np.linalg.inv(np.array([[1, 0, 0, 0.0], [0, 1, 0, 0], [0, 0, 1, 2.5], [0, 0, 0, 1.0]])).astype(np.float32)
        disp = 0.5 + torch.rand(B, 1, H, W, generator=g) * 6     # i.i.d. depths + a big forward step: flow jumps pixel to pixel
    fmap = torch.randn(B, 128, H, W, generator=g)
    cur = torch.randn(B, 128, H, W, generator=g)
    d, f, m, c = tcs.warp_with_cost(disp.cuda(), fmap.cuda(), cuda(T), cuda(K), cuda(Kinv), cuda(base), cur_fmap=cur.cuda(),
                                    deterministic=True)
    rd, rf, rm = orc.warp(disp.numpy(), fmap.numpy(), T, K, Kinv, base)
    assert_exact(host(m), rm, what="splat mask (%s)" % kind)
    assert 0.05 < rm.mean() < 1.0
    assert_close(host(d), rd, rtol=1e-5, atol=2e-6, what="warped disparity (%s)" % kind)
    assert_close(host(f), rf, rtol=1e-5, atol=2e-6, what="warped features (%s)" % kind)


def test_warp_is_deterministic(tcs):
    """The list formulation adds a target's contributions in source order: bitwise repeatable (the reference's
    atomic scatter is not)."""
    g = load_golden("warp_small")
    args = [cuda(g[k]) for k in ("disp", "fmap", "rel_T", "K", "K_inv", "baseline")]
    a = tcs.warp_with_cost(*args, cur_fmap=cuda(g["cur_fmap"]), deterministic=True)
    b = tcs.warp_with_cost(*args, cur_fmap=cuda(g["cur_fmap"]), deterministic=True)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert_exact(host(a[2]), g["warped_mask"], what="gather splat mask vs reference")
    assert_close(host(a[0]), g["warped_disp"], rtol=1e-5, atol=1e-5, what="gather warped disparity vs reference")
    assert_close(host(a[1]), g["warped_fmap"], rtol=1e-5, atol=1e-5, what="gather warped features vs reference")
    assert_close(host(a[3]), g["cost"], rtol=1e-5, atol=2e-6, what="gather matching cost vs reference")
    # a larger frame with i.i.d. depths: the lists are filled in whatever order the atomics land, the sort undoes it
    B, H, W = 2, 60, 128
    K, Kinv, T, _, base = camera(B, H, W, 9)
    gen = torch.Generator().manual_seed(5)
    disp = (0.5 + torch.rand(B, 1, H, W, generator=gen) * 8).cuda()
    fmap = torch.randn(B, 128, H, W, generator=gen).cuda()
    cur = torch.randn(B, 128, H, W, generator=gen).cuda()
    runs = [tcs.warp_with_cost(disp, fmap, cuda(T), cuda(K), cuda(Kinv), cuda(base), cur_fmap=cur, deterministic=True)
            for _ in range(3)]
    for r in runs[1:]:
        for x, y in zip(runs[0], r):
            assert torch.equal(x, y)


def test_warp_carry_skips_the_transposition_and_changes_nothing(tcs):
    """Frame t's cost kernel hands fmap1_t on transposed (WarpCarry); frame t+1's warp of that very tensor reads the rows
    instead of transposing again.  Same bits as the list formulation without a carry; a stale or foreign carry is ignored."""
    B, H, W = 2, 40, 96
    K, Kinv, T, _, base = camera(B, H, W, 7)
    g = torch.Generator().manual_seed(3)
    disp = (0.5 + torch.rand(B, 1, H, W, generator=g) * 6).cuda()
    f_prev = torch.randn(B, 128, H, W, generator=g).cuda()
    f_cur = torch.randn(B, 128, H, W, generator=g).cuda()
    f_next = torch.randn(B, 128, H, W, generator=g).cuda()
    cam = (cuda(T), cuda(K), cuda(Kinv), cuda(base))
    c0, c1 = tcs.WarpCarry(), tcs.WarpCarry()
    # frame t: warps f_prev, cost against f_cur, carry_out <- f_cur transposed
    a = tcs.warp_with_cost(disp, f_prev, *cam, cur_fmap=f_cur, want_fmap=False, carry_out=c0)
    assert c0.tensor is f_cur and tuple(c0.rows.shape) == (B * H * W, 128)
    rows = c0.rows.view(B, H, W, 128)
    perm = torch.tensor([lane + 32 * q for lane in range(32) for q in range(4)])       # position 4*lane + q <-> channel lane + 32 q
    assert torch.equal(rows, f_cur.permute(0, 2, 3, 1)[..., perm])
    # frame t+1: warps f_cur (the carried tensor)
    ref = tcs.warp_with_cost(disp, f_cur, *cam, cur_fmap=f_next, want_fmap=False, deterministic=True)
    got = tcs.warp_with_cost(disp, f_cur, *cam, cur_fmap=f_next, want_fmap=False, carry_in=c0, carry_out=c1)
    for x, y in zip(ref, got):
        assert (x is None and y is None) or torch.equal(x, y)
    rd, rf, rm = orc.warp(host(disp), host(f_cur), T, K, Kinv, base)
    assert_exact(host(got[2]), rm, what="splat mask (carry)")
    assert_close(host(got[3]), orc.matching_cost(host(f_next), rf, rm), rtol=1e-5, atol=2e-6, what="matching cost (carry)")
    # a carry keyed on another tensor, or on a tensor modified since, must not be used
    other = f_cur.clone()
    o1 = tcs.warp_with_cost(disp, other, *cam, cur_fmap=f_next, want_fmap=False, carry_in=c0)
    s1 = tcs.warp_with_cost(disp, other, *cam, cur_fmap=f_next, want_fmap=False)
    assert torch.equal(o1[2], s1[2])
    f_cur.mul_(2.0)                                    # bumps the version: the rows are stale now
    assert not c0.matches(f_cur)
    stale = tcs.warp_with_cost(disp, f_cur, *cam, cur_fmap=f_next, want_fmap=False, carry_in=c0, deterministic=True)
    fresh = tcs.warp_with_cost(disp, f_cur, *cam, cur_fmap=f_next, want_fmap=False, deterministic=True)
    for x, y in zip(stale, fresh):
        assert (x is None and y is None) or torch.equal(x, y)
    with pytest.raises(ValueError):
        tcs.warp_with_cost(disp, f_cur, *cam, cur_fmap=f_next, want_fmap=True, carry_out=c1)


def test_warp_identity_pose_keeps_everything(tcs):
    """Zero motion: every pixel lands on itself, the splat is the identity and the mask is all ones."""
    B, H, W = 1, 24, 48
    K, Kinv, _, _, base = camera(B, H, W, 1)
    K[:, 0, 2], K[:, 1, 2] = 24.0, 12.0            # exactly representable intrinsics keep x -> x exact
    Kinv = np.linalg.inv(K.astype(np.float64)).astype(np.float32)
    T = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
    g = torch.Generator().manual_seed(2)
    disp = 1.0 + torch.rand(B, 1, H, W, generator=g) * 4
    fmap = torch.randn(B, 128, H, W, generator=g)
    d, f, m = tcs.warp(disp.cuda(), fmap.cuda(), cuda(T), cuda(K), cuda(Kinv), cuda(base))
    rd, rf, rm = orc.warp(disp.numpy(), fmap.numpy(), T, K, Kinv, base)
    assert_exact(host(m), rm, what="mask")
    assert_close(host(d), rd, rtol=1e-5, atol=2e-6, what="disp")
    assert_close(host(f), rf, rtol=1e-5, atol=2e-6, what="fmap")


@pytest.mark.parametrize("case", WARP_CASES)
def test_backward_grid_and_hidden_states_golden(tcs, case):
    g = load_golden(case)
    grid = tcs.get_backward_grid(cuda(g["disp_init"]), cuda(g["rel_T_inv"]), cuda(g["K"]), cuda(g["K_inv"]), cuda(g["baseline"]))
    assert_exact(host(grid) == -1, g["backward_grid"] == -1, what="behind-the-camera entries vs reference")
    assert_close(host(grid), g["backward_grid"], rtol=1e-5, atol=1e-4, what="backward grid vs reference")
    assert_exact(host(grid), orc.backward_grid(g["disp_init"], g["rel_T_inv"], g["K"], g["K_inv"], g["baseline"]),
                 what="backward grid vs oracle (same arithmetic, bit for bit)")
    for i in range(3):
        gi = cuda(g["grid%d" % i])
        out = tcs.sample_planar(cuda(g["net%d" % i]), gi)
        assert_close(host(out), g["warped_net%d" % i], rtol=1e-5, atol=2e-6, what="hidden state %d vs reference" % i)
        out2 = tcs.bilinear_sampler(cuda(g["net%d" % i]), gi.permute(0, 2, 3, 1))
        assert torch.equal(out, out2)
        if i < 2:
            assert_close(host(tcs.halve_grid(gi)), g["grid%d" % (i + 1)], rtol=1e-5, atol=1e-5, what="halved grid %d" % i)


def test_backward_grid_behind_camera_is_minus_one(tcs):
    B, H, W = 1, 16, 32
    K, Kinv, _, _, base = camera(B, H, W, 1)
    T = np.tile(np.eye(4, dtype=np.float32), (B, 1, 1))
    T[:, 2, 3] = -100.0                              # everything ends up behind the camera
    disp = torch.full((B, 1, H, W), 4.0)
    grid = tcs.get_backward_grid(disp.cuda(), cuda(T), cuda(K), cuda(Kinv), cuda(base))
    assert (grid == -1).all()
    assert_exact(host(grid), orc.backward_grid(disp.numpy(), T, K, Kinv, base))


@pytest.mark.parametrize("B,C,H,W", [(1, 128, 136, 240), (2, 128, 68, 120), (1, 16, 34, 60), (1, 3, 7, 9)])
def test_bilinear_sample_and_halve_full_size(tcs, B, C, H, W):
    g = torch.Generator().manual_seed(4)
    img = torch.tanh(torch.randn(B, C, H, W, generator=g))
    grid = torch.stack([torch.rand(B, H, W, generator=g) * (W + 4) - 2, torch.rand(B, H, W, generator=g) * (H + 4) - 2], 1)
    grid.view(-1)[::41] = -1.0
    grid[:, :, 0, 0] = float("nan")
    out = tcs.sample_planar(img.cuda(), grid.cuda())
    assert_close(host(out), orc.bilinear_sample(img.numpy(), grid.numpy()), rtol=1e-5, atol=1e-6, what="bilinear sample")
    if H >= 4:
        clean = torch.nan_to_num(grid, nan=0.0)
        assert_exact(host(tcs.halve_grid(clean.cuda())), orc.grid_halve(clean.numpy()), what="halved grid")


# ---------------------------------------------------------------------------------------------------------
# error behaviour: loud failures, no fallback
# ---------------------------------------------------------------------------------------------------------

def test_cpu_tensors_are_rejected(tcs):
    f = torch.randn(1, 64, 2, 16)
    with pytest.raises(TypeError):
        tcs.CorrBlock1D(f, f)
    with pytest.raises(TypeError):
        tcs.get_backward_grid(torch.zeros(1, 1, 2, 2), torch.eye(4)[None], torch.eye(3)[None], torch.eye(3)[None], torch.ones(1, 1))


def test_bad_shapes_raise(tcs):
    f = torch.randn(1, 100, 2, 16).cuda()           # C not a multiple of 64
    with pytest.raises(RuntimeError):
        tcs.CorrBlock1D(f, f)
    f = torch.randn(1, 64, 2, 16).cuda()
    with pytest.raises(ValueError):
        tcs.CorrBlock1D(f, f, radius=20)
    blk = tcs.CorrBlock1D(f, f)
    with pytest.raises(ValueError):
        blk(torch.zeros(1, 1, 3, 16).cuda())
    with pytest.raises(RuntimeError):                # warp needs C in {128,...,512}
        tcs.warp(torch.ones(1, 1, 2, 16).cuda(), f, torch.eye(4)[None].cuda(), torch.eye(3)[None].cuda(),
                 torch.eye(3)[None].cuda(), torch.ones(1, 1).cuda())


# ---------------------------------------------------------------------------------------------------------
# "next" row: the per-GRU-iteration 3x3 stencils, one kernel each
# ---------------------------------------------------------------------------------------------------------

def test_stencils_golden_bit_exact(tcs):
    """Against the reference's own outputs (tests/golden/stencils_small.npz): exact, including the boolean mask."""
    g = load_golden("stencils_small")
    disp, grad = cuda(g["disp"]), cuda(g["grad"])
    grads, edge = tcs.disp2disp_gradient_xy(disp)
    assert edge.dtype == torch.bool and grads.shape == g["grads"].shape
    assert_exact(host(grads), g["grads"], what="disp2disp_gradient_xy")
    assert np.array_equal(host(edge), g["edge_mask"])
    assert_exact(host(tcs.disp2disp_grad_candidates(disp, level=1)), g["cands1"], what="grad candidates level 1")
    assert_exact(host(tcs.disp2disp_grad_candidates(disp, level=2)), g["cands2"], what="grad candidates level 2")
    prop, matrix = tcs.propagate_disparity(grad, disp)
    assert_exact(host(prop), g["prop"], what="propagate_disparity")
    assert_exact(host(matrix), g["matrix"], what="propagate_disparity matrix")


@pytest.mark.parametrize("N,H,W", [(8, 136, 240), (1, 1, 7), (3, 5, 1), (2, 96, 312)])
def test_stencils_full_size(tcs, N, H, W):
    gen = torch.Generator().manual_seed(N * 1000 + W)
    disp = torch.rand(N, 1, H, W, generator=gen) * (W / 8 + 1)
    disp.view(-1)[::7] = 0.0
    grad = torch.randn(N, 2, H, W, generator=gen)
    grads, edge = tcs.disp2disp_gradient_xy(disp.cuda())
    rg, re = orc.disp_gradient_xy(disp.numpy())
    assert_exact(host(grads), rg, what="gradient_xy %dx%dx%d" % (N, H, W))
    assert np.array_equal(host(edge), re)
    for level in (1, 2):
        assert_exact(host(tcs.disp2disp_grad_candidates(disp.cuda(), level=level)), orc.disp_grad_candidates(disp.numpy(), level),
                     what="grad candidates level %d" % level)
    for level in (3, 4):       # factors of 3: a product is no longer exact, the reference's own rounding is an ulp away
        assert_close(host(tcs.disp2disp_grad_candidates(disp.cuda(), level=level)), orc.disp_grad_candidates(disp.numpy(), level),
                     rtol=1e-5, atol=1e-5, what="grad candidates level %d" % level)
    prop, matrix = tcs.propagate_disparity(grad.cuda(), disp.cuda())
    rp, rm = orc.disp_propagate(grad.numpy(), disp.numpy())
    assert_exact(host(prop), rp, what="propagate")
    assert_exact(host(matrix), rm, what="propagate matrix")


def test_stencils_reject_bad_arguments(tcs):
    d = torch.zeros(1, 1, 4, 4, device="cuda")
    with pytest.raises(ValueError):
        tcs.disp2disp_grad_candidates(d, level=5)
    with pytest.raises(ValueError):
        tcs.disp2disp_gradient_xy(torch.zeros(1, 2, 4, 4, device="cuda"))
    with pytest.raises(ValueError):
        tcs.propagate_disparity(torch.zeros(1, 3, 4, 4, device="cuda"), d)
    with pytest.raises(TypeError):
        tcs.disp2disp_gradient_xy(torch.zeros(1, 1, 4, 4))


def test_convex_upsample_golden_and_full_size(tcs):
    """TCStereo.upsample_flow (tc_stereo.py:75-88): against the reference's output, then against the oracle at 540p."""
    g = load_golden("stencils_small")
    up = tcs.convex_upsample(cuda(-g["disp"]), cuda(g["up_mask"]), 4, True)
    assert up.shape == g["up"].shape
    assert_close(host(up), g["up"], rtol=1e-5, atol=1e-5, what="upsample_flow vs reference")
    gen = torch.Generator().manual_seed(12)
    for (N, D, H, W, f, scale) in [(2, 1, 136, 240, 4, True), (1, 2, 9, 13, 8, False), (1, 1, 5, 6, 2, True)]:
        flow = torch.randn(N, D, H, W, generator=gen) * 20
        mask = torch.randn(N, 9 * f * f, H, W, generator=gen) * 4
        out = tcs.convex_upsample(flow.cuda(), mask.cuda(), f, scale)
        ref = orc.convex_upsample(flow.numpy(), mask.numpy(), f, scale)
        assert_close(host(out), ref, rtol=1e-5, atol=2e-5, what="upsample %dx%dx%dx%d f=%d" % (N, D, H, W, f))
    with pytest.raises(ValueError):
        tcs.convex_upsample(torch.zeros(1, 1, 4, 4, device="cuda"), torch.zeros(1, 9 * 9, 4, 4, device="cuda"), 3)


# ---------------------------------------------------------------------------------------------------------
# a6: relative pose (geo_utils.py:148-155)
# ---------------------------------------------------------------------------------------------------------

def test_relative_pose_matches_torch(tcs):
    """T2 @ inv(T1) in one launch with no host sync, against the reference's expression evaluated in fp64 (the yardstick)
    and in fp32 on the GPU (what the reference runs): rigid world2cam poses, a general affine matrix that needs pivoting,
    a single [4,4] pair."""
    g = torch.Generator().manual_seed(3)
    B = 5
    T1 = torch.eye(4).repeat(B, 1, 1)
    T2 = torch.eye(4).repeat(B, 1, 1)
    for T in (T1, T2):
        q, _ = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))
        T[:, :3, :3] = q
        T[:, :3, 3] = torch.randn(B, 3, generator=g) * 2
    T1[4] = torch.tensor([[0.0, 2.0, 0.0, 1.0], [1.0, 0.0, 0.5, 0.0], [0.0, 0.0, 3.0, -1.0], [0.0, 0.0, 0.0, 1.0]])   # zero leading pivot
    got = tcs.cal_relative_transformation(T1.cuda(), T2.cuda())
    want64 = (T2.double() @ torch.linalg.inv(T1.double())).numpy()
    assert_close(host(got), want64, rtol=1e-6, atol=1e-6, what="relative pose vs fp64")
    want32 = torch.matmul(T2.cuda(), torch.linalg.inv(T1.cuda()))
    assert_close(host(got), host(want32), rtol=1e-5, atol=2e-6, what="relative pose vs torch fp32 on the GPU")
    one = tcs.cal_relative_transformation(T1[0].cuda(), T2[0].cuda())
    assert one.shape == (4, 4) and torch.equal(one, got[0])
    with pytest.raises(ValueError):
        tcs.cal_relative_transformation(T1.cuda(), T2[:2].cuda())


@pytest.mark.parametrize("B,H,W,Cs", [(2, 136, 240, (128, 128, 128)), (1, 15, 21, (8, 8, 8)), (1, 12, 16, (20, 7, 33)), (3, 5, 4, (16, 16, 16))])
def test_hidden_state_warp_one_launch_is_bit_identical_to_the_chain(tcs, B, H, W, Cs):
    """tcs_warp_hidden_states (tc_stereo.py:159-163 in one launch, grids halved on the fly) against the chain of
    tcs_bilinear_sample / tcs_grid_halve calls it replaces: same operations in the same order, so every bit agrees -
    odd sizes (floor halving), channel counts off the 16-channel groups, out-of-range and -1 grid entries."""
    g = torch.Generator().manual_seed(H * W)
    grid = torch.stack([torch.rand(B, H, W, generator=g) * (W + 6) - 3, torch.rand(B, H, W, generator=g) * (H + 6) - 3], 1)
    grid.view(-1)[::13] = -1.0
    nets = [torch.randn(B, c, H >> l, W >> l, generator=g).cuda() for l, c in enumerate(Cs)]
    got = tcs.warp_hidden_states(nets, grid.cuda())
    gg, want = grid.cuda(), []
    for l, net in enumerate(nets):
        want.append(tcs.sample_planar(net, gg))
        if l < 2:
            gg = tcs.halve_grid(gg)
    for l in range(3):
        assert got[l].shape == want[l].shape
        assert torch.equal(got[l], want[l]), "level %d differs" % l


@pytest.mark.parametrize("B,H,W1,W2", [(1, 20, 312, 312), (2, 5, 312, 312), (1, 6, 400, 312), (1, 4, 104, 296)])
def test_pitched_levels_give_the_same_bits_as_dense_ones(tcs, monkeypatch, B, H, W1, W2):
    """Widths with W2 % 16 == 8 (the KITTI shape's 312) are built with a row pitch of the next multiple of 16, zeros in
    the padding, which puts them on the lookup's predicate-free kernels.  Everything that reads the levels must give the
    bits it gives on dense rows: the levels themselves, the lookup (out-of-range, integer and non-finite coordinates
    included), the fused lookup + 1x1, argmax_disp and the cost volume (also when W1 > W2, where w1 reaches past W2)."""
    g = torch.Generator().manual_seed(W1 + W2)
    f1 = torch.randn(B, 256, H, W1, generator=g).cuda()
    f2 = (torch.roll(f1[..., :W2], -3, dims=3) + 0.3 * torch.randn(B, 256, H, W2, generator=g).cuda()) if W1 >= W2 else torch.randn(B, 256, H, W2, generator=g).cuda()
    coords = make_coords(B, H, W1, 3)
    coords.view(-1)[2] = float("nan")
    coords = coords.cuda()
    w = torch.randn(64, 36, generator=g).cuda() * 0.2
    pitched = tcs.CorrBlock1D(f1, f2)
    assert pitched.W2p == (W2 + 15) // 16 * 16 and pitched.W2p != W2 and pitched._levels[0].shape[-1] == W2
    monkeypatch.setenv("TCS_B200_LEVEL_PITCH", "0")
    dense = tcs.CorrBlock1D(f1, f2)
    assert dense.W2p == W2
    for l in range(4):
        assert torch.equal(pitched._levels[l], dense._levels[l]), "level %d" % l
        lv = pitched._levels[l]
        phys = torch.as_strided(lv, (B, H, W1, pitched.W2p >> l), lv.stride(), lv.storage_offset())
        assert float(phys[..., W2 >> l:].abs().max()) == 0.0, "level %d: the padding columns must be exact zeros" % l
    assert torch.equal(pitched(coords), dense(coords))
    monkeypatch.setenv("TCS_B200_ENCODE_TC", "0")            # pitched rows also qualify for the tensor-core 1x1: same kernel for both here
    assert torch.equal(pitched.lookup_encoded(coords, w), dense.lookup_encoded(coords, w))
    monkeypatch.delenv("TCS_B200_ENCODE_TC")
    for a, b in zip(pitched.argmax_disp(), dense.argmax_disp()):
        assert torch.equal(a, b)
    assert torch.equal(pitched.get_cost_volume(), dense.get_cost_volume())
    monkeypatch.delenv("TCS_B200_LEVEL_PITCH")
    again = tcs.CorrBlock1D.from_levels([lv.contiguous() for lv in dense._levels])     # copies into pitched, zero-padded rows
    assert again.W2p == pitched.W2p and torch.equal(again(coords), dense(coords))


def test_second_device_in_the_same_process(tcs):
    """Kernel attributes (dynamic shared memory above 48 KB, carve-out hints) are per device: the build (fused and two-step),
    the warp (both formulations), the lookups (pyramid, fused encode, tensor-core alternate) must work on cuda:1 after they ran
    on cuda:0 in one process, and give the same bits there."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    g = torch.Generator().manual_seed(8)
    f1 = torch.randn(1, 256, 6, 320, generator=g)
    f2 = torch.randn(1, 256, 6, 320, generator=g)
    s1 = torch.randn(1, 256, 6, 64, generator=g)
    coords = make_coords(1, 6, 320, 5)
    K, Kinv, T, _, base = camera(1, 6, 64, 2)
    disp = 0.5 + torch.rand(1, 1, 6, 64, generator=g) * 4
    w = torch.randn(64, 36, generator=g)
    outs = []
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        d = torch.device(dev)
        blk = tcs.CorrBlock1D(f1.to(d), f2.to(d))                                   # two-step build (W2 > 240)
        small = tcs.CorrBlock1D(f1[..., :240].contiguous().to(d), f2[..., :240].contiguous().to(d))   # fused build
        alt = tcs.CorrBlock1D(f1.to(d), f2.to(d), mode="alternate")
        res = [blk(coords.to(d)), small(coords[..., :240].contiguous().to(d)), alt(coords.to(d)), blk.lookup_encoded(coords.to(d), w.to(d))]
        cam = [torch.from_numpy(x).to(d) for x in (T, K, Kinv, base)]
        for det in (False, True):
            res += [x for x in tcs.warp_with_cost(disp.to(d), s1.to(d), *cam, cur_fmap=s1.flip(3).contiguous().to(d), deterministic=det) if x is not None]
        torch.cuda.synchronize(d)
        outs.append([r.cpu() for r in res])
    for i, (a, b, c) in enumerate(zip(*outs)):
        if i in (4, 5, 7):       # warped disparity / features / cost of the atomic scatter: unordered adds
            assert torch.allclose(a, b, rtol=1e-5, atol=2e-6) and torch.allclose(a, c, rtol=1e-5, atol=2e-6)
        else:
            assert torch.equal(a, b) and torch.equal(a, c), "output %d differs between devices" % i


# ---------------------------------------------------------------------------------------------------------
# training: the cost-volume initialisation loss on level 0
# ---------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_init_loss_kernel_matches_the_oracle_on_the_golden_case(tcs, precision):
    """tcs_b200.init_loss (one kernel each way on the pyramid's level 0) against the oracle on the case whose loss the reference's
    own init_loss produced; the interpolation of `valid` is torch's on this device (see the oracle's docstring)."""
    import torch.nn.functional as F
    g = load_golden("init_loss_small")
    k, thr = int(g["k"]), float(g["threshold"])
    f1, f2 = cuda(g["fmap1"]).requires_grad_(True), cuda(g["fmap2"]).requires_grad_(True)
    flow, valid = cuda(g["flow_gt"]), cuda(g["valid"])
    blk = tcs.DifferentiableCorrBlock1D(f1, f2, precision=precision)
    cv = blk.get_cost_volume()
    assert isinstance(cv, tcs.LazyCostVolume) and cv.size(1) == g["cost_volume"].shape[1]
    loss, metrics = tcs.init_loss(cv, flow, valid, k=k, scale=0.25, threshold=thr)
    vi = F.interpolate(valid, scale_factor=0.25, mode="bilinear", align_corners=True)
    o = orc.init_loss(host(cv.materialize()), g["flow_gt"], g["valid"], k=k, scale=0.25, threshold=thr, valid_interp=host(vi))
    assert abs(metrics["init_gt_loss"] - float(o["gt_loss"])) <= 2e-6
    assert abs(metrics["init_nm_loss"] - float(o["nm_loss"])) <= 2e-6
    assert abs(metrics["init_loss"] - float(o["loss"])) <= 2e-6
    assert abs(metrics["forward_mask_rate"] - float(o["forward_mask_rate"])) <= 1e-6
    assert abs(metrics["init_loss"] - float(g["loss"])) <= 1e-3, "far from the reference's own value"
    # the gradient that reaches the volume: d loss / d level 0 against the oracle's d loss / d cost_volume (transposed, w2 <= w1)
    (dvol,) = torch.autograd.grad(loss, blk._vol, retain_graph=True)
    B, D, H, W = g["cost_volume"].shape
    tri = np.arange(D).reshape(1, D, 1, 1) <= np.arange(W).reshape(1, 1, 1, W)
    assert_close(host(dvol).transpose(0, 3, 1, 2) * tri, o["grad_cost_volume"] * tri, rtol=1e-5, atol=1e-9, what="d loss / d volume")
    assert np.all((host(dvol).transpose(0, 3, 1, 2) * ~tri) == 0), "a gradient outside w2 <= w1"
    loss.backward()
    assert torch.isfinite(f1.grad).all() and float(f1.grad.abs().max()) > 0 and float(f2.grad.abs().max()) > 0


def test_init_loss_refuses_a_plain_tensor(tcs):
    with pytest.raises(TypeError):
        tcs.init_loss(torch.zeros(1, 8, 2, 8, device="cuda"), torch.zeros(1, 1, 8, 32, device="cuda"), torch.ones(1, 1, 8, 32, device="cuda"))


@pytest.mark.parametrize("B,H,W1,W2", [(2, 9, 240, 240), (1, 5, 248, 248), (1, 4, 100, 78)])
def test_odd_levels_are_not_stored_until_asked_for(tcs, monkeypatch, B, H, W1, W2):
    """Every radius-4, 4-level lookup kernel reads levels 0 and 2 only (it re-pools 1 and 3), so the build leaves the odd levels
    unwritten and unallocated; lookups, the fused 1x1, argmax and the cost volume do not need them, and whatever does ask for one
    gets exactly what the build would have stored (dense, pitched and odd widths)."""
    g = torch.Generator().manual_seed(W2)
    f1, f2 = torch.randn(B, 128, H, W1, generator=g).cuda(), torch.randn(B, 128, H, W2, generator=g).cuda()
    coords = make_coords(B, H, W1, 4).cuda()
    w = torch.randn(64, 36, device="cuda") * 0.2
    lazy = tcs.CorrBlock1D(f1, f2)
    assert lazy._levels.pending and lazy._levels.raw(1) is None and lazy._levels.raw(3) is None
    out, enc, amax = lazy(coords), lazy.lookup_encoded(coords, w), lazy.argmax_disp()
    cv = lazy.get_cost_volume()
    assert lazy._levels.pending, "a consumer that does not need the odd levels materialised them"
    monkeypatch.setenv("TCS_B200_LAZY_ODD_LEVELS", "0")
    full = tcs.CorrBlock1D(f1, f2)
    assert not full._levels.pending
    assert torch.equal(out, full(coords)) and torch.equal(enc, full.lookup_encoded(coords, w)) and torch.equal(cv, full.get_cost_volume())
    for a, b in zip(amax, full.argmax_disp()):
        assert torch.equal(a, b)
    pyr = lazy.corr_pyramid                                            # the reference's attribute: asks for every level
    assert not lazy._levels.pending
    for l in range(4):
        assert torch.equal(lazy._levels[l], full._levels[l]), "level %d" % l
        assert lazy._levels[l].stride() == full._levels[l].stride()
        assert pyr[l].shape == (B * H * W1, 1, 1, W2 >> l)
    # the general-radius kernel reads every level: such a block stores them all
    monkeypatch.delenv("TCS_B200_LAZY_ODD_LEVELS")
    r3 = tcs.CorrBlock1D(f1, f2, radius=3)
    assert not r3._levels.pending
