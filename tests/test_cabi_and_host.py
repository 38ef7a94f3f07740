"""CPU-side checks: the C-ABI library loads and exports exactly what include/tcs_b200.h declares, the host
logic (sharding, shapes, drop-in rebinding) behaves, and the 2-rank gloo path reduces metrics."""
import ctypes
import os
import re
import subprocess
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tcs_b200.h")


def header_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tcs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import tcs_b200
    lib = ctypes.CDLL(tcs_b200._lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libtcs_b200.so does not export %s" % n
    assert sorted(tcs_b200._lib.SIGNATURES) == names, "ctypes signatures and header disagree"
    assert lib.tcs_abi_version() == tcs_b200._lib.ABI_VERSION == 10


def test_argument_errors_are_reported_without_a_gpu():
    """Validation happens before any CUDA call, so the status codes can be checked on a CPU-only host."""
    import tcs_b200
    lib = tcs_b200._lib.load()
    assert lib.tcs_corr_lookup(None, None, None, None, None, 0, None, 1, 1, 1, 16, 4, 4, 0, None) == -1
    assert b"null" in lib.tcs_last_error()
    assert lib.tcs_corr_prepass(ctypes.c_void_p(256), ctypes.c_void_p(256), None, None, 1, 100, 1, 1, 0, None) == -2
    assert lib.tcs_warp_scratch_bytes(0, 256, 4, 4) == 0
    n = lib.tcs_warp_scratch_bytes(2, 256, 136, 240)
    assert n >= 2 * 136 * 240 * 260 * 4 and n % 256 == 0
    with pytest.raises(tcs_b200._lib.TcsError):
        tcs_b200._lib.call("tcs_grid_halve", None, None, 1, 4, 4, None)
    # tcs_warp_forward: the carried transposition is produced by the cost-only call and consumed by the list formulation
    P = ctypes.c_void_p(4096)                          # never dereferenced: the checks come first
    warp_args = lambda out_fmap, fmap_t, cur_t, flags: (P, P, P, P, P, P, P, P, out_fmap, P, P, fmap_t, cur_t, P, 1, 128, 4, 4, flags, None)
    assert lib.tcs_warp_forward(*warp_args(P, None, P, 0)) == -1 and b"cur_t_out" in lib.tcs_last_error()
    assert lib.tcs_warp_forward(*warp_args(None, P, None, 0)) == -1 and b"fmap_t" in lib.tcs_last_error()
    assert lib.tcs_warp_forward(*warp_args(None, ctypes.c_void_p(4100), None, 2)) == -3        # TCS_E_ALIGN
    assert lib.tcs_disp_grad_candidates(P, P, 1, 4, 4, 5, None) == -2 and b"level" in lib.tcs_last_error()
    assert lib.tcs_convex_upsample(P, P, P, 1, 1, 4, 4, 3, 1, None) == -2 and b"factor" in lib.tcs_last_error()


def test_no_cpu_fallback():
    import tcs_b200
    f = torch.randn(1, 64, 2, 16)
    with pytest.raises(TypeError):
        tcs_b200.CorrBlock1D(f, f)
    with pytest.raises(TypeError):
        tcs_b200.warp(torch.ones(1, 1, 2, 16), torch.randn(1, 128, 2, 16), torch.eye(4)[None], torch.eye(3)[None],
                      torch.eye(3)[None], torch.ones(1, 1))
    src = ""
    pkg = os.path.join(ROOT, "temporally-consistent-stereo-matching_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src += open(os.path.join(pkg, fn)).read()
    assert "import oracle" not in src and "from oracle" not in src and "tcs_oracle" not in src, "the product must not touch the oracle"


def test_missing_library_fails_loudly(tmp_path):
    code = ("import importlib, sys; sys.path.insert(0, %r);"
            "m = importlib.import_module('temporally-consistent-stereo-matching_b200._lib');"
            "m.LIB_PATH = %r; m._lib = None; m.load()") % (ROOT, str(tmp_path / "nope.so"))
    # import of the package itself loads the real library; point the loader elsewhere afterwards
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "not built" in r.stderr


def test_sharding_covers_every_sequence_once():
    import tcs_b200
    for n in (1, 2, 4, 8):
        seen = []
        for r in range(n):
            seen += tcs_b200.shard_sequences(64, n, r)
        assert sorted(seen) == list(range(64))
        assert all(len(tcs_b200.shard_sequences(64, n, r)) == 64 // n for r in range(n))
    with pytest.raises(ValueError):
        tcs_b200.shard_sequences(4, 2, 2)


def test_feature_shapes_of_the_baseline_configs():
    from tcs_b200 import sequence
    assert sequence.feature_shape(540, 960) == (136, 240)
    assert sequence.feature_shape(480, 640) == (120, 160)
    assert sequence.feature_shape(375, 1242) == (96, 312)
    assert sequence.feature_shape(1080, 1920) == (272, 480)


def test_relative_pose_round_trip():
    from tcs_b200 import sequence
    prev = torch.stack([sequence.synthetic_pose(3, s) for s in range(4)])
    cur = torch.stack([sequence.synthetic_pose(4, s) for s in range(4)])
    fwd, inv = sequence.relative_pose(prev, cur)
    eye = torch.eye(4).expand(4, 4, 4)
    assert torch.allclose(fwd @ inv, eye, atol=1e-6)
    assert torch.allclose(fwd @ prev, cur, atol=1e-6)        # world2cam convention (geo_utils.py:148-155)


def test_dropin_rebinds_and_restores_names():
    import tcs_b200
    fake = types.ModuleType("core.tc_stereo")
    sentinel = object()
    for n in ("CorrBlock1D", "warp", "get_backward_grid", "bilinear_sampler", "cal_relative_transformation"):
        setattr(fake, n, sentinel)
    new = tcs_b200.install(fake, precision="bf16")
    assert fake.warp is tcs_b200.warp and fake.get_backward_grid is tcs_b200.get_backward_grid
    assert issubclass(fake.CorrBlock1D, tcs_b200.CorrBlock1D) and fake.CorrBlock1D.__name__ == "CorrBlock1D"
    assert set(new) == {"CorrBlock1D", "warp", "get_backward_grid", "bilinear_sampler", "cal_relative_transformation"}
    tcs_b200.uninstall(fake)
    assert fake.warp is sentinel and fake.CorrBlock1D is sentinel
    with pytest.raises(AttributeError):
        tcs_b200.install(types.ModuleType("empty"))


def test_dropin_patches_motion_encoder():
    import tcs_b200
    tcs_mod = types.ModuleType("core.tc_stereo")
    for n in ("CorrBlock1D", "warp", "get_backward_grid", "bilinear_sampler", "cal_relative_transformation"):
        setattr(tcs_mod, n, object())
    upd = types.ModuleType("core.update")

    class BasicMotionEncoder(torch.nn.Module):
        def forward(self, flow, corr):
            return "original"
    upd.BasicMotionEncoder = BasicMotionEncoder
    original = BasicMotionEncoder.forward
    tcs_b200.install(tcs_mod, fuse_motion_encoder=upd)
    assert BasicMotionEncoder.forward is not original
    assert tcs_mod.CorrBlock1D.__call__ is not tcs_b200.CorrBlock1D.__call__     # returns deferred lookups
    tcs_b200.uninstall(tcs_mod, upd)
    assert BasicMotionEncoder.forward is original


def test_dropin_patches_the_iteration_stencils():
    import tcs_b200
    tcs_mod = types.ModuleType("core.tc_stereo")
    for n in ("CorrBlock1D", "warp", "get_backward_grid", "bilinear_sampler", "cal_relative_transformation"):
        setattr(tcs_mod, n, object())
    orig_grad = tcs_mod.disp2disp_gradient_xy = lambda disp: "gradient"
    upd = types.ModuleType("core.update")
    orig_cands = upd.disp2disp_grad_candidates = lambda disp, level=1: "candidates"

    class DispRefine(torch.nn.Module):
        def propagate_disparity(self, disparity_grad, disparity_map):
            return "propagate"
    upd.DispRefine = DispRefine
    orig_prop = DispRefine.propagate_disparity

    class TCStereo(torch.nn.Module):
        def upsample_flow(self, flow, mask, scale=True):
            return "upsample"
    tcs_mod.TCStereo = TCStereo
    orig_up = TCStereo.upsample_flow
    tcs_b200.install(tcs_mod, stencils=upd)
    assert TCStereo.upsample_flow is not orig_up
    assert tcs_mod.disp2disp_gradient_xy is tcs_b200.disp2disp_gradient_xy
    assert upd.disp2disp_grad_candidates is tcs_b200.disp2disp_grad_candidates
    assert DispRefine.propagate_disparity is not orig_prop
    with pytest.raises(TypeError):                       # no CPU path: a CPU tensor is refused, not silently computed
        DispRefine().propagate_disparity(torch.zeros(1, 2, 4, 4), torch.zeros(1, 1, 4, 4))
    tcs_b200.uninstall(tcs_mod, upd)
    assert tcs_mod.disp2disp_gradient_xy is orig_grad and upd.disp2disp_grad_candidates is orig_cands
    assert DispRefine.propagate_disparity is orig_prop and TCStereo.upsample_flow is orig_up


def test_warp_carry_is_keyed_on_the_storage_and_its_version():
    """The carried transposition may only be used for the very memory it was made from, unmodified (host logic only);
    the model hands the tensor back as `fmap1.detach()` (tc_stereo.py:227,242): a new object, same storage and version."""
    import tcs_b200
    c = tcs_b200.WarpCarry()
    f = torch.zeros(1, 8, 2, 4)
    assert not c.matches(f)                                  # empty
    c.reserve(f)
    assert tuple(c.rows.shape) == (8, 8) and not c.matches(f)   # reserved, nothing recorded yet
    c.record(f)
    assert c.matches(f) and c.matches(f.detach()) and c.matches(f.view(1, 8, 2, 4))
    assert not c.matches(f[:, :4]) and not c.matches(f.view(1, 4, 4, 4))   # same pointer, another shape
    assert not c.matches(f.clone())                          # another tensor with the same contents
    f.add_(1.0)                                              # modified in place: the rows are stale
    assert not c.matches(f)
    c.version = f._version
    assert c.matches(f) and not c.matches(f.double()) and not c.matches(f.permute(0, 1, 3, 2))
    c.reserve(torch.zeros(2, 8, 2, 4))                       # another shape: reallocated and invalidated
    assert tuple(c.rows.shape) == (16, 8) and not c.matches(f)


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
import tcs_b200
from tcs_b200 import sequence
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"], rank=int(os.environ["RANK"]), world_size=2)
mine = tcs_b200.shard_sequences(10, 2, dist.get_rank())
frames = float(len(mine) * 5)
tot = sequence.reduce_metrics([frames, float(sum(mine))])
assert tot == [50.0, 45.0], tot
dist.barrier()
dist.destroy_process_group()
print("ok", dist.get_rank() if False else os.environ["RANK"])
"""


def test_two_rank_gloo_metric_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER % ROOT)
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0 and "ok" in out, out


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints ONE JSON line with the contract's keys (CPU port of the path)."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--height", "128", "--width", "192", "--iters", "2", "--seqs-per-gpu", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["product_so_loaded"] == [], "the reference arm must not load the product library"
    assert d["config"]["seqs_per_gpu"] == 2                      # same batch per step as the B200 arm's --seqs-per-gpu
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_launch_count_model():
    from tcs_b200 import sequence
    assert sequence.launches_per_frame(32, False) == 1 + 4 + 1 + 1 + 32   # fused build; geometry, weights, splat, finalize; grid; hidden-state warp; lookups
    assert sequence.launches_per_frame(32, True) == 1 + 1 + 32
    assert sequence.launches_per_frame(32, False, fused_build=False) == sequence.launches_per_frame(32, False) + 2
    # list formulation on a carried transposition: geometry, weights + count, row sums, offsets, fill, sort, cost
    assert sequence.launches_per_frame(32, False, warp_lists=True) == 1 + 7 + 1 + 1 + 32
    assert sequence.launches_per_frame(32, False, hidden_levels=2) == 1 + 4 + 1 + 2 + 1 + 32      # other depths chain gathers and halvings


def test_bench_reference_gpu_arm_degrades_without_a_gpu():
    """`bench.py --impl reference-gpu` needs a CUDA device; without one it says so in one JSON line and exits 0."""
    import json
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference-gpu", "--steps", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert d["impl"] == "reference-gpu" and "unavailable" in d


def test_unstored_odd_levels_materialise_as_the_pooled_level_below():
    """corr._Levels (host logic, no GPU): levels 1 and 3 that the build did not store are allocated and pooled from the level
    below on first touch - (even + odd) * 0.5 on the physical rows, i.e. F.avg_pool2d's values (ref: corr.py:21-23), pitch and zero
    padding included; even levels never trigger it."""
    import torch
    import torch.nn.functional as F
    from tcs_b200 import corr
    g = torch.Generator().manual_seed(0)
    for W2, P in ((40, 40), (24, 32), (77, 77)):                      # dense, pitched (zeros past W2), odd width
        B, H, W1 = 2, 3, 5
        phys0 = torch.zeros(B, H, W1, P)
        phys0[..., :W2] = torch.randn(B, H, W1, W2, generator=g)
        lv0 = phys0[..., :W2]
        phys2 = torch.zeros(B, H, W1, P >> 2)
        lv1_ref = F.avg_pool2d(lv0.reshape(-1, 1, 1, W2), [1, 2], stride=[1, 2]).reshape(B, H, W1, W2 >> 1)
        lv2_ref = F.avg_pool2d(lv1_ref.reshape(-1, 1, 1, W2 >> 1), [1, 2], stride=[1, 2]).reshape(B, H, W1, W2 >> 2)
        phys2[..., :W2 >> 2] = lv2_ref
        levels = corr._Levels([lv0, None, phys2[..., :W2 >> 2], None])
        levels.pending = True
        assert levels[0] is lv0 and levels[2].shape[-1] == W2 >> 2 and levels[-2] is levels.raw(2)
        assert levels.pending and levels.raw(1) is None, "an even level materialised the odd ones"
        lv1 = levels[1]
        assert not levels.pending and levels.raw(3) is not None
        assert torch.equal(lv1, lv1_ref) and lv1.stride(2) == P >> 1
        lv3_ref = F.avg_pool2d(lv2_ref.reshape(-1, 1, 1, W2 >> 2), [1, 2], stride=[1, 2]).reshape(B, H, W1, W2 >> 3)
        assert torch.equal(levels[3], lv3_ref) and levels[3].stride(2) == P >> 3
        full1 = torch.as_strided(lv1, (B, H, W1, P >> 1), lv1.stride(), lv1.storage_offset())
        assert float(full1[..., W2 >> 1:].abs().sum()) == 0.0, "padding columns must stay zero"
        assert [tuple(x.shape) for x in levels] == [(B, H, W1, W2 >> l) for l in range(4)]
