"""The UNMODIFIED reference model on the B200 with and without the drop-in: ms per frame of TCStereo.forward.

    python tools/real_model_bench.py [--height 544 --width 960 --iters 32]        (under gpurun; writes gpurun_out/real_model_timing.json)

BASELINE config 1 (single 540x960 pair padded to 544x960, batch 1, 32 GRU iterations, random-init weights): a first frame
(argmax initialisation) and a temporal frame (pose warp of the previous disparity / features / hidden states), for
  reference            baseline/_ref as shipped (its own soft-splat kernel through oracle/cupy_shim.py)
  reference -O         the same with its assert statements (host syncs) stripped: tcs_b200.strip_asserts, i.e. `python -O`
  dropin               tcs_b200.install(core.tc_stereo)
  dropin fused         + fuse_cost, fuse_motion_encoder, stencils
  dropin fused -O      + strip_asserts
  ... + graphed iteration modules   + tcs_b200.graph_modules(model): the four learned blocks of the GRU iteration as CUDA graphs
cuDNN in its default mode (TF32 allowed), as a user of the reference would run it.  Test infrastructure (imports oracle/).
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import ref_model  # noqa: E402


def cpu_reference(ref, args):
    """The reference's own TCStereo.forward on the host (all threads); its cupy splat kernel cannot run there and is replaced by
    oracle/torch_port.splat (index_add_).  One warm-up and two timed calls per kind of frame: a call takes tens of seconds."""
    import time
    model = ref_model.make_model("cpu", mixed_precision=False)
    imgs, K, poses, base = ref_model.synthetic_sequence(2, args.height, args.width, device="cpu")
    undo = ref_model.use_cpu_splat(ref, threaded=True)
    res = {"threads": torch.get_num_threads(), "cores": os.cpu_count()}
    try:
        with torch.no_grad():
            o0 = model(imgs[0][0], imgs[0][1], iters=2, test_mode=True)
            params = {"K": K, "T": poses[1], "previous_T": poses[0], "last_disp": o0["flow_q"], "last_net_list": o0["net_list"],
                      "fmap1": o0["fmap1"], "baseline": base}
            for name, fn in (("first_frame_ms", lambda: model(imgs[0][0], imgs[0][1], iters=args.iters, test_mode=True)),
                             ("temporal_frame_ms", lambda: model(imgs[1][0], imgs[1][1], iters=args.iters, test_mode=True, params=dict(params)))):
                fn()
                ts = []
                for _ in range(2):
                    t0 = time.perf_counter()
                    fn()
                    ts.append(1e3 * (time.perf_counter() - t0))
                res[name] = min(ts)
    finally:
        if callable(undo):
            undo()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=544)
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--iters", type=int, default=32)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--mixed-precision", action="store_true", help="the shipped scripts' --mixed_precision: learned blocks under fp16 autocast")
    ap.add_argument("--cpu-reference", action="store_true", help="also time the reference's TCStereo.forward on the host cores (SURVEY.md 8d(i))")
    args = ap.parse_args()
    import tcs_b200
    ref = ref_model.load()
    model = ref_model.make_model("cuda", mixed_precision=args.mixed_precision)
    imgs, K, poses, base = ref_model.synthetic_sequence(2, args.height, args.width, device="cuda")
    mods = (ref.geo, ref.update, ref.corr, ref.tc_stereo, ref.utils)

    def frames():
        with torch.no_grad():
            o0 = model(imgs[0][0], imgs[0][1], iters=8, test_mode=True)         # a sane state for the temporal frame
            params = {"K": K, "T": poses[1], "previous_T": poses[0], "last_disp": o0["flow_q"], "last_net_list": o0["net_list"],
                      "fmap1": o0["fmap1"], "baseline": base}
            first = lambda: model(imgs[0][0], imgs[0][1], iters=args.iters, test_mode=True)
            temporal = lambda: model(imgs[1][0], imgs[1][1], iters=args.iters, test_mode=True, params=dict(params))
            res = {}
            for name, fn in (("first_frame_ms", first), ("temporal_frame_ms", temporal)):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                ts = []
                for _ in range(args.reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    fn()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                res[name] = statistics.median(ts)
            return res

    out = {"config": {"height": args.height, "width": args.width, "iters": args.iters, "batch": 1, "reps": args.reps, "mixed_precision": args.mixed_precision,
                      "gpu": torch.cuda.get_device_name(0)}}
    fused = dict(fuse_cost=True, fuse_motion_encoder=ref.update, stencils=ref.update)
    for name, kw, strip, graphs in (("reference", None, False, False), ("reference -O", None, True, False), ("dropin", {}, False, False),
                                    ("dropin fused", fused, False, False), ("dropin fused -O", fused, True, False),
                                    ("dropin fused -O + graphed iteration modules", fused, True, True),
                                    ("dropin fused -O + graphed iteration modules + fused completor stems", fused, True, "stems")):
        if strip:
            tcs_b200.strip_asserts(*mods)
        if kw is not None:
            tcs_b200.install(ref.tc_stereo, **kw)
        if graphs:
            tcs_b200.graph_modules(model, strip=False)
        if graphs == "stems":
            tcs_b200.fuse_completor_stems(model.disp_completor)
        try:
            out[name] = frames()
        finally:
            if graphs == "stems":
                tcs_b200.unfuse_completor_stems(model.disp_completor)
            if graphs:
                tcs_b200.ungraph_modules(model, restore=False)
            if kw is not None:
                tcs_b200.uninstall(ref.tc_stereo, ref.update)
            if strip:
                tcs_b200.restore_asserts()
        print(name, out[name], flush=True)
    if args.cpu_reference:
        out["reference on the host cores"] = cpu_reference(ref, args)
        print("reference on the host cores", out["reference on the host cores"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "real_model_timing%s.json" % ("_amp" if args.mixed_precision else "")), "w"), indent=1)


if __name__ == "__main__":
    main()
