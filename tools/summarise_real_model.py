"""profiles/r02_real_model.md from gpurun_out/real_model.json (written by tests/test_gpu_real_model.py on the B200)."""
import json
import sys

d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/real_model.json"))
f = lambda x: "%.2e" % x
CFG = ["dropin", "dropin_fused", "dropin_fp32", "dropin_fused_fp32"]
out = ["# Round 2: the UNMODIFIED reference model on the B200, with and without `tcs_b200.install()`", "",
       "Source: `tests/test_gpu_real_model.py` (`-m gpu`), numbers from `gpurun_out/real_model.json` of the run that also produced",
       "`profiles/r02_launches.md`.  Reference side = `baseline/_ref` (the reference's own files, unmodified) on `cuda:0`, including its own",
       "soft-splat CUDA kernel string compiled by NVRTC through `oracle/cupy_shim.py`.  Random-init `TCStereo` (seed 1234), images U(0,255),",
       "synthetic poses; convolutions in true fp32 (`torch.backends.cudnn.allow_tf32 = False`: with cuDNN's default TF32 the reference differs from",
       "its own re-run by 0.18 px after 32 iterations and no drift number means anything).", "",
       "Drift = mean |difference| against the reference's output, in px: `flow_q` at 1/4 resolution / `flow` at full resolution.",
       "Floor = the same measure for the REFERENCE ITSELF with N(0, sigma) added to its own correlation volume (sigma = 1e-7: three seeds, what another",
       "fp32 summation order does; sigma = 3e-6: one seed, the stated bound of the fp16x3 tensor-core build); the table gives the largest sample.",
       "Configurations: `dropin` = `install(core.tc_stereo)` (tensor-core fp16x3 build); `dropin_fused` = + `fuse_cost`, `fuse_motion_encoder`,",
       "`stencils`; `*_fp32` = the CUDA-core fp32 build.", ""]


def row(name, v):
    return "| %s | %s | %s |" % (name, f(v["flow_q"]), f(v["flow"]))


for key, title in (("pair_544x960", "BASELINE config 1: one 544x960 pair (540x960 padded as evaluate_stereo.py:179 does), first frame, 32 iterations"),
                   ("frame1_480x640_identical_state_8iters", "480x640, second frame from the reference's own first-frame state, 8 iterations"),
                   ("frame1_480x640_identical_state_32iters", "the same at 32 iterations (the random-init model has diverged: see the reference's own re-run)")):
    if key not in d:
        continue
    v = d[key]
    out += ["## " + title, "", "| | flow_q drift | flow drift |", "|---|---|---|", row("floor (largest of 4 noise samples)", v["floors"]["max"])]
    if "reference_rerun" in v:
        out.append(row("reference vs its own re-run (its atomic splat is unordered)", v["reference_rerun"]))
    for c in CFG:
        if c in v:
            extra = ""
            if "argmax_mask_flips" in v[c]:
                extra = "  (argmax mask: %d flips, density %.3f)" % (v[c]["argmax_mask_flips"], v[c]["argmax_mask_density"])
            out.append(row("`%s`%s" % (c, extra), v[c]))
    if "reference_mean_abs_flow" in v:
        out.append("")
        out.append("Reference mean |flow| of this frame: %s px." % f(v["reference_mean_abs_flow"]))
    out.append("")
for key in ("sequence_3x480x640_8iters", "sequence_3x480x640_32iters"):
    if key not in d:
        continue
    v = d[key]
    out += ["## 3-frame 480x640 temporal sequence, each arm carrying its own state, %s iterations" % key.split("_")[-1].replace("iters", ""), "",
            "flow_q / flow drift per frame (frame 0 = argmax initialisation, frames 1-2 = pose warp + hidden-state warp):", "",
            "| | frame 0 | frame 1 | frame 2 |", "|---|---|---|---|",
            "| reference mean abs(flow), px | " + " | ".join(f(x) for x in v["reference_mean_abs_flow"]) + " |",
            "| floor (largest of 4 noise samples) | " + " | ".join("%s / %s" % (f(t["flow_q"]), f(t["flow"])) for t in v["floors"]["max"]) + " |"]
    for i, smp in enumerate(v["floors"]["samples"]):
        s, seed = v["floors"]["sigma_seed"][i]
        out.append("| floor sample sigma=%g seed=%d | " % (s, seed) + " | ".join("%s / %s" % (f(t["flow_q"]), f(t["flow"])) for t in smp) + " |")
    for c in CFG:
        out.append("| `%s` | " % c + " | ".join("%s / %s" % (f(t["flow_q"]), f(t["flow"])) for t in v[c]) + " |")
    out.append("")
if "graphed_modules_2x480x640_8iters" in d:
    v = d["graphed_modules_2x480x640_8iters"]
    out += ["## `graph_modules`: the GRU iteration's learned blocks replayed as CUDA graphs (2 frames of 480x640, 8 iterations, second pass = replays only)", "",
            "| | frame 0 | frame 1 |", "|---|---|---|",
            "| floor (largest of 4 noise samples) | " + " | ".join("%s / %s" % (f(t["flow_q"]), f(t["flow"])) for t in v["floor"]) + " |",
            "| graphed drop-in vs the reference | " + " | ".join("%s / %s" % (f(t["flow_q"]), f(t["flow"])) for t in v["graphed"]) + " |",
            "| graphed vs the same drop-in run eagerly, max abs d flow | " + " | ".join(f(x) for x in v["graphed_vs_eager_max_abs_flow"]) + " |", ""]
w = {k: v for k, v in d.items() if k.startswith("warp_vs_reference_kernel")}
if w:
    out += ["## Row a8: `tcs_warp_forward` against the reference's own splat kernel (full size, C = 256)", "",
            "(i) torch-CPU geometry (bit-identical to the kernels' and the oracle's) + the reference's CUDA kernel for the scatter; (ii) everything of the",
            "reference on the GPU (cuBLAS-rounded geometry: an ulp of a target coordinate is 1.5e-5 of a bilinear weight).", "",
            "| shape, formulation | (i) masks differing | (i) max abs d disp | (i) max abs d fmap | (ii) mask flips | (ii) disp / fmap entries beyond 1e-4 | the reference (ii) vs the reference (i), disp beyond 1e-4 |",
            "|---|---|---|---|---|---|---|"]
    for k, v in sorted(w.items()):
        a, b = v["cpu_geometry_gpu_reference_splat"], v["all_reference_on_gpu"]
        out.append("| %s | %d | %s | %s | %.1e | %.2e / %.2e | %.2e |" % (k.replace("warp_vs_reference_kernel_", "").replace("_det0", ", atomic scatter").replace("_det1", ", sorted lists"),
                   a["mask_mismatches"], f(a["max_abs_disp"]), f(a["max_abs_fmap"]), b["mask_flip_fraction"], b["disp_outlier_fraction_1e-4"],
                   b["fmap_outlier_fraction_1e-4"], b["reference_gpu_vs_reference_cpu_geometry_disp_outliers"]))
    out.append("")
open("profiles/r02_real_model.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
