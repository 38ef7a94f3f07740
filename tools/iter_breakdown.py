"""Where a graphed frame's time goes: slope/intercept over the iteration count, and the four graphs' replay times."""
import os, sys, warnings, statistics
warnings.simplefilter("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tcs_b200
from oracle import ref_model
ref = ref_model.load()
model = ref_model.make_model("cuda")
imgs, K, poses, base = ref_model.synthetic_sequence(1, 544, 960, device="cuda")
tcs_b200.install(ref.tc_stereo, fuse_cost=True, fuse_motion_encoder=ref.update, stencils=ref.update)
tcs_b200.strip_asserts(ref.geo, ref.update, ref.corr, ref.tc_stereo, ref.utils)
h = tcs_b200.graph_modules(model, strip=False)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)
with torch.no_grad():
    res = {it: t(lambda: model(imgs[0][0], imgs[0][1], iters=it, test_mode=True)) for it in (4, 8, 16, 32)}
    print("ms per frame by iterations:", res, "slope ms/iter", (res[32] - res[8]) / 24, "intercept", res[8] - 8 * (res[32] - res[8]) / 24)
    for n, g in h.items():
        for key, c in g.captured.items():
            print(n, "graph replay ms", round(t(lambda: c.graph.replay(), 20), 3), "inputs MB", round(sum(x.numel() * 4 for x in c.inputs) / 1e6, 1))
