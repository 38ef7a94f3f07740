import os, sys, warnings
warnings.simplefilter("ignore")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tcs_b200
from oracle import ref_model
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
ref = ref_model.load()
model = ref_model.make_model("cuda")
imgs, K, poses, base = ref_model.synthetic_sequence(1, 480, 640, device="cuda")
tcs_b200.install(ref.tc_stereo, fuse_cost=True, stencils=ref.update)
rec = {}
orig_up = ref.tc_stereo.TCStereo.upsample_flow
def spy(self, flow, mask, scale=True):
    rec.setdefault(cur[0], []).append((flow.clone(), mask.clone()))
    return orig_up(self, flow, mask, scale)
ref.tc_stereo.TCStereo.upsample_flow = spy
cur = ["eager"]
with torch.no_grad():
    e = model(imgs[0][0], imgs[0][1], iters=8, test_mode=True)
    e = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in e.items()}
    h = tcs_b200.graph_modules(model)
    for tag in ("g1", "g2"):
        cur[0] = tag
        g = model(imgs[0][0], imgs[0][1], iters=8, test_mode=True)
        print(tag, "flow_q max diff", (g["flow_q"] - e["flow_q"]).abs().max().item(), "flow max diff", (g["flow"] - e["flow"]).abs().max().item())
        for i, ((f0, m0), (f1, m1)) in enumerate(zip(rec["eager"], rec[tag])):
            print("  upsample call", i, "flow diff", (f0 - f1).abs().max().item(), "mask diff", (m0 - m1).abs().max().item(), "mask abs max", m0.abs().max().item())
print({n: (x.replays, len(x.captured)) for n, x in h.items()})
