"""Per-GPU sequence loop of the hot path and the sequence -> GPU partitioning.

Template: the reference's evaluation loop (evaluate_stereo.py:167-198) which carries
(flow_q, net_list, fmap1, previous_T) from frame to frame, and the order of operations of
TCStereo.forward (core/tc_stereo.py:114-177): correlation block -> first-frame argmax OR temporal warp +
matching cost -> backward grid + hidden-state warp -> one pyramid lookup per GRU iteration.

The learned blocks between those calls (encoders, disparity completion, GRUs) are out of scope; their
outputs are inputs here (feature maps, the completed disparity, the per-iteration coordinates), so that a
"frame" in this module is exactly the hot path's share of a frame.  B independent sequences are batched
along the batch axis; sequences never cross devices.
"""
import math

import torch

from .corr import CorrBlock1D
from . import geo


def shard_sequences(num_sequences, world_size, rank):
    """Sequence s belongs to rank s mod world_size (whole sequences per GPU, no data-path collective)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank %d / world_size %d" % (rank, world_size))
    return list(range(rank, num_sequences, world_size))


def reduce_metrics(values, device=None):
    """Final metric reduction: SUM-all_reduce of a short fp64 vector (the only collective of the path)."""
    import torch.distributed as dist

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().tolist()


def feature_shape(height, width, downsample=4, divis_by=32):
    """Image size -> 1/4-resolution feature size after padding to a multiple of 32
    (evaluate_stereo.py:179 InputPadder(divis_by=32); n_downsample=2)."""
    ph = (height + divis_by - 1) // divis_by * divis_by
    pw = (width + divis_by - 1) // divis_by * divis_by
    return ph // downsample, pw // downsample


def synthetic_intrinsics(batch, height, width, device, scale=0.25):
    """TartanAir-style pinhole camera scaled to feature resolution (evaluate_stereo.py:138-142,
    tc_stereo.py:121-123: K * diag(s, s, 1) and its inverse)."""
    f = 0.5 * width
    K = torch.tensor([[f, 0.0, 0.5 * width], [0.0, f, 0.5 * height], [0.0, 0.0, 1.0]], dtype=torch.float64)
    Ks = K * torch.tensor([scale, scale, 1.0], dtype=torch.float64).view(3, 1)
    Ks_inv = torch.linalg.inv(Ks)
    rep = lambda m: m.to(torch.float32).unsqueeze(0).repeat(batch, 1, 1).contiguous().to(device)
    return rep(Ks), rep(Ks_inv)


def synthetic_pose(frame, seq_id=0):
    """world2cam 4x4 of a camera advancing 0.05 m per frame along +z with 0.2 degrees of yaw per frame
    and a small per-sequence lateral offset."""
    yaw = math.radians(0.2 * frame)
    c, s = math.cos(yaw), math.sin(yaw)
    cam2world = torch.tensor([[c, 0.0, s, 0.002 * (seq_id % 7) * frame],
                              [0.0, 1.0, 0.0, 0.0],
                              [-s, 0.0, c, 0.05 * frame],
                              [0.0, 0.0, 0.0, 1.0]], dtype=torch.float64)
    return torch.linalg.inv(cam2world).to(torch.float32)


def relative_pose(prev_T, cur_T):
    """previous->current camera transform and its inverse, both [B,4,4] fp32, computed on the host in
    fp64 (ref: geo_utils.py:148-155; tc_stereo.py:127,159)."""
    p = prev_T.double()
    c = cur_T.double()
    fwd = c @ torch.linalg.inv(p)
    return fwd.float().contiguous(), torch.linalg.inv(fwd).float().contiguous()   # inv() returns column-major


def hot_path_frame(fmap1, fmap2, coords_seq, state=None, disp_init=None, rel_T=None, rel_T_inv=None, K=None,
                   K_inv=None, baseline=None, num_levels=4, radius=4, precision=None, mode=None, per_sample_mean=True,
                   carry_in=None, carry_out=None):
    """The hot path's share of one frame for B batched sequences (order of core/tc_stereo.py:114-177).

      fmap1, fmap2  [B,C,H,W] features of the current stereo pair
      coords_seq    [iters,B,1,H,W] x-coordinate queried by each GRU iteration (tc_stereo.py:176-177)
      state         None on a first frame, else (last_disp [B,1,H,W], last_fmap1 [B,C,H,W], last_net_list or None)
      disp_init     [B,1,H,W] completed disparity the backward grid is built from (tc_stereo.py:159);
                    defaults to the warped disparity (get_backward_grid clips it at 0.01 itself)
      rel_T, rel_T_inv  previous->current and current->previous camera transforms [B,4,4]
      carry_out     a geo.WarpCarry that receives fmap1 transposed for the next frame's warp (a by-product of the cost);
      carry_in      the WarpCarry filled by the previous frame: the warp then runs its list formulation on those rows
    Returns a dict: the last lookup, the sparse initialisation (disp, cost, mask), the warped hidden states.
    Every device operation inside is a libtcs_b200 kernel (plus one memset)."""
    corr_fn = CorrBlock1D(fmap1, fmap2, num_levels=num_levels, radius=radius, precision=precision, mode=mode)
    warped_net = None
    if state is None:
        sparse_disp, cost, mask = corr_fn.argmax_disp()
    else:
        last_disp, last_fmap1, last_net_list = state
        sparse_disp, _, mask, cost = geo.warp_with_cost(last_disp, last_fmap1, rel_T, K, K_inv, baseline,
                                                        cur_fmap=fmap1, per_sample_mean=per_sample_mean, want_fmap=False,
                                                        carry_in=carry_in, carry_out=carry_out,
                                                        deterministic=carry_out is not None)   # a carrying run is
        # the list formulation from its first warp on (which has to transpose for itself): bitwise repeatable throughout
        if last_net_list is not None:
            grid = geo.get_backward_grid(disp_init if disp_init is not None else sparse_disp,   # the kernel clips at 0.01
                                         rel_T_inv, K, K_inv, baseline)
            warped_net = geo.warp_hidden_states(last_net_list, grid)
    out = None
    for it in range(coords_seq.shape[0]):
        out = corr_fn(coords_seq[it])
    return {"corr": out, "sparse_disp": sparse_disp, "cost": cost, "mask": mask, "warped_net": warped_net,
            "corr_fn": corr_fn}


def launches_per_frame(iters, first_frame, hidden_levels=3, mode="pyramid", num_levels=4, fused_build=True, warp_lists=False,
                       alt_tc=True):
    """Number of libtcs_b200 kernel launches hot_path_frame issues (memset nodes not counted).  warp_lists: the warp
    runs its list formulation on a carried transposition (a HotPathRunner's frames after the second)."""
    # fused build | prepass x2 + build | alternate: prepass x2 (tensor-core lookups) or prepass x2 + pools (CUDA-core lookups)
    n = (1 if fused_build else 3) if mode == "pyramid" else (2 if alt_tc else 2 + (num_levels - 1))
    if first_frame:
        n += 1 if mode == "pyramid" else 2                         # argmax (+ an on-demand level-0 build)
        if mode != "pyramid":
            n += 2                                                 # its prepasses
    else:
        # scatter: geometry, weights, splat, cost | lists: geometry, weights + count, row sums, offsets, fill, sort, cost
        # + grid; the hidden-state warp: one launch for the model's three levels, else gathers + halves
        n += (7 if warp_lists else 4) + 1 + (1 if hidden_levels == 3 else hidden_levels + (hidden_levels - 1))
    return n + iters


class HotPathRunner:
    """Carries the temporal state of B batched sequences from frame to frame (evaluate_stereo.py:192-197)."""

    def __init__(self, num_levels=4, radius=4, precision=None, mode=None, per_sample_mean=True):
        self.kw = dict(num_levels=num_levels, radius=radius, precision=precision, mode=mode,
                       per_sample_mean=per_sample_mean)   # independent sequences must not share the splat mean
        self.reset()

    def reset(self):
        self.last_disp = None
        self.last_fmap1 = None
        self.last_net_list = None
        self._carry = [geo.WarpCarry(), geo.WarpCarry()]     # fmap1 transposed for the next frame, double-buffered
        self._frame = 0

    def frame(self, fmap1, fmap2, coords_seq, net_list=None, **camera):
        state = None if self.last_disp is None else (self.last_disp, self.last_fmap1, self.last_net_list)
        out = hot_path_frame(fmap1, fmap2, coords_seq, state=state, **camera, **self.kw,
                             carry_in=self._carry[(self._frame + 1) % 2], carry_out=self._carry[self._frame % 2])
        self._frame += 1
        W = fmap1.shape[3]
        xs = torch.arange(W, device=fmap1.device, dtype=torch.float32).view(1, 1, 1, W)
        self.last_disp = (xs - coords_seq[-1]).clamp_min(0)      # disp = coords0 - coords1; flow_q is clipped at 0
        self.last_fmap1 = fmap1
        self.last_net_list = net_list
        return out
