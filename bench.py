#!/usr/bin/env python
"""Throughput of the TC-Stereo cost-volume hot path on B200: stereo frames/s at 540x960, 32 GRU iterations.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun by the driver)
    python bench.py --impl reference ...                      (the reference's own code for the path on the host cores)
    python [-O] bench.py --impl reference-gpu ...             (the reference's own PyTorch code for the path on the B200)

A step = the hot path's share of ONE temporal frame for every sequence this GPU owns (B sequences batched
along the batch axis): the fused tcgen05 correlation build of all 4 levels (normalisation and hi/lo split on chip), forward
warp of the previous disparity/features + matching cost, backward grid + the 3-level hidden-state warp, and 32 pyramid
lookups (--mode alternate: 32 tensor-core on-the-fly lookups, no pyramid).  Inputs (feature maps, per-iteration coordinates, hidden states, poses) are synthetic and already
resident in HBM for `value`; `e2e` feeds the same step from pinned HOST buffers through the public Python API
(H2D of the feature maps and D2H of the results inside the timed region).  Sequences are independent, so
N GPUs run N*B sequences with no data-path collective (weak scaling; --sequences S shards S sequences in total: strong
scaling, BASELINE config 4); NCCL is used for the barrier, the max-over-ranks of the timings and the final all_reduce of the
frame count.  One JSON line is printed by rank 0.  The reference arms (--impl reference: the reference's own code on the host
cores; --impl reference-gpu: the same code on the B200) never load libtcs_b200.so.
"""
import argparse
import json
import math
import os
import subprocess
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "stereo frames/s at 540x960 (32 iters), cost-volume hot path"
C = 256
LEVELS = 4
RADIUS = 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--seqs-per-gpu", type=int, default=8)
    ap.add_argument("--sequences", type=int, default=0,
                    help="BASELINE config 4: this many sequences in TOTAL, sharded over the ranks with shard_sequences() (strong "
                         "scaling; e.g. --sequences 64 --height 480 --width 640); 0 = --seqs-per-gpu on every rank (weak scaling)")
    ap.add_argument("--height", type=int, default=540)
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--iters", type=int, default=32)
    ap.add_argument("--precision", default="fp16x3")
    ap.add_argument("--mode", default="pyramid", choices=["pyramid", "alternate"])
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-warp-carry", dest="warp_carry", action="store_false",
                    help="do not hand each frame's transposed features to the next frame's warp (atomic scatter instead)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-gpu-reference", action="store_true", help="do not time the reference's own PyTorch path on the GPU")
    ap.add_argument("--want-fmap", action="store_true", help="plain drop-in warp: materialise the 256-channel warped features")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    return ap.parse_args()


def feature_hw(h, w):
    return ((h + 31) // 32 * 32) // 4, ((w + 31) // 32 * 32) // 4


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------------------

def make_slot(B, H, W, iters, seed, device, pin=False):
    """One frame's inputs for B sequences: what the learned blocks around the hot path would deliver."""
    g = torch.Generator().manual_seed(seed)
    f1 = torch.randn(B, C, H, W, generator=g)
    f2 = torch.roll(f1, -5, dims=3) + 0.3 * torch.randn(B, C, H, W, generator=g)
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    disp = 0.5 + torch.rand(B, 1, H, W, generator=g) * (W / 16.0)
    walk = torch.cumsum(0.05 * torch.randn(iters, B, 1, H, W, generator=g), 0)
    coords = xs - (disp[None] + walk)                                   # coords1 = coords0 - disp
    last_disp = (xs - coords[-1]).clamp_min(0)                          # what the next frame warps
    nets = [torch.tanh(torch.randn(B, 128, H >> i, W >> i, generator=g)) for i in range(3)]
    host = {"f1": f1, "f2": f2}
    if pin:
        host = {k: v.pin_memory() for k, v in host.items()}
    dev = {"f1": f1.to(device), "f2": f2.to(device), "coords": coords.to(device), "last_disp": last_disp.to(device),
           "nets": [n.to(device) for n in nets]}
    return host, dev


def synthetic_intrinsics(batch, height, width, device, scale=0.25):
    """TartanAir-style pinhole camera scaled to feature resolution (evaluate_stereo.py:138-142, tc_stereo.py:121-123).
    The workload definition lives here, not in the product package: the reference arms must not load libtcs_b200.so."""
    f = 0.5 * width
    K = torch.tensor([[f, 0.0, 0.5 * width], [0.0, f, 0.5 * height], [0.0, 0.0, 1.0]], dtype=torch.float64)
    Ks = K * torch.tensor([scale, scale, 1.0], dtype=torch.float64).view(3, 1)
    rep = lambda m: m.to(torch.float32).unsqueeze(0).repeat(batch, 1, 1).contiguous().to(device)
    return rep(Ks), rep(torch.linalg.inv(Ks))


def synthetic_pose(frame, seq_id=0):
    """world2cam 4x4: 0.05 m per frame along +z, 0.2 degrees of yaw per frame, a per-sequence lateral offset."""
    yaw = math.radians(0.2 * frame)
    c, s = math.cos(yaw), math.sin(yaw)
    cam2world = torch.tensor([[c, 0.0, s, 0.002 * (seq_id % 7) * frame], [0.0, 1.0, 0.0, 0.0], [-s, 0.0, c, 0.05 * frame],
                              [0.0, 0.0, 0.0, 1.0]], dtype=torch.float64)
    return torch.linalg.inv(cam2world).to(torch.float32)


def camera(B, H, W, device, frame):
    K, K_inv = synthetic_intrinsics(B, 4 * H, 4 * W, device)
    prev = torch.stack([synthetic_pose(frame - 1, s) for s in range(B)]).double()
    cur = torch.stack([synthetic_pose(frame, s) for s in range(B)]).double()
    fwd = cur @ torch.linalg.inv(prev)                                   # geo_utils.py:148-155
    inv = torch.linalg.inv(fwd)
    return {"K": K, "K_inv": K_inv, "rel_T": fwd.float().contiguous().to(device), "rel_T_inv": inv.float().contiguous().to(device),
            "baseline": torch.full((B, 1), 0.25, device=device)}


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), sampled through NVML every
    2 ms from a thread (the timed region is tens of milliseconds: nvidia-smi -lms is too coarse for it)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.sm, self.bits, self.power, self.mx, self.err = index, [], 0, [], None, None
        self._stop = threading.Event()
        self.t = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.index < len(ids) and ids[self.index].isdigit():
                    phys = int(ids[self.index])
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = "nvml unavailable: %s" % (e,)
            return

        def loop():
            n = 0
            while not self._stop.is_set():
                try:
                    self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    self.bits |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                    if n % 8 == 0:                         # the power query is the slow one: keep the clock samples dense
                        self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
                except Exception as e:  # noqa: BLE001
                    self.err = str(e)
                    return
                n += 1
                time.sleep(0.0005)

        self.t = threading.Thread(target=loop, daemon=True)
        self.t.start()

    def stop(self):
        self._stop.set()
        if self.t is not None:
            self.t.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": [self.err or "no samples"]}
        reasons = sorted(name for bit, name in self.REASONS.items() if self.bits & bit)
        return {"sm_mhz": statistics.median(self.sm), "sm_min_mhz": min(self.sm), "sm_max_mhz": self.mx, "reasons": reasons,
                "samples": len(self.sm), "power_w_max": max(self.power) if self.power else None}


# ------------------------------------------------------------------------------------------------------------
# reference arms: the reference's own code for the path (baseline/_ref through oracle/ref_model.py), or, when the
# reference did not travel, the torch-CPU port of its call sequence (oracle/torch_port.py).  None of the product.
# ------------------------------------------------------------------------------------------------------------

def workload_text(args, B):
    return "%dx%d temporal frames, %d lookup iters, 4-level pyramid r=4, C=256, %d sequences per GPU batched" % (
        args.height, args.width, args.iters, B)


def cpu_frames_per_second(B, H, W, iters, budget_s, steps=None, warmup=1):
    """One step = the hot path's share of one temporal frame for B batched sequences, on the host cores."""
    from oracle import ref_model
    try:   # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, RuntimeError):
        pass
    _, dev = make_slot(B, H, W, iters, 7, "cpu")
    _, prev = make_slot(B, H, W, 1, 24, "cpu")
    cam = camera(B, H, W, "cpu", 1)
    state = (prev["last_disp"], prev["f1"], prev["nets"])
    kind = "port"
    if ref_model.reference_root() is not None:
        ref = ref_model.load()
        ref_model.use_cpu_splat(ref, threaded=True)
        kind = "reference"

        def one():
            with torch.no_grad():
                ref_model.hot_path_frame(ref, dev["f1"], dev["f2"], dev["coords"], state, cam["rel_T"], cam["rel_T_inv"],
                                         cam["K"], cam["K_inv"], cam["baseline"])
    else:
        from oracle import torch_port as tp

        def one():
            with torch.no_grad():
                tp.frame(dev["f1"], dev["f2"], dev["coords"], state=state, rel_T=cam["rel_T"], rel_T_inv=cam["rel_T_inv"],
                         K=cam["K"], K_inv=cam["K_inv"], baseline=cam["baseline"])

    for _ in range(warmup):
        one()
    n, t0 = 0, time.perf_counter()
    while True:
        one()
        n += 1
        el = time.perf_counter() - t0
        if (steps is not None and n >= steps) or (steps is None and (el >= budget_s or n >= 50)):
            break
    return B * n / el, n, el, kind


CPU_NOTE = {"reference": "the reference's own core/corr.py, geo_utils.py, utils.py calls (baseline/_ref) in TCStereo.forward's order; "
                         "its cupy splat kernel cannot run on a CPU and is replaced by index_add_ (oracle/torch_port.py)",
            "port": "oracle/torch_port.py (torch-CPU restatement of the reference's call sequence; baseline/_ref absent)"}


def product_so_loaded():
    """Names of this repository's native libraries mapped into the process (the reference arms must report none)."""
    try:
        with open("/proc/self/maps") as f:
            return sorted({os.path.basename(l.split()[-1]) for l in f if "libtcs_b200" in l})
    except OSError:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.seqs_per_gpu
    H, W = feature_hw(args.height, args.width)
    fps, n, el, kind = cpu_frames_per_second(B, H, W, args.iters, 0, steps=args.steps, warmup=max(1, min(args.warmup, 2)))
    cores = torch.get_num_threads()
    sample = "%d steps of %d batched sequences (%dx%d, %d lookups each) on %d host threads; %s" % (
        n, B, args.height, args.width, args.iters, cores, CPU_NOTE[kind])
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / n, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args, B), "feature_hw": [H, W], "seqs_per_gpu": B},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "product_so_loaded": product_so_loaded(),
    }))


def run_reference_gpu(args):
    """the reference's own PyTorch code for the path on the B200 (cuBLAS SGEMM behind einsum,
    4 avg_pool2d, 4 grid_sample per lookup, ~40 elementwise launches + its own NVRTC-compiled splat kernel in warp) on the
    same batched inputs as the B200 arm, eager (it cannot be graph-captured: host syncs), CUDA-event timed per phase.
    Run under `python -O` to strip its assert-induced host syncs."""
    from oracle import ref_model
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if ref_model.reference_root() is None or not torch.cuda.is_available():
        print(json.dumps({"impl": "reference-gpu", "unavailable": "baseline/_ref not installed or no CUDA device"}))
        return
    ref = ref_model.load()
    device = torch.device("cuda", 0)
    B, iters = args.seqs_per_gpu, args.iters
    H, W = feature_hw(args.height, args.width)
    slots = [make_slot(B, H, W, iters, 1234 + 17 * s, device)[1] for s in range(2)]
    cam = camera(B, H, W, device, 1)
    marks = []

    def step(k, timed):
        s, o = slots[k % 2], slots[1 - k % 2]
        ev = {}

        def tick(label):
            if timed:
                ev[label] = torch.cuda.Event(enable_timing=True)
                ev[label].record()
        with torch.no_grad():
            out = ref_model.hot_path_frame(ref, s["f1"], s["f2"], s["coords"], (o["last_disp"], o["f1"], o["nets"]), cam["rel_T"],
                                           cam["rel_T_inv"], cam["K"], cam["K_inv"], cam["baseline"], timers=tick)
        if timed:
            marks.append(ev)
        return out

    for k in range(max(args.warmup, 4)):
        step(k, False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(args.steps):
        out = step(k, True)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    # medians: the first timed steps of an eager torch program still pay allocator growth and library start-up
    ph = {n: statistics.median(m[a].elapsed_time(m[b]) for m in marks)
          for n, a, b in (("build_ms", "start", "build"), ("warp_ms", "build", "warp"), ("lookups_ms", "warp", "lookups"))}
    total_ms = statistics.median(m["start"].elapsed_time(m["lookups"]) for m in marks) * len(marks)
    fps = B * args.steps / (total_ms * 1e-3)              # from the median step: an eager torch program's wall clock wanders by +-15 %
    print(json.dumps({
        "impl": "reference-gpu", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 4), "ms_per_step": total_ms / args.steps, "wall_ms_per_step_mean": 1e3 * wall / args.steps,
        "higher_is_better": True, "dtype": "f32", "data": "synthetic", "asserts": bool(__debug__),
        "config": {"workload": workload_text(args, B), "feature_hw": [H, W], "seqs_per_gpu": B,
                   "what": "reference core/corr.py + geo_utils.py + utils.py + softsplat.py (own CUDA kernel via NVRTC) on cuda:0, eager"},
        "phases": ph, "checksum": float(out["corr"].double().sum().item()), "product_so_loaded": product_so_loaded(),
    }))


def gpu_reference_legs(args):
    """Both modes of the reference-on-GPU leg as subprocesses (the -O mode needs its own interpreter)."""
    res = {}
    base = [os.path.join(ROOT, "bench.py"), "--impl", "reference-gpu", "--steps", "10", "--warmup", "4",
            "--seqs-per-gpu", str(args.seqs_per_gpu), "--height", str(args.height), "--width", str(args.width), "--iters", str(args.iters)]
    for name, flags in (("as_is", []), ("python_O", ["-O"])):
        try:
            env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
            r = subprocess.run([sys.executable] + flags + base, capture_output=True, text=True, timeout=600, env=env)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            res[name] = json.loads(line[-1]) if line else {"unavailable": (r.stderr or "no output")[-300:]}
        except Exception as e:  # noqa: BLE001
            res[name] = {"unavailable": str(e)[:300]}
    return res


# ------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------

def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    import tcs_b200
    from tcs_b200 import sequence

    B, iters = args.seqs_per_gpu, args.iters
    if args.sequences > 0:          # whole sequences per GPU: sequence s lives on rank s mod N
        B = len(sequence.shard_sequences(args.sequences, world, rank))
        if B == 0:
            raise SystemExit("rank %d owns no sequence: --sequences must be >= the number of GPUs" % rank)
    H, W = feature_hw(args.height, args.width)
    fused_build = (args.mode == "pyramid" and args.precision != "fp32" and W <= 240 and W % 4 == 0
                   and os.environ.get("TCS_B200_FUSED_BUILD", "1") != "0")
    K_steps, W_steps = args.steps, max(args.warmup, 3)
    pk = peaks()

    slots_host, slots = [], []
    for s in range(2):
        h, d = make_slot(B, H, W, iters, 1234 + 17 * s + 1000 * rank, device, pin=not args.skip_e2e)
        slots_host.append(h)
        slots.append(d)
    cam = camera(B, H, W, device, 1)
    kw = dict(num_levels=LEVELS, radius=RADIUS, precision=args.precision, mode=args.mode)
    live = [{}, {}]

    def phase_build(s):
        live[s]["blk"] = tcs_b200.CorrBlock1D(slots[s]["f1"], slots[s]["f2"], **kw)

    carries = [tcs_b200.WarpCarry(), tcs_b200.WarpCarry()]   # each frame's fmap1, transposed by its own cost kernel for the next

    def phase_warp(s):
        o = slots[1 - s]
        carry = args.warp_carry and not args.want_fmap
        d, _, m, c = tcs_b200.warp_with_cost(o["last_disp"], o["f1"], cam["rel_T"], cam["K"], cam["K_inv"], cam["baseline"],
                                             cur_fmap=slots[s]["f1"], per_sample_mean=True, want_fmap=args.want_fmap,
                                             carry_in=carries[1 - s] if carry else None,
                                             carry_out=carries[s] if carry else None)
        grid = tcs_b200.get_backward_grid(d, cam["rel_T_inv"], cam["K"], cam["K_inv"], cam["baseline"])
        live[s]["init"] = (d, c, m)
        live[s]["nets"] = tcs_b200.warp_hidden_states(o["nets"], grid)

    def phase_lookup(s):
        blk, coords = live[s]["blk"], slots[s]["coords"]
        out = None
        for it in range(iters):
            out = blk(coords[it])
        live[s]["corr"] = out

    phases = (phase_build, phase_warp, phase_lookup)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: a first frame (argmax initialisation), then temporal frames, eagerly
    first = tcs_b200.hot_path_frame(slots[0]["f1"], slots[0]["f2"], slots[0]["coords"], state=None, **kw)
    del first
    for k in range(W_steps):
        for ph in phases:
            ph(k % 2)
    torch.cuda.synchronize()

    graphs = None
    if not args.no_graph:
        pool = torch.cuda.graph_pool_handle()
        graphs = [[None] * 3 for _ in range(2)]
        for s in range(2):
            for i, ph in enumerate(phases):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    ph(s)
                graphs[s][i] = g
        for s in range(2):                     # one untimed replay of everything
            for g in graphs[s]:
                g.replay()
        torch.cuda.synchronize()

    # ---- timed region: exactly K steps, device-timed, phase boundaries marked with events
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K_steps)]
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    barrier()
    for k in range(K_steps):
        s = k % 2
        for i in range(3):
            ev[k][i].record()
            if graphs is not None:
                graphs[s][i].replay()
            else:
                phases[i](s)
        ev[k][3].record()
    barrier()
    clocks = sampler.stop()
    total_ms = ev[0][0].elapsed_time(ev[-1][3])
    phase_ms = [sum(ev[k][i].elapsed_time(ev[k][i + 1]) for k in range(K_steps)) / K_steps for i in range(3)]
    tmax = torch.tensor([total_ms] + phase_ms, dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms, phase_ms = tmax[0].item(), tmax[1:].tolist()
    # frames of the whole job: every rank's sequences x steps, summed with the path's only collective (final metric reduce)
    frames = int(round(sequence.reduce_metrics([float(B * K_steps)], device=device)[0]))
    value = frames / (total_ms * 1e-3)
    checksum = float(live[(K_steps - 1) % 2]["corr"].double().sum().item())

    # ---- e2e: same step through the public API, inputs in pinned host memory, results read back
    e2e = None
    if not args.skip_e2e:
        copy_streams = [torch.cuda.Stream(), torch.cuda.Stream()]   # one DMA queue per feature map (55 vs 38 GB/s)
        main = torch.cuda.current_stream()
        # three device staging buffers: previous frame (read by the warp), current frame, next frame (in flight)
        stage = [{"f1": torch.empty_like(slots[0]["f1"]), "f2": torch.empty_like(slots[0]["f2"])} for _ in range(3)]
        res_host = [{"corr": torch.empty((B, LEVELS * (2 * RADIUS + 1), H, W), dtype=torch.float32).pin_memory(),
                     "init": [torch.empty((B, 1, H, W), dtype=torch.float32).pin_memory() for _ in range(3)]} for _ in range(2)]
        n_e2e = 2 + K_steps
        copied = [[torch.cuda.Event() for _ in range(2)] for _ in range(n_e2e + 1)]
        done = [torch.cuda.Event() for _ in range(n_e2e + 1)]
        h2d = 2 * slots[0]["f1"].numel() * 4
        d2h = (res_host[0]["corr"].numel() + 3 * res_host[0]["init"][0].numel()) * 4

        def upload(k):
            for j, name in enumerate(("f1", "f2")):
                cs = copy_streams[j]
                with torch.cuda.stream(cs):
                    if k >= 2:
                        cs.wait_event(done[k - 2])              # step k-2 was the last reader of stage[k % 3]
                    stage[k % 3][name].copy_(slots_host[k % 2][name], non_blocking=True)
                    copied[k][j].record(cs)

        def e2e_step(k):
            s = k % 2
            if k + 1 < n_e2e:
                upload(k + 1)                                   # next frame's H2D overlaps this frame's kernels
            main.wait_event(copied[k][0])
            main.wait_event(copied[k][1])
            o = slots[1 - s]
            cur, prev = stage[k % 3], stage[(k - 1) % 3]
            out = tcs_b200.hot_path_frame(cur["f1"], cur["f2"], slots[s]["coords"],
                                          state=(o["last_disp"], prev["f1"], o["nets"]), rel_T=cam["rel_T"],
                                          rel_T_inv=cam["rel_T_inv"], K=cam["K"], K_inv=cam["K_inv"], baseline=cam["baseline"],
                                          carry_in=e2e_carry[(k - 1) % 3] if args.warp_carry else None,
                                          carry_out=e2e_carry[k % 3] if args.warp_carry else None, **kw)
            r = res_host[s]
            r["corr"].copy_(out["corr"], non_blocking=True)
            for dst, name in zip(r["init"], ("sparse_disp", "cost", "mask")):
                dst.copy_(out[name], non_blocking=True)
            done[k].record(main)

        e2e_carry = [tcs_b200.WarpCarry().reserve(slots[0]["f1"]) for _ in range(3)]    # one per staging buffer
        stage[2]["f1"].copy_(slots[1]["f1"])                    # "frame -1" features for the first warp
        # copy-only probe: what the host -> device link gives this rank with nothing else going on (all ranks at once)
        barrier()
        pe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        pe[0].record(main)
        for cs in copy_streams:
            cs.wait_event(pe[0])
        for rep in range(4):
            for j, name in enumerate(("f1", "f2")):
                with torch.cuda.stream(copy_streams[j]):
                    stage[rep % 2][name].copy_(slots_host[rep % 2][name], non_blocking=True)
        for cs in copy_streams:
            main.wait_stream(cs)
        pe[1].record(main)
        barrier()
        probe_gbs = 4 * h2d / (pe[0].elapsed_time(pe[1]) * 1e-3) / 1e9
        upload(0)
        for k in range(2):                                       # warm-up of the e2e loop itself
            e2e_step(k)
        barrier()
        t0 = time.perf_counter()
        for k in range(2, n_e2e):
            e2e_step(k)
        barrier()
        el = time.perf_counter() - t0
        tt = torch.tensor([el], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        rates = torch.tensor([K_steps * h2d / el / 1e9, probe_gbs], dtype=torch.float64, device=device)
        allr = [torch.zeros_like(rates) for _ in range(world)]
        if world > 1:
            dist.all_gather(allr, rates)
        else:
            allr = [rates]
        e2e = {"value": frames / tt.item(), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "h2d_gbs_per_rank_in_loop": [round(r[0].item(), 2) for r in allr],
               "h2d_gbs_per_rank_copy_only": [round(r[1].item(), 2) for r in allr],
               "h2d_gbs_total_copy_only": round(sum(r[1].item() for r in allr), 1),
               "device_ms_per_step_without_copies": total_ms / K_steps,
               "note": "pinned host fmaps -> H2D (two copy streams, triple-buffered staging) -> hot_path_frame (eager public API) -> D2H "
                       "of lookup + init.  Bound by the host->device link: compare h2d_gbs_per_rank_in_loop with the copy-only probe "
                       "(all ranks copying at once, no kernels); the kernels need device_ms_per_step_without_copies"}

    # ---- rooflines
    npix = B * H * W
    lookup_bytes = 308 * npix                                   # 4 coord + 4*10*4 taps + 4*9*4 out per pixel
    lookup_ms = phase_ms[2] / iters
    traffic, traffic_src = None, None
    tp_path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp_path) and (H, W) == (136, 240):
        tj = json.load(open(tp_path))
        # dram__bytes_read.sum + dram__bytes_write.sum per launch from an `ncu --set full` capture of THIS command's
        # lookups in the middle of a step (other phases' data and earlier calls' output planes in L2), not of an isolated launch
        traffic = tj.get("corr_lookup_in_step_bytes_per_launch_B%d" % B, tj.get("corr_lookup_bytes_per_launch_B%d" % B))
        traffic_src = tj.get("source")
    roofline = {"kernel": "corr_lookup_r4x4o_kernel", "bound": "hbm", "achieved": lookup_bytes / (lookup_ms * 1e-3) / 1e9,
                "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": lookup_bytes / (lookup_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": pk["source"], "launches_per_step": iters,
                "algorithmic_bytes_per_launch": lookup_bytes}
    if args.mode == "alternate":
        # alternate mode: the dominant kernel is the tensor-core on-the-fly lookup.  Algorithmic work: per pixel and
        # call 4 levels x 10 distinct taps x a 256-long dot product = 20.5 kFLOP; the kernel issues the whole band of the row block
        # (128 x ~256 x 256 MACs per tile and pass) to get them, which is what `executed_tflops_estimate` counts for a 256-column band
        passes = 3 if args.precision.endswith("x3") else 1
        alg_flops = npix * LEVELS * 10 * 2.0 * C
        exe_flops = 2.0 * npix * min(W, 256) * C * passes
        operand_bytes = 2.0 * npix * C * 2 * (2 if passes == 3 else 1)
        roofline = {"kernel": "corr_lookup_alt_tc_kernel", "bound": "tensor", "achieved": alg_flops / (lookup_ms * 1e-3) / 1e12,
                    "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": alg_flops / (lookup_ms * 1e-3) / 1e12 / pk["bf16_tflops"],
                    "traffic": None, "peak_source": pk["source"], "launches_per_step": iters, "algorithmic_flops_per_launch": alg_flops,
                    "executed_tflops_estimate": exe_flops / (lookup_ms * 1e-3) / 1e12,
                    "operand_stream_gbs": operand_bytes / (lookup_ms * 1e-3) / 1e9,
                    "note": "each call re-streams the 16-bit operands from HBM (operand_stream_gbs) and rebuilds every tile's band in TMEM"}
    build_flops = 2.0 * npix * W * C
    # levels 1 and 3 are not stored any more (every radius-4 lookup kernel re-pools them from levels 0 and 2): the bytes the build
    # has to move are the two fp32 maps in and levels 0 and 2 out.  SURVEY.md 8d's figure (all four levels) is kept beside it.
    lazy_odd = args.mode == "pyramid" and os.environ.get("TCS_B200_LAZY_ODD_LEVELS", "1") != "0" and args.precision != "fp32"
    build_bytes_all_levels = 2 * npix * C * 4 + npix * W * 4 * (1 + 0.5 + 0.25 + 0.125)
    build_bytes = 2 * npix * C * 4 + npix * W * 4 * ((1 + 0.25) if lazy_odd else (1 + 0.5 + 0.25 + 0.125))
    warp_bytes = npix * (4 + 1024 + 1024 + 4 + 4 + 4 + 1344)   # disp + prev fmap + cur fmap in; disp', mask, cost out; hidden gather
    phases_out = {
        "build_ms": phase_ms[0], "warp_ms": phase_ms[1], "lookups_ms": phase_ms[2],
        "build": {"note": ("fused normalise + split + tcgen05 build" if fused_build else "2 pre-passes + tcgen05 build") + ", fp32 fmaps in, fp32 levels out" + (" (levels 0 and 2 stored; 1 and 3 re-pooled by the lookup)" if lazy_odd else ""),
                  "algorithmic_bytes": build_bytes, "algorithmic_bytes_all_four_levels": build_bytes_all_levels,
                  "hbm_gbs": build_bytes / (phase_ms[0] * 1e-3) / 1e9, "hbm_frac": build_bytes / (phase_ms[0] * 1e-3) / 1e9 / pk["hbm_gbs"],
                  "tflops": build_flops / (phase_ms[0] * 1e-3) / 1e12, "tensor_frac": build_flops / (phase_ms[0] * 1e-3) / 1e12 / pk["bf16_tflops"]},
        "warp": {"algorithmic_bytes": warp_bytes, "hbm_gbs": warp_bytes / (phase_ms[1] * 1e-3) / 1e9,
                 "hbm_frac": warp_bytes / (phase_ms[1] * 1e-3) / 1e9 / pk["hbm_gbs"]},
    }

    used_graphs = graphs is not None
    cpu, gpu_ref = None, None
    if rank == 0 and world == 1 and not args.skip_gpu_reference:
        live.clear()
        graphs = None
        torch.cuda.empty_cache()
        gpu_ref = gpu_reference_legs(args)
    if rank == 0 and world == 1 and not args.skip_cpu:
        fps, n, el, kind = cpu_frames_per_second(B, H, W, iters, args.cpu_seconds)
        cpu = {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": kind,
               "sample": "%d steps of %d batched sequences (%dx%d, %d lookups) in %.1f s; %s" % (n, B, args.height, args.width, iters, el, CPU_NOTE[kind])}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K_steps, "warmup": W_steps,
            "ms_per_step": total_ms / K_steps, "higher_is_better": True, "scaling": "strong" if args.sequences > 0 else "weak", "vs_baseline": None,
            "dtype": {"fp16x3": "f16x3->f32", "bf16x3": "bf16x3->f32", "bf16": "bf16->f32", "fp16": "f16->f32", "fp32": "f32"}[args.precision],
            "data": "synthetic",
            "config": {"workload": workload_text(args, B),
                       "feature_hw": [H, W], "seqs_per_gpu": B, "precision": args.precision, "mode": args.mode,
                       "cuda_graphs": used_graphs, "fused_build": fused_build, "warp": ("scatter + warped-feature store (plain drop-in)" if args.want_fmap else "lists on carried transposition" if args.warp_carry else "scatter, cost only"), "l2": "inputs larger than L2 (%.0f MB of feature maps per step, 2 alternating slots)" % (2 * npix * C * 4 / 1e6),
                       "total_sequences": args.sequences if args.sequences > 0 else world * B,
                       "parallelism": "sequences sharded per GPU (sequence s -> rank s mod N), no data-path collective; one all_reduce of the frame count"},
            "e2e": e2e, "gpu_launches": K_steps * sequence.launches_per_frame(iters, False, mode=args.mode, fused_build=fused_build,
                                                                                    warp_lists=args.warp_carry and not args.want_fmap),
            "clocks": clocks, "roofline": roofline, "phases": phases_out, "cpu_baseline": cpu, "gpu_reference": gpu_ref, "checksum": checksum,
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-gpu":
        run_reference_gpu(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
