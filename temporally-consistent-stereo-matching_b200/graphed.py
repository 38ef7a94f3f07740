"""CUDA-graph replay of the learned blocks that TCStereo.forward calls once per GRU iteration.

Once the hot path is a handful of kernels, what bounds the model on a B200 at batch 1 is launch overhead: each of the 32
iterations of core/tc_stereo.py:175-200 issues ~150 small cuDNN / elementwise launches through
`self.update_block`, `self.disp_grad_refine`, `self.disp_refine` and `self.hiddenstate_update` (core/update.py).  The
loop itself is inside the reference's forward and stays there, unmodified; `graph_modules(model)` makes each of those four
module calls ONE graph launch instead:

    tcs_b200.install(core.tc_stereo, fuse_cost=True, stencils=core.update)     # no host-side tensor building left in the modules
    tcs_b200.graph_modules(model)                                              # also strips the asserts of the modules' source files
    out = model(image1, image2, iters=32, test_mode=True, params=...)          # TCStereo.forward, as shipped

Per call signature (argument structure, shapes, dtypes, flags, autocast state) the module's forward is captured once into a
torch.cuda.CUDAGraph on static copies of its tensor arguments; later calls copy the arguments in (skipped for an argument that
is the same, unmodified tensor as last time: the per-frame context features) and replay.  Inference only.  The outputs are the
graph's static tensors: they are overwritten by the module's next call, which is how TCStereo.forward uses them (every output
is consumed, or copied by an arithmetic op, before the same module runs again; the hidden states it returns for the next frame
are read by that frame's warp before its first iteration).  `ungraph_modules(model)` restores the modules.
"""
import torch

from .corr import LazyLookup

GRAPHED = ("update_block", "disp_grad_refine", "disp_refine", "hiddenstate_update")
_static_outputs = set()      # id() of every graph's output tensors: a replay rewrites them WITHOUT touching their version counter


def _flatten(obj, tensors):
    """Nested args -> hashable spec with tensor slots; tensors are appended to `tensors`."""
    if isinstance(obj, LazyLookup):
        obj = obj.materialize()                       # the graph bakes pointers in: a lookup into this frame's pyramid stays outside
    if isinstance(obj, torch.Tensor):
        tensors.append(obj)
        return ("T", len(tensors) - 1)
    if isinstance(obj, (list, tuple)):
        return ("L" if isinstance(obj, list) else "U", tuple(_flatten(o, tensors) for o in obj))
    if isinstance(obj, dict):
        return ("D", tuple((k, _flatten(v, tensors)) for k, v in sorted(obj.items())))
    if obj is None or isinstance(obj, (bool, int, float, str)):
        return ("C", obj)
    raise TypeError("graph_modules: unsupported argument type %r" % (type(obj),))


def _rebuild(spec, tensors):
    kind, val = spec
    if kind == "T":
        return tensors[val]
    if kind == "L":
        return [_rebuild(s, tensors) for s in val]
    if kind == "U":
        return tuple(_rebuild(s, tensors) for s in val)
    if kind == "D":
        return {k: _rebuild(s, tensors) for k, s in val}
    return val


class _Captured:
    def __init__(self):
        self.graph, self.inputs, self.outputs, self.last = None, None, None, None


class GraphedForward:
    """Callable that replaces `module.forward` (instance attribute) with capture-once / replay."""

    def __init__(self, module, name):
        self.module, self.name = module, name
        self.forward = type(module).forward.__get__(module)      # the class's own forward, bound
        self.captured = {}
        self.replays = 0

    def _capture(self, spec, tensors):
        if torch.is_grad_enabled() and any(t.requires_grad for t in tensors):
            raise RuntimeError("graph_modules is inference only: call the model under torch.no_grad()")
        c = _Captured()
        c.inputs = [t.detach().clone() for t in tensors]
        args, kwargs = _rebuild(spec, c.inputs)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                             # library warm-up (cuDNN plans, workspaces) outside the capture
            for _ in range(2):
                self.forward(*_rebuild(spec, c.inputs)[0], **_rebuild(spec, c.inputs)[1])
        torch.cuda.current_stream().wait_stream(side)
        # One private memory pool PER GRAPH.  In a shared pool a later capture's OUTPUT may be given a block that an earlier
        # capture used for an intermediate; the model replays the graphs in an order of its own (disp_refine's up_mask is read
        # after hiddenstate_update ran again), so the earlier graph would overwrite it: measured, the mask came back as garbage.
        c.graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(c.graph):
                c.outputs = self.forward(*args, **kwargs)
        except RuntimeError as e:
            raise RuntimeError("graph_modules: %s.forward cannot be captured (%s).  It must not synchronise or build tensors on the host: "
                               "install the drop-in with stencils=core.update and let graph_modules strip the asserts." % (self.name, e)) from e
        c.last = [None] * len(tensors)
        keep = []
        _flatten(c.outputs, keep)
        _static_outputs.update(id(t) for t in keep)               # (the graph keeps them alive, so the ids stay theirs)
        return c

    def __call__(self, *args, **kwargs):
        tensors = []
        spec = _flatten((args, kwargs), tensors)
        key = (spec, tuple((tuple(t.shape), t.dtype, t.device) for t in tensors), torch.is_autocast_enabled())
        c = self.captured.get(key)
        if c is None:
            c = self.captured[key] = self._capture(spec, tensors)
        for i, (dst, src) in enumerate(zip(c.inputs, tensors)):
            if src.data_ptr() == dst.data_ptr():
                continue                                          # the caller handed the static buffer itself back
            seen = c.last[i]
            if seen is not None and seen[0] is src and seen[1] == src._version and id(src) not in _static_outputs:
                continue                                          # same tensor object, unmodified since it was copied in (another
                                                                  # graph's output is the same object every time but not the same data)
            dst.copy_(src)
            c.last[i] = (src, src._version)                       # keeps it alive: its storage cannot be recycled under the same identity
        c.graph.replay()
        self.replays += 1
        return _fresh_containers(c.outputs)


def _fresh_containers(obj):
    """The same tensors in new lists / tuples / dicts: a caller that edits a returned list must not edit the graph's own."""
    if isinstance(obj, list):
        return [_fresh_containers(o) for o in obj]
    if isinstance(obj, tuple):
        return tuple(_fresh_containers(o) for o in obj)
    if isinstance(obj, dict):
        return {k: _fresh_containers(v) for k, v in obj.items()}
    return obj


def graph_modules(model, names=GRAPHED, strip=True):
    """Replace the forward of model.<name> for every name by a CUDA-graph replay.  Returns {name: GraphedForward}."""
    import sys
    from . import dropin
    out = {}
    mods = set()
    for n in names:
        m = getattr(model, n)
        out[n] = GraphedForward(m, n)
        mods.add(sys.modules[type(m).__module__])
    if strip:
        dropin.strip_asserts(*mods)                               # `assert not torch.isnan(x).any()` is a host sync: not capturable
    for n in names:
        getattr(model, n).forward = out[n]                        # instance attribute: nn.Module.__call__ finds it before the class's
    return out


def ungraph_modules(model, names=GRAPHED, restore=True):
    from . import dropin
    for n in names:
        m = getattr(model, n)
        if "forward" in m.__dict__:
            del m.forward
    if restore:
        dropin.restore_asserts()
