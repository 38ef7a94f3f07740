"""Operator plumbing shared by the deferred results of the drop-in (corr.LazyLookup, dropin.LazyWarpedFmap).

A deferred result implements `materialize()` and `__torch_function__`; torch functions reach it through the latter.
Python operators and attribute reads are not torch functions, so they are routed there (operators) or answered by the
materialised tensor (attributes) — a consumer the fusion does not know about simply sees a tensor."""
import torch


class LazyTensorOps:
    def __getattr__(self, name):                 # only reached when normal lookup fails
        if name.startswith("_"):
            raise AttributeError(name)
        fn = getattr(torch, name, None)
        if callable(fn) and callable(getattr(torch.Tensor, name, None)):      # x.sum(...) == torch.sum(x, ...)
            return lambda *a, **k: fn(self, *a, **k)
        return getattr(self.materialize(), name)

    dtype = torch.float32

    def __add__(self, o): return torch.add(self, o)
    def __radd__(self, o): return torch.add(o, self)
    def __sub__(self, o): return torch.sub(self, o)
    def __rsub__(self, o): return torch.sub(o, self)
    def __mul__(self, o): return torch.mul(self, o)
    def __rmul__(self, o): return torch.mul(o, self)
    def __truediv__(self, o): return torch.div(self, o)
    def __rtruediv__(self, o): return torch.div(o, self)
    def __neg__(self): return torch.neg(self)
    def __getitem__(self, idx): return self.materialize()[idx]
    def __len__(self): return self.shape[0]
