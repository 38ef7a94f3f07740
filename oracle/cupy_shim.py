"""TEST INFRASTRUCTURE — a stand-in for the five `cupy` names the reference's soft-splat wrapper uses, so that
the reference's OWN CUDA kernel string runs on the GPU box where cupy is not installed.

core/utils/splatting/softsplat.py (reference) builds its kernel source with pure-Python preprocessing
(`cuda_kernel`, softsplat.py:27-216), then calls

    cupy.cuda.compile_with_cache(source, options).get_function(name)(grid=, block=, args=, stream=)   (:219-227, :336-345)

with `cupy.int32` / `cupy.float32` scalars (:17-24) under `@cupy.memoize(for_each_device=True)` (:219).  cupy
(`cupy_cuda117==10.6.0`, requirements.txt:1) cannot be installed offline, but everything it does for those calls is
"NVRTC-compile a string, load the module, launch": this file does exactly that through cuda-python
(`cuda.bindings.nvrtc` / `cuda.bindings.driver`, present in the image).  The kernel text, its launch geometry and its
arguments all stay the reference's; nothing of the splat arithmetic is restated here.

`install()` puts this module into `sys.modules['cupy']`; it must run before `core.utils.geo_utils` is imported.
Only tests/, bench.py's reference legs and __graft_entry__.smoke() may use it (it is the checker, not the product).
"""
import ctypes
import sys
import types

import numpy as np

int32 = np.int32
float32 = np.float32

_modules = []          # keep loaded CUmodules alive


def memoize(for_each_device=False):
    def deco(fn):
        cache = {}

        def wrapper(*args):
            dev = None
            if for_each_device:
                import torch
                dev = torch.cuda.current_device()
            key = (dev,) + args
            if key not in cache:
                cache[key] = fn(*args)
            return cache[key]

        wrapper.__wrapped__ = fn
        return wrapper
    return deco


def _check(res):
    err, rest = res[0], res[1:]
    if int(err) != 0:
        raise RuntimeError("CUDA/NVRTC call failed: %r" % (err,))
    if not rest:
        return None
    return rest[0] if len(rest) == 1 else rest


def compile_source(source, options=(), arch=None):
    """NVRTC: source string -> cubin bytes for `arch` (e.g. 'sm_100'); raises with the compiler log on failure."""
    from cuda.bindings import nvrtc
    if arch is None:
        import torch
        major, minor = torch.cuda.get_device_capability()
        arch = "sm_%d%d" % (major, minor)
    opts = [b"--gpu-architecture=" + arch.encode()]
    for o in options:                       # the reference passes '-I <dir>' with a blank inside one option string
        o = o.strip()
        if o.startswith("-I"):
            o = "-I" + o[2:].strip()
        opts.append(o.encode())
    prog = _check(nvrtc.nvrtcCreateProgram(source.encode(), b"reference_kernel.cu", 0, [], []))
    res = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    if int(res[0]) != 0:
        n = _check(nvrtc.nvrtcGetProgramLogSize(prog))
        log = b" " * n
        _check(nvrtc.nvrtcGetProgramLog(prog, log))
        raise RuntimeError("NVRTC failed on the reference's kernel:\n" + log.decode(errors="replace"))
    n = _check(nvrtc.nvrtcGetCUBINSize(prog))
    cubin = b" " * n
    _check(nvrtc.nvrtcGetCUBIN(prog, cubin))
    nvrtc.nvrtcDestroyProgram(prog)
    return cubin


class _Function:
    def __init__(self, fn, name):
        self.fn, self.name = fn, name
        self.launches = 0

    def __call__(self, grid, block, args, stream=None, shared_mem=0):
        from cuda.bindings import driver
        vals, types_ = [], []
        for a in args:
            if isinstance(a, np.int32):
                vals.append(int(a)); types_.append(ctypes.c_int)
            elif isinstance(a, np.float32):
                vals.append(float(a)); types_.append(ctypes.c_float)
            elif isinstance(a, int):                 # tensor.data_ptr()
                vals.append(a); types_.append(ctypes.c_void_p)
            else:
                raise TypeError("unsupported kernel argument %r" % (type(a),))
        g = tuple(grid) + (1,) * (3 - len(grid))
        b = tuple(block) + (1,) * (3 - len(block))
        s = getattr(stream, "ptr", 0) if stream is not None else 0
        _check(driver.cuLaunchKernel(self.fn, g[0], g[1], g[2], b[0], b[1], b[2], shared_mem, s,
                                     (tuple(vals), tuple(types_)), 0))
        self.launches += 1


class _Module:
    def __init__(self, cubin):
        from cuda.bindings import driver
        import torch
        torch.cuda.current_stream()          # makes sure torch's primary context is current on this thread
        self.mod = _check(driver.cuModuleLoadData(cubin))
        _modules.append(self)

    def get_function(self, name):
        from cuda.bindings import driver
        return _Function(_check(driver.cuModuleGetFunction(self.mod, name.encode())), name)


def _compile_with_cache(source, options=(), arch=None, **_):
    return _Module(compile_source(source, options, arch))


def _get_cuda_path():
    import os
    return os.environ.get("CUDA_HOME") or os.environ.get("CUDA_PATH") or "/usr/local/cuda"


cuda = types.SimpleNamespace(compile_with_cache=_compile_with_cache, get_cuda_path=_get_cuda_path)


def install():
    """Make `import cupy` resolve to this module (idempotent)."""
    sys.modules["cupy"] = sys.modules[__name__]
    return sys.modules[__name__]
