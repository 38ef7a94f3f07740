"""CorrBlock1D — the reference's correlation-block interface on top of libtcs_b200.so.

Mirrors core/corr.py of the reference (constructor :8-31, __call__ :33-52, corr :54-62, get_cost_volume
:64-65, argmax_disp :67-79): same names, argument meaning and return shapes, so that
`core.tc_stereo.CorrBlock1D = CorrBlock1D` is a drop-in.  All arithmetic happens in the CUDA library; this
file validates arguments, owns the device buffers (torch's caching allocator) and passes raw pointers.
Inference only (no autograd), CUDA tensors only, no fallback.
"""
import os

import torch

from . import _lib
from .lazy import LazyTensorOps

_DEFAULT_PRECISION = os.environ.get("TCS_B200_PRECISION", "fp16x3")
_DEFAULT_MODE = os.environ.get("TCS_B200_CORR_MODE", "pyramid")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check_fmap(name, t):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA tensor (libtcs_b200 has no CPU path)" % name)
    if t.dim() != 4:
        raise ValueError("%s must be [B, C, H, W], got %s" % (name, tuple(t.shape)))
    if t.requires_grad and torch.is_grad_enabled():
        # in the reference, gradients flow through the volume into fnet (corr.py:60) and get_cost_volume feeds the
        # training loss (train_stereo.py:385): a training run on these kernels would silently train with them cut
        raise RuntimeError("%s requires grad, but libtcs_b200 is inference only (no backward kernels): run under "
                           "torch.no_grad() with test_mode=True, or uninstall() the drop-in for training" % name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _pad16(n_floats):
    return (n_floats + 3) & ~3


def normalized_operands(fmap, precision="bf16x3", want_hi=True, want_lo=None, want_n32=False, kblocked=False):
    """L2-normalise [B,C,H,W] fp32 features over C and re-lay them out channels-last (ref: corr.py:58-59).

    Returns (hi, lo, n32): 16-bit [B,H,W,C] operands of the tensor-core build (lo only for the x3 modes)
    and/or the fp32 normalised features.  kblocked: the 16-bit operands K-block-major, [B,H,C/64,W,64] (what the
    tensor-core alternate lookup streams)."""
    fmap = _check_fmap("fmap", fmap)
    B, C, H, W = fmap.shape
    fp16 = precision in ("fp16", "fp16x3")
    if want_lo is None:
        want_lo = precision in ("bf16x3", "fp16x3")
    dt = torch.float16 if fp16 else torch.bfloat16
    shape = (B, H, C // 64, W, 64) if kblocked else (B, H, W, C)
    hi = torch.empty(shape, dtype=dt, device=fmap.device) if want_hi else None
    lo = torch.empty(shape, dtype=dt, device=fmap.device) if (want_hi and want_lo) else None
    n32 = torch.empty((B, H, W, C), dtype=torch.float32, device=fmap.device) if want_n32 else None
    prec = _lib.PRECISIONS["fp16" if fp16 else "bf16"]
    with torch.cuda.device(fmap.device):
        if kblocked:
            if want_n32 or not want_hi:
                raise ValueError("kblocked operands are the 16-bit ones")
            _lib.call("tcs_corr_prepass_kblocked", fmap.data_ptr(), hi.data_ptr(), lo.data_ptr() if lo is not None else None,
                      B, C, H, W, prec, _stream())
            return hi, lo, None
        _lib.call("tcs_corr_prepass", fmap.data_ptr(), hi.data_ptr() if hi is not None else None,
                  lo.data_ptr() if lo is not None else None, n32.data_ptr() if n32 is not None else None,
                  B, C, H, W, prec, _stream())
    return hi, lo, n32


def level_pitch(W2, num_levels=4, radius=4):
    """Row pitch (floats) of level 0 for a W2-wide volume: W2 itself, or the next multiple of 16 when that puts the rows of
    levels 0 and 2 on 16-byte boundaries (the lookup's predicate-free kernels; KITTI's 312 -> 320).  Needs W2 % 8 == 0 (no
    pooled entry then mixes real and padding columns) and the standard 4 levels / radius 4."""
    if W2 % 16 != 0 and W2 % 8 == 0 and num_levels == 4 and radius == 4 and os.environ.get("TCS_B200_LEVEL_PITCH", "1") != "0":
        return (W2 + 15) & ~15
    return W2


def _pool_level_into(src, dst):
    """dst = avg_pool2d(src, [1,2]) (ref: corr.py:21-23) on the PHYSICAL rows of two level views (their pitch may exceed their
    width: the padding columns are zeros and pool to zeros): (even + odd) * 0.5, the expression the build epilogue uses, so the
    bits are the ones the build would have stored."""
    B, H, W1, _ = src.shape
    ps = torch.as_strided(src, (B, H, W1, src.stride(2)), src.stride(), src.storage_offset())
    pd = torch.as_strided(dst, (B, H, W1, dst.stride(2)), dst.stride(), dst.storage_offset())
    n = min(pd.shape[3], ps.shape[3] // 2)
    pd[..., :n] = (ps[..., 0:2 * n:2] + ps[..., 1:2 * n:2]) * 0.5
    if n < pd.shape[3]:
        pd[..., n:] = 0


class _Levels(list):
    """The levels of a pyramid, [B,H,W1,W2>>l] each.  `pending`: the build did not store levels 1 and 3 — the row-aligned
    lookup kernels re-pool them from levels 0 and 2 on the fly (csrc/corr_lookup.cu), so a third of the pyramid's bytes never
    has to be written.  Anything that does touch an odd level (an index, a slice, an iteration: corr_pyramid, the tests) gets
    it pooled from the level below first; raw(l) is for the kernels' pointer arguments."""
    pending = False

    def raw(self, l):
        return list.__getitem__(self, l)

    def fill(self):
        if self.pending:
            self.pending = False
            with torch.no_grad():
                for l in range(1, len(self), 2):
                    src = list.__getitem__(self, l - 1)
                    if list.__getitem__(self, l) is None:          # same pitch rule as alloc_pyramid: level l rows are pitch >> l wide
                        B, H, W1, Ws = src.shape
                        phys = torch.empty((B, H, W1, src.stride(2) >> 1), dtype=torch.float32, device=src.device)
                        list.__setitem__(self, l, phys[..., :Ws >> 1])
                    _pool_level_into(src, list.__getitem__(self, l))

    def __getitem__(self, i):
        if self.pending and not (isinstance(i, int) and (i % len(self)) % 2 == 0):
            self.fill()
        return list.__getitem__(self, i)

    def __iter__(self):
        self.fill()
        return list.__iter__(self)


def alloc_pyramid(B, H, W1, W2, num_levels, device, pitch=None, zero=False, skip_odd=False):
    """One flat fp32 buffer holding every level [B,H,W1,W2>>l], each 128-byte aligned (the lookup's 32-byte
    loads need 32) and padded so the lookup may read up to the next 16-byte boundary past a level's end.
    pitch: row pitch of level 0 (level l: pitch >> l); the returned levels are then views of the first W2>>l columns of each
    row, and whoever fills them must leave zeros in the rest (tcs_corr_build does; zero=True for copies).
    skip_odd: levels 1 and 3 get no storage (None in the returned list; _Levels.fill allocates them if ever asked for)."""
    P = pitch or W2
    sizes = [B * H * W1 * (P >> l) for l in range(num_levels)]
    offs, total = [], 0
    for l, s in enumerate(sizes):
        offs.append(total)
        if not (skip_odd and l % 2 == 1):
            total += (_pad16(s) + 4 + 31) & ~31
    flat = (torch.zeros if (zero and P != W2) else torch.empty)(total, dtype=torch.float32, device=device)
    levels = [None if (skip_odd and l % 2 == 1) else flat[o:o + s].view(B, H, W1, P >> l)[..., :W2 >> l]
              for l, (o, s) in enumerate(zip(offs, sizes))]
    return flat, levels


FUSED_MAX_W2 = 240
FUSED_MAX_W1 = 256


def build_pyramid(fmap1, fmap2, num_levels=4, precision="bf16x3", fused=None, pitch=None, odd_levels=True):
    """All pyramid levels of the 1-D all-pairs cosine correlation (ref: corr.py:54-62 + :15-23).

    odd_levels=False (4 levels, tensor-core build): levels 1 and 3 are left unwritten (`levels.pending`) for a consumer that
    re-pools them from levels 0 and 2 - every radius-4 lookup kernel does; see _Levels.
    precision: 'bf16' | 'bf16x3' | 'fp16' | 'fp16x3' (tcgen05 tensor cores) or 'fp32' (CUDA cores).
    fused: None = the single fused kernel (normalise + split + UMMA + pyramid, no operand round trip through HBM)
    whenever the shape allows it (W2 <= 240, W1 <= 256, both multiples of 4), else the pre-pass + build pair;
    True / False force one of the two.  TCS_B200_FUSED_BUILD=0 makes None mean "never fused"."""
    fmap1 = _check_fmap("fmap1", fmap1)
    fmap2 = _check_fmap("fmap2", fmap2)
    B, C, H, W1 = fmap1.shape
    B2, C2, H2, W2 = fmap2.shape
    if (B, C, H) != (B2, C2, H2):
        raise ValueError("fmap1 %s and fmap2 %s disagree on B, C or H" % (tuple(fmap1.shape), tuple(fmap2.shape)))
    if not 1 <= num_levels <= _lib.MAX_LEVELS:
        raise ValueError("num_levels must be in [1, %d]" % _lib.MAX_LEVELS)
    if precision != "fp32" and precision not in _lib.PRECISIONS:
        raise ValueError("unknown precision %r" % (precision,))
    fits = 8 <= W2 <= FUSED_MAX_W2 and W1 <= FUSED_MAX_W1 and W1 % 4 == 0 and W2 % 4 == 0 and C % 32 == 0
    if fused is None:
        fused = fits and precision != "fp32" and os.environ.get("TCS_B200_FUSED_BUILD", "1") != "0"
    if fused or precision == "fp32":
        pitch = None                                # only the pre-pass + tcgen05 build writes pitched rows
    skip_odd = (not odd_levels) and num_levels == 4 and precision != "fp32"
    flat, levels = alloc_pyramid(B, H, W1, W2, num_levels, fmap1.device, pitch=pitch, skip_odd=skip_odd)
    levels = _Levels(levels)
    ptrs = [levels.raw(l).data_ptr() if (l < num_levels and levels.raw(l) is not None) else None for l in range(4)]
    levels.pending = skip_odd
    with torch.cuda.device(fmap1.device):
        if precision == "fp32":
            _, _, a32 = normalized_operands(fmap1, want_hi=False, want_n32=True)
            _, _, b32 = normalized_operands(fmap2, want_hi=False, want_n32=True)
            _lib.call("tcs_corr_build_fp32", a32.data_ptr(), b32.data_ptr(), *ptrs, B, H, W1, W2, C, num_levels, _stream())
        else:
            if fused:
                _lib.call("tcs_corr_build_fused", fmap1.data_ptr(), fmap2.data_ptr(), *ptrs, B, H, W1, W2, C, num_levels,
                          _lib.PRECISIONS[precision], _stream())
                return flat, levels
            a_hi, a_lo, _ = normalized_operands(fmap1, precision)
            b_hi, b_lo, _ = normalized_operands(fmap2, precision)
            _lib.call("tcs_corr_build", a_hi.data_ptr(), a_lo.data_ptr() if a_lo is not None else None,
                      b_hi.data_ptr(), b_lo.data_ptr() if b_lo is not None else None, *ptrs,
                      B, H, W1, W2, C, num_levels, _lib.PRECISIONS[precision], pitch or 0, _stream())
    return flat, levels


def _coords_plane(coords, B, H, W1):
    """Channel 0 of [B,>=1,H,W1] coords as (tensor, pointer, batch stride) without copying when possible."""
    if not coords.is_cuda:
        raise TypeError("coords must be a CUDA tensor")
    if coords.dim() != 4 or coords.shape[0] != B or coords.shape[2] != H or coords.shape[3] != W1:
        raise ValueError("coords must be [B, >=1, H, W], got %s for B=%d H=%d W=%d" % (tuple(coords.shape), B, H, W1))
    c = coords[:, :1]
    if c.dtype != torch.float32:
        c = c.float()
    if not (c.stride(3) == 1 and c.stride(2) == W1):
        c = c.contiguous()
    return c, c.data_ptr(), (c.stride(0) if B > 1 else H * W1)


_packed_cache = {}
# Below this many pixels per call (3 frames of 136x240) the tensor-core form's per-CTA set-up (TMEM allocation, barrier, 9 KB of
# weights) costs more than its GEMM saves: measured 28 vs 24 us at 1 x 136x240, 48 vs 66 us at 8 x 136x240 (tools/time_encode.py).
_ENCODE_TC_MIN_PIXELS = 3 * 136 * 240


def _packed_encoder_weights(weight, bias, w2d, b1d):
    """The B operand of tcs_corr_lookup_encode_tc for this (weight, bias): packed on the device by one small kernel, cached until
    either tensor is modified in place (version counter) or replaced (storage pointer).  A later call on ANOTHER stream waits for
    the packing kernel's event first (the cache is process-wide, the packing ran on whichever stream missed)."""
    key = (weight.data_ptr(), weight._version, None if bias is None else (bias.data_ptr(), bias._version), weight.device.index)
    hit = _packed_cache.get(key)
    stream = torch.cuda.current_stream(weight.device)
    if hit is None:
        if len(_packed_cache) > 16:
            _packed_cache.clear()
        packed = torch.empty(int(_lib.load().tcs_corr_encode_packed_bytes()), dtype=torch.uint8, device=weight.device)
        _lib.call("tcs_corr_encode_pack_weights", w2d.data_ptr(), b1d.data_ptr() if b1d is not None else None, packed.data_ptr(), _stream())
        if torch.cuda.is_current_stream_capturing():
            return packed                  # graph-pool memory: packed again by every replay, never handed to an eager caller
        done = torch.cuda.Event()
        done.record(stream)
        hit = _packed_cache[key] = (packed, done, stream, weight, bias)   # the tensors are kept alive: their pointers stay theirs
    elif hit[2] != stream:
        stream.wait_event(hit[1])
    return hit[0]


class LazyLookup(LazyTensorOps):
    """corr_fn(coords) not yet evaluated.  BasicMotionEncoder's patched forward calls .encode(convc1) and never
    materialises the 36 tap planes; every other use (torch functions, the reference's isnan asserts) goes through
    __torch_function__ and sees the ordinary lookup result."""

    def __init__(self, block, coords):
        self.block, self.coords, self._value = block, coords, None

    def materialize(self):
        if self._value is None:
            # the base class's lookup: a configured block's own __call__ is what returned this deferred object
            self._value = CorrBlock1D.__call__(self.block, self.coords)
        return self._value

    def encode(self, conv, relu=True):
        if self._value is not None:
            # something already forced the 36 tap planes (without `python -O` the reference's isnan/isinf assert at
            # update.py:155 does, every iteration): finish with the ordinary 1x1 instead of looking everything up twice
            out = conv(self._value)
            return torch.relu_(out) if relu else out
        return self.block.lookup_encoded(self.coords, conv.weight, conv.bias, relu=relu)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        unwrap = lambda a: a.materialize() if isinstance(a, LazyLookup) else a
        args = tuple(unwrap(a) for a in args)
        kwargs = {k: unwrap(v) for k, v in (kwargs or {}).items()}
        return func(*args, **kwargs)

    @property
    def shape(self):
        b = self.block
        return torch.Size((b.B, b.num_levels * (2 * b.radius + 1), b.H, b.W1))


class CorrBlock1D:
    """ref: core/corr.py:7-79.  `mode='alternate'` never materialises the volume (subsystem 3)."""

    def __init__(self, fmap1, fmap2, num_levels=4, radius=4, thres=0.2, precision=None, mode=None):
        self.num_levels = num_levels
        self.thres = thres          # stored and unused, as in the reference (argmax uses 0.3, corr.py:73)
        self.radius = radius
        self.precision = precision or _DEFAULT_PRECISION
        self.mode = mode or _DEFAULT_MODE
        if self.mode not in ("pyramid", "alternate"):
            raise ValueError("mode must be 'pyramid' or 'alternate'")
        if not 0 <= radius <= _lib.MAX_RADIUS:
            raise ValueError("radius must be in [0, %d]" % _lib.MAX_RADIUS)
        fmap1 = _check_fmap("fmap1", fmap1)
        fmap2 = _check_fmap("fmap2", fmap2)
        self.B, self.C, self.H, self.W1 = fmap1.shape
        self.W2 = fmap2.shape[3]
        self.device = fmap1.device
        self.fmap1 = fmap1          # what the build read (the model's fmap1 itself when it is fp32 and contiguous)
        self._cost_volume = None
        self.W2p = self.W2
        if self.mode == "pyramid":
            # 4 levels at radius 4 always take the lookup kernels that read only levels 0 and 2 (corr_lookup_r4x4*_kernel)
            lazy_odd = num_levels == 4 and radius == 4 and os.environ.get("TCS_B200_LAZY_ODD_LEVELS", "1") != "0"
            self._flat, self._levels = build_pyramid(fmap1, fmap2, num_levels, self.precision,
                                                     pitch=level_pitch(self.W2, num_levels, radius), odd_levels=not lazy_odd)
            self.W2p = self._levels[0].stride(2)            # the row pitch the build really used (== W2 when dense)
            self._fmaps = None
        else:
            self._flat, self._levels = None, None
            self._fmaps = (fmap1, fmap2)   # argmax_disp / get_cost_volume need level 0 on demand
            # tensor-core formulation (4 levels, radius 4): the 16-bit operands once per frame, the band of each row's
            # block rebuilt in TMEM by every call; anything else: dot products at the taps on the CUDA cores
            self._alt_tc = (num_levels == 4 and radius == 4 and self.C % 64 == 0 and self.W2 >= 16
                            and self.precision in _lib.PRECISIONS and os.environ.get("TCS_B200_ALT_TC", "1") != "0")
            if self._alt_tc:
                self._a_hi, self._a_lo, _ = normalized_operands(fmap1, self.precision, kblocked=True)
                self._b_hi, self._b_lo, _ = normalized_operands(fmap2, self.precision, kblocked=True)
            else:
                _, _, self._a32 = normalized_operands(fmap1, want_hi=False, want_n32=True)
                _, _, b32 = normalized_operands(fmap2, want_hi=False, want_n32=True)
                self._b32 = [b32]
                with torch.cuda.device(self.device):
                    for l in range(1, num_levels):
                        prev = self._b32[-1]
                        nxt = torch.empty((self.B, self.H, prev.shape[2] // 2, self.C), dtype=torch.float32, device=self.device)
                        _lib.call("tcs_fmap_pool_w", prev.data_ptr(), nxt.data_ptr(), self.B, self.H, prev.shape[2], self.C, _stream())
                        self._b32.append(nxt)

    @classmethod
    def from_levels(cls, levels, radius=4):
        """A block over an existing pyramid (levels[l] [B,H,W1,W2>>l] fp32 CUDA): lookup / argmax / cost volume
        of volumes built elsewhere, e.g. the reference's own pyramid in the parity tests."""
        self = cls.__new__(cls)
        lv0 = levels[0]
        self.B, self.H, self.W1, self.W2 = lv0.shape
        self.C, self.fmap1 = None, None
        self.num_levels, self.radius, self.thres = len(levels), radius, 0.2
        self.precision, self.mode, self.device = "external", "pyramid", lv0.device
        self.W2p = level_pitch(self.W2, self.num_levels, radius)
        self._flat, self._levels = alloc_pyramid(self.B, self.H, self.W1, self.W2, self.num_levels, self.device, pitch=self.W2p, zero=True)
        for dst, src in zip(self._levels, levels):
            if tuple(dst.shape) != tuple(src.shape):
                raise ValueError("level shape %s, expected %s" % (tuple(src.shape), tuple(dst.shape)))
            dst.copy_(src)
        self._fmaps, self._cost_volume = None, None
        return self

    # -- reference attributes -----------------------------------------------------------------------
    @property
    def corr_pyramid(self):
        """Levels shaped like the reference's list entries, [B*H*W1, 1, 1, W2>>l] (corr.py:18-23)."""
        if self._levels is None:
            raise RuntimeError("mode='alternate' does not materialise corr_pyramid")
        return [lv.view(self.B * self.H * self.W1, 1, 1, lv.shape[3]) for lv in self._levels]

    @property
    def cost_volume(self):
        return self.get_cost_volume()

    def _level0(self):
        """Level 0 for argmax_disp / get_cost_volume, which may treat a pitched level as a dense volume `pitch` wide: its
        zero columns past W2 are masked like every other w2 > w1 — as long as no w1 reaches them (W1 <= W2, the model's
        case); otherwise a dense copy."""
        if self._levels is not None:
            lv = self._levels[0]
            return lv.contiguous() if (lv.stride(2) != self.W2 and self.W1 > self.W2) else lv
        _, levels = build_pyramid(self._fmaps[0], self._fmaps[1], 1, self.precision)
        return levels[0]

    def _level_ptr(self, l):
        """Level l's address for a kernel argument (an unwritten odd level stays unwritten: see _Levels)."""
        lv = self._levels
        t = lv.raw(l) if isinstance(lv, _Levels) else lv[l]
        return None if t is None else t.data_ptr()

    def _pitch_arg(self):
        """The row pitch for the C-ABI (0 = dense), read off the level-0 view so that a replaced level is seen."""
        p = self._levels[0].stride(2)
        return 0 if p == self.W2 else p

    # -- reference methods ----------------------------------------------------------------------------
    def __call__(self, coords):
        """ref: corr.py:33-52.  coords [B,>=1,H,W] -> [B, num_levels*(2r+1), H, W] fp32."""
        c, cptr, cstride = _coords_plane(coords, self.B, self.H, self.W1)
        out = torch.empty((self.B, self.num_levels * (2 * self.radius + 1), self.H, self.W1),
                          dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            if self.mode == "pyramid":
                ptrs = [self._level_ptr(l) if l < self.num_levels else None for l in range(4)]
                _lib.call("tcs_corr_lookup", *ptrs, cptr, cstride, out.data_ptr(),
                          self.B, self.H, self.W1, self.W2, self.num_levels, self.radius, self._pitch_arg(), _stream())
            elif self._alt_tc:
                _lib.call("tcs_corr_lookup_alt_tc", self._a_hi.data_ptr(), self._a_lo.data_ptr() if self._a_lo is not None else None,
                          self._b_hi.data_ptr(), self._b_lo.data_ptr() if self._b_lo is not None else None, cptr, cstride,
                          out.data_ptr(), self.B, self.H, self.W1, self.W2, self.C, _lib.PRECISIONS[self.precision], _stream())
            else:
                ptrs = [self._b32[l].data_ptr() if l < self.num_levels else None for l in range(4)]
                _lib.call("tcs_corr_lookup_alt", self._a32.data_ptr(), *ptrs, cptr, cstride, out.data_ptr(),
                          self.B, self.H, self.W1, self.W2, self.C, self.num_levels, self.radius, _stream())
        return out

    def lookup_encoded(self, coords, weight, bias=None, relu=True):
        """relu(conv1x1(self(coords))) in one kernel: the lookup fused with the motion encoder's first layer
        (ref: core/update.py:97,104, BasicMotionEncoder.convc1 + F.relu).  weight [Cout, L*(2r+1)] or the conv's
        [Cout, L*(2r+1), 1, 1]; -> [B, Cout, H, W] fp32.  Pyramid mode, 4 levels, radius 4."""
        if self.mode != "pyramid" or self.num_levels != 4 or self.radius != 4:
            raise NotImplementedError("lookup_encoded needs mode='pyramid', num_levels=4, radius=4")
        c, cptr, cstride = _coords_plane(coords, self.B, self.H, self.W1)
        w = weight.detach()
        if w.dim() == 4:
            w = w.reshape(w.shape[0], -1)
        if not w.is_cuda or w.dim() != 2 or w.shape[1] != 36:
            raise ValueError("weight must be a CUDA tensor [Cout, 36] (or [Cout, 36, 1, 1]), got %s" % (tuple(weight.shape),))
        w = w.float().contiguous()
        bb = bias.detach().float().contiguous() if bias is not None else None
        cout = w.shape[0]
        out = torch.empty((self.B, cout, self.H, self.W1), dtype=torch.float32, device=self.device)
        ptrs = [self._level_ptr(l) for l in range(4)]
        with torch.cuda.device(self.device):
            tc = os.environ.get("TCS_B200_ENCODE_TC", "auto")          # "0" never, "1" whenever possible, default: when it pays
            if cout == 64 and self.W2p % 16 == 0 and tc != "0" and (tc == "1" or self.B * self.H * self.W1 >= _ENCODE_TC_MIN_PIXELS):
                # tensor-core form: the weights split / swizzled once per (weight, bias) version
                packed = _packed_encoder_weights(weight, bias, w, bb)
                _lib.call("tcs_corr_lookup_encode_tc", *ptrs, cptr, cstride, packed.data_ptr(), out.data_ptr(),
                          self.B, self.H, self.W1, self.W2, 1 if relu else 0, self._pitch_arg(), _stream())
            else:
                _lib.call("tcs_corr_lookup_encode", *ptrs, cptr, cstride, w.data_ptr(), bb.data_ptr() if bb is not None else None,
                          out.data_ptr(), self.B, self.H, self.W1, self.W2, 4, 4, cout, 1 if relu else 0, self._pitch_arg(), _stream())
        return out

    def lazy(self, coords):
        """A deferred lookup for a consumer that can fuse it (see dropin.install(..., fuse_motion_encoder=...)).
        Anything else that touches it as a tensor gets the ordinary lookup."""
        return LazyLookup(self, coords)

    def get_cost_volume(self):
        """ref: corr.py:25-31,64-65.  [B, W2, H, W1], zero where w2 > w1.  Built on first use."""
        if self._cost_volume is None:
            lvl0 = self._level0()
            P = lvl0.stride(2)                      # a pitched level 0 is a dense volume P wide whose columns past W2 are zero
            out = torch.empty((self.B, P, self.H, self.W1), dtype=torch.float32, device=self.device)
            with torch.cuda.device(self.device):
                _lib.call("tcs_corr_cost_volume", lvl0.data_ptr(), out.data_ptr(), self.B, self.H, self.W1, P, _stream())
            self._cost_volume = out if P == self.W2 else out[:, :self.W2].contiguous()
        return self._cost_volume

    def argmax_disp(self, thres=0.3):
        """ref: corr.py:67-79.  -> (sparse_disp, main_cost, mask), each [B,1,H,W1] fp32."""
        lvl0 = self._level0()
        outs = [torch.empty((self.B, 1, self.H, self.W1), dtype=torch.float32, device=self.device) for _ in range(3)]
        with torch.cuda.device(self.device):
            # (a pitched level 0: its zero columns past W2 are all masked, w2 >= W2 > w1, like every other w2 > w1)
            _lib.call("tcs_corr_argmax", lvl0.data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
                      self.B, self.H, self.W1, lvl0.stride(2), float(thres), _stream())
        return tuple(outs)

    @staticmethod
    def corr(fmap1, fmap2, precision=None):
        """ref: corr.py:54-62.  -> [B, H, W1, 1, W2] fp32."""
        _, levels = build_pyramid(fmap1, fmap2, 1, precision or _DEFAULT_PRECISION)
        lv = levels[0]
        return lv.view(lv.shape[0], lv.shape[1], lv.shape[2], 1, lv.shape[3])
