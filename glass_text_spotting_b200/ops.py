"""Tensor-level wrappers over the C ABI (include/glass_b200.h).

PyTorch is used only for device memory and streams: every function here takes/returns CUDA
tensors, passes raw pointers to ``libglass_b200.so`` and enqueues on the current stream.
There is no fallback: a missing extension or a non-CUDA tensor raises.
"""
import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import lib as _lib

MODE_SPLIT = 0    # fp16x3 split: fp32-grade precision (parity mode; default)
MODE_FAST = 1     # single fp16 pass (fast mode, 11-bit operands)
ACT_SCALE = 16.0  # == kActScale in csrc/common.cuh: split planes hold ACT_SCALE * x

# k-blocks (64 wide) accumulated inside the tensor core between drains to fp32 registers (0 = library default)
KB_PER_CHUNK = int(__import__("os").environ.get("GLASS_KB_PER_CHUNK", "0"))
# strided convs gather into the row order of their padded output plane (flat GEMM, TMA-store epilogue); 0 = dense rows
GATHER_PADDED = int(__import__("os").environ.get("GLASS_GATHER_PADDED", "1"))
# 0 = auto, 1 = never pair CTAs, 2 = always use tcgen05 cta_group::2 CTA pairs (when the tile width allows)
PAIR_MODE = int(__import__("os").environ.get("GLASS_PAIR_MODE", "0"))
# 0 = auto (3x3 convs stage one activation block per tap row), 1 = every tap loads its own tile
TAP_MODE = int(__import__("os").environ.get("GLASS_TAP_MODE", "0"))
# 0 = auto (TMA-store epilogue for flat layers), 1 = direct-store epilogue everywhere, 2 = insist on TMA (tests)
EPI_MODE = int(__import__("os").environ.get("GLASS_EPI_MODE", "0"))

# GLASS_DEBUG_SAT=1: every GEMM counts the outputs that saturate the split-fp16 storage range (|y| > 3750) into one device
# counter, read with saturation_count() -- a checkpoint with a hot activation then shows up instead of degrading silently
DEBUG_SAT = __import__("os").environ.get("GLASS_DEBUG_SAT", "0") == "1"
_SAT = {}


def _sat_counter() -> torch.Tensor:
    dev = torch.cuda.current_device()
    if dev not in _SAT:
        _SAT[dev] = torch.zeros((1,), dtype=torch.int32, device=f"cuda:{dev}")
    return _SAT[dev]


def saturation_count(reset: bool = False) -> int:
    """Outputs clamped by the split-fp16 storage format since the last reset (only counted under GLASS_DEBUG_SAT=1 /
    ops.DEBUG_SAT = True)."""
    c = _sat_counter()
    v = int(c.item())
    if reset:
        c.zero_()
    return v


# bench.py's per-kernel timing: when a list, every conv_gemm launch appends (start event, end event, algorithmic FLOPs)
PROFILE = None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("glass_text_spotting_b200 ops need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


BORDER_SHARED = 0x100   # == GLASS_BORDER_SHARED (include/glass_b200.h)


class Act:
    """An fp32 activation [n,c,h,w] stored as split-fp16 padded NHWC (see include/glass_b200.h).

    ``buf`` is one fp16 tensor [2, n(+1), hp, wp, cp]: plane 0 = hi, plane 1 = lo.  Borders and pad
    channels are zero and are never written with anything but zeros by any kernel.

    ``shared`` = shared-border planes (GLASS_BORDER_SHARED): only LEADING zero rows / columns, hp = h + border,
    wp = w + border; a row's right neighbour is the next row's leading zero column and the row below a plane is the next
    plane's leading zero row, so the buffer carries one more (all-zero) plane after the last image.
    """

    def __init__(self, n: int, c: int, h: int, w: int, border: int = 1, cp: Optional[int] = None,
                 device="cuda", buf: Optional[torch.Tensor] = None, shared: bool = False):
        self.n, self.c, self.h, self.w, self.border, self.shared = n, c, h, w, border, shared
        self.cp = cp if cp is not None else round_up(c, 64)
        assert self.cp % 8 == 0 and self.cp >= c
        hi = 0 if shared else border
        self.hp, self.wp = h + border + hi, w + border + hi
        shape = (2, n + (1 if shared else 0), self.hp, self.wp, self.cp)
        if buf is None:
            buf = torch.zeros(shape, dtype=torch.float16, device=device)
        else:
            # ``buf`` may have spare capacity along n (Workspace): this Act is then a view of its first n images
            assert buf.dtype == torch.float16 and buf.is_contiguous() and buf.shape[0] == 2 and buf.shape[1] >= shape[1] \
                and tuple(buf.shape[2:]) == shape[2:]
        self.buf = buf

    @property
    def border_code(self) -> int:
        """the `border` argument of the C ABI for this activation's planes"""
        return self.border | (BORDER_SHARED if self.shared else 0)

    @property
    def hi(self) -> torch.Tensor:
        return self.buf[0, : self.n]

    @property
    def lo(self) -> torch.Tensor:
        return self.buf[1, : self.n]

    @property
    def rows(self) -> int:
        return self.n * self.hp * self.wp

    def interior(self) -> torch.Tensor:
        """[2, n, h, w, cp] view of the pixels (tests)"""
        b = self.border
        return self.buf[:, : self.n, b: b + self.h, b: b + self.w]

    @staticmethod
    def from_nchw(x: torch.Tensor, border: int = 1, cp: Optional[int] = None, shared: bool = False) -> "Act":
        n, c, h, w = x.shape
        a = Act(n, c, h, w, border, cp, x.device, shared=shared)
        x = x.contiguous().float()
        _lib.check(_lib.load().glass_pack_nchw(_ptr(x), n, c, h, w, _ptr(a.hi), _ptr(a.lo), a.cp, a.border_code, _stream()))
        return a

    def to_nchw(self) -> torch.Tensor:
        out = torch.empty((self.n, self.c, self.h, self.w), dtype=torch.float32, device=self.buf.device)
        _lib.check(_lib.load().glass_unpack_nchw(_ptr(self.hi), _ptr(self.lo), self.n, self.c, self.h, self.w,
                                                 self.cp, self.border_code, _ptr(out), _stream()))
        return out


class F32Map:
    """fp32 padded NHWC feature map [n, h+2b, w+2b, ld] (the conv kernel's fp32 output; RoIAlign input)."""

    def __init__(self, n, c, h, w, border=1, ld=None, device="cuda"):
        self.n, self.c, self.h, self.w, self.border = n, c, h, w, border
        self.ld = ld if ld is not None else round_up(c, 64)
        self.hp, self.wp = h + 2 * border, w + 2 * border
        self.buf = torch.zeros((n, self.hp, self.wp, self.ld), dtype=torch.float32, device=device)

    def to_nchw(self) -> torch.Tensor:
        out = torch.empty((self.n, self.c, self.h, self.w), dtype=torch.float32, device=self.buf.device)
        _lib.check(_lib.load().glass_nhwc_f32_to_nchw(_ptr(self.buf), self.n, self.c, self.h, self.w, self.ld,
                                                      self.border, _ptr(out), _stream()))
        return out


class PackedWeight:
    """Weights of one conv / Linear packed K-major [n_p, taps*cin_p] as fp16 hi/lo, plus the folded
    per-channel epilogue (scale, bias).  Built by ``packing.pack_conv`` / ``pack_linear``."""

    def __init__(self, w: torch.Tensor, scale: torch.Tensor, bias: torch.Tensor, cout: int, cin: int,
                 kh: int, kw: int, stride: Tuple[int, int], pad: Tuple[int, int], cin_p: int):
        self.w = w  # fp16 [2, n_p, taps*cin_p]
        self.scale, self.bias = scale, bias
        self.cout, self.cin, self.kh, self.kw = cout, cin, kh, kw
        self.stride, self.pad, self.cin_p = stride, pad, cin_p
        self.n_p = w.shape[1]


def conv_gemm(a_hi: torch.Tensor, a_lo: Optional[torch.Tensor], rows_a: int, k_per_tap: int,
              tap_shift: Sequence[int], w: PackedWeight, m_geom: Tuple[int, int, int, int],
              out: Optional[Act] = None, out_f32: Optional[torch.Tensor] = None, ld_f32: int = 0,
              out_geom: Optional[Tuple[int, int, int]] = None, ld_out: int = 0,
              out_hi: Optional[torch.Tensor] = None, out_lo: Optional[torch.Tensor] = None,
              residual: Optional[Act] = None, res_shift: int = 0, relu_pre: bool = False, relu_post: bool = False,
              mode: int = MODE_SPLIT, n_store: int = 0, a_ld: int = 0, a_col0: int = 0, a_inner: int = 0,
              m_count: Optional[Tuple[torch.Tensor, int]] = None, res_geom: Optional[Tuple[int, int, int]] = None,
              valid_pixels: Optional[int] = None) -> None:
    """Raw launch of glass_conv_gemm.  m_geom = (imgs, h, w, border) of the M space;
    out_geom = (hp, wp, border) of the output rows (defaults to ``out``'s geometry).
    ``m_count`` = (int32 device scalar, rows per count): only rows < count * rows_per_count are computed (the live word
    count stays on the device; m_geom then describes the capacity)."""
    p = _lib.ConvGemmParams()
    p.a_hi, p.a_lo, p.rows_a, p.a_ld = _ptr(a_hi), _ptr(a_lo), rows_a, a_ld
    p.k_per_tap, p.ntaps = k_per_tap, len(tap_shift)
    for i, s in enumerate(tap_shift):
        p.tap_shift[i] = s
    p.b_hi, p.b_lo, p.n, p.mode = _ptr(w.w[0]), _ptr(w.w[1]), w.n_p, mode
    p.m_imgs, p.m_h, p.m_w, p.m_border = m_geom
    p.scale, p.bias = _ptr(w.scale), _ptr(w.bias)  # scale also undoes the operand pre-scales
    p.relu_pre, p.relu_post = int(relu_pre), int(relu_post)
    if out is not None:
        out_hi, out_lo = out.hi, out.lo
        out_geom = (out.hp, out.wp, out.border_code)
        ld_out = out.cp
    if residual is not None:
        p.res_hi, p.res_lo = _ptr(residual.hi), _ptr(residual.lo)
        if res_geom is not None:   # the residual's memory seen in another row geometry (pixel-grouped rows)
            p.res_hp, p.res_wp, p.res_border, p.res_shift = res_geom[0], res_geom[1], res_geom[2], res_shift
        else:
            p.res_hp, p.res_wp, p.res_border, p.res_shift = residual.hp, residual.wp, residual.border_code, res_shift
            if ld_out == 0:
                ld_out = residual.cp
            assert residual.cp == ld_out, "residual and output must share the channel stride"
    p.out_hi, p.out_lo, p.out_f32 = _ptr(out_hi), _ptr(out_lo), _ptr(out_f32)
    p.out_hp, p.out_wp, p.out_border = out_geom
    p.ld_out, p.ld_f32, p.n_store = ld_out, ld_f32, n_store
    p.kb_per_chunk = KB_PER_CHUNK or getattr(w, "kb_per_chunk", 0)   # env override > per-weight choice > library default
    p.pair_mode = PAIR_MODE if (PAIR_MODE != 2 or w.n_p % 32 == 0) else 0
    p.tap_mode = TAP_MODE
    p.epi_mode = EPI_MODE
    if DEBUG_SAT:
        p.sat_count = _ptr(_sat_counter())
    p.a_col0, p.a_inner = a_col0, a_inner
    if m_count is not None:
        p.m_count_dev, p.m_rows_per_count = _ptr(m_count[0]), int(m_count[1])
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        _launch_gemm(p)
        e1.record(torch.cuda.current_stream())
        b_lo = p.m_border & 0xff
        b_hi = 0 if (p.m_border & BORDER_SHARED) else b_lo
        m_valid = p.m_imgs * (p.m_h - b_lo - b_hi) * (p.m_w - b_lo - b_hi)
        if valid_pixels is not None:     # pixel-grouped launches: the M space is groups of padded pixels
            m_valid = valid_pixels / getattr(w, "grouped_p", 1)
        # algorithmic FLOPs of this launch + the GEMM's shape (tools/layer_profile.py).  With a device-side M count the
        # launch is sized for the capacity but computes only the live rows: the FLOPs are scaled by live / capacity when
        # the profile is read (profile_flops), after the stream has been synchronised.
        live = None
        if m_count is not None:
            live = (m_count[0], float(m_count[1]) / float(p.m_imgs * p.m_h * p.m_w))
        PROFILE.append((e0, e1, 2.0 * m_valid * getattr(w, "grouped_p", 1) * w.cout * w.cin * w.kh * w.kw,
                        dict(m=p.m_imgs * p.m_h * p.m_w, n=w.n_p, k=k_per_tap * len(tap_shift), taps=len(tap_shift),
                             a_ld=a_ld, res=residual is not None, f32=out_f32 is not None), live))
        return
    _launch_gemm(p)


def profile_flops(rec) -> float:
    """Algorithmic FLOPs of one PROFILE record (call after a synchronise): capacity FLOPs x the live fraction of M."""
    flops, live = rec[2], rec[4]
    if live is None:
        return flops
    return flops * min(1.0, float(live[0].item()) * live[1])


# Launch plans (glass_plan_create / glass_plan_launch): the model's buffers are persistent workspaces, so every step
# repeats the same parameter blocks -- the host-side derivation (validation, 4 TMA descriptors, geometry) is done once
# per distinct block and the steady state pays one cudaLaunchKernelEx per GEMM.  GLASS_PLAN_CACHE=0 disables it.
_PLANS = {}
_PLAN_CACHE = __import__("os").environ.get("GLASS_PLAN_CACHE", "1") != "0"
_PLAN_CACHE_MAX = 4096


def _launch_gemm(p) -> None:
    L = _lib.load()
    if not _PLAN_CACHE:
        _lib.check(L.glass_conv_gemm(C.byref(p), _stream()))
        return
    key = (torch.cuda.current_device(), bytes(p))
    plan = _PLANS.get(key)
    if plan is None:
        if len(_PLANS) >= _PLAN_CACHE_MAX:
            clear_plans()
        h = C.c_void_p()
        _lib.check(L.glass_plan_create(C.byref(p), C.byref(h)))
        plan = _PLANS[key] = h.value
    _lib.check(L.glass_plan_launch(plan, _stream()))


def clear_plans() -> None:
    L = _lib.load()
    for h in _PLANS.values():
        L.glass_plan_destroy(h)
    _PLANS.clear()


def conv2d(x: Act, w: PackedWeight, relu: bool = False, residual: Optional[Act] = None, res_shift: int = 0,
           relu_pre: bool = False, out: Optional[Act] = None, f32: Optional[F32Map] = None, want_act: bool = True,
           mode: int = MODE_SPLIT, gather_buf: Optional[torch.Tensor] = None,
           n_dev: Optional[torch.Tensor] = None) -> Optional[Act]:
    """conv (+ folded norm) (+ReLU) (+residual) on a split-fp16 activation.

    stride-1 'same' convs run as shifted-row implicit GEMM straight from ``x``; everything else goes
    through one tap-gather pass.  ``relu`` = ReLU after the residual add (d2 bottleneck order),
    ``relu_pre`` = ReLU before it (CNN_V1_1 order).  ``n_dev``: int32 device scalar = live images / words (x.n is then
    the capacity the launch is sized for)."""
    if getattr(w, "grouped_p", 0):
        if x.wp % w.grouped_p == 0 and res_shift == 0 and not x.shared:
            return _conv2d_grouped(x, w, relu, residual, relu_pre, out, mode, n_dev)
        w = w.fallback
    if getattr(w, "compact_cp", 0):
        return _conv2d_compact(x, w, relu, residual, res_shift, relu_pre, out, mode, n_dev)
    assert x.cp == w.cin_p, (x.cp, w.cin_p)
    sh, sw = w.stride
    ph, pw = w.pad
    ho = (x.h + 2 * ph - w.kh) // sh + 1
    wo = (x.w + 2 * pw - w.kw) // sw + 1
    if want_act and out is None:
        out = Act(x.n, w.cout, ho, wo, 1, w.n_p, x.buf.device, shared=x.shared)
    if out is not None:
        assert (out.n, out.h, out.w, out.cp) == (x.n, ho, wo, w.n_p)
    geom = (out.hp, out.wp, out.border_code) if out is not None else (f32.hp, f32.wp, f32.border)
    kwargs = dict(out=out, out_f32=None if f32 is None else f32.buf, ld_f32=0 if f32 is None else f32.ld,
                  out_geom=geom, residual=residual, res_shift=res_shift, relu_pre=relu_pre, relu_post=relu,
                  mode=mode)
    flat = (sh == 1 and sw == 1 and ho == x.h and wo == x.w and w.kh % 2 == 1 and w.kw % 2 == 1
            and x.border >= ph and x.border >= pw)
    if flat:
        shifts = [(r - ph) * x.wp + (s - pw) for r in range(w.kh) for s in range(w.kw)]
        if out is not None:
            assert out.shared == x.shared, "a flat conv keeps its input's plane geometry"
        conv_gemm(x.hi, x.lo, x.rows, x.cp, shifts, w, (x.n, x.hp, x.wp, x.border_code),
                  m_count=None if n_dev is None else (n_dev, x.hp * x.wp), **kwargs)
    else:
        taps = w.kh * w.kw
        # a split-fp16 output plane: gather into ITS row order, so that the GEMM is flat (TMA-store epilogue); a caller's
        # dense gather buffer or an fp32 output keep the dense M space
        db = out.border_code if (out is not None and f32 is None and GATHER_PADDED) else 0
        if gather_buf is not None and gather_buf.shape[1] != gather_rows(x.n, ho, wo, db):
            db = 0
        rows = gather_rows(x.n, ho, wo, db)
        g = gather_taps(x, w.kh, w.kw, sh, sw, ph, pw, ho, wo, out=gather_buf, n_dev=n_dev, dst_border=db)
        # the gathered matrix is already tap-major: one "tap" of width taps*cp
        if db:
            conv_gemm(g[0], g[1], rows, taps * x.cp, [0], w, (x.n, out.hp, out.wp, db),
                      m_count=None if n_dev is None else (n_dev, out.hp * out.wp), **kwargs)
        else:
            conv_gemm(g[0], g[1], rows, taps * x.cp, [0], w, (x.n, ho, wo, 0),
                      m_count=None if n_dev is None else (n_dev, ho * wo), **kwargs)
    return out


def _conv2d_compact(x: Act, w: PackedWeight, relu, residual, res_shift, relu_pre, out, mode, n_dev=None) -> Act:
    """stride-1 'same' 1x1 / 3x3 conv on a narrow activation (cp = 8/16/32): a 64-wide k-block spans 64/cp
    consecutive pixels, so the three s-taps of a row are read by one (cp <= 16) or two (cp = 32) TMA boxes."""
    cp = w.compact_cp
    assert x.cp == cp and w.stride == (1, 1) and x.border >= w.pad[0] and (w.kh, w.kw) in ((1, 1), (3, 3)) and not x.shared
    ppk = 64 // cp
    nj = (w.kw + ppk - 1) // ppk
    shifts = [(r - w.pad[0]) * x.wp - w.pad[1] + j * ppk for r in range(w.kh) for j in range(nj)]
    if out is None:
        out = Act(x.n, w.cout, x.h, x.w, 1, w.n_p, x.buf.device)
    assert (out.n, out.h, out.w, out.cp) == (x.n, x.h, x.w, w.n_p)
    conv_gemm(x.hi, x.lo, x.rows, 64, shifts, w, (x.n, x.hp, x.wp, x.border), out=out, residual=residual,
              res_shift=res_shift, relu_pre=relu_pre, relu_post=relu, mode=mode, a_ld=cp,
              m_count=None if n_dev is None else (n_dev, x.hp * x.wp))
    return out


def _conv2d_grouped(x: Act, w: PackedWeight, relu, residual, relu_pre, out, mode, n_dev=None) -> Act:
    """Pixel-grouped implicit GEMM (include/glass_b200.h, a_inner > 0; packing.pack_conv_grouped): P pixels per GEMM
    row, then the border of the output (which the all-valid M space overwrites) is re-zeroed.  A residual of the output's
    own geometry is read through the same grouped row view (its zero border adds nothing to the rows re-zeroed anyway)."""
    P, cp = w.grouped_p, w.grouped_cp
    assert x.cp == cp and x.border == 1 and not x.shared and x.wp % P == 0 and x.rows % P == 0, (x.cp, cp, x.wp, P)
    if out is None:
        out = Act(x.n, w.cout, x.h, x.w, 1, w.grouped_cout_p, x.buf.device)
    assert (out.n, out.h, out.w, out.cp, out.border) == (x.n, x.h, x.w, w.grouped_cout_p, 1)
    if residual is not None:
        assert (residual.n, residual.h, residual.w, residual.cp, residual.border) == (out.n, out.h, out.w, out.cp, 1)
    rows = x.rows // P
    kwin = w.cin_p  # K per tap row (window, padded to 64)
    left = 1 if w.kw == 3 else 0
    # window of group g, tap row r starts at pixel P*g - left + (r - pad)*wp = row (g - left + (r - pad)*wp/P), column
    # left * (P - 1) * cp of the overlapping-row tensor
    shifts = [(r - w.pad[0]) * (x.wp // P) - left for r in range(w.kh)]
    col0 = left * (P - 1) * cp
    conv_gemm(x.hi, x.lo, rows, kwin, shifts, w, (1, rows, 1, 0), out_hi=out.hi, out_lo=out.lo, out_geom=(rows, 1, 0),
              ld_out=P * out.cp, relu_pre=relu_pre, relu_post=relu, mode=mode, a_ld=P * cp, a_col0=col0,
              a_inner=col0 + kwin, m_count=None if n_dev is None else (n_dev, x.hp * x.wp // P),
              residual=residual, res_geom=(rows, 1, 0), valid_pixels=x.n * x.h * x.w,
              n_store=w.grouped_n_store if w.grouped_n_store != w.n_p else 0)
    _lib.check(_lib.load().glass_zero_border(_ptr(out.hi), _ptr(out.lo), out.n, out.h, out.w, out.cp, _ptr(n_dev),
                                             _stream()))
    return out


def linear(a: torch.Tensor, w: PackedWeight, relu: bool = False, want_split: bool = True, want_f32: bool = False,
           mode: int = MODE_SPLIT, out: Optional[torch.Tensor] = None, out_f32: Optional[torch.Tensor] = None,
           m_count: Optional[Tuple[torch.Tensor, int]] = None):
    """y = a @ W^T (*scale + bias) for a split-fp16 matrix ``a`` [2, rows, k] (k == w.cin_p).
    Returns (split [2, rows, n_p] or None, fp32 [rows, n_p] or None); ``out`` / ``out_f32`` are caller-owned result
    buffers of those shapes (persistent workspaces: no allocation, stable pointers for the plan cache)."""
    assert a.dim() == 3 and a.shape[0] == 2 and a.shape[2] == w.cin_p and a[0].is_contiguous()
    rows = a.shape[1]
    o = of = None
    if want_split:
        o = out if out is not None else torch.empty((2, rows, w.n_p), dtype=torch.float16, device=a.device)
        assert tuple(o.shape) == (2, rows, w.n_p) and o[0].is_contiguous()
    if want_f32:
        of = out_f32 if out_f32 is not None else torch.empty((rows, w.n_p), dtype=torch.float32, device=a.device)
        assert tuple(of.shape) == (rows, w.n_p) and of.is_contiguous()
    conv_gemm(a[0], a[1], rows, w.cin_p, [0], w, (1, rows, 1, 0),
              out_hi=None if o is None else o[0], out_lo=None if o is None else o[1], out_f32=of, ld_f32=w.n_p,
              out_geom=(rows, 1, 0), ld_out=w.n_p, relu_post=relu, mode=mode, m_count=m_count)
    return o, of


def maxpool2d(x: Act, k: Tuple[int, int], s: Tuple[int, int], p: Tuple[int, int], out: Optional[Act] = None,
              n_dev: Optional[torch.Tensor] = None) -> Act:
    ho = (x.h + 2 * p[0] - k[0]) // s[0] + 1
    wo = (x.w + 2 * p[1] - k[1]) // s[1] + 1
    if out is None:
        out = Act(x.n, x.c, ho, wo, 1, x.cp, x.buf.device, shared=x.shared)
    assert (out.n, out.h, out.w, out.cp) == (x.n, ho, wo, x.cp)
    _lib.check(_lib.load().glass_maxpool(_ptr(x.hi), _ptr(x.lo), x.n, x.h, x.w, x.cp, x.border_code, k[0], k[1], s[0],
                                         s[1], p[0], p[1], ho, wo, _ptr(out.hi), _ptr(out.lo), out.border_code,
                                         _ptr(n_dev), _stream()))
    return out


def gather_rows(n: int, ho: int, wo: int, dst_border: int = 0) -> int:
    """Rows of the gathered matrix: dense n*ho*wo, or the rows of the padded / shared-border OUTPUT planes."""
    lo, hi = (dst_border & 0xFF), (0 if dst_border & BORDER_SHARED else dst_border & 0xFF)
    return n * (ho + lo + hi) * (wo + lo + hi)


def gather_taps(x: Act, kh: int, kw: int, sh: int, sw: int, ph: int, pw: int, ho: int, wo: int,
                out: Optional[torch.Tensor] = None, n_dev: Optional[torch.Tensor] = None,
                dst_border: int = 0) -> torch.Tensor:
    """im2col of a split activation: rows [2, n*ho*wo, kh*kw*cp] (tap-major K); with ``dst_border`` (the border code of
    the conv's OUTPUT plane) the row of output pixel (y, x) sits at that plane's flattened position, so that the GEMM's M
    space is the output plane and the layer finishes through the TMA-store epilogue (border rows are not written: whatever
    they hold is multiplied and thrown away, the epilogue writes zeros there)."""
    rows = gather_rows(x.n, ho, wo, dst_border)
    if out is None:
        out = torch.zeros((2, rows, kh * kw * x.cp), dtype=torch.float16, device=x.buf.device)
    assert tuple(out.shape) == (2, rows, kh * kw * x.cp) and out[0].is_contiguous(), (tuple(out.shape), rows)
    _lib.check(_lib.load().glass_gather_taps(_ptr(x.hi), _ptr(x.lo), x.n, x.h, x.w, x.cp, x.border_code, kh, kw, sh, sw,
                                             ph, pw, ho, wo, _ptr(out[0]), _ptr(out[1]), dst_border, _ptr(n_dev),
                                             _stream()))
    return out


def stem_s2d(img: torch.Tensor, mean: Sequence[float], std: Sequence[float], out: torch.Tensor) -> torch.Tensor:
    """raw fp32 NCHW [n,3,h,w] -> normalised space-to-depth map, split planes [2, n, h/2+2b, w/2+2b, 16] (b-pixel zero
    border, b = 1 or 2 read off ``out``'s shape, which must already be zero in ``out``)."""
    n, c, h, w = img.shape
    assert c == 3 and img.dtype == torch.float32 and img.is_contiguous()
    border = (out.shape[2] - h // 2) // 2
    assert border in (1, 2) and tuple(out.shape) == (2, n, h // 2 + 2 * border, w // 2 + 2 * border, 16)
    assert out.dtype == torch.float16 and out.is_contiguous()
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[1.0 / float(v) for v in std])
    _lib.check(_lib.load().glass_stem_s2d(_ptr(img), n, h, w, m, s, _ptr(out[0]), _ptr(out[1]), border, _stream()))
    return out


def roi_align_rotated(feats: List, rois: torch.Tensor, output_size: Tuple[int, int],
                      scales: Sequence[float], sampling_ratio: int, min_level: int = 2,
                      out_f32: bool = True, out_split: Optional[Tuple[torch.Tensor, int, int, int, int, int]] = None,
                      n_rois_dev: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Multi-level rotated RoIAlign over fp32 maps (F32Map) or split-fp16 activations (Act).
    rois fp32 [R,6] (batch, cx, cy, w, h, angle).  Returns fp32
    [R, ph, pw, C] (NHWC order) when out_f32; out_split = (buf[2,...], hp, wp, border, coff, ld)."""
    assert rois.dtype == torch.float32 and rois.dim() == 2 and rois.shape[1] == 6 and rois.is_contiguous()
    p = _lib.RoiAlignParams()
    p.num_levels = len(feats)
    f0 = feats[0]
    split_in = isinstance(f0, Act)
    p.feat_is_split = int(split_in)
    for i, f in enumerate(feats):
        p.feat_h[i], p.feat_w[i], p.spatial_scale[i] = f.h, f.w, float(scales[i])
        if split_in:
            p.feat[i], p.feat_lo[i] = _ptr(f.hi), _ptr(f.lo)
        else:
            p.feat[i] = _ptr(f.buf)
    ld0 = f0.cp if split_in else f0.ld
    assert all((f.cp if split_in else f.ld) == ld0 and f.border == f0.border and f.c == f0.c for f in feats)
    p.feat_border, p.feat_ld, p.channels, p.min_level = f0.border, ld0, f0.c, min_level
    p.rois, p.n_rois = _ptr(rois), rois.shape[0]
    p.n_rois_dev = _ptr(n_rois_dev)
    p.pooled_h, p.pooled_w, p.sampling_ratio = output_size[0], output_size[1], sampling_ratio
    out = None
    if out_f32:
        out = torch.zeros((rois.shape[0], output_size[0], output_size[1], f0.c), dtype=torch.float32,
                          device=rois.device)
        p.out_f32 = _ptr(out)
    if out_split is not None:
        buf, hp, wp, border, coff, ld = out_split
        p.out_hi, p.out_lo = _ptr(buf[0]), _ptr(buf[1])
        p.out_hp, p.out_wp, p.out_border, p.out_coff, p.ld_out = hp, wp, border, coff, ld
    _lib.check(_lib.load().glass_roi_align_rotated(C.byref(p), _stream()))
    return out


def image_roi_align_rotated(img: torch.Tensor, pad_hw: Tuple[int, int], mean, std, rois: torch.Tensor,
                            output_size: Tuple[int, int], sampling_ratio: int, out_f32: bool = False,
                            out_act: Optional[Act] = None, n_rois_dev: Optional[torch.Tensor] = None,
                            workspace: Optional[torch.Tensor] = None):
    """``workspace``: optional uint8 / float32 CUDA scratch of n*h*w*16 bytes (see include/glass_b200.h)."""
    n, c, h, w = img.shape
    assert c == 3 and img.dtype == torch.float32 and img.is_contiguous()
    p = _lib.ImageRoiAlignParams()
    p.img, p.n, p.h, p.w, p.h_pad, p.w_pad = _ptr(img), n, h, w, pad_hw[0], pad_hw[1]
    for i in range(3):
        p.mean[i] = float(mean[i])
        p.inv_std[i] = 1.0 / float(std[i])
    p.rois, p.n_rois, p.n_rois_dev = _ptr(rois), rois.shape[0], _ptr(n_rois_dev)
    p.pooled_h, p.pooled_w, p.sampling_ratio = output_size[0], output_size[1], sampling_ratio
    out = None
    if out_f32:
        out = torch.zeros((rois.shape[0], 3, output_size[0], output_size[1]), dtype=torch.float32, device=img.device)
        p.out_f32 = _ptr(out)
    if out_act is not None:
        assert (out_act.h, out_act.w) == tuple(output_size) and out_act.n >= rois.shape[0]
        p.out_hi, p.out_lo, p.out_border, p.ld_out = _ptr(out_act.hi), _ptr(out_act.lo), out_act.border, out_act.cp
    if workspace is not None:
        p.workspace, p.workspace_bytes = _ptr(workspace), workspace.numel() * workspace.element_size()
    _lib.check(_lib.load().glass_image_roi_align_rotated(C.byref(p), _stream()))
    return out


# ------------------------------------------------------------------------------------------ detector decisions
def rpn_topk_decode(pred: torch.Tensor, num_anchors: int, stride: int, cell_anchors, weights: Sequence[float],
                    topk: int, level: int, num_levels: int, out_boxes: torch.Tensor, out_scores: torch.Tensor,
                    workspace: Optional[torch.Tensor] = None) -> None:
    """pred fp32 [n,h,w,ld] (fused RPN 1x1 output); cell_anchors: list of (w, h, angle) per anchor.
    Writes slots [level*topk, (level+1)*topk) of out_boxes [n, L*topk, 5] / out_scores [n, L*topk]."""
    n, h, w, ld = pred.shape
    assert pred.dtype == torch.float32 and pred.is_contiguous()
    p = _lib.RpnTopkParams()
    p.pred, p.n_img, p.h, p.w, p.ld, p.num_anchors, p.stride = _ptr(pred), n, h, w, ld, num_anchors, stride
    for a, (aw, ah, aa) in enumerate(cell_anchors):
        p.anchor_w[a], p.anchor_h[a], p.anchor_angle[a] = aw, ah, aa
    for j in range(5):
        p.weights[j] = float(weights[j])
    p.topk, p.level, p.num_levels = topk, level, num_levels
    p.out_boxes, p.out_scores = _ptr(out_boxes), _ptr(out_scores)
    need = _lib.load().glass_rpn_topk_workspace_bytes(n, h, w, num_anchors)
    if workspace is None:
        workspace = torch.empty((need,), dtype=torch.uint8, device=pred.device)
    p.workspace, p.workspace_bytes = _ptr(workspace), workspace.numel()
    _lib.check(_lib.load().glass_rpn_topk_decode(C.byref(p), _stream()))


def nms_rotated(boxes: torch.Tensor, scores: torch.Tensor, iou_thresh: float, max_keep: int,
                group: Optional[torch.Tensor] = None, group_size: int = 0, m_dev: Optional[torch.Tensor] = None,
                img_hw: Optional[torch.Tensor] = None, clip: bool = False, filter_empty: bool = False,
                score_thresh: float = float("-inf"), workspace: Optional[torch.Tensor] = None, out=None):
    """boxes fp32 [n,m,5], scores [n,m] -> (boxes [n,max_keep,5], scores [n,max_keep], index i32, count i32[n])."""
    n, m, _ = boxes.shape
    assert boxes.dtype == torch.float32 and boxes.is_contiguous() and scores.is_contiguous()
    dev = boxes.device
    L = _lib.load()
    need = L.glass_nms_workspace_bytes(n, m)
    if workspace is None:
        workspace = torch.empty((need,), dtype=torch.uint8, device=dev)
    if out is None:
        out = (torch.empty((n, max_keep, 5), dtype=torch.float32, device=dev),
               torch.empty((n, max_keep), dtype=torch.float32, device=dev),
               torch.empty((n, max_keep), dtype=torch.int32, device=dev),
               torch.empty((n,), dtype=torch.int32, device=dev))
    p = _lib.NmsParams()
    p.boxes, p.scores, p.group, p.group_size, p.m_dev = _ptr(boxes), _ptr(scores), _ptr(group), group_size, _ptr(m_dev)
    p.n_img, p.m, p.img_hw, p.clip, p.filter_empty = n, m, _ptr(img_hw), int(clip), int(filter_empty)
    p.score_thresh, p.iou_thresh, p.max_keep = score_thresh, iou_thresh, max_keep
    p.out_boxes, p.out_scores, p.out_index, p.out_count = (_ptr(t) for t in out)
    p.workspace, p.workspace_bytes = _ptr(workspace), workspace.numel()
    _lib.check(L.glass_nms_rotated(C.byref(p), _stream()))
    return out


def box_decode(pred: torch.Tensor, proposals: torch.Tensor, counts: Optional[torch.Tensor], n_img: int, per_img: int,
               weights: Sequence[float]):
    """pred fp32 [n_img*per_img, ld>=11]; proposals fp32 [n_img*per_img, 5] -> (boxes, scores, orientations)."""
    assert pred.dtype == torch.float32 and pred.is_contiguous() and proposals.is_contiguous()
    dev = pred.device
    boxes = torch.empty((n_img, per_img, 5), dtype=torch.float32, device=dev)
    scores = torch.empty((n_img, per_img), dtype=torch.float32, device=dev)
    orient = torch.empty((n_img, per_img, 2), dtype=torch.float32, device=dev)
    w = (C.c_float * 5)(*[float(v) for v in weights])
    _lib.check(_lib.load().glass_box_decode(_ptr(pred), pred.shape[1], _ptr(proposals), _ptr(counts), n_img, per_img, w,
                                            _ptr(boxes), _ptr(scores), _ptr(orient), _stream()))
    return boxes, scores, orient


# ------------------------------------------------------------------------------------------ recognizer head
def gc_attention(f: Act, y: Act, n_words: int, w, n_dev: Optional[torch.Tensor] = None) -> None:
    """MultiAspectGCAttention pooling + channel_add MLP + broadcast add (concat channel order); w: dict of
    fp32 device tensors w_mask[512], b_mask(float), w1t[512,256], b1, ln_g, ln_b, w2t[256,512], b2."""
    assert f.cp == 512 and y.cp == 512 and (f.h, f.w, f.border_code) == (y.h, y.w, y.border_code)
    p = _lib.GcAttentionParams()
    p.f_hi, p.f_lo, p.y_hi, p.y_lo = _ptr(f.hi), _ptr(f.lo), _ptr(y.hi), _ptr(y.lo)
    p.n_words, p.h, p.w, p.border, p.channels = n_words, f.h, f.w, f.border_code, 512
    p.w_mask, p.b_mask = _ptr(w["w_mask"]), float(w["b_mask"])
    p.w1t, p.b1, p.ln_g, p.ln_b, p.w2t, p.b2 = (_ptr(w[k]) for k in ("w1t", "b1", "ln_g", "ln_b", "w2t", "b2"))
    p.n_words_dev = _ptr(n_dev)
    _lib.check(_lib.load().glass_gc_attention(C.byref(p), _stream()))


def hmean_rows(x: Act, n: int, out: torch.Tensor, out_f32: Optional[torch.Tensor] = None,
               n_dev: Optional[torch.Tensor] = None) -> None:
    """mean over H: x [n,c,h,w] -> rows out [2, n*w, cp]."""
    assert tuple(out.shape) == (2, n * x.w, x.cp)
    _lib.check(_lib.load().glass_hmean_rows(_ptr(x.hi), _ptr(x.lo), n, x.h, x.w, x.cp, x.border_code, _ptr(out[0]),
                                            _ptr(out[1]), _ptr(out_f32), _ptr(n_dev), _stream()))


def lstm_bidir(gates_in: torch.Tensor, whh_t: torch.Tensor, n_seq: int, T: int, out: torch.Tensor,
               out_f32: Optional[torch.Tensor] = None, n_dev: Optional[torch.Tensor] = None) -> None:
    assert gates_in.dtype == torch.float32 and tuple(gates_in.shape) == (n_seq * T, 2048) and gates_in.is_contiguous()
    assert tuple(whh_t.shape) == (2, 256, 1024) and whh_t.is_contiguous() and tuple(out.shape) == (2, n_seq * T, 512)
    _lib.check(_lib.load().glass_lstm_bidir(_ptr(gates_in), _ptr(whh_t), n_seq, T, 256, _ptr(out[0]), _ptr(out[1]),
                                            _ptr(out_f32), _ptr(n_dev), _stream()))


def aster_decode(xproj: torch.Tensor, pctx: torch.Tensor, n_words: int, T: int, steps: int, num_classes: int, w,
                 probs: torch.Tensor, first_eos: torch.Tensor, logits: Optional[torch.Tensor] = None,
                 alphas: Optional[torch.Tensor] = None, n_dev: Optional[torch.Tensor] = None) -> None:
    """Greedy attention decoding.  xproj fp32 [n_words*T, 256] = xEmbed(x); pctx fp32 [n_words*T, 768] =
    x . W_ih[:, 256:]^T; w["emb_gi"] [num_classes, 768] = W_ih[:, :256] . Emb + b_ih (see include/glass_b200.h)."""
    assert xproj.dtype == torch.float32 and xproj.is_contiguous() and tuple(xproj.shape) == (n_words * T, 256)
    assert pctx.dtype == torch.float32 and tuple(pctx.shape) == (n_words * T, 768) and pctx.is_contiguous()
    assert tuple(w["emb_gi"].shape) == (num_classes, 768) and w["emb_gi"].is_contiguous()
    p = _lib.AsterParams()
    p.xproj, p.pctx, p.n_words, p.n_words_dev = _ptr(xproj), _ptr(pctx), n_words, _ptr(n_dev)
    p.T, p.steps, p.num_classes, p.dim = T, steps, num_classes, 256
    assert w["wh_frag"].dtype == torch.int32 and w["wh_frag"].numel() == 64 * 16 * 2 * 32 * 4 and w["bh"].numel() == 1024
    p.wh_frag, p.bh, p.we, p.be, p.emb_gi = _ptr(w["wh_frag"]), _ptr(w["bh"]), _ptr(w["we"]), float(w["be"]), _ptr(w["emb_gi"])
    p.wo_t, p.bo, p.temperature = _ptr(w["wo_t"]), _ptr(w["bo"]), float(w["temperature"])
    p.probs, p.logits, p.alphas, p.first_eos = _ptr(probs), _ptr(logits), _ptr(alphas), _ptr(first_eos)
    _lib.check(_lib.load().glass_aster_decode(C.byref(p), _stream()))


def aster_finalize(probs: torch.Tensor, first_eos: torch.Tensor, word_start: torch.Tensor, n_img: int, steps: int,
                   num_classes: int) -> None:
    assert word_start.dtype == torch.int32 and word_start.numel() == n_img + 1
    _lib.check(_lib.load().glass_aster_finalize(_ptr(probs), _ptr(first_eos), _ptr(word_start), n_img, steps,
                                                num_classes, _stream()))


def resize_bilinear_u8(img_hwc: torch.Tensor, out_hw: Tuple[int, int], flip_channels: bool = False) -> torch.Tensor:
    """uint8 HWC CUDA image -> fp32 CHW, bilinear (align_corners=False) resize to out_hw
    (GlassRunner._image_to_tensor, glass/inference/glass_runner.py:123-148)."""
    assert img_hwc.dtype == torch.uint8 and img_hwc.dim() == 3 and img_hwc.shape[2] == 3 and img_hwc.is_contiguous()
    h, w, _ = img_hwc.shape
    out = torch.empty((3, out_hw[0], out_hw[1]), dtype=torch.float32, device=img_hwc.device)
    _lib.check(_lib.load().glass_resize_bilinear_u8(_ptr(img_hwc), h, w, int(flip_channels), _ptr(out), out_hw[0],
                                                    out_hw[1], _stream()))
    return out


# ---------------------------------------------------------------------------------------------- detector -> recognizer
def pack_rois(boxes: torch.Tensor, counts: torch.Tensor, rois: torch.Tensor, word_start: torch.Tensor,
              total: torch.Tensor) -> None:
    """boxes fp32 [n, m, 5] + counts int32 [n] -> rois fp32 [n*m, 6], word_start int32 [n+1], total int32 [1]
    (glass_pack_rois: the detections become the recognizer's RoI list without leaving the device)."""
    n, m, _ = boxes.shape
    assert boxes.is_contiguous() and counts.dtype == torch.int32 and tuple(rois.shape) == (n * m, 6)
    assert word_start.dtype == torch.int32 and word_start.numel() == n + 1 and total.dtype == torch.int32
    _lib.check(_lib.load().glass_pack_rois(_ptr(boxes), _ptr(counts), n, m, _ptr(rois), _ptr(word_start), _ptr(total),
                                           _stream()))


def pack_detections(boxes: torch.Tensor, scores: torch.Tensor, orient: Optional[torch.Tensor], counts: torch.Tensor,
                    probs: torch.Tensor, word_start: torch.Tensor, rec: torch.Tensor) -> torch.Tensor:
    """-> rec fp32 [n, m, 10 + steps*classes] (glass_pack_detections), every row written (zeros past the counts)."""
    n, m, _ = boxes.shape
    steps, classes = probs.shape[1], probs.shape[2]
    assert tuple(rec.shape) == (n, m, 10 + steps * classes) and rec.is_contiguous() and probs.is_contiguous()
    assert boxes.is_contiguous() and scores.is_contiguous() and (orient is None or orient.is_contiguous())
    _lib.check(_lib.load().glass_pack_detections(_ptr(boxes), _ptr(scores), _ptr(orient), _ptr(counts), _ptr(probs),
                                                 _ptr(word_start), n, m, steps, classes, _ptr(rec), _stream()))
    return rec


def prepack_weights(w: torch.Tensor, scale: torch.Tensor):
    """fp32 CUDA matrix [n, k] + epilogue scale [n] -> (fp16 [2, n, k] hi/lo planes, scale with the pre-scales folded in)
    through glass_prepack_weights."""
    assert w.is_cuda and w.dtype == torch.float32 and w.dim() == 2 and w.is_contiguous()
    n, k = w.shape
    planes = torch.empty((2, n, k), dtype=torch.float16, device=w.device)
    scale = scale.to(device=w.device, dtype=torch.float32).contiguous().clone()
    _lib.check(_lib.load().glass_prepack_weights(_ptr(w), n, k, _ptr(planes[0]), _ptr(planes[1]), _ptr(scale), _stream()))
    return planes, scale


# ---------------------------------------------------------------------------------------------- post-processing
def text_scores(probs: torch.Tensor, stop_index: int = 1, want_steps: bool = False,
                n_dev: Optional[torch.Tensor] = None):
    """pred_text_prob fp32 [n, steps, classes] -> word score [n] (+ per-step argmax int32 and max prob [n, steps])."""
    assert probs.dim() == 3 and probs.dtype == torch.float32 and probs.is_contiguous()
    n, steps, classes = probs.shape
    score = torch.empty((n,), dtype=torch.float32, device=probs.device)
    idx = torch.empty((n, steps), dtype=torch.int32, device=probs.device) if want_steps else None
    maxp = torch.empty((n, steps), dtype=torch.float32, device=probs.device) if want_steps else None
    if n:
        _lib.check(_lib.load().glass_text_scores(_ptr(probs), n, steps, classes, stop_index, _ptr(score), _ptr(idx),
                                                 _ptr(maxp), _ptr(n_dev), _stream()))
    return (score, idx, maxp) if want_steps else score


def postprocess_merge(boxes: torch.Tensor, scores: torch.Tensor, counts: Optional[torch.Tensor] = None,
                      text_scores_: Optional[torch.Tensor] = None, min_box_dim: float = 2.0, valid_score: float = 0.15,
                      detect_threshold: float = 0.25, text_threshold: float = 0.25, merge_ioa_thresh: float = 0.3,
                      pairs_height_ratio_thresh: float = 0.35, max_angle_diff: float = 15.0,
                      minimal_ioa_thresh: float = 0.01, nms_iou: float = 0.99, max_iters: int = 1000,
                      want_polygons: bool = True):
    """Device-side PostProcessorRotatedBoxes / PostProcessorAcademic (see include/glass_b200.h).
    boxes fp32 [n_img, m, 5], scores [n_img, m], counts int32 [n_img] -> dict of padded outputs."""
    assert boxes.dim() == 3 and boxes.shape[2] == 5 and boxes.dtype == torch.float32 and boxes.is_contiguous()
    n_img, m = boxes.shape[0], boxes.shape[1]
    assert scores.shape == (n_img, m) and scores.dtype == torch.float32 and scores.is_contiguous()
    dev = boxes.device
    out = {"boxes": torch.empty((n_img, m, 5), dtype=torch.float32, device=dev),
           "scores": torch.empty((n_img, m), dtype=torch.float32, device=dev),
           "polygons": torch.empty((n_img, m, 4, 2), dtype=torch.float32, device=dev) if want_polygons else None,
           "index": torch.empty((n_img, m), dtype=torch.int32, device=dev),
           "count": torch.empty((n_img,), dtype=torch.int32, device=dev),
           "iters": torch.empty((n_img,), dtype=torch.int32, device=dev)}
    p = _lib.PostprocessParams()
    p.boxes, p.scores, p.text_scores, p.counts = _ptr(boxes), _ptr(scores), _ptr(text_scores_), _ptr(counts)
    if counts is not None:
        assert counts.dtype == torch.int32 and counts.numel() == n_img
    if text_scores_ is not None:
        assert text_scores_.shape == (n_img, m) and text_scores_.dtype == torch.float32 and text_scores_.is_contiguous()
    p.n_img, p.m = n_img, m
    p.min_box_dim, p.valid_score, p.detect_threshold, p.text_threshold = min_box_dim, valid_score, detect_threshold, text_threshold
    p.merge_ioa_thresh, p.pairs_height_ratio_thresh, p.max_angle_diff = merge_ioa_thresh, pairs_height_ratio_thresh, max_angle_diff
    p.minimal_ioa_thresh, p.nms_iou, p.max_iters = minimal_ioa_thresh, nms_iou, max_iters
    p.out_boxes, p.out_scores, p.out_polygons = _ptr(out["boxes"]), _ptr(out["scores"]), _ptr(out["polygons"])
    p.out_index, p.out_count, p.out_iters = _ptr(out["index"]), _ptr(out["count"]), _ptr(out["iters"])
    if n_img:
        _lib.check(_lib.load().glass_postprocess_merge(C.byref(p), _stream()))
    return out


# ---------------------------------------------------------------------------------------------- mask branch
def mask_finalize(logits: torch.Tensor, k_words: int, h: int, w: int, out: torch.Tensor) -> torch.Tensor:
    """logits fp32 [>= k*(h+2)*(w+2), ld] (columns dy*2+dx) -> out fp32 [k, 1, 2h, 2w] = sigmoid, sub-pixels scattered."""
    assert logits.dtype == torch.float32 and logits.is_contiguous() and logits.shape[0] >= k_words * (h + 2) * (w + 2)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() == k_words * 4 * h * w
    _lib.check(_lib.load().glass_mask_finalize(_ptr(logits), logits.shape[1], k_words, h, w, _ptr(out), _stream()))
    return out


def paste_masks_rotated(masks: torch.Tensor, boxes: torch.Tensor, image_shape: Tuple[int, int], threshold: float = 0.5,
                        want_soft: bool = False):
    """masks fp32 [k, m, m], boxes fp32 [k, 5] -> bool [k, H, W] (+ fp32 sampled values when want_soft)."""
    assert masks.dim() == 3 and masks.shape[1] == masks.shape[2] and masks.dtype == torch.float32 and masks.is_contiguous()
    assert boxes.shape == (masks.shape[0], 5) and boxes.dtype == torch.float32 and boxes.is_contiguous()
    k, m = masks.shape[0], masks.shape[1]
    h, w = int(image_shape[0]), int(image_shape[1])
    out = torch.empty((k, h, w), dtype=torch.uint8, device=masks.device)
    soft = torch.empty((k, h, w), dtype=torch.float32, device=masks.device) if want_soft else None
    if k:
        _lib.check(_lib.load().glass_paste_masks_rotated(_ptr(masks), _ptr(boxes), k, m, h, w, float(threshold), _ptr(out),
                                                         _ptr(soft), _stream()))
    return (out.bool(), soft) if want_soft else out.bool()
