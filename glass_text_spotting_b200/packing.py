"""Weight pre-packing: detectron2-style ``state_dict`` tensors -> the K-major split-fp16 layout the
tcgen05 GEMM consumes (SURVEY.md A.10 lists the key names).  Done once at load time."""
from typing import Optional, Tuple

import torch

from .ops import ACT_SCALE, PackedWeight, round_up


def split16(x: torch.Tensor) -> torch.Tensor:
    """fp32 tensor (already pre-scaled) -> fp16 [2, ...] (hi, lo) with x ~= hi + lo (22 mantissa bits)."""
    x = x.float().clamp(-60000.0, 60000.0)
    hi = x.to(torch.float16)
    lo = (x - hi.float()).to(torch.float16)
    return torch.stack((hi, lo), 0).contiguous()


def split_act(x: torch.Tensor) -> torch.Tensor:
    """An activation matrix in the library's storage convention: planes hold ACT_SCALE * x."""
    return split16(x.float() * ACT_SCALE)


def _row_shift(w: torch.Tensor) -> torch.Tensor:
    """Per-row power-of-two exponent s = floor(log2(512 / amax)) clamped to [-24, 24] (0 for an all-zero row), taken from
    the float's own exponent (frexp) so that the host rule and glass_prepack_weights agree bit for bit."""
    amax = w.abs().amax(dim=1)
    f, e = torch.frexp(amax.clamp_min(1e-37))          # amax = f * 2^e, f in [0.5, 1)
    s = torch.where(f == 0.5, 10 - e, 9 - e).clamp(-24, 24).to(torch.float32)
    return torch.where(amax > 0, s, torch.zeros_like(s))


def _pack_rows(w: torch.Tensor, scale: torch.Tensor, device="cpu"):
    """Split a weight matrix [n, k] with a per-row power-of-two pre-scale (rows land in [256, 512) so
    both planes stay in fp16's normal range) and fold 2^-s / ACT_SCALE into the epilogue scale.  On a CUDA device the
    packing itself runs through the C ABI (glass_prepack_weights); the host path below is its restatement (CPU-side
    tests, tools) and produces the same bits."""
    if torch.device(device).type == "cuda":
        from . import ops
        return ops.prepack_weights(w.float().contiguous().to(device), scale)
    s = _row_shift(w)
    return split16(w * torch.exp2(s).unsqueeze(1)), scale * torch.exp2(-s) / ACT_SCALE


def fold_bn(bn_w, bn_b, bn_mean, bn_var, eps: float = 1e-5, conv_bias: Optional[torch.Tensor] = None):
    """eval-mode BatchNorm after a conv -> per-channel (scale, bias) applied in the GEMM epilogue."""
    scale = bn_w.double() / torch.sqrt(bn_var.double() + eps)
    bias = bn_b.double() - bn_mean.double() * scale
    if conv_bias is not None:
        bias = bias + conv_bias.double() * scale
    return scale.float(), bias.float()


def pack_conv(weight: torch.Tensor, scale: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
              stride: Tuple[int, int] = (1, 1), pad: Tuple[int, int] = (0, 0), cin_p: Optional[int] = None,
              n_align: int = 64, device="cuda") -> PackedWeight:
    """weight fp32 [cout, cin, kh, kw] -> [2, n_p, kh*kw*cin_p] fp16 (K = tap-major, channel-minor)."""
    cout, cin, kh, kw = weight.shape
    cin_p = cin_p if cin_p is not None else round_up(cin, 64)
    n_p = round_up(cout, n_align)
    w = torch.zeros((n_p, kh * kw, cin_p), dtype=torch.float32)
    w[:cout, :, :cin] = weight.detach().float().permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    s = torch.ones(n_p, dtype=torch.float32)
    b = torch.zeros(n_p, dtype=torch.float32)
    if scale is not None:
        s[:cout] = scale.detach().float()
    if bias is not None:
        b[:cout] = bias.detach().float()
    wp, s = _pack_rows(w.reshape(n_p, kh * kw * cin_p), s, device)
    return PackedWeight(wp.to(device), s.to(device), b.to(device), cout, cin, kh, kw, tuple(stride), tuple(pad), cin_p)


def pack_conv_compact(weight: torch.Tensor, cp_in: int, scale: Optional[torch.Tensor] = None,
                      bias: Optional[torch.Tensor] = None, n_align: int = 16, device="cuda") -> PackedWeight:
    """stride-1 'same' conv (1x1 or 3x3) whose INPUT activation is narrow (cp_in = 8/16/32 channels per pixel).
    K layout per tap-row r: nj k-blocks of 64 = (64/cp_in pixels) x cp_in channels starting at pixel x - pad;
    entries beyond the kw taps / cin channels are zero (see ops._conv2d_compact)."""
    cout, cin, kh, kw = weight.shape
    assert cp_in in (8, 16, 32) and cin <= cp_in and (kh, kw) in ((1, 1), (3, 3))
    ppk = 64 // cp_in
    nj = (kw + ppk - 1) // ppk
    n_p = round_up(cout, n_align)
    w = torch.zeros((n_p, kh, nj * ppk, cp_in), dtype=torch.float32)
    w[:cout, :, :kw, :cin] = weight.detach().float().permute(0, 2, 3, 1)
    s = torch.ones(n_p, dtype=torch.float32)
    b = torch.zeros(n_p, dtype=torch.float32)
    if scale is not None:
        s[:cout] = scale.detach().float()
    if bias is not None:
        b[:cout] = bias.detach().float()
    wp, s = _pack_rows(w.reshape(n_p, kh * nj * 64), s, device)
    pw = PackedWeight(wp.to(device), s.to(device), b.to(device), cout, cin, kh, kw, (1, 1), ((kh - 1) // 2, (kw - 1) // 2),
                      64)
    pw.compact_cp = cp_in
    return pw


def pack_conv_grouped(weight: torch.Tensor, cp_in: int, group: int, scale: Optional[torch.Tensor] = None,
                      bias: Optional[torch.Tensor] = None, device="cuda", n_align: int = 16) -> PackedWeight:
    """Pixel-grouped packing of a stride-1 'same' conv (1x1 or 3x3) on a narrow activation (cp_in channels per pixel):
    one GEMM row = ``group`` consecutive pixels, K per tap row = the window of group + kw - 1 pixels (zero padded to a
    multiple of 64), weights = the block-Toeplitz matrix [group * cout_p, kh * Kwin] whose row (p, co) holds
    weight[co, :, r, s] at window pixel p + s.  Output row = group * cout_p contiguous elements = the NHWC rows of the
    group's pixels, so cout_p (cout rounded up to 16) must be the output activation's channel stride.  ``n_align`` pads
    the weight rows with zeros (5 pixels x 16 channels = 80 -> 128: a tile width the TMA-store epilogue takes; the store
    clips the pad columns, ``grouped_n_store`` = the real row width)."""
    cout, cin, kh, kw = weight.shape
    assert cin <= cp_in and cp_in % 8 == 0 and (kh, kw) in ((1, 1), (3, 3)) and group >= 1
    cout_p = round_up(cout, 16)
    win = group + kw - 1
    kwin = round_up(win * cp_in, 64)
    assert kwin % cp_in == 0
    w = torch.zeros((group, cout_p, kh, kwin // cp_in, cp_in), dtype=torch.float32)
    wt = weight.detach().float().permute(0, 2, 3, 1)  # [cout, kh, kw, cin]
    for p in range(group):
        for s_ in range(kw):
            w[p, :cout, :, p + s_, :cin] = wt[:, :, s_, :]
    n_real = group * cout_p
    n_p = round_up(n_real, n_align)
    s = torch.ones(cout_p, dtype=torch.float32)
    b = torch.zeros(cout_p, dtype=torch.float32)
    if scale is not None:
        s[:cout] = scale.detach().float()
    if bias is not None:
        b[:cout] = bias.detach().float()
    rows = torch.zeros((n_p, kh * kwin), dtype=torch.float32)
    rows[:n_real] = w.reshape(n_real, kh * kwin)
    s_rows, b_rows = torch.ones(n_p, dtype=torch.float32), torch.zeros(n_p, dtype=torch.float32)
    s_rows[:n_real], b_rows[:n_real] = s.repeat(group), b.repeat(group)
    wp, s_all = _pack_rows(rows, s_rows, device)
    pw = PackedWeight(wp.to(device), s_all.to(device), b_rows.to(device), cout, cin, kh, kw, (1, 1),
                      ((kh - 1) // 2, (kw - 1) // 2), kwin)
    pw.grouped_p, pw.grouped_cp, pw.grouped_cout_p, pw.grouped_n_store = group, cp_in, cout_p, n_real
    return pw


def pack_linear(weight: torch.Tensor, bias: Optional[torch.Tensor] = None, k_p: Optional[int] = None,
                n_align: int = 64, device="cuda") -> PackedWeight:
    """nn.Linear weight [out, in] -> packed as a 1x1 'conv' over rows."""
    out_f, in_f = weight.shape
    k_p = k_p if k_p is not None else round_up(in_f, 64)
    w = torch.zeros((round_up(out_f, n_align), k_p), dtype=torch.float32)
    w[:out_f, :in_f] = weight.detach().float()
    s = torch.ones(w.shape[0], dtype=torch.float32)
    b = torch.zeros(w.shape[0], dtype=torch.float32)
    if bias is not None:
        b[:out_f] = bias.detach().float()
    wp, s = _pack_rows(w, s, device)
    return PackedWeight(wp.to(device), s.to(device), b.to(device), out_f, in_f, 1, 1, (1, 1), (0, 0), k_p)


def pack_stem_s2d(weight: torch.Tensor, scale, bias, device="cuda") -> PackedWeight:
    """BasicStem conv1 [64,3,7,7] (stride 2, pad 3) as a 4x4 stride-1 conv over the space-to-depth map of
    ops.stem_s2d: K = (i*4 + q)*16 + (dy*2+dx)*3 + c with w[co, c, 2i+dy-1, 2q+dx-1] (zero outside the 7x7 support);
    i, q index the s2d rows / columns Y-2..Y+1, X-2..X+1."""
    cout = weight.shape[0]
    wt = weight.detach().float()
    w = torch.zeros((round_up(cout, 64), 4, 4, 16), dtype=torch.float32)
    for i in range(4):
        for dy in range(2):
            r = 2 * i + dy - 1
            if not 0 <= r < 7:
                continue
            for q in range(4):
                for dx in range(2):
                    s_ = 2 * q + dx - 1
                    if 0 <= s_ < 7:
                        base = (dy * 2 + dx) * 3
                        w[:cout, i, q, base:base + 3] = wt[:, :, r, s_]
    s = torch.ones(w.shape[0])
    b = torch.zeros(w.shape[0])
    s[:cout] = scale.detach().float()
    b[:cout] = bias.detach().float()
    wp, s = _pack_rows(w.reshape(w.shape[0], 256), s, device)
    return PackedWeight(wp.to(device), s.to(device), b.to(device), cout, 3, 7, 7, (2, 2), (3, 3), 64)


DEC_SW = 64.0   # == DEC_SW in csrc/recognizer.cu: power-of-two pre-scale of the decoder's fp16-split weights


def pack_decoder_h_weights(ws: torch.Tensor, bs: torch.Tensor, whh: torch.Tensor, bhh: torch.Tensor, device="cuda"):
    """The decoder's products with the hidden state -- sEmbed [256, 256] and GRU W_hh [768, 256] (nn.Linear / nn.GRU
    layout [out, in]) -- as ONE 1024 x 256 matrix in the fragment order of mma.sync.m16n8k16's A operand, split into fp16
    hi / lo planes of DEC_SW * w (include/glass_b200.h, GlassAsterParams.wh_frag).  Returns (int32 [64,16,2,32,4], fp32
    bias [1024])."""
    w = torch.cat((ws.detach().float(), whh.detach().float()), 0)            # [1024, 256]
    assert tuple(w.shape) == (1024, 256)
    planes = split16(w * DEC_SW)                                             # fp16 [2, 1024, 256]
    # dims: plane, m-tile (64), r2 (row g / g + 8), g (8), k-step (16), c2 (k 2q.. / 2q + 8..), q (4), e (2 halfs)
    f = planes.view(2, 64, 2, 8, 16, 2, 4, 2).permute(1, 4, 0, 3, 6, 5, 2, 7).contiguous()   # mt, ks, plane, g, q, c2, r2, e
    frag = f.view(64, 16, 2, 32, 4, 2).view(torch.int32).view(64, 16, 2, 32, 4)
    bias = torch.cat((bs.detach().float(), bhh.detach().float()), 0)
    return frag.contiguous().to(device), bias.contiguous().to(device)
