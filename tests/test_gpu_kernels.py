"""GPU parity tests of the individual kernels, called through the C ABI (ctypes) and compared with
the CPU oracle (oracle/d2_ops.c for the detectron2 operators; torch fp32 CPU conv/linear -- which is
what the oracle's nets are made of -- for the tcgen05 GEMM).  Tolerance for floating point is the
north-star's rtol=1e-3 / atol=1e-4 (relative to the tensor's scale)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-3, 1e-4


def _close(got: torch.Tensor, ref: torch.Tensor, name=""):
    got, ref = got.detach().cpu().float(), ref.detach().cpu().float()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    scale = max(ref.abs().max().item(), 1e-6)
    err = (got - ref).abs()
    tol = ATOL * max(scale, 1.0) + RTOL * ref.abs()
    bad = (err > tol).sum().item()
    assert bad == 0, f"{name}: {bad}/{err.numel()} out of tolerance, max err {err.max().item():.3e}, scale {scale:.3e}"


@pytest.fixture(scope="module")
def ops(glass_lib):
    from glass_text_spotting_b200 import ops
    return ops


def test_pack_unpack_roundtrip(ops):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 70, 9, 13, generator=g) * 3
    a = ops.Act.from_nchw(x.cuda())
    assert a.cp == 128
    y = a.to_nchw().cpu()
    assert (y - x).abs().max().item() <= 2 ** -21 * x.abs().max().item()
    # borders and pad channels stay zero
    assert a.buf[:, :, 0].abs().max().item() == 0 and a.buf[..., 70:].abs().max().item() == 0


def test_maxpool_matches_torch(ops):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 64, 16, 32, generator=g)
    a = ops.Act.from_nchw(x.cuda())
    xs = a.to_nchw().cpu()  # the split representation is the kernel's input
    for k, s, p in [((3, 3), (2, 2), (1, 1)), ((2, 2), (2, 2), (0, 0)), ((2, 2), (2, 1), (0, 1))]:
        got = ops.maxpool2d(a, k, s, p).to_nchw().cpu()
        ref = F.max_pool2d(xs, k, s, p)
        assert torch.equal(got, ref), (k, s, p)


@pytest.mark.parametrize("sb", [1, 2])
@pytest.mark.parametrize("shape", [(2, 64, 96), (1, 32, 160), (3, 96, 64)])
def test_stem_space_to_depth(ops, shape, sb):
    """The 7x7/s2/p3 stem as a 4x4/s1 conv over the normalised space-to-depth map (no im2col matrix), with a
    non-trivial std and folded scale / bias; borders of the map stay zero.  Border 1 (the product's layout: the map has the
    output plane's geometry, flat GEMM, TMA stores; the taps' 2-pixel reach wraps onto neighbouring border cells) and
    border 2 (direct stores) must agree bit for bit -- checked through the same reference."""
    from glass_text_spotting_b200 import packing
    n, h, w_ = shape
    g = torch.Generator().manual_seed(n * 1000 + h)
    img = torch.randint(0, 256, (n, 3, h, w_), generator=g).float()
    mean, std = (103.53, 116.28, 123.675), (57.375, 57.12, 58.395)
    wt = torch.randn(64, 3, 7, 7, generator=g) * 0.05
    scale = 1.0 + 0.1 * torch.randn(64, generator=g)
    bias = 0.1 * torch.randn(64, generator=g)
    hp2, wp2 = h // 2 + 2 * sb, w_ // 2 + 2 * sb
    s2d = torch.zeros((2, n, hp2, wp2, 16), dtype=torch.float16, device="cuda")
    ops.stem_s2d(img.cuda(), mean, std, out=s2d)
    assert s2d[:, :, :sb].abs().max().item() == 0 and s2d[:, :, :, -sb:].abs().max().item() == 0
    assert s2d[..., 12:].abs().max().item() == 0
    pw = packing.pack_stem_s2d(wt, scale, bias)
    out = ops.Act(n, 64, h // 2, w_ // 2)
    ops.conv_gemm(s2d[0], s2d[1], n * hp2 * wp2, 64, [(i - 2) * wp2 - 2 for i in range(4)], pw, (n, hp2, wp2, sb),
                  out=out, relu_post=True, a_ld=16)
    x = (img - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    ref = F.relu(F.conv2d(x, wt, stride=2, padding=3) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1))
    _close(out.to_nchw(), ref, "stem_s2d")
    assert out.buf[:, :, 0].abs().max().item() == 0 and out.buf[:, :, :, -1].abs().max().item() == 0


CONV_CASES = [
    # name, n, cin, cout, h, w, k, stride, pad, relu, residual
    ("1x1_64_256", 2, 64, 256, 24, 40, 1, 1, 0, False, False),
    ("3x3_64_64_relu", 2, 64, 64, 24, 40, 3, 1, 1, True, False),
    ("3x3_128_128_res_relu", 1, 128, 128, 33, 16, 3, 1, 1, True, True),
    ("1x1_256_512_s2", 2, 256, 512, 32, 32, 1, 2, 0, False, False),
    ("3x3_256_256_big", 3, 256, 256, 64, 64, 3, 1, 1, True, True),
    ("1x1_512_2048", 1, 512, 2048, 16, 16, 1, 1, 0, True, False),
    ("3x3_16_32_padded_channels", 2, 16, 32, 20, 20, 3, 1, 1, True, False),
    ("2x2_s21_256", 2, 256, 256, 16, 33, 2, (2, 1), 0, True, False),
    ("2x1_s21_256", 2, 256, 256, 8, 32, (2, 1), (2, 1), 0, True, False),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_gemm_matches_torch_fp32(ops, case):
    from glass_text_spotting_b200 import packing
    name, n, cin, cout, h, w, k, stride, pad, relu, use_res = case
    k = (k, k) if isinstance(k, int) else k
    stride = (stride, stride) if isinstance(stride, int) else stride
    pad = (pad, pad) if isinstance(pad, int) else pad
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, *k, generator=g) / math.sqrt(cin * k[0] * k[1])
    scale = 1.0 + 0.1 * torch.randn(cout, generator=g)
    bias = 0.1 * torch.randn(cout, generator=g)
    a = ops.Act.from_nchw(x.cuda())
    pw = packing.pack_conv(wt, scale, bias, stride, pad)
    ref = F.conv2d(x, wt, stride=stride, padding=pad) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    res_act = None
    if use_res:
        r = torch.randn(ref.shape, generator=g)
        res_act = ops.Act.from_nchw(r.cuda())
        ref = ref + r
    if relu:
        ref = F.relu(ref)
    f32 = ops.F32Map(n, cout, ref.shape[2], ref.shape[3], ld=pw.n_p)
    out = ops.conv2d(a, pw, relu=relu, residual=res_act, f32=f32)
    _close(out.to_nchw(), ref, name + "/split")
    _close(f32.to_nchw(), ref, name + "/f32")
    # zero border is never written
    assert out.buf[:, :, 0].abs().max().item() == 0 and out.buf[:, :, :, 0].abs().max().item() == 0
    assert f32.buf[:, 0].abs().max().item() == 0


COMPACT_CASES = [
    # name, cin, cp_in, cout, k
    ("c3_cp8_3x3", 3, 8, 16, 3), ("c16_cp16_3x3", 16, 16, 32, 3), ("c32_cp32_3x3", 32, 32, 64, 3),
    ("c32_cp32_1x1", 32, 32, 64, 1), ("c5_cp8_1x1", 5, 8, 48, 1),
]


@pytest.mark.parametrize("case", COMPACT_CASES, ids=[c[0] for c in COMPACT_CASES])
def test_conv_gemm_compact_channels(ops, case):
    """Narrow activations (8/16/32 channels per pixel): overlapping-row TMA, 64/cp pixels per k-block."""
    from glass_text_spotting_b200 import packing
    name, cin, cp_in, cout, k = case
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    x = torch.randn(3, cin, 21, 37, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    scale = 1.0 + 0.1 * torch.randn(cout, generator=g)
    bias = 0.1 * torch.randn(cout, generator=g)
    a = ops.Act.from_nchw(x.cuda(), cp=cp_in)
    pw = packing.pack_conv_compact(wt, cp_in, scale, bias)
    out = ops.conv2d(a, pw, relu=True)
    ref = F.relu(F.conv2d(x, wt, padding=k // 2) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1))
    assert out.cp == pw.n_p and out.cp % 16 == 0
    _close(out.to_nchw(), ref, name)
    assert out.buf[:, :, 0].abs().max().item() == 0 and out.buf[:, :, :, -1].abs().max().item() == 0


@pytest.mark.parametrize("case", [("g3x3_c3_P5", 3, 8, 16, 3, 5, (18, 33)), ("g3x3_c16_P2", 16, 16, 32, 3, 2, (20, 38)),
                                  ("g3x3_c32_P2", 32, 32, 64, 3, 2, (14, 22)), ("g1x1_c32_P2", 32, 32, 64, 1, 2, (14, 22)),
                                  ("g3x3_c3_P1", 3, 8, 16, 3, 1, (9, 11))], ids=lambda c: c[0])
def test_conv_gemm_pixel_grouped(ops, case):
    """Pixel-grouped mode: P pixels per GEMM row, block-Toeplitz weights, border re-zeroed afterwards.  Widths are
    chosen so that the padded width w + 2 is a multiple of P."""
    from glass_text_spotting_b200 import packing
    name, cin, cp_in, cout, k, P, (h, w_) = case
    assert (w_ + 2) % P == 0
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    x = torch.randn(3, cin, h, w_, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    scale = 1.0 + 0.1 * torch.randn(cout, generator=g)
    bias = 0.1 * torch.randn(cout, generator=g)
    a = ops.Act.from_nchw(x.cuda(), cp=cp_in)
    pw = packing.pack_conv_grouped(wt, cp_in, P, scale, bias)
    out = ops.conv2d(a, pw, relu=True)
    ref = F.relu(F.conv2d(x, wt, padding=k // 2) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1))
    assert out.cp == pw.grouped_cout_p
    _close(out.to_nchw(), ref, name)
    b = out.buf
    assert b[:, :, 0].abs().max().item() == 0 and b[:, :, -1].abs().max().item() == 0
    assert b[:, :, :, 0].abs().max().item() == 0 and b[:, :, :, -1].abs().max().item() == 0


@pytest.mark.parametrize("res", [False, True])
def test_conv_gemm_pixel_pairs_64_channels(ops, res):
    """The 64 -> 64 channel 3x3 convs as pixel-pair GEMMs (N = 128 tiles), incl. the residual read through the same grouped
    row view and the device-side word count."""
    from glass_text_spotting_b200 import packing
    g = torch.Generator().manual_seed(64 + res)
    x = torch.randn(5, 64, 14, 22, generator=g)
    r = torch.randn(5, 64, 14, 22, generator=g)
    wt = torch.randn(64, 64, 3, 3, generator=g) / math.sqrt(64 * 9)
    scale = 1.0 + 0.1 * torch.randn(64, generator=g)
    bias = 0.1 * torch.randn(64, generator=g)
    pw = packing.pack_conv_grouped(wt, 64, 2, scale, bias)
    pw.fallback = packing.pack_conv(wt, scale, bias, (1, 1), (1, 1))
    a = ops.Act.from_nchw(x.cuda())
    ra = ops.Act.from_nchw(r.cuda()) if res else None
    out = ops.conv2d(a, pw, relu=True, residual=ra)
    ref = F.conv2d(x, wt, padding=1) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    ref = F.relu(ref + r) if res else F.relu(ref)
    _close(out.to_nchw(), ref, f"pairs64 res={res}")
    b = out.buf
    assert b[:, :, 0].abs().max().item() == 0 and b[:, :, -1].abs().max().item() == 0
    assert b[:, :, :, 0].abs().max().item() == 0 and b[:, :, :, -1].abs().max().item() == 0
    # same launch sized for 5 words with 3 live ones
    out2 = ops.Act(5, 64, 14, 22)
    ops.conv2d(a, pw, relu=True, residual=ra, out=out2, n_dev=torch.tensor([3], dtype=torch.int32).cuda())
    assert torch.equal(out2.to_nchw()[:3], out.to_nchw()[:3])
    # odd padded width: falls back to the plain tap-row GEMM
    x3 = torch.randn(1, 64, 9, 11, generator=g)
    o3 = ops.conv2d(ops.Act.from_nchw(x3.cuda()), pw, relu=True)
    _close(o3.to_nchw(), F.relu(F.conv2d(x3, wt, padding=1) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)), "fallback")


@pytest.mark.parametrize("hw", [(16, 33), (8, 32), (5, 7)])
def test_shared_border_planes_equal_padded_planes(ops, hw):
    """Shared-border planes (GLASS_BORDER_SHARED: one leading zero row / column per plane, the next row / plane provides
    the trailing one) give the same bits as fully padded planes through every kernel that addresses a plane: pack / unpack,
    3x3 conv (+ residual, both epilogues), max-pool, tap gather, GC attention, H-mean."""
    from glass_text_spotting_b200 import packing
    g = torch.Generator().manual_seed(hw[0] * 100 + hw[1])
    n = 5
    x = torch.randn(n, 256, *hw, generator=g)
    r = torch.randn(n, 256, *hw, generator=g)
    w = packing.pack_conv(torch.randn(256, 256, 3, 3, generator=g) * 0.02, torch.rand(256, generator=g) + 0.5,
                          torch.randn(256, generator=g), (1, 1), (1, 1))
    outs = {}
    for shared in (False, True):
        a, ra = ops.Act.from_nchw(x.cuda(), shared=shared), ops.Act.from_nchw(r.cuda(), shared=shared)
        assert torch.equal(a.to_nchw().cpu(), ops.Act.from_nchw(x.cuda()).to_nchw().cpu())
        y = ops.conv2d(a, w, relu=True, residual=ra)
        assert y.shared == shared
        res = [y.to_nchw()]
        b = y.border
        assert y.buf[:, :, :b].abs().max().item() == 0 and y.buf[:, :, :, :b].abs().max().item() == 0   # leading borders
        if shared:
            assert y.buf[:, n].abs().max().item() == 0                                                  # the extra plane
        if hw[0] % 2 == 0:
            res.append(ops.maxpool2d(y, (2, 2), (2, 1), (0, 1)).to_nchw())
            gt = ops.gather_taps(y, 2, 2, 2, 1, 0, 0, (hw[0] - 2) // 2 + 1, hw[1] - 1)
            res.append(gt.clone())
        seq = torch.empty((2, n * hw[1], 256), dtype=torch.float16, device="cuda")
        ops.hmean_rows(y, n, seq)
        res.append(seq)
        outs[shared] = res
    for u, v in zip(outs[False], outs[True]):
        assert torch.equal(u, v)


def test_conv_gemm_fast_mode_is_fp16_grade(ops):
    from glass_text_spotting_b200 import packing
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 128, 16, 16, generator=g)
    wt = torch.randn(128, 128, 3, 3, generator=g) / math.sqrt(128 * 9)
    a = ops.Act.from_nchw(x.cuda())
    pw = packing.pack_conv(wt, None, None, (1, 1), (1, 1))
    out = ops.conv2d(a, pw, mode=ops.MODE_FAST).to_nchw().cpu()
    ref = F.conv2d(x, wt, padding=1)
    rel = (out - ref).norm() / ref.norm()
    assert 1e-5 < rel < 3e-3, rel


def test_fpn_upsample_residual(ops):
    """lateral 1x1 conv + nearest-2x upsampled coarser map, fused in the epilogue (d2 FPN top-down)."""
    from glass_text_spotting_b200 import packing
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 128, 16, 24, generator=g)
    top = torch.randn(2, 64, 8, 12, generator=g)
    wt = torch.randn(64, 128, 1, 1, generator=g) / math.sqrt(128)
    a, t = ops.Act.from_nchw(x.cuda()), ops.Act.from_nchw(top.cuda())
    pw = packing.pack_conv(wt)
    out = ops.conv2d(a, pw, residual=t, res_shift=1)
    ref = F.conv2d(x, wt) + F.interpolate(top, scale_factor=2.0, mode="nearest")
    _close(out.to_nchw(), ref, "fpn_topdown")


def test_linear_gemm(ops):
    from glass_text_spotting_b200 import packing
    g = torch.Generator().manual_seed(7)
    a = torch.randn(100, 12544, generator=g)
    w = torch.randn(2048, 12544, generator=g) / math.sqrt(12544)
    b = torch.randn(2048, generator=g) * 0.1
    pw = packing.pack_linear(w, b)
    o, of = ops.linear(packing.split_act(a).cuda(), pw, relu=True, want_f32=True)
    ref = F.relu(a @ w.t() + b)
    _close(of, ref, "fc1/f32")
    _close((o[0].float() + o[1].float()) / ops.ACT_SCALE, ref, "fc1/split")


@pytest.mark.parametrize("n_seq", [100, 37, 21, 1])
def test_lstm_bidir_matches_torch(ops, n_seq):
    """The persistent recurrent kernel against torch.nn.LSTM(bidirectional) given the same input projections;
    T = 32, H = 256 as in BiLSTMBlockV2 (recognizer_encoder.py:136-144)."""
    mode = "stream"
    T, H = 32, 256
    g = torch.Generator().manual_seed(n_seq)
    lstm = torch.nn.LSTM(H, H, bidirectional=True, batch_first=True)
    with torch.no_grad():
        for p_ in lstm.parameters():
            p_.copy_(torch.randn(p_.shape, generator=g) * 0.06)
    x = torch.randn(n_seq, T, H, generator=g)
    with torch.no_grad():
        ref, _ = lstm(x)
        gates = torch.cat([x @ lstm.weight_ih_l0.t() + lstm.bias_ih_l0 + lstm.bias_hh_l0,
                           x @ lstm.weight_ih_l0_reverse.t() + lstm.bias_ih_l0_reverse + lstm.bias_hh_l0_reverse], 2)
        whh_t = torch.stack((lstm.weight_hh_l0.t().contiguous(), lstm.weight_hh_l0_reverse.t().contiguous()), 0)
    out = torch.zeros((2, n_seq * T, 2 * H), dtype=torch.float16, device="cuda")
    f32 = torch.zeros((n_seq * T, 2 * H), dtype=torch.float32, device="cuda")
    ops.lstm_bidir(gates.reshape(n_seq * T, 8 * H).contiguous().cuda(), whh_t.contiguous().cuda(), n_seq, T, out, f32)
    _close(f32.view(n_seq, T, 2 * H), ref, f"lstm/{mode}/f32")
    _close(((out[0].float() + out[1].float()) / ops.ACT_SCALE).view(n_seq, T, 2 * H), ref, f"lstm/{mode}/split")


# ------------------------------------------------------------------------------ rotated RoIAlign
def _random_rois(g, n, img=1024.0, batch=1):
    cx = torch.rand(n, generator=g) * img
    cy = torch.rand(n, generator=g) * img
    w = torch.exp(torch.rand(n, generator=g) * (math.log(512) - math.log(16)) + math.log(16))
    h = w * (0.1 + 0.9 * torch.rand(n, generator=g))
    a = torch.rand(n, generator=g) * 360 - 180
    b = torch.randint(0, batch, (n,), generator=g).float()
    return torch.stack((b, cx, cy, w, h, a), 1).contiguous()


def _f32map_from_nchw(ops, x):
    n, c, h, w = x.shape
    f = ops.F32Map(n, c, h, w, border=1, ld=c)
    f.buf[:, 1:-1, 1:-1, :] = x.permute(0, 2, 3, 1).cuda()
    return f


def test_roi_align_rotated_d2_kat(ops):
    """detectron2 tests/layers/test_roi_align_rotated.py::test_forward_output_0_90_180_270."""
    x = torch.arange(25, dtype=torch.float32).reshape(1, 1, 5, 5).repeat(1, 4, 1, 1)
    f = _f32map_from_nchw(ops, x)
    base = torch.tensor([[4.5, 5.0, 5.5, 6.0], [7.0, 7.5, 8.0, 8.5], [9.5, 10.0, 10.5, 11.0], [12.0, 12.5, 13.0, 13.5]])
    for i in range(4):
        rois = torch.tensor([[0, 2, 2, 2, 2, 90.0 * i]], dtype=torch.float32).cuda()
        out = ops.roi_align_rotated([f], rois, (4, 4), [1.0], 0).cpu()  # [1,4,4,C]
        want = torch.rot90(base, -i)
        for c in range(4):
            assert torch.allclose(out[0, :, :, c], want, atol=1e-5), (i, c)


@pytest.mark.parametrize("cfg", ["box_pooler", "recog_pooler"])
def test_roi_align_rotated_matches_oracle(ops, cfg):
    from oracle import d2_ops
    g = torch.Generator().manual_seed(11)
    if cfg == "box_pooler":
        sizes, scales, out_size, sampling, n = [64, 32, 16, 8, 4], [1 / 4, 1 / 8, 1 / 16, 1 / 32, 1 / 64], (7, 7), 2, 96
        feats = [torch.randn(2, 256, s, s, generator=g) for s in sizes]
        rois = _random_rois(g, n, img=256.0, batch=2)
        rois[:, 3:5] *= 0.6
    else:
        scales, out_size, sampling, n = [1 / 4], (8, 32), 0, 24
        feats = [torch.randn(1, 256, 64, 64, generator=g)]
        rois = _random_rois(g, n, img=256.0, batch=1)
        rois[:, 3:5] *= 0.5
    ref = d2_ops.roi_pooler(feats, [rois[rois[:, 0] == b][:, 1:] for b in range(feats[0].shape[0])], out_size,
                            scales, sampling)
    order = torch.cat([torch.nonzero(rois[:, 0] == b).squeeze(1) for b in range(feats[0].shape[0])])
    maps = [_f32map_from_nchw(ops, f) for f in feats]
    got = ops.roi_align_rotated(maps, rois[order].contiguous().cuda(), out_size, scales, sampling)
    _close(got.permute(0, 3, 1, 2), ref, cfg)


@pytest.mark.parametrize("cfg", ["box_pooler", "recog_pooler", "narrow"])
def test_roi_align_rotated_split_input_matches_oracle(ops, cfg):
    """The production path: split-fp16 feature maps in (8 channels per lane), fp32 and split rows out.  The oracle
    pools the values the split planes represent (Act.to_nchw), so the only differences are summation order."""
    from oracle import d2_ops
    g = torch.Generator().manual_seed(21)
    if cfg == "box_pooler":
        sizes, scales, out_size, sampling, n, c = [64, 32, 16, 8, 4], [1 / 4, 1 / 8, 1 / 16, 1 / 32, 1 / 64], (7, 7), 2, 96, 256
        batch = 2
    elif cfg == "recog_pooler":
        sizes, scales, out_size, sampling, n, c, batch = [64], [1 / 4], (8, 32), 0, 24, 256, 1
    else:  # fewer channels than lanes * 8: the upper lanes are idle
        sizes, scales, out_size, sampling, n, c, batch = [32, 16], [1 / 4, 1 / 8], (5, 3), 2, 17, 72, 1
    acts = [ops.Act.from_nchw((torch.randn(batch, c, s, s, generator=g) * 2).cuda()) for s in sizes]
    feats = [a.to_nchw().cpu() for a in acts]
    rois = _random_rois(g, n, img=256.0, batch=batch)
    rois[:, 3:5] *= 0.55
    # a few RoIs hanging over the border / fully outside exercise the "sample outside the map" path
    rois[0, 1:3] = torch.tensor([-3.0, 10.0])
    rois[1, 1:3] = torch.tensor([400.0, 400.0])
    ref = d2_ops.roi_pooler(feats, [rois[rois[:, 0] == b][:, 1:] for b in range(batch)], out_size, scales, sampling)
    order = torch.cat([torch.nonzero(rois[:, 0] == b).squeeze(1) for b in range(batch)])
    out_act = ops.Act(n, c, out_size[0], out_size[1], border=1)
    got = ops.roi_align_rotated(acts, rois[order].contiguous().cuda(), out_size, scales, sampling,
                                out_split=(out_act.buf, out_act.hp, out_act.wp, out_act.border, 0, out_act.cp))
    _close(got.permute(0, 3, 1, 2), ref, cfg + "/f32")
    _close(out_act.to_nchw(), ref, cfg + "/split")
    assert out_act.buf[:, :, 0].abs().max().item() == 0  # the zero border is never written


def test_image_roi_align_matches_oracle(ops):
    from oracle import d2_ops
    g = torch.Generator().manual_seed(12)
    img = torch.randint(0, 256, (1, 3, 200, 232), generator=g).float()
    mean, std = (103.53, 116.28, 123.675), (1.0, 1.0, 1.0)
    norm = torch.zeros(1, 3, 224, 256)
    norm[:, :, :200, :232] = img - torch.tensor(mean).view(1, 3, 1, 1)
    rois = _random_rois(g, 6, img=220.0)
    rois[:, 3:5] *= 0.3
    ref = d2_ops.roi_pooler([norm], [rois[:, 1:]], (128, 128), [1.0], 2)
    act = ops.Act(6, 3, 128, 128)
    got = ops.image_roi_align_rotated(img.cuda(), (224, 256), mean, std, rois.cuda(), (128, 128), 2, out_f32=True,
                                      out_act=act)
    _close(got, ref, "image_pooler/f32")
    _close(act.to_nchw(), ref, "image_pooler/split")
    # the pre-normalised float4 path (workspace given) gathers the same values: bit-identical
    ws = torch.empty((1, 200, 232, 4), dtype=torch.float32, device="cuda")
    act2 = ops.Act(6, 3, 128, 128)
    got2 = ops.image_roi_align_rotated(img.cuda(), (224, 256), mean, std, rois.cuda(), (128, 128), 2, out_f32=True,
                                       out_act=act2, workspace=ws)
    assert torch.equal(got2, got) and torch.equal(act2.buf, act.buf)


@pytest.mark.parametrize("cin,cout,k,res,relu,n_img,hw", [
    (64, 256, 1, True, True, 8, (56, 72)),      # res2 conv3: HBM-bound, CTA pairs, 4 parts per tile
    (64, 256, 1, False, False, 8, (56, 72)),    # shortcut: no residual
    (256, 64, 1, False, True, 4, (40, 56)),     # one part per tile
    (128, 512, 1, True, True, 2, (24, 40)),     # two N tiles
    (128, 128, 3, True, True, 3, (16, 33)),     # 3x3 (tap-row mode) + residual, odd plane
    (64, 64, 3, False, True, 1, (7, 9)),        # tiny: a single partial tile
    (128, 512, 1, True, True, 8, (56, 72)),     # ring of three box pairs wrapping over ~4 tiles x 4 parts per CTA
    (64, 192, 1, True, False, 8, (56, 72)),     # three parts per tile: ring position differs from tile to tile
    (64, 128, 1, False, True, 8, (56, 72)),     # no residual: ring of two pairs, two parts per tile
    (256, 256, 1, True, True, 8, (56, 72)),     # four k-blocks (the FPN laterals' shape)
])
def test_tma_epilogue_equals_direct_store_epilogue(ops, monkeypatch, cin, cout, k, res, relu, n_img, hw):
    """The TMA-store epilogue (staging boxes + cp.async.bulk.tensor stores, TMA-loaded residual) writes the same bits as
    the direct-store epilogue, never touches pad channels, and leaves the zero border zero."""
    from glass_text_spotting_b200 import packing
    g = torch.Generator().manual_seed(cin + cout + k)
    x = ops.Act.from_nchw(torch.randn(n_img, cin, *hw, generator=g).cuda())
    w = packing.pack_conv(torch.randn(cout, cin, k, k, generator=g) * 0.05, torch.rand(cout, generator=g) + 0.5,
                          torch.randn(cout, generator=g), (1, 1), (k // 2, k // 2))
    r = ops.Act.from_nchw(torch.randn(n_img, cout, *hw, generator=g).cuda()) if res else None
    outs = []
    for mode in (1, 2):
        monkeypatch.setattr(ops, "EPI_MODE", mode)
        o = ops.Act(n_img, cout, hw[0], hw[1])
        ops.conv2d(x, w, relu=relu, residual=r, out=o)
        torch.cuda.synchronize()
        outs.append(o.buf.clone())
    assert torch.equal(outs[0], outs[1])
    b = outs[1]
    assert b[:, :, 0].abs().max().item() == 0 and b[:, :, -1].abs().max().item() == 0
    assert b[:, :, :, 0].abs().max().item() == 0 and b[:, :, :, -1].abs().max().item() == 0


@pytest.mark.parametrize("cin,cout,k,stride,pad,hw,shared", [
    (256, 128, 1, (2, 2), (0, 0), (48, 64), False),     # res3 conv1: stride in the 1x1
    (256, 512, 1, (2, 2), (0, 0), (48, 64), False),     # its shortcut: two N tiles
    (256, 256, 2, (2, 1), (0, 0), (16, 33), True),      # conv4_1: k2 s(2,1) into shared-border planes
    (256, 256, 2, (2, 1), (0, 0), (8, 32), True),       # CNN_V1_1 conv1's shape
])
def test_strided_conv_gathers_into_the_output_plane_order(ops, monkeypatch, cin, cout, k, stride, pad, hw, shared):
    """Strided convs gather their taps into the row order of the padded OUTPUT plane, which makes the GEMM flat (TMA-store
    epilogue); same bits as the dense gather + direct-store epilogue, zero border untouched."""
    from glass_text_spotting_b200 import packing
    g = torch.Generator().manual_seed(cin + cout + k)
    n = 5
    x = ops.Act.from_nchw(torch.randn(n, cin, *hw, generator=g).cuda(), shared=shared)
    w = packing.pack_conv(torch.randn(cout, cin, k, k, generator=g) * 0.05, torch.rand(cout, generator=g) + 0.5,
                          torch.randn(cout, generator=g), stride, pad)
    ho, wo = (hw[0] + 2 * pad[0] - k) // stride[0] + 1, (hw[1] + 2 * pad[1] - k) // stride[1] + 1
    outs = []
    for padded in (0, 1):
        monkeypatch.setattr(ops, "GATHER_PADDED", padded)
        o = ops.Act(n, cout, ho, wo, shared=shared)
        ops.conv2d(x, w, relu=True, out=o)
        torch.cuda.synchronize()
        outs.append(o.buf.clone())
    assert torch.equal(outs[0], outs[1])
    b = outs[1]
    assert b[:, :, 0].abs().max().item() == 0 and b[:, :, :, 0].abs().max().item() == 0


def test_saturation_counter_debug_flag(ops, monkeypatch):
    """the split-fp16 storage clamps |y| > 3750 silently; under the debug flag the GEMM counts such outputs."""
    from glass_text_spotting_b200 import packing
    g = torch.Generator().manual_seed(5)
    x = ops.Act.from_nchw(torch.randn(1, 64, 12, 16, generator=g).cuda())
    w = torch.randn(64, 64, 1, 1, generator=g) * 0.1
    monkeypatch.setattr(ops, "DEBUG_SAT", True)
    ops.saturation_count(reset=True)
    for mode in (1, 2):
        monkeypatch.setattr(ops, "EPI_MODE", mode)
        ops.conv2d(x, packing.pack_conv(w))
        assert ops.saturation_count() == 0
        hot = packing.pack_conv(w, torch.full((64,), 1.0), torch.cat((torch.tensor([5000.0, -9000.0]), torch.zeros(62))))
        ops.conv2d(x, hot)
        assert ops.saturation_count(reset=True) == 2 * 12 * 16      # two hot channels at every pixel
