"""CPU tests of the host-side logic (no GPU): weight packing layouts, structures, post-processing --
each checked against the oracle / plain torch on the same inputs."""

import pytest
import torch
import torch.nn.functional as F


def _unsplit(w):
    """PackedWeight -> effective fp32 weight rows [n_p, K] and epilogue (scale, bias) with pre-scales undone."""
    from glass_text_spotting_b200.ops import ACT_SCALE
    return (w.w[0].float() + w.w[1].float()), w.scale * ACT_SCALE, w.bias


def test_pack_conv_is_tap_major_channel_minor():
    from glass_text_spotting_b200 import packing
    g = torch.Generator().manual_seed(0)
    wt = torch.randn(20, 7, 3, 3, generator=g)
    scale, bias = torch.rand(20, generator=g) + 0.5, torch.randn(20, generator=g)
    pw = packing.pack_conv(wt, scale, bias, (1, 1), (1, 1), device="cpu")
    assert pw.n_p == 64 and pw.cin_p == 64 and tuple(pw.w.shape) == (2, 64, 9 * 64)
    rows, s, b = _unsplit(pw)
    # emulate the implicit GEMM on CPU: im2col with K = (tap, channel) and compare with conv2d
    x = torch.randn(2, 7, 6, 5, generator=g)
    xp = F.pad(x, (1, 1, 1, 1))
    cols = torch.zeros(2, 6, 5, 9, 64)
    for r in range(3):
        for c in range(3):
            cols[:, :, :, r * 3 + c, :7] = xp[:, :, r:r + 6, c:c + 5].permute(0, 2, 3, 1)
    out = (cols.reshape(-1, 9 * 64) @ rows.t()) * s + b
    ref = F.conv2d(x, wt, padding=1) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    assert torch.allclose(out[:, :20].view(2, 6, 5, 20).permute(0, 3, 1, 2), ref, rtol=1e-4, atol=1e-4)
    assert out[:, 20:].abs().max() == 0  # pad output channels stay exactly zero


def test_pack_rows_prescale_is_a_power_of_two_and_keeps_22_bits():
    from glass_text_spotting_b200 import packing
    g = torch.Generator().manual_seed(1)
    w = torch.randn(5, 64, generator=g) * torch.tensor([1e-3, 1e-1, 1.0, 30.0, 0.0]).view(5, 1)
    pw = packing.pack_linear(w, None, device="cpu", n_align=16)
    rows, s, _ = _unsplit(pw)
    rec = rows[:5] * s[:5].view(5, 1)
    assert (rec - w).abs().max() <= 2 ** -21 * w.abs().max()
    ratio = s[:4]
    assert torch.all(torch.log2(ratio) == torch.log2(ratio).round())
    assert pw.w[0][:4].abs().amax(1).min() >= 256 and pw.w[0].abs().max() < 512


def test_pack_conv_compact_matches_conv2d():
    from glass_text_spotting_b200 import packing
    g = torch.Generator().manual_seed(2)
    for cin, cp, k in [(3, 8, 3), (16, 16, 3), (32, 32, 3), (32, 32, 1)]:
        wt = torch.randn(24, cin, k, k, generator=g)
        pw = packing.pack_conv_compact(wt, cp, device="cpu")
        rows, s, b = _unsplit(pw)
        ppk = 64 // cp
        nj = (k + ppk - 1) // ppk
        assert tuple(pw.w.shape) == (2, 32, k * nj * 64) and pw.compact_cp == cp
        x = torch.randn(1, cin, 5, 9, generator=g)
        pad = k // 2
        # flat padded NHWC rows with cp channels, plus slack so every 64-wide window is in range
        xp = torch.zeros(5 + 2, 9 + 2, cp)
        xp[1:-1, 1:-1, :cin] = x[0].permute(1, 2, 0)
        flat = torch.cat((xp.reshape(-1), torch.zeros(64)))
        wp = 9 + 2
        out = torch.zeros(5, 9, 32)
        for y in range(5):
            for xx in range(9):
                m = (y + 1) * wp + (xx + 1)
                acc = torch.zeros(32)
                t = 0
                for r in range(k):
                    for j in range(nj):
                        row = m + (r - pad) * wp - pad + j * ppk
                        a = flat[row * cp: row * cp + 64]
                        acc += rows[:, t * 64:(t + 1) * 64] @ a
                        t += 1
                out[y, xx] = acc * s + b
        ref = F.conv2d(x, wt, padding=pad)[0].permute(1, 2, 0)
        assert torch.allclose(out[..., :24], ref, rtol=1e-4, atol=1e-4), (cin, cp, k)


def test_rotated_boxes_match_oracle_helpers():
    from glass_text_spotting_b200.structures import RotatedBoxes
    from oracle import d2_ops
    g = torch.Generator().manual_seed(3)
    t = torch.rand(50, 5, generator=g) * torch.tensor([300, 200, 120, 60, 720]) - torch.tensor([20, 20, 0, 0, 360])
    t[::5, 4] = torch.rand(10, generator=g) * 2 - 1  # near-horizontal boxes get clipped
    a, b = RotatedBoxes(t.clone()), t.clone()
    a.clip((150, 260))
    d2_ops.clip_rotated_(b, (150, 260))
    assert torch.equal(a.tensor, b)
    assert torch.equal(a.nonempty(), d2_ops.nonempty_rotated(b))
    a.scale(1.7, 0.6)
    d2_ops.scale_rotated_(b, 1.7, 0.6)
    assert torch.allclose(a.tensor, b, rtol=1e-6, atol=1e-6)


def test_detector_postprocess_matches_oracle():
    from glass_text_spotting_b200.modeling.glass_rcnn import detector_postprocess
    from glass_text_spotting_b200.structures import Instances, RotatedBoxes
    from oracle import model as om
    g = torch.Generator().manual_seed(4)
    boxes = torch.rand(12, 5, generator=g) * torch.tensor([256, 192, 80, 30, 360]) - torch.tensor([0, 0, 0, 0, 180])
    boxes[3, 2:4] = 0  # empty box is dropped
    scores = torch.rand(12, generator=g)
    inst = Instances((192, 256), pred_boxes=RotatedBoxes(boxes.clone()), scores=scores,
                     pred_text_prob=torch.rand(12, 26, 97, generator=g))
    got = detector_postprocess(inst, 384, 640)
    want = om.GlassOracle.postprocess({"pred_boxes": boxes.clone(), "scores": scores}, (192, 256), 384, 640)
    assert torch.allclose(got.pred_boxes.tensor, want["pred_boxes"], atol=1e-5)
    assert torch.equal(got.scores, want["scores"]) and got.image_size == (384, 640)
    assert got.pred_text_prob.shape[0] == len(got) == 11


def test_image_list_pads_with_pixel_mean():
    from glass_text_spotting_b200.structures import ImageList
    a, b = torch.ones(3, 30, 50), 2 * torch.ones(3, 40, 33)
    il = ImageList.from_tensors([a, b], 32, pad_value=(5.0, 6.0, 7.0))
    assert tuple(il.tensor.shape) == (2, 3, 64, 64) and il.image_sizes == [(30, 50), (40, 33)]
    assert torch.equal(il.tensor[0, :, :30, :50], a) and torch.equal(il.tensor[1, :, :40, :33], b)
    assert torch.all(il.tensor[0, 1, 30:, :] == 6.0) and torch.all(il.tensor[1, 2, :, 33:] == 7.0)


def test_instances_indexing_and_len():
    from glass_text_spotting_b200.structures import Instances, RotatedBoxes
    inst = Instances((10, 10), pred_boxes=RotatedBoxes(torch.arange(20.).view(4, 5)), scores=torch.arange(4.))
    sub = inst[torch.tensor([True, False, True, False])]
    assert len(inst) == 4 and len(sub) == 2 and torch.equal(sub.scores, torch.tensor([0., 2.]))
    with pytest.raises(AssertionError):
        inst.set("bad", torch.zeros(3))


def test_rotated_cell_anchors_match_oracle():
    from glass_text_spotting_b200.modeling.rpn import rotated_cell_anchors
    from oracle import d2_ops
    for size in (16, 32, 64, 128, 256):
        mine = torch.tensor(rotated_cell_anchors(float(size), (0.2, 0.5, 1.0), (-90, -45, 0, 45)), dtype=torch.float32)
        ref = d2_ops.rotated_cell_anchors([size], [0.2, 0.5, 1.0], [-90, -45, 0, 45])
        assert torch.equal(mine, ref[:, 2:])


def test_backbone_flop_count_matches_survey():
    from glass_text_spotting_b200.modeling.backbone import B200ResNetFPN
    assert abs(B200ResNetFPN.flops_per_image(1024, 1024) / 1e9 - 279.94) < 0.1   # 161.16 + 118.78


def test_d2_ops_surface_validates_on_the_host():
    """glass_text_spotting_b200.d2_ops keeps detectron2's names and argument checks; CPU tensors are refused before any
    library call (no fallback path)."""
    import pytest
    from glass_text_spotting_b200 import d2_ops
    from glass_text_spotting_b200.structures import RotatedBoxes
    for name in ("ROIPooler", "ROIAlignRotated", "roi_align_rotated_forward", "nms_rotated", "batched_nms_rotated",
                 "box_iou_rotated", "pairwise_iou_rotated", "pairwise_ioa_rotated"):
        assert hasattr(d2_ops, name)
    p = d2_ops.ROIPooler(output_size=7, scales=[1 / 4, 1 / 8, 1 / 16, 1 / 32, 1 / 64], sampling_ratio=2,
                         pooler_type="ROIAlignRotated")
    assert (p.min_level, p.max_level, p.output_size) == (2, 6, (7, 7))
    assert d2_ops.ROIPooler([8, 32], (1 / 4,), 0, "ROIAlignRotated").output_size == (8, 32)      # recognizers_hybrid_head.py:464
    with pytest.raises(ValueError, match="Unknown pooler type"):
        d2_ops.ROIPooler(7, [0.25], 2, pooler_type="ROIAlignV2")
    with pytest.raises(AssertionError, match="power of 2"):
        d2_ops.ROIPooler(7, [0.3], 2)
    with pytest.raises(AssertionError, match="pyramid"):
        d2_ops.ROIPooler(7, [1 / 4, 1 / 16], 2)
    b = torch.tensor([[5.0, 5.0, 4.0, 2.0, 0.0]])
    with pytest.raises(AssertionError, match="lists"):
        p(torch.zeros(1, 4, 8, 8), [RotatedBoxes(b)])
    for call in (lambda: d2_ops.box_iou_rotated(b, b), lambda: d2_ops.nms_rotated(b, torch.ones(1), 0.5),
                 lambda: d2_ops.pairwise_ioa_rotated(b, b),
                 lambda: d2_ops.roi_align_rotated_forward(torch.zeros(1, 4, 8, 8), torch.zeros(1, 6), 1.0, 2, 2, 2),
                 lambda: p([torch.zeros(1, 4, s, s) for s in (64, 32, 16, 8, 4)], [RotatedBoxes(b)])):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()
    fmt = d2_ops.convert_boxes_to_pooler_format([RotatedBoxes(b), RotatedBoxes(torch.cat((b, b + 1)))])
    assert fmt.shape == (3, 6) and fmt[:, 0].tolist() == [0.0, 1.0, 1.0]


def test_glass_rcnn_module_signatures_match_the_reference_call_sites():
    """glass_rcnn.py:88, :93, :96 call the sub-modules positionally; the B200 modules must accept exactly that."""
    import inspect
    from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN
    from glass_text_spotting_b200.modeling.roi_heads import B200GlassROIHeads
    from glass_text_spotting_b200.modeling.rpn import B200RotatedRPN
    assert list(inspect.signature(B200RotatedRPN.forward).parameters) == ["self", "images", "features", "gt_instances"]
    assert list(inspect.signature(B200GlassROIHeads.forward).parameters) == ["self", "images", "features", "proposals", "targets"]
    assert list(inspect.signature(B200GlassROIHeads.forward_with_given_boxes).parameters) == ["self", "images", "features", "instances"]
    assert list(inspect.signature(B200GlassRCNN.inference).parameters)[:4] == ["self", "batched_inputs", "detected_instances", "do_postprocess"]
    assert B200GlassRCNN.__call__ is B200GlassRCNN.forward and B200GlassROIHeads.__call__ is B200GlassROIHeads.forward


def test_decoder_fragment_packing_matches_the_mma_operand_layout():
    """packing.pack_decoder_h_weights: register `reg` of lane (g, q) of (m-tile, k-step, plane) must hold
    (row g + 8*(reg & 1), k = 2q + 8*(reg >> 1) + {0, 1}) of its 16 x 16 tile -- mma.sync.m16n8k16's A operand."""
    import random
    from glass_text_spotting_b200 import packing
    g = torch.Generator().manual_seed(0)
    ws, whh = torch.randn(256, 256, generator=g) * 0.06, torch.randn(768, 256, generator=g) * 0.06
    frag, bias = packing.pack_decoder_h_weights(ws, torch.ones(256), whh, torch.zeros(768), device="cpu")
    assert frag.dtype == torch.int32 and tuple(frag.shape) == (64, 16, 2, 32, 4) and bias[:256].eq(1).all() and bias[256:].eq(0).all()
    planes = packing.split16(torch.cat((ws, whh), 0) * packing.DEC_SW)
    f16 = frag.view(torch.float16).view(64, 16, 2, 32, 4, 2)
    rnd = random.Random(1)
    for _ in range(3000):
        mt, ks, pl, lane, reg, e = (rnd.randrange(n) for n in (64, 16, 2, 32, 4, 2))
        row, k = mt * 16 + (lane >> 2) + 8 * (reg & 1), ks * 16 + 2 * (lane & 3) + 8 * (reg >> 1) + e
        assert f16[mt, ks, pl, lane, reg, e] == planes[pl, row, k]


def test_gather_rows_follow_the_output_plane():
    """ops.gather_rows: dense rows, the rows of padded planes, and of shared-border planes (only leading borders)."""
    from glass_text_spotting_b200 import ops
    assert ops.gather_rows(3, 8, 32, 0) == 3 * 8 * 32
    assert ops.gather_rows(3, 8, 32, 1) == 3 * 10 * 34
    assert ops.gather_rows(3, 8, 32, 1 | ops.BORDER_SHARED) == 3 * 9 * 33
    a = ops.Act.__new__(ops.Act)   # geometry only: no device memory needed for the row count
    for shared in (False, True):
        a.n, a.h, a.w, a.border, a.shared = 5, 16, 33, 1, shared
        hp = 16 + (1 if shared else 2)
        wp = 33 + (1 if shared else 2)
        assert ops.gather_rows(5, 16, 33, a.border | (ops.BORDER_SHARED if shared else 0)) == 5 * hp * wp
