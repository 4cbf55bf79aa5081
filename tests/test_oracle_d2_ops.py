"""Pins oracle/d2_ops.{c,py} with detectron2's upstream known-answer tests
([d2-recall] of detectron2 v0.6 tests/) and torchvision
angle-0 equivalences.  CPU only."""
import math

import numpy as np
import pytest
import torch
import torchvision

from oracle import d2_ops


def test_roi_align_rotated_0_90_180_270():
    # d2 tests/layers/test_roi_align_rotated.py::test_forward_output_0_90_180_270
    inp = torch.arange(25, dtype=torch.float32).reshape(1, 1, 5, 5)
    base = np.array([[4.5, 5.0, 5.5, 6.0], [7.0, 7.5, 8.0, 8.5], [9.5, 10.0, 10.5, 11.0], [12.0, 12.5, 13.0, 13.5]])
    for i in range(4):
        rois = torch.tensor([[0, 2, 2, 2, 2, 90.0 * i]])
        out = d2_ops.roi_align_rotated(inp, rois, (4, 4), 1.0, 0)
        expect = np.rot90(base, -i)
        assert np.allclose(out[0, 0].numpy(), expect, atol=1e-5), i


@pytest.mark.parametrize("sampling_ratio", [0, 2])
def test_roi_align_rotated_angle0_matches_torchvision(sampling_ratio):
    g = torch.Generator().manual_seed(1)
    feat = torch.randn(2, 5, 23, 31, generator=g)
    n = 40
    cx = torch.rand(n, generator=g) * 140 - 10
    cy = torch.rand(n, generator=g) * 110 - 10
    w = torch.rand(n, generator=g) * 60 + 1
    h = torch.rand(n, generator=g) * 40 + 1
    b = torch.randint(0, 2, (n,), generator=g).float()
    rois = torch.stack([b, cx, cy, w, h, torch.zeros(n)], 1)
    out = d2_ops.roi_align_rotated(feat, rois, (7, 5), 0.25, sampling_ratio)
    xyxy = torch.stack([b, cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
    ref = torchvision.ops.roi_align(feat, xyxy, (7, 5), 0.25, sampling_ratio, aligned=True)
    assert torch.allclose(out, ref, atol=2e-5, rtol=1e-5)


def test_roi_align_rotated_empty():
    out = d2_ops.roi_align_rotated(torch.zeros(1, 3, 4, 4), torch.zeros(0, 6), (2, 2), 1.0, 2)
    assert out.shape == (0, 3, 2, 2)


IOU_KATS = [
    ([0.5, 0.5, 1.0, 1.0, 0.0], [0.25, 0.5, 0.5, 1.0, 0.0], 0.5),
    ([1, 1, math.sqrt(2), math.sqrt(2), 45], [1, 1, 2, 2, 0], 0.5),
    ([1, 1, 2 * math.sqrt(2), 2 * math.sqrt(2), -45], [1, 1, 2, 2, 0], 0.5),
    ([5, 5, 10.0, 6.0, 55], [5, 5, 10.0, 6.0, -35], 36.0 / 84.0),
    ([565, 565, 10, 10.0, 0], [565, 565, 10, 8.3, 0], 0.83),
    ([299.5, 417.370422, 600.0, 364.259186, 27.1828], [299.5, 417.370422, 600.0, 364.259155, 27.1828],
     364.259155 / 364.259186),
]


@pytest.mark.parametrize("b1,b2,expect", IOU_KATS)
def test_box_iou_rotated_kats(b1, b2, expect):
    # d2 tests/structures/test_rotated_boxes.py
    iou = d2_ops.box_iou_rotated(torch.tensor([b1]), torch.tensor([b2]))
    assert abs(iou.item() - expect) < 1e-4


def test_box_iou_rotated_degenerate():
    z = torch.tensor([[1.0, 1.0, 0.0, 3.0, 10.0]])
    assert d2_ops.box_iou_rotated(z, torch.tensor([[1.0, 1.0, 2.0, 2.0, 0.0]])).item() == 0.0


def _random_boxes(n, g):
    x0 = torch.rand(n, generator=g) * 100
    y0 = torch.rand(n, generator=g) * 100
    w = torch.rand(n, generator=g) * 40 + 1
    h = torch.rand(n, generator=g) * 40 + 1
    return torch.stack([x0, y0, x0 + w, y0 + h], 1)


@pytest.mark.parametrize("rot", [0, 90, 180])
def test_nms_rotated_matches_torchvision(rot):
    # d2 tests/layers/test_nms_rotated.py
    g = torch.Generator().manual_seed(3)
    boxes = _random_boxes(200, g)
    scores = torch.rand(200, generator=g)
    r = torch.zeros(200, 5)
    r[:, 0] = (boxes[:, 0] + boxes[:, 2]) / 2
    r[:, 1] = (boxes[:, 1] + boxes[:, 3]) / 2
    r[:, 2] = boxes[:, 2] - boxes[:, 0]
    r[:, 3] = boxes[:, 3] - boxes[:, 1]
    if rot == 90:
        r[:, 2], r[:, 3] = r[:, 3].clone(), r[:, 2].clone()
    r[:, 4] = rot
    for thr in [0.2, 0.5, 0.8]:
        keep_ref = torchvision.ops.nms(boxes, scores, thr)
        keep = d2_ops.nms_rotated(r, scores, thr)
        assert torch.equal(keep, keep_ref), (rot, thr)


def test_batched_nms_rotated_matches_torchvision():
    g = torch.Generator().manual_seed(4)
    boxes = _random_boxes(150, g)
    scores = torch.rand(150, generator=g)
    idxs = torch.randint(0, 4, (150,), generator=g)
    r = torch.zeros(150, 5)
    r[:, 0] = (boxes[:, 0] + boxes[:, 2]) / 2
    r[:, 1] = (boxes[:, 1] + boxes[:, 3]) / 2
    r[:, 2] = boxes[:, 2] - boxes[:, 0]
    r[:, 3] = boxes[:, 3] - boxes[:, 1]
    keep_ref = torchvision.ops.batched_nms(boxes, scores, idxs, 0.5)
    keep = d2_ops.batched_nms_rotated(r, scores, idxs, 0.5)
    assert torch.equal(keep, keep_ref)
    assert d2_ops.batched_nms_rotated(torch.zeros(0, 5), torch.zeros(0), torch.zeros(0), 0.5).numel() == 0


def test_rrpn_anchor_generator_kat():
    # d2 tests/modeling/test_anchor_generator.py::test_rrpn_anchor_generator
    anchors = d2_ops.rotated_grid_anchors([(1, 2)], [4], [[32, 64]], [[0.25, 1, 4]], [[0, 45]])[0]
    expect = torch.tensor([
        [0, 0, 64, 16, 0], [0, 0, 64, 16, 45], [0, 0, 32, 32, 0], [0, 0, 32, 32, 45],
        [0, 0, 16, 64, 0], [0, 0, 16, 64, 45], [0, 0, 128, 32, 0], [0, 0, 128, 32, 45],
        [0, 0, 64, 64, 0], [0, 0, 64, 64, 45], [0, 0, 32, 128, 0], [0, 0, 32, 128, 45],
        [4, 0, 64, 16, 0], [4, 0, 64, 16, 45], [4, 0, 32, 32, 0], [4, 0, 32, 32, 45],
        [4, 0, 16, 64, 0], [4, 0, 16, 64, 45], [4, 0, 128, 32, 0], [4, 0, 128, 32, 45],
        [4, 0, 64, 64, 0], [4, 0, 64, 64, 45], [4, 0, 32, 128, 0], [4, 0, 32, 128, 45]], dtype=torch.float32)
    assert torch.allclose(anchors, expect)


def _get_deltas_rotated(src, dst, weights):
    # inverse of apply_deltas (d2 Box2BoxTransformRotated.get_deltas), test helper
    wx, wy, ww, wh, wa = weights
    dx = wx * (dst[:, 0] - src[:, 0]) / src[:, 2]
    dy = wy * (dst[:, 1] - src[:, 1]) / src[:, 3]
    dw = ww * torch.log(dst[:, 2] / src[:, 2])
    dh = wh * torch.log(dst[:, 3] / src[:, 3])
    da = dst[:, 4] - src[:, 4]
    da = (da + 180.0) % 360.0 - 180.0
    da = da * wa * math.pi / 180.0
    return torch.stack((dx, dy, dw, dh, da), 1)


def test_box2box_rotated_roundtrip():
    # d2 tests/modeling/test_box2box_transform.py::test_reconstruction (rotated)
    g = torch.Generator().manual_seed(5)
    for weights in [(1, 1, 1, 1, 2), (10, 10, 5, 5, 10)]:
        src = torch.rand(50, 5, generator=g) * 50 + 1
        dst = torch.rand(50, 5, generator=g) * 50 + 1
        src[:, 4] = torch.rand(50, generator=g) * 360 - 180
        dst[:, 4] = torch.rand(50, generator=g) * 360 - 180
        deltas = _get_deltas_rotated(src, dst, weights)
        rec = d2_ops.apply_deltas_rotated(deltas, src, weights)
        assert torch.allclose(rec[:, :4], dst[:, :4], atol=1e-3)
        diff = (rec[:, 4] - dst[:, 4] + 180.0) % 360.0 - 180.0
        assert diff.abs().max() < 1e-3


def test_clip_only_near_axis_aligned():
    t = torch.tensor([[5.0, 5.0, 20.0, 8.0, 0.5], [5.0, 5.0, 20.0, 8.0, 30.0], [95.0, 50.0, 20.0, 8.0, 180.2]])
    c = d2_ops.clip_rotated_(t.clone(), (60, 100))
    # box 0: |angle|<=1 -> clipped to x in [0,15], y in [1,9]
    assert torch.allclose(c[0, :4], torch.tensor([7.5, 5.0, 15.0, 8.0]))
    # box 1: untouched
    assert torch.equal(c[1], t[1])
    # box 2: angle normalised to -179.8 -> not clipped
    assert abs(c[2, 4].item() + 179.8) < 1e-4 and torch.equal(c[2, :4], t[2, :4])


def test_level_assignment():
    b = torch.tensor([[0, 0, 224.0, 224.0, 0], [0, 0, 112.0, 112.0, 0], [0, 0, 10.0, 10.0, 0], [0, 0, 2000.0, 2000.0, 0],
                      [0, 0, 448.0, 448.0, 10]])
    lv = d2_ops.assign_boxes_to_levels(b, 2, 6)
    assert lv.tolist() == [2, 1, 0, 4, 3]
