"""The detectron2 operator surface (glass_text_spotting_b200/d2_ops.py) against the oracle's restatement
of the same detectron2 operators (oracle/d2_ops.py + oracle/d2_ops.c, pinned by detectron2's upstream known-answer
tests in tests/test_oracle_d2_ops.py): same names, NCHW in / NCHW out, index outputs exact."""
import math

import pytest
import torch

from parity_common import close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def d2(glass_lib):
    from glass_text_spotting_b200 import d2_ops
    return d2_ops


def _boxes(g, n, img=512.0, clustered=True):
    cx = torch.rand(n, generator=g) * img
    cy = torch.rand(n, generator=g) * img
    w = torch.exp(torch.rand(n, generator=g) * math.log(16.0)) * 12.0
    h = w * (0.15 + 0.85 * torch.rand(n, generator=g))
    a = torch.rand(n, generator=g) * 360 - 180
    b = torch.stack((cx, cy, w, h, a), 1)
    if clustered:   # jittered copies so that many pairs overlap heavily
        k = n // 3
        b[k:2 * k] = b[:k] + torch.randn(k, 5, generator=g) * torch.tensor([2.0, 2.0, 1.5, 1.0, 4.0])
        b[:, 2:4] = b[:, 2:4].clamp_min(1.0)
    return b.float().contiguous()


def test_box_iou_rotated_matches_oracle_and_d2_kats(d2):
    from oracle import d2_ops as od
    g = torch.Generator().manual_seed(41)
    b1, b2 = _boxes(g, 150), _boxes(g, 97)
    b2[:40] = b1[:40] + torch.randn(40, 5, generator=g) * 1.5
    b2[:, 2:4] = b2[:, 2:4].clamp_min(0.5)
    got = d2.box_iou_rotated(b1.cuda(), b2.cuda()).cpu()
    want = od.box_iou_rotated(b1, b2)
    assert got.shape == (150, 97) and float(want.max()) > 0.5
    assert float((got - want).abs().max()) <= 1e-5
    # detectron2 tests/layers/test_rotated_boxes.py known answers
    kat1 = torch.tensor([[0.5, 0.5, 1.0, 1.0, 0.0]])
    kat2 = torch.tensor([[0.25, 0.5, 0.5, 1.0, 0.0], [0.5, 0.5, 1.0, 1.0, 45.0], [0.5, 0.5, 1.0, 1.0, 360.0]])
    iou = d2.pairwise_iou_rotated(kat1.cuda(), kat2.cuda()).cpu()
    assert torch.allclose(iou, torch.tensor([[0.5, 0.707107, 1.0]]), atol=1e-5)
    # degenerate / empty
    assert tuple(d2.box_iou_rotated(b1[:0].cuda(), b2.cuda()).shape) == (0, 97)
    z = d2.box_iou_rotated(torch.tensor([[5.0, 5.0, 0.0, 3.0, 10.0]]).cuda(), b2[:4].cuda())
    assert float(z.abs().max()) == 0.0


def test_pairwise_ioa_rotated_matches_oracle(d2):
    from oracle import postprocess as pp
    g = torch.Generator().manual_seed(42)
    b = _boxes(g, 120)
    got = d2.pairwise_ioa_rotated(b.cuda(), b.cuda()).cpu()
    want = pp.pairwise_ioa_rotated(b, b)
    assert float((got - want).abs().max()) <= 2e-5
    assert float((torch.diagonal(got) - 1).abs().max()) <= 1e-4
    from glass_text_spotting_b200.structures import RotatedBoxes
    with pytest.raises(AttributeError):   # like the reference: tensors only (glass/structures/boxes.py:31)
        d2.pairwise_ioa_rotated(RotatedBoxes(b.cuda()), RotatedBoxes(b.cuda()))


@pytest.mark.parametrize("n,thr", [(1, 0.5), (63, 0.3), (64, 0.5), (65, 0.5), (1000, 0.3), (3000, 0.7), (8192, 0.5)])
def test_nms_rotated_returns_every_survivor_like_the_oracle(d2, n, thr):
    from oracle import d2_ops as od
    g = torch.Generator().manual_seed(100 + n)
    b = _boxes(g, n, img=300.0 + n ** 0.5 * 20)
    s = torch.rand(n, generator=g)
    s[n // 2:] = s[: n - n // 2].clone()   # ties: the stable sort decides
    want = od.nms_rotated(b, s, thr)
    got = d2.nms_rotated(b.cuda(), s.cuda(), thr)
    assert got.dtype == torch.int64 and got.is_cuda
    assert torch.equal(got.cpu(), want), (len(got), len(want))
    assert 0 < len(want) <= n and (n < 1000 or len(want) < n)


def test_nms_rotated_angle0_is_torchvision_nms(d2):
    import torchvision
    g = torch.Generator().manual_seed(43)
    n = 500
    xy = torch.rand(n, 2, generator=g) * 200
    wh = torch.rand(n, 2, generator=g) * 60 + 4
    s = torch.rand(n, generator=g)
    rb = torch.cat((xy + wh / 2, wh, torch.zeros(n, 1)), 1)
    want = torchvision.ops.nms(torch.cat((xy, xy + wh), 1), s, 0.5)
    got = d2.nms_rotated(rb.cuda(), s.cuda(), 0.5).cpu()
    assert torch.equal(got, want)
    assert d2.nms_rotated(rb[:0].cuda(), s[:0].cuda(), 0.5).numel() == 0


def test_batched_nms_rotated_matches_oracle(d2):
    from oracle import d2_ops as od
    g = torch.Generator().manual_seed(44)
    n = 1500
    b, s = _boxes(g, n), torch.rand(n, generator=g)
    idxs = torch.randint(0, 5, (n,), generator=g)
    want = od.batched_nms_rotated(b, s, idxs, 0.5)
    got = d2.batched_nms_rotated(b.cuda(), s.cuda(), idxs.cuda(), 0.5).cpu()
    assert torch.equal(got, want)
    # suppression never crosses categories: per-category NMS gives the same kept set
    per = torch.cat([torch.nonzero(idxs == c).squeeze(1)[od.nms_rotated(b[idxs == c], s[idxs == c], 0.5)] for c in range(5)])
    assert sorted(per.tolist()) == sorted(got.tolist())


def _rois(g, n, img, batch):
    cx = torch.rand(n, generator=g) * img
    cy = torch.rand(n, generator=g) * img
    w = torch.exp(torch.rand(n, generator=g) * (math.log(img / 2) - math.log(8)) + math.log(8))
    h = w * (0.1 + 0.9 * torch.rand(n, generator=g))
    a = torch.rand(n, generator=g) * 360 - 180
    return [torch.stack((cx, cy, w, h, a), 1)[i::batch].contiguous() for i in range(batch)]


def test_roi_pooler_multilevel_matches_oracle(d2):
    from glass_text_spotting_b200.structures import RotatedBoxes
    from oracle import d2_ops as od
    g = torch.Generator().manual_seed(45)
    sizes, scales = [64, 32, 16, 8, 4], [1 / 4, 1 / 8, 1 / 16, 1 / 32, 1 / 64]
    feats = [torch.randn(2, 256, s, s, generator=g) for s in sizes]
    boxes = _rois(g, 80, 256.0, 2)
    for b in boxes:          # spread the sizes over the pyramid's levels
        b[::3, 2:4] *= 5.0
        b[1::7, 2:4] *= 0.3
    want = od.roi_pooler(feats, boxes, (7, 7), scales, 2)
    pooler = d2.ROIPooler(output_size=7, scales=scales, sampling_ratio=2, pooler_type="ROIAlignRotated")
    got = pooler([f.cuda() for f in feats], [RotatedBoxes(b.cuda()) for b in boxes])
    assert tuple(got.shape) == (80, 256, 7, 7)
    close(got, want, "ROIPooler 5 levels")
    lv = od.assign_boxes_to_levels(torch.cat(boxes), 2, 6)
    assert len(set(lv.tolist())) >= 3, "the fixture must exercise several levels"
    # no boxes -> [0, C, oh, ow] zeros (d2 poolers.py)
    empty = pooler([f.cuda() for f in feats], [RotatedBoxes(torch.zeros(0, 5).cuda())] * 2)
    assert tuple(empty.shape) == (0, 256, 7, 7)
    with pytest.raises(ValueError):
        d2.ROIPooler(7, scales, 2, pooler_type="ROIAlignV2")


@pytest.mark.parametrize("c,out_hw,sampling", [(3, (32, 48), 2), (512, (8, 32), 0), (6, (4, 4), 0)])
def test_roi_align_rotated_forward_nchw(d2, c, out_hw, sampling):
    """Single level through torch.ops.detectron2.roi_align_rotated_forward's signature: 3 channels (the reference's
    image pooler, recognizers_hybrid_head.py:495-500), 512 (two 256-channel passes), 6 (padded to 8)."""
    from oracle import d2_ops as od
    g = torch.Generator().manual_seed(46 + c)
    x = torch.randn(2, c, 40, 56, generator=g)
    boxes = _rois(g, 12, 150.0, 2)
    rois = torch.cat([torch.cat((torch.full((len(b), 1), float(i)), b), 1) for i, b in enumerate(boxes)])
    want = od.roi_align_rotated(x, rois, out_hw, 0.25, sampling)
    got = d2.roi_align_rotated_forward(x.cuda(), rois.cuda(), 0.25, out_hw[0], out_hw[1], sampling)
    assert tuple(got.shape) == (12, c) + out_hw
    close(got, want, f"roi_align_rotated_forward C={c}")
    layer = d2.ROIAlignRotated(out_hw, 0.25, sampling)
    assert torch.equal(layer(x.cuda(), rois.cuda()), got)


def test_cpu_tensors_are_refused(d2):
    b = torch.tensor([[5.0, 5.0, 4.0, 2.0, 0.0]])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d2.box_iou_rotated(b, b)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d2.nms_rotated(b, torch.ones(1), 0.5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d2.roi_align_rotated_forward(torch.zeros(1, 4, 8, 8), torch.zeros(1, 6), 1.0, 2, 2, 2)


@pytest.mark.parametrize("rot", [0, 90, 180])
def test_nms_rotated_d2_kat_rotations(d2, rot):
    """detectron2 tests/layers/test_nms_rotated.py: all boxes rotated by 90 degrees (w/h swapped) or 180 keep the same
    indices as torchvision.ops.nms / batched_nms on the axis-aligned boxes (same fixture as tests/test_oracle_d2_ops.py)."""
    import torchvision
    g = torch.Generator().manual_seed(3)
    n = 200
    x0, y0 = torch.rand(n, generator=g) * 100, torch.rand(n, generator=g) * 100
    w, h = torch.rand(n, generator=g) * 40 + 1, torch.rand(n, generator=g) * 40 + 1
    boxes = torch.stack([x0, y0, x0 + w, y0 + h], 1)
    scores = torch.rand(n, generator=g)
    r = torch.zeros(n, 5)
    r[:, 0], r[:, 1] = (boxes[:, 0] + boxes[:, 2]) / 2, (boxes[:, 1] + boxes[:, 3]) / 2
    r[:, 2], r[:, 3] = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
    if rot == 90:
        r[:, 2], r[:, 3] = r[:, 3].clone(), r[:, 2].clone()
    r[:, 4] = rot
    for thr in [0.2, 0.5, 0.8]:
        assert torch.equal(d2.nms_rotated(r.cuda(), scores.cuda(), thr).cpu(), torchvision.ops.nms(boxes, scores, thr)), (rot, thr)
    idxs = torch.randint(0, 4, (n,), generator=g)
    assert torch.equal(d2.batched_nms_rotated(r.cuda(), scores.cuda(), idxs.cuda(), 0.5).cpu(),
                       torchvision.ops.batched_nms(boxes, scores, idxs, 0.5))
