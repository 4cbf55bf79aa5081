"""Independent implementations for the ROTATED halves of the detectron2 operators the oracle restates (SURVEY.md A.4,
A.6), which torchvision does not have: OpenCV's rotated-rectangle intersection for the rotated IoU, and torch's own
bilinear sampler (``F.grid_sample``) driven by the published sampling-point formula for RoIAlignRotated."""
import math

import numpy as np
import torch
import torch.nn.functional as F


def test_rotated_iou_matches_opencv_rotated_rectangle_intersection():
    import cv2
    from oracle import d2_ops
    g = torch.Generator().manual_seed(12)
    n = 300
    a = torch.stack((torch.rand(n, generator=g) * 100, torch.rand(n, generator=g) * 100, 5 + torch.rand(n, generator=g) * 60,
                     5 + torch.rand(n, generator=g) * 30, torch.rand(n, generator=g) * 360 - 180), 1)
    b = a + torch.randn(n, 5, generator=g) * torch.tensor([8.0, 8.0, 6.0, 4.0, 25.0])
    b[:, 2:4] = b[:, 2:4].clamp_min(2.0)
    got = torch.diagonal(d2_ops.box_iou_rotated(a, b))
    want = []
    for r1, r2 in zip(a.tolist(), b.tolist()):
        # OpenCV's RotatedRect angle turns clockwise in image coordinates, detectron2's counter-clockwise
        q1, q2 = ((r1[0], r1[1]), (r1[2], r1[3]), -r1[4]), ((r2[0], r2[1]), (r2[2], r2[3]), -r2[4])
        kind, pts = cv2.rotatedRectangleIntersection(q1, q2)
        inter = 0.0
        if kind != cv2.INTERSECT_NONE and pts is not None and len(pts) >= 3:
            inter = cv2.contourArea(cv2.convexHull(pts, returnPoints=True))
        a1, a2 = r1[2] * r1[3], r2[2] * r2[3]
        want.append(inter / (a1 + a2 - inter))
    want = torch.tensor(want, dtype=torch.float32)
    assert float(want.max()) > 0.7 and float((want == 0).float().mean()) < 0.5     # the fixture overlaps for real
    assert torch.allclose(got, want, atol=2e-3), float((got - want).abs().max())


def _roi_align_rotated_by_grid_sample(x, rois, out_hw, scale, sampling):
    """RoIAlignRotated from the published definition: bin (ph, pw) averages sampling x sampling points laid out on the
    box's own axes, rotated about its centre, each read with bilinear interpolation (pixel centres at integers after the
    half-pixel shift).  Valid for RoIs that stay inside the map (no border clamping implemented here)."""
    n, c, h, w = x.shape
    oh, ow = out_hw
    out = []
    for r in rois.tolist():
        bi, cx, cy, rw, rh, ang = r
        cx, cy, rw, rh = cx * scale - 0.5, cy * scale - 0.5, rw * scale, rh * scale
        th = math.radians(ang)
        cs, sn = math.cos(th), math.sin(th)
        iy = (torch.arange(oh * sampling, dtype=torch.float64) + 0.5) / (oh * sampling)      # fractions along the height
        ix = (torch.arange(ow * sampling, dtype=torch.float64) + 0.5) / (ow * sampling)
        yy = (-rh / 2 + iy * rh).view(-1, 1).expand(oh * sampling, ow * sampling)
        xx = (-rw / 2 + ix * rw).view(1, -1).expand(oh * sampling, ow * sampling)
        ys = yy * cs - xx * sn + cy
        xs = yy * sn + xx * cs + cx
        grid = torch.stack((2 * xs / (w - 1) - 1, 2 * ys / (h - 1) - 1), -1).float()[None]      # align_corners=True
        samp = F.grid_sample(x[int(bi):int(bi) + 1], grid, mode="bilinear", padding_mode="zeros", align_corners=True)
        out.append(F.avg_pool2d(samp, sampling)[0])
    return torch.stack(out)


def test_roi_align_rotated_matches_grid_sample_formulation():
    from oracle import d2_ops
    g = torch.Generator().manual_seed(13)
    x = torch.randn(2, 8, 48, 64, generator=g)
    n = 40
    # boxes in image coordinates (scale 1/4): kept well inside the 192 x 256 image so that no sample leaves the map
    w = 16 + torch.rand(n, generator=g) * 60
    h = 8 + torch.rand(n, generator=g) * 30
    cx = 70 + torch.rand(n, generator=g) * 116
    cy = 60 + torch.rand(n, generator=g) * 72
    ang = torch.rand(n, generator=g) * 360 - 180
    rois = torch.stack((torch.randint(0, 2, (n,), generator=g).float(), cx, cy, w, h, ang), 1)
    for out_hw, sampling in [((7, 7), 2), ((8, 32), 2), ((3, 5), 3)]:
        got = d2_ops.roi_align_rotated(x, rois, out_hw, 0.25, sampling)
        want = _roi_align_rotated_by_grid_sample(x, rois, out_hw, 0.25, sampling)
        assert got.shape == want.shape
        assert torch.allclose(got, want, rtol=1e-4, atol=2e-5), (out_hw, float((got - want).abs().max()))
    assert np.isfinite(got.numpy()).all()


def test_rotated_nms_matches_greedy_nms_on_opencv_ious():
    """Greedy NMS written from its definition on top of the OpenCV IoUs (neither the IoU nor the loop shared with the
    oracle) keeps the same boxes in the same order, for arbitrary angles; pairs whose IoU is within 5e-3 of the threshold
    are excluded from the fixture (the two IoU implementations agree to ~2e-3)."""
    import cv2
    from oracle import d2_ops
    g = torch.Generator().manual_seed(14)
    n, thr = 160, 0.4
    base = torch.stack((torch.rand(n, generator=g) * 120, torch.rand(n, generator=g) * 120, 8 + torch.rand(n, generator=g) * 50,
                        6 + torch.rand(n, generator=g) * 25, torch.rand(n, generator=g) * 360 - 180), 1)
    base[n // 2:] = base[: n - n // 2] + torch.randn(n - n // 2, 5, generator=g) * torch.tensor([3.0, 3.0, 3.0, 2.0, 10.0])
    base[:, 2:4] = base[:, 2:4].clamp_min(3.0)
    scores = torch.rand(n, generator=g)

    def iou(r1, r2):
        kind, pts = cv2.rotatedRectangleIntersection(((r1[0], r1[1]), (r1[2], r1[3]), -r1[4]), ((r2[0], r2[1]), (r2[2], r2[3]), -r2[4]))
        inter = cv2.contourArea(cv2.convexHull(pts, returnPoints=True)) if kind != cv2.INTERSECT_NONE and pts is not None and len(pts) >= 3 else 0.0
        return inter / (r1[2] * r1[3] + r2[2] * r2[3] - inter)
    rows = base.tolist()
    m = np.array([[iou(rows[i], rows[j]) if i != j else 1.0 for j in range(n)] for i in range(n)])
    ambiguous = set(np.argwhere(np.abs(m - thr) < 5e-3).reshape(-1).tolist())
    keep_idx = [i for i in range(n) if i not in ambiguous]
    assert len(keep_idx) > 100
    b, s = base[keep_idx], scores[keep_idx]
    mm = m[np.ix_(keep_idx, keep_idx)]
    order = torch.sort(s, descending=True, stable=True).indices.tolist()
    kept, dead = [], set()
    for i in order:
        if i in dead:
            continue
        kept.append(i)
        dead.update(j for j in order if mm[i, j] > thr and j != i)
    got = d2_ops.nms_rotated(b, s, thr).tolist()
    assert got == kept and 10 < len(kept) < len(keep_idx)
