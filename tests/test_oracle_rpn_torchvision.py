"""Independent cross-check of the oracle's detectron2-recalled proposal selection (row a4):
``find_top_rrpn_proposals`` = per-level top-k -> concat -> clip -> non-empty -> NMS per level -> top-k.  torchvision's
``RegionProposalNetwork.filter_proposals`` is an independent implementation of the same published procedure for
axis-aligned boxes; at angle 0 the rotated procedure must select the same proposals in the same order."""
import torch
from torchvision.models.detection.rpn import RegionProposalNetwork


def test_rrpn_selection_at_angle0_matches_torchvision_filter_proposals():
    from oracle import d2_ops
    g = torch.Generator().manual_seed(9)
    n_img, hw = 2, (200, 320)
    per_level = [1800, 600, 150, 40, 12]           # anchors per level (levels 0-1 exceed the pre-NMS top-k)
    pre_topk, post_topk, thr = 500, 60, 0.7
    props, logits = [], []
    for lvl, a in enumerate(per_level):
        size = 12.0 * 2 ** lvl
        cx = torch.rand(n_img, a, generator=g) * (hw[1] + 40) - 20          # some boxes hang over the border: clipped
        cy = torch.rand(n_img, a, generator=g) * (hw[0] + 40) - 20
        w = size * (0.5 + torch.rand(n_img, a, generator=g))
        h = size * (0.5 + torch.rand(n_img, a, generator=g)) * 0.6
        props.append(torch.stack((cx, cy, w, h, torch.zeros(n_img, a)), 2))
        logits.append(torch.randn(n_img, a, generator=g))
    got = d2_ops.find_top_rrpn_proposals(props, logits, [hw] * n_img, thr, pre_topk, post_topk)

    rpn = RegionProposalNetwork(None, None, 0.7, 0.3, 256, 0.5, dict(training=pre_topk, testing=pre_topk),
                                dict(training=post_topk, testing=post_topk), thr, score_thresh=0.0).eval()
    allp = torch.cat(props, 1)
    xyxy = torch.stack((allp[..., 0] - allp[..., 2] / 2, allp[..., 1] - allp[..., 3] / 2,
                        allp[..., 0] + allp[..., 2] / 2, allp[..., 1] + allp[..., 3] / 2), 2)
    boxes_tv, scores_tv = rpn.filter_proposals(xyxy, torch.cat(logits, 1).reshape(-1, 1), [hw] * n_img, per_level)
    for i in range(n_img):
        b, s = got[i]
        assert len(b) == len(boxes_tv[i]) == post_topk
        mine = torch.stack((b[:, 0] - b[:, 2] / 2, b[:, 1] - b[:, 3] / 2, b[:, 0] + b[:, 2] / 2, b[:, 1] + b[:, 3] / 2), 1)
        assert torch.allclose(mine, boxes_tv[i], atol=1e-3), float((mine - boxes_tv[i]).abs().max())
        assert torch.allclose(torch.sigmoid(s), scores_tv[i], atol=1e-6)
        assert bool((s[:-1] >= s[1:]).all())


def test_rotated_box_decode_at_angle0_matches_torchvision_boxcoder():
    """Box2BoxTransformRotated.apply_deltas with a zero angle delta == torchvision's BoxCoder.decode
    (same published parameterisation: centre shift scaled by the size, log-size with the log(1000/16) clamp)."""
    from torchvision.models.detection._utils import BoxCoder
    from oracle import d2_ops
    g = torch.Generator().manual_seed(10)
    n = 400
    cx, cy = torch.rand(n, generator=g) * 500, torch.rand(n, generator=g) * 400
    w, h = 4 + torch.rand(n, generator=g) * 200, 4 + torch.rand(n, generator=g) * 100
    anchors = torch.stack((cx, cy, w, h, torch.zeros(n)), 1)
    deltas = torch.randn(n, 5, generator=g) * torch.tensor([2.0, 2.0, 3.0, 3.0, 0.0])
    deltas[:20, 2:4] = 40.0                                                     # far beyond the clamp
    for weights in [(1.0, 1.0, 1.0, 1.0, 2.0), (10.0, 10.0, 5.0, 5.0, 10.0)]:   # RPN / box head (glass_pretrain.yaml:66, 95)
        got = d2_ops.apply_deltas_rotated(deltas, anchors, weights)
        coder = BoxCoder(weights[:4])
        xyxy = torch.stack((cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2), 1)
        want = coder.decode_single(deltas[:, :4], xyxy)
        mine = torch.stack((got[:, 0] - got[:, 2] / 2, got[:, 1] - got[:, 3] / 2, got[:, 0] + got[:, 2] / 2,
                            got[:, 1] + got[:, 3] / 2), 1)
        assert torch.allclose(mine, want, rtol=1e-5, atol=1e-2), float((mine - want).abs().max())
        assert torch.equal(got[:, 4], torch.zeros(n))


def test_rotated_anchor_grid_at_angle0_matches_torchvision_anchor_generator():
    """RotatedAnchorGenerator (w = sqrt(s^2 / r), h = r w; order y, x, then ratio) with the single angle 0
    == torchvision's AnchorGenerator for sizes / ratios whose anchors are integral (torchvision rounds its base anchors)."""
    from torchvision.models.detection.anchor_utils import AnchorGenerator
    from torchvision.models.detection.image_list import ImageList
    from oracle import d2_ops
    grids, strides = [(6, 8), (3, 4)], [8, 16]           # a 48 x 64 image: torchvision infers the strides as size // grid
    got = d2_ops.rotated_grid_anchors(grids, strides, [[32.0], [64.0]], [[0.25, 1.0, 4.0]], [[0.0]])
    ag = AnchorGenerator(sizes=((32,), (64,)), aspect_ratios=((0.25, 1.0, 4.0),) * 2)
    tv = ag(ImageList(torch.zeros(1, 3, 48, 64), [(48, 64)]), [torch.zeros(1, 1, gh, gw) for gh, gw in grids])[0]
    off = 0
    for lvl, a in enumerate(got):
        n = a.shape[0]
        assert n == grids[lvl][0] * grids[lvl][1] * 3
        mine = torch.stack((a[:, 0] - a[:, 2] / 2, a[:, 1] - a[:, 3] / 2, a[:, 0] + a[:, 2] / 2, a[:, 1] + a[:, 3] / 2), 1)
        assert torch.equal(a[:, 4], torch.zeros(n))
        assert torch.allclose(mine, tv[off: off + n], atol=1e-4), (lvl, float((mine - tv[off: off + n]).abs().max()))
        off += n
    assert off == tv.shape[0]
