"""GPU edge cases of the hot path: no detections, ragged image sizes inside one batch, sizes that are not
multiples of 32, the NMS at its maximum candidate count, padded / empty proposal slots."""
import pytest
import torch

from parity_common import close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model_and_oracle(glass_lib):
    from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN
    from oracle import model as om
    K = 4
    cfg = om.HotPathConfig(max_detections_override=K)
    img = om.synthetic_image(11, 160, 224)
    o = om.build_oracle(seed=3, calib_images=[img], cfg=cfg)
    return B200GlassRCNN(o.state_dict(), detections_per_image=K), o, K


def test_no_detections_returns_empty_instances(model_and_oracle):
    from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN
    _, o, _ = model_and_oracle
    m = B200GlassRCNN(o.state_dict(), score_thresh=1.1)   # nothing can pass a probability threshold > 1
    img = torch.randint(0, 256, (3, 128, 160)).float()
    out = m([{"image": img}, {"image": img}])
    assert len(out) == 2
    for r in out:
        inst = r["instances"]
        assert len(inst) == 0 and tuple(inst.pred_text_prob.shape) == (0, 26, 97) and tuple(inst.pred_boxes.tensor.shape) == (0, 5)


def test_ragged_batch_matches_oracle_on_the_same_canvas(model_and_oracle):
    """Images of different (non-multiple-of-32) sizes batched together are padded to one canvas with the pixel
    mean (normalised padding = 0, like ImageList.from_tensors); every image must then match the oracle run on
    that same padded canvas with its own (unpadded) size for clipping."""
    m, o, K = model_and_oracle
    g = torch.Generator().manual_seed(21)
    a = torch.randint(0, 256, (3, 150, 201), generator=g).float()
    b = torch.randint(0, 256, (3, 97, 230), generator=g).float()
    both = m.inference([{"image": a}, {"image": b}], do_postprocess=False)
    with torch.no_grad():
        canvas, sizes = o.preprocess_image([a, b])
        assert tuple(canvas.shape[-2:]) == (160, 256)
        for i in range(2):
            feats = o.backbone(canvas[i:i + 1])
            pboxes, _ = o.rpn(feats, sizes[i])
            want = o.box_branch(feats, pboxes, sizes[i])
            got = both[i]   # do_postprocess=False -> raw list[Instances] (glass_rcnn.py:100-101)
            assert got.image_size == sizes[i] and len(got) == len(want["pred_boxes"])
            close(got.pred_boxes.tensor, want["pred_boxes"], f"ragged boxes img{i}", rtol=5e-3, atol=5e-3)
            close(got.scores, want["scores"], f"ragged scores img{i}", rtol=5e-3, atol=1e-4)


def test_odd_size_image_matches_oracle_detections(model_and_oracle):
    m, o, K = model_and_oracle
    from oracle import model as om
    img = om.synthetic_image(13, 150, 201)
    with torch.no_grad():
        want = o.inference([{"image": img}])[0]["instances"]
    got = m([{"image": img}])[0]["instances"]
    assert len(got) == len(want["pred_boxes"])
    close(got.pred_boxes.tensor, want["pred_boxes"], "odd-size boxes", rtol=5e-3, atol=5e-3)
    close(got.scores, want["scores"], "odd-size scores", rtol=5e-3, atol=1e-4)


def test_nms_at_maximum_candidate_count(glass_lib):
    from glass_text_spotting_b200 import ops
    from oracle import d2_ops
    g = torch.Generator().manual_seed(31)
    n = 8192
    boxes = torch.rand(n, 5, generator=g) * torch.tensor([2000, 2000, 80, 40, 360]) - torch.tensor([0, 0, -4, -2, 180])
    scores = torch.rand(n, generator=g)
    keep = d2_ops.nms_rotated(boxes, scores, 0.5)[:128]
    ob, os_, oi, oc = ops.nms_rotated(boxes[None].cuda().contiguous(), scores[None].cuda().contiguous(), 0.5, 128)
    assert int(oc[0]) == 128 and torch.equal(oi[0].cpu().long(), keep)
    with pytest.raises(RuntimeError, match="8192"):
        ops.nms_rotated(torch.zeros(1, 8193, 5).cuda(), torch.zeros(1, 8193).cuda(), 0.5, 10)


def test_box_branch_ignores_padded_proposal_slots(model_and_oracle):
    """Slots beyond the per-image proposal count never become detections, whatever they contain."""
    m, _, _ = model_and_oracle
    from glass_text_spotting_b200 import ops
    g = torch.Generator().manual_seed(41)
    feats = {f"p{k}": ops.Act.from_nchw(torch.randn(1, 256, -(-128 // 2 ** k), -(-160 // 2 ** k), generator=g).cuda())
             for k in range(2, 7)}
    props = torch.zeros(1, 100, 5)
    props[0, :10] = torch.tensor([[40., 40., 30., 12., 10.]]) + torch.arange(10.).view(10, 1) * torch.tensor([8., 6., 0, 0, 5.])
    props[0, 10:] = float("nan")
    hw = torch.tensor([[128., 160.]]).cuda()
    det = m.roi_heads.forward_box(feats, props.cuda(), torch.tensor([10], dtype=torch.int32).cuda(), hw)
    k = int(det["count"][0])
    assert 0 < k <= 10 and int(det["index"][0, :k].max()) < 10
    assert torch.isfinite(det["pred_boxes"][0, :k]).all()
