"""GPU parity of the ROI head against the oracle, stage-wise teacher-forced:
the box branch gets the oracle's features + proposals, the recognizer the oracle's detections."""
import math

import pytest
import torch

from parity_common import calibrate_bn, close

pytestmark = pytest.mark.gpu


def _feats(g, n, h, w, scale=1.0):
    return {f"p{k}": scale * torch.randn(n, 256, math.ceil(h / 2 ** k), math.ceil(w / 2 ** k), generator=g)
            for k in range(2, 7)}


def _boxes(g, n, h, w, lo=12.0, hi=120.0):
    cx = torch.rand(n, generator=g) * w
    cy = torch.rand(n, generator=g) * h
    bw = torch.exp(torch.rand(n, generator=g) * (math.log(hi) - math.log(lo)) + math.log(lo))
    bh = bw * (0.15 + 0.5 * torch.rand(n, generator=g))
    a = torch.rand(n, generator=g) * 360 - 180
    return torch.stack((cx, cy, bw, bh, a), 1)


def test_box_branch_matches_oracle(glass_lib):
    from glass_text_spotting_b200 import ops
    from glass_text_spotting_b200.modeling.roi_heads import B200GlassROIHeads
    from oracle import model as om
    o = om.build_oracle(seed=0)
    with torch.no_grad():  # spread the scores so that some candidates pass / fail the 0.05 threshold and NMS bites
        o.roi_heads.box_predictor.cls_score.weight.mul_(40.0)
        o.roi_heads.box_predictor.bbox_pred.weight.mul_(20.0)
    g = torch.Generator().manual_seed(4)
    h, w, n, per = 192, 256, 2, 100
    feats = _feats(g, n, h, w)
    props = torch.stack([_boxes(g, per, h, w) for _ in range(n)])
    props[1, 60:] = 0  # image 1 has only 60 proposals
    counts = torch.tensor([100, 60], dtype=torch.int32)
    heads = B200GlassROIHeads(o.state_dict())
    taps = {}
    hw = torch.tensor([[h, w]] * n, dtype=torch.float32).cuda()
    det = heads.forward_box({k: ops.Act.from_nchw(v.cuda()) for k, v in feats.items()}, props.cuda().contiguous(),
                            counts.cuda(), hw, taps)
    torch.cuda.synchronize()
    for i in range(n):
        c = int(counts[i])
        ot = {}
        with torch.no_grad():
            want = o.box_branch({k: v[i:i + 1] for k, v in feats.items()}, props[i, :c], (h, w), ot)
        rows = slice(i * per, i * per + c)
        pooled = (taps["box_pooled"][0, rows].float() + taps["box_pooled"][1, rows].float()).cpu() / ops.ACT_SCALE
        close(pooled.view(c, 7, 7, 256).permute(0, 3, 1, 2), ot["box_pooled"], f"img{i} box_pooled")
        xh = (taps["box_head_out"][0, rows].float() + taps["box_head_out"][1, rows].float()).cpu() / ops.ACT_SCALE
        close(xh, ot["box_head_out"], f"img{i} box_head_out")
        pred = taps["box_pred"][rows].cpu()
        close(pred[:, 0:2], ot["cls_logits"], f"img{i} cls_logits")
        close(pred[:, 2:7], ot["box_deltas"], f"img{i} box_deltas")
        close(pred[:, 7:11], ot["orient_logits"], f"img{i} orient_logits")
        # decisions on the ORACLE's logits (teacher forcing): identical kept set
        pr = torch.zeros(per, 16)
        pr[:c, 0:2], pr[:c, 2:7], pr[:c, 7:11] = ot["cls_logits"], ot["box_deltas"], ot["orient_logits"]
        d2 = heads.box_inference(pr.cuda(), props[i:i + 1].cuda().contiguous(), counts[i:i + 1].cuda(), hw[i:i + 1])
        k = int(d2["count"][0])
        assert k == want["pred_boxes"].shape[0], (k, want["pred_boxes"].shape)
        assert torch.equal(d2["index"][0, :k].cpu().long(), want["kept_proposal_idx"])
        close(d2["pred_boxes"][0, :k], want["pred_boxes"], f"img{i} det boxes")
        close(d2["scores"][0, :k], want["scores"], f"img{i} det scores", atol=1e-6)
        close(d2["orientations"][0, :k], want["orientations"], f"img{i} orientations", atol=1e-6)
        # and end to end from our own logits: same count, boxes within tolerance
        k2 = int(det["count"][i])
        assert k2 == k
        close(det["pred_boxes"][i, :k2], want["pred_boxes"], f"img{i} det boxes (own logits)")


def test_recognizer_branch_matches_oracle(glass_lib):
    from glass_text_spotting_b200 import ops
    from glass_text_spotting_b200.modeling.roi_heads import B200GlassROIHeads
    from oracle import d2_ops
    from oracle import model as om
    o = om.build_oracle(seed=1)
    g = torch.Generator().manual_seed(7)
    h, w, n = 192, 256, 2
    images = torch.randint(0, 256, (n, 3, h, w), generator=g).float()
    mean = torch.tensor(o.cfg.pixel_mean).view(1, 3, 1, 1)
    norm = images - mean
    feats = _feats(g, n, h, w)
    boxes = [_boxes(g, 3, h, w, 30.0, 150.0), _boxes(g, 2, h, w, 30.0, 150.0)]
    rh = o.roi_heads
    # calibrate the BatchNorms of the local CNN and the recognizer CNN on these very inputs
    with torch.no_grad():
        crops_all = torch.cat([d2_ops.roi_pooler([norm[i:i + 1]], [boxes[i]], (128, 128), [1.0], 2) for i in range(n)])
        calibrate_bn(rh.hybrid_net, lambda: rh.hybrid_net(crops_all))
        fused_in = []
        for i in range(n):
            gmap = rh.recognizer_feature_fusion(feats["p2"][i:i + 1], feats["p3"][i:i + 1])
            G = d2_ops.roi_pooler([gmap], [boxes[i]], (8, 32), [0.25], 0)
            L = rh.hybrid_net(crops_all[sum(len(b) for b in boxes[:i]): sum(len(b) for b in boxes[:i + 1])])
            fused_in.append(rh.fusion_net(torch.cat((L, G), 1)))
        fi = torch.cat(fused_in)
        calibrate_bn(rh.recognizer_head.backbone, lambda: rh.recognizer_head.backbone(fi))
    heads = B200GlassROIHeads(o.state_dict())
    rois = torch.cat([torch.cat((torch.full((len(b), 1), float(i)), b), 1) for i, b in enumerate(boxes)]).contiguous()
    word_start = torch.tensor([0, 3, 5], dtype=torch.int32)
    taps = {}
    probs = heads.forward_recognizer(images.cuda(), (h, w), {k: ops.Act.from_nchw(v.cuda()) for k, v in feats.items()},
                                     rois.cuda(), word_start.cuda(), n, taps)
    torch.cuda.synchronize()
    off = 0
    for i in range(n):
        k = len(boxes[i])
        ot = {}
        with torch.no_grad():
            want = o.recognizer_branch(norm[i:i + 1], {kk: v[i:i + 1] for kk, v in feats.items()}, boxes[i], ot)
        sl = slice(off, off + k)
        close(taps["p2p3"].to_nchw()[i:i + 1], ot["p2p3"], f"img{i} p2p3")
        close(taps["crops"].to_nchw()[sl], ot["local_crops"], f"img{i} crops")
        fused = taps["fused"].to_nchw()[sl]
        close(fused[:, 256:], ot["global_feats"], f"img{i} global_feats")
        # free-running through the 31 convs of the local CNN (and everything after it from the device's own features):
        # scale-relative atol, literal misses reported; the stage-wise literal checks are in test_gpu_fullsize_parity.py
        close(fused[:, :256], ot["local_feats"], f"img{i} local_feats", scaled=True)
        close(taps["fusion_out"].to_nchw()[sl], ot["fusion_out"], f"img{i} fusion_out", scaled=True)
        close(taps["recog_cnn"].to_nchw()[sl], ot["recog_cnn"], f"img{i} recog_cnn", scaled=True)
        close(taps["encoder_out"].view(-1, 32, 256)[sl], ot["encoder_out"], f"img{i} encoder_out")
        steps = ot["decoder_steps"]
        close(taps["decoder_logits"][sl, :steps], ot["decoder_logits"][:, :steps], f"img{i} decoder_logits")
        close(taps["decoder_alpha"][sl, :steps], ot["decoder_alpha"][:, :steps], f"img{i} decoder_alpha", atol=1e-5)
        close(probs[sl], want, f"img{i} pred_text_prob", atol=1e-5)
        off += k


@pytest.mark.parametrize("i", range(4))
def test_box_inference_matches_reference_golden(glass_lib, i):
    """glass_box_decode + glass_nms_rotated (B200GlassROIHeads.box_inference) against golden vectors from the
    reference's OWN RotatedFastRCNNOutputs.inference (tests/golden/box_inference.pt): kept proposals and their order
    exactly, boxes / scores / orientations to fp32 rounding."""
    import os
    from golden_common import make_box_inference_inputs
    from glass_text_spotting_b200 import ops
    c = torch.load(os.path.join(os.path.dirname(__file__), "golden", "box_inference.pt"), weights_only=False)["cases"][i]
    logits, deltas, orient, proposals = make_box_inference_inputs(c["seed"], c["r"], c["hw"])
    r = c["r"]
    pred = torch.zeros(r, 16)
    pred[:, 0:2], pred[:, 2:7], pred[:, 7:11] = logits, deltas, orient
    hw = torch.tensor([list(c["hw"])], dtype=torch.float32).cuda()
    cand_b, cand_s, cand_o = ops.box_decode(pred.cuda().contiguous(), proposals.cuda().contiguous(), None, 1, r,
                                            (10.0, 10.0, 5.0, 5.0, 10.0))
    boxes, scores, index, count = ops.nms_rotated(cand_b, cand_s, 0.35, 100, img_hw=hw, clip=True, filter_empty=False,
                                                  score_thresh=0.05)
    k = int(count[0])
    assert k == len(c["kept"])
    # the reference drops non-finite rows first, so its kept indices count only the valid rows before them; the device
    # keeps every proposal in its slot (invalid ones get score -inf) and reports original indices
    from oracle import d2_ops
    valid = torch.isfinite(d2_ops.apply_deltas_rotated(deltas, proposals, (10.0, 10.0, 5.0, 5.0, 10.0))).all(1) & \
        torch.isfinite(torch.softmax(logits, -1)).all(1)
    compact = torch.cumsum(valid.long(), 0) - 1
    got_idx = index[0, :k].cpu().long()
    assert valid[got_idx].all()
    assert torch.equal(compact[got_idx], c["kept"])
    assert torch.allclose(boxes[0, :k].cpu(), c["pred_boxes"], rtol=1e-5, atol=1e-4)
    assert torch.allclose(scores[0, :k].cpu(), c["scores"], rtol=1e-5, atol=1e-6)
    got_o = cand_o[0][index[0, :k].long()].cpu()
    assert torch.equal(got_o[:, 0], c["orientations"][:, 0])
    assert torch.allclose(got_o[:, 1], c["orientations"][:, 1], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("i", [0, 1, 3])
def test_recognizer_branch_matches_reference_golden(glass_lib, i):
    """B200GlassROIHeads.forward_recognizer against golden vectors written by the reference's OWN
    ``MaskRotatedRecognizerHybridHead._forward_recognizer`` + ``RecognizerRCNNHeadV3`` (tests/golden/recognizer_branch.pt,
    tools/make_golden_recognizer_branch.py): the wiring of rows a9-a16 on the device, incl. the early break (case 3)."""
    import os
    from golden_common import force_eos_bias, make_recognizer_branch_inputs, seeded_fill
    from glass_text_spotting_b200 import ops
    from glass_text_spotting_b200.modeling.roi_heads import B200GlassROIHeads
    from oracle import model as om
    c = torch.load(os.path.join(os.path.dirname(__file__), "golden", "recognizer_branch.pt"), weights_only=False)["cases"][i]
    o = om.GlassOracle()     # only a container of correctly named parameters here; the golden is the checker
    for j, name in enumerate(("recognizer_feature_fusion", "hybrid_net", "fusion_net", "recognizer_head")):
        seeded_fill(getattr(o.roi_heads, name), 500 + 10 * c["seed"] + j)
    if c["eos"]:
        with torch.no_grad():
            force_eos_bias(o.roi_heads.recognizer_head.decoder.recognizer)
            o.roi_heads.recognizer_head.decoder.recognizer.decoder.fc.bias[0] += 2.2
    image, p2, p3, boxes = make_recognizer_branch_inputs(c["seed"], c["k"])
    # the golden image is already normalised: mean 0 / std 1 make the fused normalisation the identity
    heads = B200GlassROIHeads(o.state_dict(), pixel_mean=(0.0, 0.0, 0.0), pixel_std=(1.0, 1.0, 1.0))
    k = c["k"]
    rois = torch.cat((torch.zeros(k, 1), boxes), 1).contiguous().cuda()
    word_start = torch.tensor([0, k], dtype=torch.int32).cuda()
    feats = {"p2": ops.Act.from_nchw(p2.cuda()), "p3": ops.Act.from_nchw(p3.cuda())}
    probs = heads.forward_recognizer(image[None].contiguous().cuda(), tuple(image.shape[-2:]), feats, rois, word_start, 1)
    torch.cuda.synchronize()
    want = c["pred_text_prob"]
    assert tuple(probs.shape) == tuple(want.shape)
    assert torch.equal((probs.cpu().sum(2) > 0).sum(1), (want.sum(2) > 0).sum(1)), "early break at a different step"
    assert torch.equal(probs.cpu().argmax(2), want.argmax(2))
    close(probs, want, f"case {i} pred_text_prob vs the reference", atol=1e-5)


@pytest.mark.parametrize("i", [0, 2])
def test_box_branch_matches_reference_golden(glass_lib, i):
    """B200GlassROIHeads.forward_box against golden vectors written by the reference's OWN ``_forward_box`` +
    ``RotatedFastRCNNOutputLayers`` + ``RotatedFastRCNNOutputs.inference`` (tests/golden/box_branch.pt,
    tools/make_golden_box_branch.py): rows a5-a8 wired together on the device, free-running from its own logits."""
    import os
    from golden_common import make_box_branch_inputs, seeded_fill
    from glass_text_spotting_b200 import ops
    from glass_text_spotting_b200.modeling.roi_heads import B200GlassROIHeads
    from oracle import model as om
    c = torch.load(os.path.join(os.path.dirname(__file__), "golden", "box_branch.pt"), weights_only=False)["cases"][i]
    o = om.GlassOracle()     # a container of correctly named parameters; the golden is the checker
    seeded_fill(o.roi_heads.box_head, 700 + c["seed"])
    seeded_fill(o.roi_heads.box_predictor, 710 + c["seed"])
    with torch.no_grad():
        o.roi_heads.box_predictor.cls_score.weight.mul_(4.0)
        o.roi_heads.box_predictor.bbox_pred.weight.mul_(0.3)
    feats, proposals, hw = make_box_branch_inputs(c["seed"], c["r"])
    heads = B200GlassROIHeads(o.state_dict(), detections_per_image=c["detections"])
    det = heads.forward_box({k: ops.Act.from_nchw(v.cuda()) for k, v in feats.items()}, proposals[None].contiguous().cuda(),
                            torch.tensor([c["r"]], dtype=torch.int32).cuda(),
                            torch.tensor([list(hw)], dtype=torch.float32).cuda())
    torch.cuda.synchronize()
    k = int(det["count"][0])
    assert k == len(c["scores"])
    close(det["pred_boxes"][0, :k], c["pred_boxes"], f"case {i} boxes vs the reference")
    close(det["scores"][0, :k], c["scores"], f"case {i} scores vs the reference", atol=1e-5)
    assert torch.equal(det["orientations"][0, :k, 0].cpu(), c["orientations"][:, 0])
    close(det["orientations"][0, :k, 1], c["orientations"][:, 1], f"case {i} orientation prob", atol=1e-5)
