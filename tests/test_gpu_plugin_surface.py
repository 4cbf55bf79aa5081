"""The detectron2 plugin surface of SURVEY.md 8(b), called the way the reference's GlassRCNN.inference calls it
(glass/modeling/meta_arch/glass_rcnn.py:82-101): ``proposal_generator(images, features, None)``,
``roi_heads(images, features, proposals, None)``, ``roi_heads.forward_with_given_boxes(images, features, instances)``
and ``inference(batched_inputs, detected_instances=...)``.  The fused path (``forward_device``) is what the parity
tests pin against the oracle; here the module-level surface must reproduce it exactly (same kernels, same order)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model_and_inputs(glass_lib):
    from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN
    from oracle import model as om
    K = 6
    cfg = om.HotPathConfig(max_detections_override=K)
    imgs = [om.synthetic_image(31, 160, 224), om.synthetic_image(32, 150, 201)]
    o = om.build_oracle(seed=5, calib_images=imgs[:1], cfg=cfg)
    model = B200GlassRCNN(o.state_dict(), detections_per_image=K, filter_small_boxes=2)
    return model, [{"image": imgs[0], "height": 320, "width": 448}, {"image": imgs[1]}]


def _snapshot(instances):
    return [{k: (v.tensor if hasattr(v, "tensor") else v).clone() for k, v in x.get_fields().items()} for x in instances]


def test_module_surface_reproduces_the_fused_path(model_and_inputs):
    model, inputs = model_and_inputs
    want = _snapshot(model.inference(inputs, do_postprocess=False))
    assert sum(len(w["scores"]) for w in want) > 0, "the seeded model must detect something for this test to bite"
    il = model.preprocess_image(inputs)
    feats = model.backbone(il.tensor)
    proposals, losses = model.proposal_generator(il, feats, None)
    assert losses == {} and len(proposals) == 2
    for p, size in zip(proposals, il.image_sizes):
        assert p.image_size == size and p.has("proposal_boxes") and p.has("objectness_logits")
        s = p.objectness_logits
        assert len(p) <= model.proposal_generator.post_nms_topk and bool((s[:-1] >= s[1:]).all())
    results, losses = model.roi_heads(il, feats, proposals, None)
    assert losses == {}
    for r, w in zip(results, want):
        assert torch.equal(r.pred_boxes.tensor, w["pred_boxes"]) and torch.equal(r.scores, w["scores"])
        assert torch.equal(r.orientations, w["orientations"]) and torch.equal(r.pred_classes, w["pred_classes"])
        assert torch.equal(r.pred_text_prob, w["pred_text_prob"])


def test_given_boxes_only_runs_the_recognizer(model_and_inputs):
    from glass_text_spotting_b200.structures import Instances, RotatedBoxes
    model, inputs = model_and_inputs
    raw = _snapshot(model.inference(inputs, do_postprocess=False))
    given = [Instances((int(i["image"].shape[-2]), int(i["image"].shape[-1])), pred_boxes=RotatedBoxes(r["pred_boxes"].cpu()),
                       pred_classes=r["pred_classes"].cpu()) for i, r in zip(inputs, raw)]
    out = model.inference(inputs, detected_instances=given, do_postprocess=False)
    for o, r in zip(out, raw):
        assert torch.equal(o.pred_boxes.tensor, r["pred_boxes"]) and torch.equal(o.pred_text_prob, r["pred_text_prob"])
        assert not o.has("scores")        # nothing but the per-RoI outputs is added (recognizers_hybrid_head.py:571-609)
    # an image without boxes next to one with boxes
    given[1] = given[1][torch.zeros(len(given[1]), dtype=torch.bool)]
    out = model.inference(inputs, detected_instances=given, do_postprocess=False)
    assert tuple(out[1].pred_text_prob.shape) == (0, 26, 97)
    assert torch.equal(out[0].pred_text_prob, raw[0]["pred_text_prob"])


def test_postprocess_options_apply_after_the_device_path(model_and_inputs):
    """MIN_BOX_DIMENSION (glass_rcnn.py:117-118) and the output-size rescale act on the raw results exactly like the
    host-side golden test (tests/test_oracle_meta_postprocess.py) says; here: consistency of the two entry points."""
    from glass_text_spotting_b200.modeling.glass_rcnn import detector_postprocess, filter_small_boxes
    model, inputs = model_and_inputs
    raw = model.inference(inputs, do_postprocess=False)
    want = [detector_postprocess(filter_small_boxes(r, 2), i.get("height", r.image_size[0]), i.get("width", r.image_size[1]))
            for r, i in zip(raw, inputs)]
    got = model(inputs)
    for g, w in zip(got, want):
        g = g["instances"]
        assert g.image_size == w.image_size and torch.equal(g.pred_boxes.tensor, w.pred_boxes.tensor)
        assert torch.equal(g.pred_text_prob, w.pred_text_prob)
        assert bool((torch.min(g.pred_boxes.tensor[:, 2], g.pred_boxes.tensor[:, 3]) > 0).all())


def test_checkpoint_without_orientation_head(model_and_inputs):
    """MODEL.ORIENTATION_ON False (configs/glass_finetune_textocr.yaml:106): no box_predictor.orientation_pred in the
    checkpoint, no ``orientations`` in the results (rotated_fast_rcnn.py:141-142, 547-549); everything else unchanged."""
    from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN
    from oracle import model as om
    model, inputs = model_and_inputs
    want = _snapshot(model.inference(inputs, do_postprocess=False))
    o = om.build_oracle(seed=5, calib_images=[inputs[0]["image"]], cfg=om.HotPathConfig(max_detections_override=6))
    sd = {k: v for k, v in o.state_dict().items() if "orientation_pred" not in k}
    assert len(sd) == len(o.state_dict()) - 2
    plain = B200GlassRCNN(sd, detections_per_image=6)
    assert plain.roi_heads.orientation_on is False and model.roi_heads.orientation_on is True
    got = plain.inference(inputs, do_postprocess=False)
    for g, w in zip(got, want):
        assert not g.has("orientations") and "orientations" in w
        assert torch.equal(g.pred_boxes.tensor, w["pred_boxes"]) and torch.equal(g.scores, w["scores"])
        assert torch.equal(g.pred_text_prob, w["pred_text_prob"])


def test_registry_adapters_reproduce_the_fused_model(glass_lib):
    """d2_adapter: modules built BY NAME from a glass_pretrain.yaml-shaped cfg through (stub) detectron2 registries,
    weights loaded afterwards (DetectionCheckpointer's order).  The fused meta-arch adapter must equal B200GlassRCNN bit
    for bit; the three component adapters, driven the way d2's stock GeneralizedRCNN drives them (NORMALISED, zero-padded
    images), must give the same detections (the normalisation happens in a different place, so the pixels differ in the
    last bit)."""
    from test_config_cpu import PRETRAIN_LIKE
    from test_d2_adapter_cpu import _registries
    from glass_text_spotting_b200 import d2_adapter as ad, weights
    from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN
    sd = weights.random_state_dict(0)
    g = torch.Generator().manual_seed(21)
    inputs = [{"image": torch.randint(0, 256, (3, 200, 280), generator=g).float()},
              {"image": torch.randint(0, 256, (3, 224, 256), generator=g).float()}]
    regs = ad.register_all(**_registries(), generalized_rcnn=True)
    want = B200GlassRCNN(sd, detections_per_image=100)(inputs)
    model = ad.build_model(PRETRAIN_LIKE, regs)
    model.load_state_dict(sd)
    got = model(inputs)
    comp = ad.ComponentRCNN(PRETRAIN_LIKE, regs)
    comp.load_state_dict(sd)
    got_c = comp.inference(inputs)
    torch.cuda.synchronize()
    for w, a, c in zip(want, got, got_c):
        w, a = w["instances"], a["instances"]
        assert len(w) == len(a) and len(w) > 0
        assert torch.equal(w.pred_boxes.tensor, a.pred_boxes.tensor) and torch.equal(w.pred_text_prob, a.pred_text_prob)
        assert len(c) == len(w)
        assert (c.pred_boxes.tensor - w.pred_boxes.tensor).abs().max().item() < 1e-2
        assert (c.pred_text_prob - w.pred_text_prob).abs().max().item() < 1e-3
