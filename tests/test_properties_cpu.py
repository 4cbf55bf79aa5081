"""Size-independent properties of the host-side glue (hypothesis): the rotated-box structure against the oracle's
restatement of detectron2's RotatedBoxes, the GlassRCNN post-processing chain, and the image sharding."""
import math

import torch
from hypothesis import given, settings, strategies as st

boxes_st = st.lists(
    st.tuples(st.floats(-50, 700), st.floats(-50, 500), st.floats(0, 300), st.floats(0, 200), st.floats(-400, 400)),
    min_size=0, max_size=12)


def _t(rows):
    return torch.tensor(rows, dtype=torch.float32).reshape(-1, 5)


@settings(max_examples=60, deadline=None)
@given(boxes_st, st.integers(16, 640), st.integers(16, 480))
def test_clip_matches_oracle_and_is_idempotent(rows, w, h):
    from glass_text_spotting_b200.structures import RotatedBoxes
    from oracle import d2_ops
    b = RotatedBoxes(_t(rows).clone())
    b.clip((h, w))
    want = d2_ops.clip_rotated_(_t(rows).clone(), (h, w))
    assert torch.equal(b.tensor, want)
    again = RotatedBoxes(b.tensor.clone())
    again.clip((h, w))
    assert torch.allclose(again.tensor, b.tensor, atol=1e-4)
    assert bool(((b.tensor[:, 4] >= -180) & (b.tensor[:, 4] < 180)).all())


@settings(max_examples=60, deadline=None)
@given(boxes_st, st.floats(0.25, 4.0), st.floats(0.25, 4.0))
def test_scale_matches_oracle_and_preserves_area_ratio(rows, sx, sy):
    from glass_text_spotting_b200.structures import RotatedBoxes
    from oracle import d2_ops
    b = RotatedBoxes(_t(rows).clone())
    b.scale(sx, sy)
    want = d2_ops.scale_rotated_(_t(rows).clone(), sx, sy)
    assert torch.allclose(b.tensor, want, rtol=1e-6, atol=1e-5)
    if sx == sy:      # isotropic scaling keeps the angle and multiplies both sides by the factor
        src = _t(rows)
        assert torch.allclose(b.tensor[:, 2:4], src[:, 2:4] * sx, rtol=1e-5, atol=1e-4)


@settings(max_examples=40, deadline=None)
@given(boxes_st, st.sampled_from([None, 1.0, 2.0, 5.0]), st.sampled_from([None, 0.05, 0.2]),
       st.tuples(st.integers(32, 400), st.integers(32, 400)))
def test_postprocess_chain_matches_oracle(rows, min_dim, inflate, out_hw):
    """B200GlassRCNN._postprocess (host glue) == the oracle's restatement of glass_rcnn.py:103-128 for any input."""
    from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN
    from glass_text_spotting_b200.structures import Instances, RotatedBoxes
    from oracle import postprocess as pp
    src = _t(rows)
    hw = (300, 400)
    model = B200GlassRCNN.__new__(B200GlassRCNN)
    model.filter_small_boxes, model.inflate_ratio = min_dim, inflate
    inst = Instances(hw, pred_boxes=RotatedBoxes(src.clone()), idx=torch.arange(len(src)))
    out = model._postprocess([inst], [{"height": out_hw[0], "width": out_hw[1]}], [hw])[0]["instances"]
    want_b, want_i = pp.glass_rcnn_postprocess(src, hw, out_hw[0], out_hw[1], min_dim, inflate)
    assert torch.equal(out.idx, want_i)
    assert torch.allclose(out.pred_boxes.tensor, want_b, rtol=1e-6, atol=1e-5)
    assert bool((out.pred_boxes.tensor[:, 2:4] > 0).all())


@settings(max_examples=100, deadline=None)
@given(st.integers(0, 500), st.integers(1, 16))
def test_shards_partition_the_dataset(n, world):
    from glass_text_spotting_b200 import parallel
    r = [parallel.shard_range(n, k, world) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
    sizes = [e - b for b, e in r]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert max(sizes) == math.ceil(n / world)


@settings(max_examples=30, deadline=None)
@given(st.lists(st.integers(0, 5), min_size=1, max_size=4), st.integers(0, 2 ** 31 - 1))
def test_pack_detections_roundtrip(counts, seed):
    """pack_detections (the fixed-size all-gather record) -> parallel.unpack_detections /
    evaluation.instances_from_packed gives back exactly the detections, for any per-image counts incl. zero."""
    from glass_text_spotting_b200 import evaluation, parallel
    from parity_common import pack_detections_reference
    steps, nc, max_det = 4, 7, 5
    g = torch.Generator().manual_seed(seed)
    n = len(counts)
    det = {"pred_boxes": torch.rand(n, max_det, 5, generator=g) * 100, "scores": torch.rand(n, max_det, generator=g),
           "orientations": torch.rand(n, max_det, 2, generator=g)}
    starts = [0]
    for c in counts:
        starts.append(starts[-1] + c)
    probs = torch.rand(starts[-1], steps, nc, generator=g)
    rec = pack_detections_reference(det, probs, counts, starts)   # the layout glass_pack_detections writes on the device
    assert tuple(rec.shape) == (n, max_det, 10 + steps * nc)
    dets = parallel.unpack_detections(rec, steps, nc)
    insts = evaluation.instances_from_packed(rec, [(64, 80)] * n, steps, nc)
    for i, c in enumerate(counts):
        assert len(dets[i]["scores"]) == c == len(insts[i])
        assert torch.equal(dets[i]["pred_boxes"], det["pred_boxes"][i, :c])
        assert torch.equal(insts[i].pred_boxes.tensor, det["pred_boxes"][i, :c])
        assert torch.equal(insts[i].scores, det["scores"][i, :c])
        assert torch.equal(insts[i].orientations, det["orientations"][i, :c])
        assert torch.equal(insts[i].pred_text_prob, probs[starts[i]: starts[i + 1]])
        assert bool((rec[i, c:] == 0).all())


@settings(max_examples=80, deadline=None)
@given(st.floats(-170, 170), st.floats(2, 300), st.floats(2, 200), st.floats(0.3, 3.0), st.floats(0.3, 3.0))
def test_scale_is_the_image_of_the_box_edges(angle, w, h, sx, sy):
    """RotatedBoxes.scale from first principles (no second implementation exists to compare with): under the map
    (x, y) -> (sx x, sy y) the new width / height are the lengths of the images of the box's width / height edge and the
    new angle is the direction of the image of the height edge (detectron2's definition)."""
    from oracle import d2_ops
    th = math.radians(angle)
    c, s = math.cos(th), math.sin(th)
    # image coordinates (y down), angle counter-clockwise: width edge along (c, -s), height edge along (s, c)
    we = (sx * c * w, -sy * s * w)
    he = (sx * s * h, sy * c * h)
    b = torch.tensor([[10.0, 20.0, w, h, angle]])
    d2_ops.scale_rotated_(b, sx, sy)
    assert abs(b[0, 0].item() - 10.0 * sx) < 1e-4 and abs(b[0, 1].item() - 20.0 * sy) < 1e-4
    assert abs(b[0, 2].item() - math.hypot(*we)) < 1e-3 * max(1.0, math.hypot(*we))
    assert abs(b[0, 3].item() - math.hypot(*he)) < 1e-3 * max(1.0, math.hypot(*he))
    want = math.degrees(math.atan2(he[0], he[1]))
    diff = (b[0, 4].item() - want + 180.0) % 360.0 - 180.0
    assert abs(diff) < 1e-3
