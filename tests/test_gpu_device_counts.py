"""The step without a host round trip: device-side word counts, glass_pack_rois /
glass_pack_detections, launch plans, weight pre-packing through the ABI, and the whole step as one CUDA graph."""
import pytest
import torch

from parity_common import pack_detections_reference

pytestmark = pytest.mark.gpu


def test_prepack_weights_equals_host_rule(glass_lib):
    """glass_prepack_weights (device) and packing._pack_rows (its host restatement) give the same bits, incl. rows whose
    largest magnitude is an exact power of two, all-zero rows and tiny rows."""
    from glass_text_spotting_b200 import ops, packing
    g = torch.Generator().manual_seed(0)
    w = torch.randn(96, 320, generator=g) * torch.exp(torch.randn(96, 1, generator=g) * 4)
    w[3] = 0
    w[5] = 0
    w[5, 7] = 0.25          # amax an exact power of two
    w[6] *= 1e-12
    w[7, :] = torch.where(w[7].abs() > 2.0, torch.full_like(w[7], 2.0), w[7])
    w[7, 0] = -2.0
    scale = torch.rand(96, generator=g) + 0.5
    hp, hs = packing._pack_rows(w, scale, "cpu")
    dp, ds = ops.prepack_weights(w.cuda().contiguous(), scale)
    assert torch.equal(dp.cpu().view(torch.int16), hp.view(torch.int16))
    assert torch.equal(ds.cpu(), hs)


def test_pack_rois_and_detections(glass_lib):
    from glass_text_spotting_b200 import ops
    g = torch.Generator().manual_seed(1)
    n, m, steps, nc = 4, 100, 26, 97
    for counts in ([100, 0, 37, 100], [0, 0, 0, 0], [1, 2, 3, 4]):
        boxes = torch.rand(n, m, 5, generator=g) * 100
        scores = torch.rand(n, m, generator=g)
        orient = torch.rand(n, m, 2, generator=g)
        c = torch.tensor(counts, dtype=torch.int32)
        starts = [0]
        for k in counts:
            starts.append(starts[-1] + k)
        rois = torch.full((n * m, 6), -7.0).cuda()
        ws = torch.zeros(n + 1, dtype=torch.int32).cuda()
        tot = torch.zeros(1, dtype=torch.int32).cuda()
        ops.pack_rois(boxes.cuda(), c.cuda(), rois, ws, tot)
        assert ws.cpu().tolist() == starts and int(tot) == starts[-1]
        want = torch.cat([torch.cat((torch.full((k, 1), float(i)), boxes[i, :k]), 1) for i, k in enumerate(counts)])
        assert torch.equal(rois.cpu()[: starts[-1]], want) and bool((rois.cpu()[starts[-1]:] == 0).all())
        probs = torch.rand(n * m, steps, nc, generator=g)        # capacity-sized, rows >= total are stale
        rec = torch.full((n, m, 10 + steps * nc), 3.0).cuda()
        ops.pack_detections(boxes.cuda(), scores.cuda(), orient.cuda(), c.cuda(), probs.cuda(), ws, rec)
        det = {"pred_boxes": boxes, "scores": scores, "orientations": orient}
        assert torch.equal(rec.cpu(), pack_detections_reference(det, probs[: starts[-1]], counts, starts))


def test_conv_gemm_device_side_m_count(glass_lib):
    """The GEMM's M extent read from device memory: rows of the live words are computed exactly as with a host-sized
    launch, rows past them are not touched; a count of zero launches nothing harmful."""
    from glass_text_spotting_b200 import ops, packing
    g = torch.Generator().manual_seed(2)
    cap, live = 9, 5
    x = torch.randn(cap, 64, 16, 33, generator=g)
    w = packing.pack_conv(torch.randn(128, 64, 3, 3, generator=g) * 0.05, torch.rand(128, generator=g) + 0.5,
                          torch.randn(128, generator=g), (1, 1), (1, 1))
    a = ops.Act.from_nchw(x.cuda())
    want = ops.conv2d(ops.Act.from_nchw(x[:live].cuda()), w, relu=True).to_nchw()
    for count in (live, 0, cap, cap + 3):
        out = ops.Act(cap, 128, 16, 33)
        out.buf.fill_(1.0)
        n_dev = torch.tensor([count], dtype=torch.int32).cuda()
        ops.conv2d(a, w, relu=True, out=out, n_dev=n_dev)
        torch.cuda.synchronize()
        k = min(count, cap)
        got = out.to_nchw()
        if k >= live:
            assert torch.equal(got[:live], want)
        # words past the live count are not computed; the tile that straddles the end may clear (zero) rows of the
        # first non-live word, nothing else is written
        assert bool((out.buf[:, k + 1:] == 1.0).all()), "words past the live count were written"
        if k < cap:
            t = out.buf[:, k]
            assert bool(((t == 1.0) | (t == 0.0)).all())


def test_plan_cache_reuses_plans(glass_lib):
    from glass_text_spotting_b200 import ops, packing
    g = torch.Generator().manual_seed(3)
    x = ops.Act.from_nchw(torch.randn(2, 64, 12, 20, generator=g).cuda())
    w = packing.pack_conv(torch.randn(64, 64, 1, 1, generator=g) * 0.1)
    out = ops.Act(2, 64, 12, 20)
    ops.clear_plans()
    ops.conv2d(x, w, out=out)
    n1 = len(ops._PLANS)
    first = out.buf.clone()
    for _ in range(3):
        ops.conv2d(x, w, out=out)
    assert len(ops._PLANS) == n1 == 1 and torch.equal(out.buf, first)
    ops.conv2d(x, w, relu=True, out=out)          # another parameter block -> another plan
    assert len(ops._PLANS) == 2


@pytest.fixture(scope="module")
def model_and_batch(glass_lib):
    from glass_text_spotting_b200 import weights
    from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN
    g = torch.Generator().manual_seed(77)
    images = torch.randint(0, 256, (2, 3, 320, 384), generator=g, dtype=torch.uint8).float().cuda()
    hw = torch.tensor([[320, 384], [300, 360]], dtype=torch.float32).cuda()
    return B200GlassRCNN(weights.random_state_dict(0)), images, hw


def test_forward_packed_equals_host_sized_path(model_and_batch):
    """Capacity-sized launches with the word count on the device == the recognizer sized on the host with the exact K
    (the plugin surface's path): same records bit for bit."""
    model, images, hw = model_and_batch
    rec, det, probs, word_start = model.forward_packed(images, hw)
    torch.cuda.synchronize()
    rec = rec.clone()
    counts = det["count"].cpu().tolist()
    assert sum(counts) > 10 and word_start.cpu().tolist()[-1] == sum(counts)
    # host-sized reference: the same detections through forward_recognizer with K known
    feats, det2 = model.detect(images, hw)
    rois = torch.cat([torch.cat((torch.full((c, 1), float(i), device="cuda"), det2["pred_boxes"][i, :c]), 1)
                      for i, c in enumerate(counts)]).contiguous()
    starts = [0]
    for c in counts:
        starts.append(starts[-1] + c)
    ws = torch.tensor(starts, dtype=torch.int32).cuda()
    probs2 = model.roi_heads.forward_recognizer(images, tuple(images.shape[-2:]), feats, rois, ws, 2)
    rec2 = model.pack_detections(det2, probs2, counts, starts)
    torch.cuda.synchronize()
    assert torch.equal(rec, rec2)


def test_graph_step_equals_eager(model_and_batch):
    model, images, hw = model_and_batch
    want = model.forward_packed(images, hw)[0].clone()
    got = model.graph_step(images, hw).clone()
    got2 = model.graph_step(images, hw).clone()          # replay
    torch.cuda.synchronize()
    assert torch.equal(got, want) and torch.equal(got2, want)
    assert model.last_graph["launches"] > 100
    # another batch through the same graph
    other = images.flip(0).contiguous()
    want_o = model.forward_packed(other, hw)[0].clone()
    got_o = model.graph_step(other, hw).clone()
    torch.cuda.synchronize()
    assert torch.equal(got_o, want_o)
