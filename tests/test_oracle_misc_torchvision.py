"""More independent cross-checks of detectron2-recalled pieces of the oracle against torchvision's implementations of
the same published components: the Mask R-CNN head (mask branch), the FPN level assignment of the RoI
pooler (A.5) and the batch padding of ImageList.from_tensors (A.1)."""
import torch


def test_mask_head_matches_torchvision_maskrcnn_head():
    """4 x (conv3x3 + ReLU) -> ConvTranspose2d(2, s2) + ReLU -> conv1x1, then sigmoid (mask_rcnn_inference)."""
    from torchvision.models.detection.mask_rcnn import MaskRCNNHeads, MaskRCNNPredictor
    from oracle import mask as om
    m = om.seeded_mask_head(3).eval()
    heads = MaskRCNNHeads(256, (256, 256, 256, 256), 1).eval()
    pred = MaskRCNNPredictor(256, 256, 1).eval()
    with torch.no_grad():
        for k in range(4):
            heads[k][0].weight.copy_(getattr(m, f"mask_fcn{k + 1}").weight)
            heads[k][0].bias.copy_(getattr(m, f"mask_fcn{k + 1}").bias)
        pred.conv5_mask.weight.copy_(m.deconv.weight)
        pred.conv5_mask.bias.copy_(m.deconv.bias)
        pred.mask_fcn_logits.weight.copy_(m.predictor.weight)
        pred.mask_fcn_logits.bias.copy_(m.predictor.bias)
        g = torch.Generator().manual_seed(4)
        x = torch.randn(5, 256, 14, 14, generator=g)
        want = torch.sigmoid(pred(heads(x)))
        got = m(x)
    assert tuple(got.shape) == (5, 1, 28, 28)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), float((got - want).abs().max())


def test_level_assignment_matches_torchvision_level_mapper():
    """floor(4 + log2(sqrt(area) / 224)) clamped to the pyramid; the two libraries place their epsilon differently, so the
    fixture stays a hair away from the level boundaries."""
    from torchvision.ops.poolers import LevelMapper
    from oracle import d2_ops
    g = torch.Generator().manual_seed(5)
    n = 4000
    side = torch.exp(torch.rand(n, generator=g) * 7.5)                        # 1 .. 1800 px
    frac = torch.log2(side / 224.0) % 1.0
    side = side[(frac > 1e-3) & (frac < 1 - 1e-3)]
    ratio = 0.2 + torch.rand(len(side), generator=g) * 3
    w, h = side * torch.sqrt(ratio), side / torch.sqrt(ratio)
    boxes5 = torch.stack((torch.zeros_like(w), torch.zeros_like(w), w, h, torch.rand(len(side), generator=g) * 360 - 180), 1)
    got = d2_ops.assign_boxes_to_levels(boxes5, 2, 6)
    mapper = LevelMapper(2, 6, canonical_scale=224, canonical_level=4)
    xyxy = torch.stack((-w / 2, -h / 2, w / 2, h / 2), 1)
    want = mapper([xyxy])
    assert torch.equal(got, want)
    assert set(got.tolist()) == {0, 1, 2, 3, 4}


def test_image_list_padding_matches_torchvision_batch_images():
    """ImageList.from_tensors(size_divisibility=32): top-left aligned, zero padded to the batch maximum rounded up to 32."""
    from torchvision.models.detection.transform import GeneralizedRCNNTransform
    from glass_text_spotting_b200.structures import ImageList
    from oracle import model as om
    g = torch.Generator().manual_seed(6)
    imgs = [torch.rand(3, 97, 130, generator=g), torch.rand(3, 150, 111, generator=g), torch.rand(3, 64, 64, generator=g)]
    t = GeneralizedRCNNTransform(1, 1, [0, 0, 0], [1, 1, 1], size_divisible=32)
    want = t.batch_images(imgs, size_divisible=32)
    got = ImageList.from_tensors(imgs, 32)
    assert torch.equal(got.tensor, want) and got.image_sizes == [(97, 130), (150, 111), (64, 64)]
    o = om.GlassOracle()
    canvas, sizes = o.preprocess_image([i * 255 for i in imgs])
    assert tuple(canvas.shape) == tuple(want.shape) and sizes == got.image_sizes
    mean = torch.tensor(o.cfg.pixel_mean).view(3, 1, 1)
    assert torch.allclose(canvas[1, :, :150, :111], imgs[1] * 255 - mean, atol=1e-4)
    assert float(canvas[1, :, 150:, :].abs().max()) == 0.0 and float(canvas[0, :, :, 130:].abs().max()) == 0.0
