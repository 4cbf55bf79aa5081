"""Independent cross-check of the oracle's detectron2-recalled backbone wiring (rows a2-a3).

detectron2 cannot be installed offline, so the oracle RESTATES its ResNet-50 + FPN from the published semantics.  The
same published architectures exist as an independent implementation in torchvision (present in this image):
``torchvision.models.resnet50`` (stem, bottleneck wiring, downsample shortcut) and
``torchvision.ops.FeaturePyramidNetwork`` + ``LastLevelMaxPool`` (lateral 1x1, nearest top-down add, 3x3 output conv,
p6 = stride-2 subsample; with a norm layer the convs lose their bias, like detectron2's FPN with ``NORM`` set).
The only architectural difference is detectron2's ``STRIDE_IN_1X1: True`` (stride on the first 1x1 of a stage's first
block instead of torchvision's 3x3) -- set on the torchvision modules by moving the stride attribute.  The oracle's
weights are loaded into the torchvision modules by name mapping and the feature maps must agree to fp32 rounding."""
import pytest
import torch
import torchvision
from torchvision.ops import FeaturePyramidNetwork
from torchvision.ops.feature_pyramid_network import LastLevelMaxPool


def _copy_conv_bn(conv, bn, src, prefix):
    conv.weight.data.copy_(src[prefix + ".weight"])
    assert conv.bias is None and (prefix + ".bias") not in src
    for k in ("weight", "bias", "running_mean", "running_var"):
        getattr(bn, k).data.copy_(src[prefix + ".norm." + k])


@pytest.fixture(scope="module")
def oracle_backbone():
    from golden_common import seeded_fill
    from oracle import nets
    bb = nets.ResNetFPN().eval()
    seeded_fill(bb, 77)
    return bb


def test_resnet50_matches_torchvision_with_stride_in_1x1(oracle_backbone):
    src = oracle_backbone.bottom_up.state_dict()
    tv = torchvision.models.resnet50(weights=None).eval()
    _copy_conv_bn(tv.conv1, tv.bn1, src, "stem.conv1")
    for li, stage in enumerate(["res2", "res3", "res4", "res5"], start=1):
        layer = getattr(tv, f"layer{li}")
        for b, blk in enumerate(layer):
            p = f"{stage}.{b}"
            _copy_conv_bn(blk.conv1, blk.bn1, src, p + ".conv1")
            _copy_conv_bn(blk.conv2, blk.bn2, src, p + ".conv2")
            _copy_conv_bn(blk.conv3, blk.bn3, src, p + ".conv3")
            if blk.downsample is not None:
                _copy_conv_bn(blk.downsample[0], blk.downsample[1], src, p + ".shortcut")
            else:
                assert (p + ".shortcut.weight") not in src
            if blk.conv2.stride == (2, 2):      # STRIDE_IN_1X1: the stride sits on conv1, not on the 3x3
                blk.conv1.stride, blk.conv2.stride = (2, 2), (1, 1)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 96, 128, generator=g) * 40
    with torch.no_grad():
        got = oracle_backbone.bottom_up(x)
        t = tv.maxpool(tv.relu(tv.bn1(tv.conv1(x))))
        want = {}
        for li, stage in enumerate(["res2", "res3", "res4", "res5"], start=1):
            t = getattr(tv, f"layer{li}")(t)
            want[stage] = t
    for k in want:
        assert got[k].shape == want[k].shape
        rel = float((got[k] - want[k]).norm() / want[k].norm())
        assert rel < 2e-6, (k, rel)
    assert tuple(got["res5"].shape) == (2, 2048, 3, 4)


def test_fpn_matches_torchvision(oracle_backbone):
    src = oracle_backbone.state_dict()
    fpn = FeaturePyramidNetwork([256, 512, 1024, 2048], 256, extra_blocks=LastLevelMaxPool(),
                                norm_layer=torch.nn.BatchNorm2d).eval()
    for i, k in enumerate([2, 3, 4, 5]):
        _copy_conv_bn(fpn.inner_blocks[i][0], fpn.inner_blocks[i][1], src, f"fpn_lateral{k}")
        _copy_conv_bn(fpn.layer_blocks[i][0], fpn.layer_blocks[i][1], src, f"fpn_output{k}")
    g = torch.Generator().manual_seed(6)
    c = {f"res{k}": torch.randn(2, ch, 96 // 2 ** k, 128 // 2 ** k, generator=g)
         for k, ch in zip([2, 3, 4, 5], [256, 512, 1024, 2048])}
    from collections import OrderedDict
    with torch.no_grad():
        want = fpn(OrderedDict((k, v) for k, v in c.items()))
        # the oracle's FPN on the same bottom-up maps (its forward recomputes them from an image, so replay the top-down)
        o = oracle_backbone
        prev = o.fpn_lateral5(c["res5"])
        got = {"p5": o.fpn_output5(prev)}
        for k in [4, 3, 2]:
            prev = getattr(o, f"fpn_lateral{k}")(c[f"res{k}"]) + torch.nn.functional.interpolate(prev, scale_factor=2.0, mode="nearest")
            got[f"p{k}"] = getattr(o, f"fpn_output{k}")(prev)
        got["p6"] = torch.nn.functional.max_pool2d(got["p5"], kernel_size=1, stride=2, padding=0)
    names = dict(zip(["p2", "p3", "p4", "p5", "p6"], list(want.keys())))
    for p, tvk in names.items():
        assert got[p].shape == want[tvk].shape, p
        assert torch.allclose(got[p], want[tvk], rtol=1e-5, atol=1e-5), (p, float((got[p] - want[tvk]).abs().max()))


def test_oracle_forward_is_its_own_replay(oracle_backbone):
    """The top-down replay used above IS ResNetFPN.forward (guards the test against drifting from the oracle)."""
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 3, 64, 96, generator=g) * 40
    with torch.no_grad():
        out = oracle_backbone(x)
        o = oracle_backbone
        prev = o.fpn_lateral5(out["res5"])
        p5 = o.fpn_output5(prev)
        prev = o.fpn_lateral4(out["res4"]) + torch.nn.functional.interpolate(prev, scale_factor=2.0, mode="nearest")
        p4 = o.fpn_output4(prev)
    assert torch.equal(out["p5"], p5) and torch.equal(out["p4"], p4)
    assert torch.equal(out["p6"], out["p5"][:, :, ::2, ::2])
