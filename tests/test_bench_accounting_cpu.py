"""The accounting behind bench.py's roofline numbers: algorithmic bytes of the RoIAlign microbench and
algorithmic FLOPs of the backbone must be what the definitions say, on cases small enough to count by hand."""
import importlib.util
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_roialign_algorithmic_bytes_by_hand():
    b = _bench()
    # a 56 x 56 px box goes to p2 (floor(4 + log2(56 / 224)) = 2, stride 4): 14 x 14 feature cells, 14 x 14 sample points
    # one cell apart that land on integer coordinates -> taps on 15 x 15 distinct cells of 256 fp32 channels
    want = 225 * 256 * 4 + 256 * 49 * 4 + 6 * 4
    for angle in (0.0, 90.0, 180.0):
        assert b._roialign_algorithmic_bytes(torch.tensor([[0, 400.0, 600.0, 56.0, 56.0, angle]])) == want
    # two identical RoIs count twice (per-RoI unique cells, as the definition says); a box off the image reads nothing
    twice = b._roialign_algorithmic_bytes(torch.tensor([[0, 400.0, 600.0, 56.0, 56.0, 0.0]] * 2))
    assert twice == 2 * want
    off = b._roialign_algorithmic_bytes(torch.tensor([[0, -500.0, -500.0, 56.0, 56.0, 0.0]]))
    assert off == 256 * 49 * 4 + 6 * 4
    # the cfg-3 set: 512 RoIs, deterministic
    _, rois = b._roialign_inputs()
    assert tuple(rois.shape) == (512, 6) and b._roialign_algorithmic_bytes(rois) == 148585472


def test_backbone_flops_by_hand():
    from glass_text_spotting_b200.modeling.backbone import B200ResNetFPN
    f = B200ResNetFPN.flops_per_image(1024, 1024)
    assert abs(f / 1e9 - 279.94) < 0.05          # SURVEY.md 8d cfg 2: ResNet 161.16 + FPN 118.78 GFLOP
    # the stem alone: 7x7x3 -> 64 at 512 x 512
    stem = 2 * 3 * 64 * 49 * 512 * 512
    assert f > stem and abs(B200ResNetFPN.flops_per_image(512, 512) * 4 - f) / f < 1e-9
