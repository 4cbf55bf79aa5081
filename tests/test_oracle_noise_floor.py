"""Documents the rounding-noise floor of the fp32 reference path itself: the oracle's ResNet-50 evaluated
in fp32 vs fp64 on identical (BatchNorm-calibrated) random weights.  The deep taps' own fp32 noise is a large
fraction of the north-star tolerance, which is why GPU parity of the backbone is asserted stage-wise
(tests/test_gpu_backbone.py) -- a free-running comparison at res5 compares two equally noisy numbers."""
import torch

from parity_common import oracle_with_calibrated_backbone


def test_fp32_reference_noise_grows_with_depth():
    g = torch.Generator().manual_seed(3)
    images = torch.randint(0, 256, (1, 3, 128, 160), generator=g).float()
    o = oracle_with_calibrated_backbone(0, images)
    x = images - torch.tensor(o.cfg.pixel_mean).view(1, 3, 1, 1)
    with torch.no_grad():
        r32 = o.backbone.bottom_up(x)
        r64 = o.backbone.bottom_up.double()(x.double())
    rel = {k: ((r32[k].double() - r64[k]).norm() / r64[k].norm()).item() for k in ["res2", "res3", "res4", "res5"]}
    assert rel["res2"] < 1e-5
    assert rel["res5"] > 4 * rel["res2"], rel          # noise is amplified stage over stage
    assert rel["res5"] < 1e-3, rel
