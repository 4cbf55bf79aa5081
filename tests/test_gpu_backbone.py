"""GPU parity: B200ResNetFPN (tcgen05 conv stack) vs the oracle's ResNet-50+FPN on identical seeded
weights (taps T1/T2), rtol=1e-3 / atol=1e-4."""
import pytest
import torch

from parity_common import close, oracle_with_calibrated_backbone

pytestmark = pytest.mark.gpu


def test_backbone_matches_oracle(glass_lib):
    from glass_text_spotting_b200.modeling.backbone import B200ResNetFPN
    g = torch.Generator().manual_seed(3)
    images = torch.randint(0, 256, (2, 3, 192, 256), generator=g).float()
    o = oracle_with_calibrated_backbone(0, images)
    mean = torch.tensor(o.cfg.pixel_mean).view(1, 3, 1, 1)
    with torch.no_grad():
        ref = o.backbone(images - mean)
    bb = B200ResNetFPN(o.state_dict())
    got = bb(images.cuda())
    torch.cuda.synchronize()
    for k in ["res2", "res3", "res4", "res5", "p5", "p4", "p3", "p2", "p6"]:
        close(got[k].to_nchw(), ref[k], k)
    # second run reuses the workspace: results identical, no new buffers
    nb = bb.ws.nbytes()
    got2 = bb(images.cuda())
    assert bb.ws.nbytes() == nb
    assert torch.equal(got2["p2"].buf, got["p2"].buf)
