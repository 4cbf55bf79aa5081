"""One small invocation of the hot path on cuda:0, checked against the oracle (test infrastructure:
only this smoke check, tests/ and bench.py's CPU legs may import ``oracle``)."""
import torch


def run() -> None:
    from . import lib, ops
    from .modeling.backbone import B200ResNetFPN
    from oracle import d2_ops
    from oracle import model as om
    lib.load()
    torch.cuda.set_device(0)
    g = torch.Generator().manual_seed(0)
    images = torch.randint(0, 256, (1, 3, 128, 160), generator=g).float()
    o = om.build_oracle(seed=0)
    mean = torch.tensor(o.cfg.pixel_mean).view(1, 3, 1, 1)
    bns = [m for m in o.backbone.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    with torch.no_grad():
        for m in bns:  # calibrate BatchNorm on this input: plain random init blows activations up
            m.train()
            m.momentum = 1.0
        o.backbone(images - mean)
        for m in bns:
            m.eval()
        ref = o.backbone(images - mean)
    bb = B200ResNetFPN(o.state_dict())
    got = bb(images.cuda())
    worst = 0.0
    for k in ["res2", "res3"]:  # strict north-star tolerance on the shallow taps
        a, b = got[k].to_nchw().cpu(), ref[k]
        err = ((a - b).abs() / (1e-4 * max(1.0, b.abs().max().item()) + 1e-3 * b.abs())).max().item()
        worst = max(worst, err)
        assert err <= 1.0, f"smoke: {k} out of tolerance ({err:.2f}x)"
    for k in ["p2", "p3", "p4", "p5", "p6"]:  # free-running deep taps: relative-L2 bound
        a, b = got[k].to_nchw().cpu(), ref[k]
        rel = ((a - b).norm() / b.norm()).item()
        assert rel < 1e-3, f"smoke: {k} relL2 {rel:.2e}"
    rois = torch.tensor([[0, 60.0, 50.0, 80.0, 30.0, 25.0], [0, 100.0, 90.0, 40.0, 20.0, -70.0]])
    feats = [got[k].to_nchw().cpu() for k in ["p2", "p3", "p4", "p5", "p6"]]  # same inputs for both sides
    want = d2_ops.roi_pooler(feats, [rois[:, 1:]], 7, [1 / 4, 1 / 8, 1 / 16, 1 / 32, 1 / 64], 2)
    pooled = ops.roi_align_rotated([got[k] for k in ["p2", "p3", "p4", "p5", "p6"]], rois.cuda(), (7, 7),
                                   [1 / 4, 1 / 8, 1 / 16, 1 / 32, 1 / 64], 2).permute(0, 3, 1, 2).cpu()
    assert torch.allclose(pooled, want, rtol=1e-3, atol=1e-4 * max(1.0, want.abs().max().item())), "smoke: RoIAlign"
    print(f"smoke ok: res2/res3 within {worst:.3f} of tolerance, p2..p6 relL2 < 1e-3, rotated RoIAlign ok, "
          f"{lib.load().glass_launch_count()} kernel launches")
