"""Shared helpers for the GPU parity tests: seeded oracle construction (test infrastructure)."""
import torch
from torch import nn

RTOL, ATOL = 1e-3, 1e-4


def close(got: torch.Tensor, ref: torch.Tensor, name="", rtol=RTOL, atol=ATOL):
    """north-star tolerance: |got-ref| <= atol*max(1,scale) + rtol*|ref| elementwise."""
    got, ref = got.detach().cpu().float(), ref.detach().cpu().float()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    scale = max(ref.abs().max().item(), 1e-6)
    err = (got - ref).abs()
    tol = atol * max(scale, 1.0) + rtol * ref.abs()
    bad = (err > tol).sum().item()
    assert bad == 0, f"{name}: {bad}/{err.numel()} out of tolerance, max err {err.max().item():.3e}, scale {scale:.3e}"
    return err.max().item()


@torch.no_grad()
def calibrate_bn(module: nn.Module, fwd):
    """BatchNorm running stats <- batch stats of one forward."""
    bns = [m for m in module.modules() if isinstance(m, nn.BatchNorm2d)]
    for m in bns:
        m.train()
        m.momentum = 1.0
    try:
        fwd()
    finally:
        for m in bns:
            m.eval()
            m.momentum = 0.1


def oracle_with_calibrated_backbone(seed: int, images: torch.Tensor):
    """Full GlassOracle with seeded weights; only the backbone's BN stats are calibrated (cheap)."""
    from oracle import model as om
    o = om.build_oracle(seed=seed)
    mean = torch.tensor(o.cfg.pixel_mean).view(1, 3, 1, 1)
    calibrate_bn(o.backbone, lambda: o.backbone(images - mean))
    return o
