"""Shared helpers for the GPU parity tests: seeded oracle construction and the tolerance check (test infrastructure).

Tolerance (BASELINE.json north_star): ``|got - ref| <= atol + rtol * |ref|`` with rtol = 1e-3, atol = 1e-4 -- the LITERAL
form, which is what ``close`` asserts by default.  A tap whose magnitude makes an absolute 1e-4 tighter than fp32
arithmetic itself can promise (the CPU oracle's own rounding noise vs an fp64 evaluation reaches it,
tests/test_oracle_noise_floor.py) may be checked with ``scaled=True``: the asserted bound then is
``atol * max(1, max|ref|) + rtol * |ref|`` and the number of elements that miss the LITERAL bound is still counted,
printed and recorded in ``REPORT`` (dumped to gpurun_out/parity_report.json by the tests' session hook) so the looser
form never hides what the literal one would have said."""
import json
import os

import torch
from torch import nn

RTOL, ATOL = 1e-3, 1e-4

# every close() call of the session: name, elements, max abs error, scale, misses of the literal / scaled bound
REPORT = []
# GLASS_PARITY_REPORT_ONLY=1: record instead of asserting (one GPU call then shows every tap's status at once)
REPORT_ONLY = os.environ.get("GLASS_PARITY_REPORT_ONLY", "0") == "1"


def close(got: torch.Tensor, ref: torch.Tensor, name="", rtol=RTOL, atol=ATOL, scaled: bool = False):
    """Assert the north-star tolerance elementwise; returns the max abs error.  See the module docstring for ``scaled``."""
    got, ref = got.detach().cpu().float(), ref.detach().cpu().float()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    scale = max(ref.abs().max().item(), 1e-6) if ref.numel() else 1.0
    err = (got - ref).abs()
    bad_literal = int((err > atol + rtol * ref.abs()).sum().item())
    bad_scaled = int((err > atol * max(scale, 1.0) + rtol * ref.abs()).sum().item())
    max_err = err.max().item() if err.numel() else 0.0
    # the worst error as a fraction of the literal bound (1.0 = at the bound): the margin a pass was made with
    worst = (err / (atol + rtol * ref.abs())).max().item() if err.numel() else 0.0
    rec = {"name": name, "numel": err.numel(), "max_err": max_err, "scale": scale, "rtol": rtol, "atol": atol,
           "worst_over_literal_bound": worst,
           "asserted": "scaled" if scaled else "literal", "literal_misses": bad_literal, "scaled_misses": bad_scaled}
    REPORT.append(rec)
    if scaled and bad_literal:
        print(f"[parity] {name}: {bad_literal}/{err.numel()} elements miss the LITERAL atol {atol:g} "
              f"(max err {max_err:.3e}, scale {scale:.3e}); asserted with atol*max(1,scale)")
    bad = bad_scaled if scaled else bad_literal
    if not REPORT_ONLY:
        assert bad == 0, (f"{name}: {bad}/{err.numel()} out of tolerance ({rec['asserted']} atol), max err {max_err:.3e}, "
                          f"scale {scale:.3e}")
    return max_err


def dump_report(path: str) -> None:
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(REPORT, f, indent=1)


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


def pack_detections_reference(det, probs: torch.Tensor, counts, starts) -> torch.Tensor:
    """The record layout of glass_pack_detections written with torch index ops: [n, max_det,
    1 + 5 + 1 + 1 + 2 + steps*classes] = (valid, box, score, class, orientation, text probs), zero rows past each count.
    Test-side statement of the layout: the CPU property tests of the unpacking code use it, the GPU test holds the kernel
    to it."""
    n, m = det["pred_boxes"].shape[0], det["pred_boxes"].shape[1]
    tp = probs.shape[1] * probs.shape[2]
    rec = torch.zeros((n, m, 10 + tp), dtype=torch.float32, device=probs.device)
    for i, c in enumerate(counts):
        if c == 0:
            continue
        rec[i, :c, 0] = 1.0
        rec[i, :c, 1:6] = det["pred_boxes"][i, :c]
        rec[i, :c, 6] = det["scores"][i, :c]
        if det.get("orientations") is not None:
            rec[i, :c, 8:10] = det["orientations"][i, :c]
        rec[i, :c, 10:] = probs[starts[i]: starts[i + 1]].reshape(c, tp)
    return rec
