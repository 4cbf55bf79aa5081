"""Random-init weight factory with detectron2 / GLASS ``state_dict`` names.

Used by bench.py and smoke(): there is no network for checkpoints, so the benchmark runs the named
architecture with seeded random weights (the reference's init rules: c2_msra / c2_xavier /
normal(0.01)); BatchNorm running stats are set so activations stay O(1) through the residual stacks.
Released ``.pth`` files load through the same names.
"""
import math
from typing import Dict

import torch


def _msra(g, cout, cin, kh, kw):
    return torch.randn(cout, cin, kh, kw, generator=g) * math.sqrt(2.0 / (cout * kh * kw))


def _xavier(g, cout, cin, kh, kw):
    b = math.sqrt(3.0 / (cin * kh * kw))
    return (torch.rand(cout, cin, kh, kw, generator=g) * 2 - 1) * b


def _bn(sd, prefix, c, g, gain=1.0):
    sd[prefix + ".weight"] = gain * (1.0 + 0.1 * torch.randn(c, generator=g))
    sd[prefix + ".bias"] = 0.1 * torch.randn(c, generator=g)
    sd[prefix + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
    sd[prefix + ".running_var"] = 1.0 + 0.1 * torch.rand(c, generator=g)


def random_backbone_state_dict(seed: int = 0, prefix: str = "backbone.") -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    p = prefix + "bottom_up.stem.conv1"
    sd[p + ".weight"] = _msra(g, 64, 3, 7, 7) / 50.0  # raw pixel scale -> O(1) activations
    _bn(sd, p + ".norm", 64, g)
    cin = 64
    for i, (nb, stage) in enumerate(zip([3, 4, 6, 3], ["res2", "res3", "res4", "res5"])):
        bott, cout = 64 * 2 ** i, 256 * 2 ** i
        for b in range(nb):
            q = f"{prefix}bottom_up.{stage}.{b}"
            if cin != cout:
                sd[q + ".shortcut.weight"] = _msra(g, cout, cin, 1, 1)
                _bn(sd, q + ".shortcut.norm", cout, g, 0.7)
            sd[q + ".conv1.weight"] = _msra(g, bott, cin, 1, 1)
            _bn(sd, q + ".conv1.norm", bott, g)
            sd[q + ".conv2.weight"] = _msra(g, bott, bott, 3, 3)
            _bn(sd, q + ".conv2.norm", bott, g)
            sd[q + ".conv3.weight"] = _msra(g, cout, bott, 1, 1)
            _bn(sd, q + ".conv3.norm", cout, g, 0.5)
            cin = cout
    for k, c in zip([2, 3, 4, 5], [256, 512, 1024, 2048]):
        sd[f"{prefix}fpn_lateral{k}.weight"] = _xavier(g, 256, c, 1, 1)
        _bn(sd, f"{prefix}fpn_lateral{k}.norm", 256, g)
        sd[f"{prefix}fpn_output{k}.weight"] = _xavier(g, 256, 256, 3, 3)
        _bn(sd, f"{prefix}fpn_output{k}.norm", 256, g)
    return sd
