"""Generate tests/golden/cnn_v1_1.pt with the REFERENCE's own CNN_V1_1 (glass/modeling/recognition/
recognizer_backbone.py:34-81, a14 of SURVEY.md 8a).  Authoring container only.  The class is built on detectron2's
``Conv2d`` wrapper and ``get_norm``, which are not installable offline: they are stubbed by their published semantics
(nn.Conv2d, then ``norm`` if given, then ``activation`` if given; ``get_norm("BN", c)`` = nn.BatchNorm2d(c)).  The
wiring -- (2,1)/s(2,1) conv + BN + ReLU, 3x3 conv + BN + ReLU, residual add of the FIRST conv's output -- is the
reference's code.

    python tools/make_golden_cnn_v1_1.py
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn.functional as F
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"


class Conv2d(nn.Conv2d):
    """detectron2.layers.wrappers.Conv2d."""

    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


def main():
    from tests.golden_common import seeded_fill

    class Registry(dict):
        def __init__(self, name):
            super().__init__()

        def register(self, obj=None):
            def deco(o):
                self[o.__name__] = o
                return o
            return deco(obj) if obj is not None else deco

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("fvcore")
    mod("fvcore.nn")
    wi = mod("fvcore.nn.weight_init", c2_msra_fill=lambda m: None, c2_xavier_fill=lambda m: None)
    sys.modules["fvcore.nn"].weight_init = wi
    mod("detectron2")
    mod("detectron2.config", configurable=lambda fn=None, **kw: fn)
    mod("detectron2.layers", Conv2d=Conv2d, ShapeSpec=object,
        get_norm=lambda norm, c: nn.BatchNorm2d(c) if norm else None)
    mod("detectron2.utils")
    mod("detectron2.utils.registry", Registry=Registry)
    spec = importlib.util.spec_from_file_location("ref_recognizer_backbone",
                                                  os.path.join(REF, "glass/modeling/recognition/recognizer_backbone.py"))
    rb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rb)
    m = rb.CNN_V1_1(types.SimpleNamespace(channels=256), conv_norm="BN").eval()
    seeded_fill(m, 17)
    g = torch.Generator().manual_seed(1017)
    x = torch.randn(3, 256, 8, 32, generator=g)
    with torch.no_grad():
        y = m(x)
    torch.save({"shape": list(y.shape), "stride": [1, 7, 1, 3], "sample": y[:, ::7, :, ::3].contiguous().clone(),
                "sum": y.double().sum().item(), "abssum": y.double().abs().sum().item(),
                "keys": sorted(m.state_dict().keys())},
               os.path.join(ROOT, "tests", "golden", "cnn_v1_1.pt"))
    print("ok", list(y.shape), y.double().abs().sum().item())


if __name__ == "__main__":
    main()
