"""glass_text_spotting_b200.d2_adapter against a stub of detectron2's Registry (detectron2 is not installable offline;
tools/make_golden.py uses the same stub to exec the reference's modules): registration under the reference's names
(glass/__init__.py:4-9), build-by-name from a glass_pretrain.yaml-shaped cfg, two-phase construction."""
import pytest
import torch

from test_config_cpu import PRETRAIN_LIKE


class Registry(dict):
    """fvcore.common.registry.Registry as detectron2 uses it: register() (decorator or call), get(name), no duplicates."""

    def __init__(self, name):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        def deco(o):
            assert o.__name__ not in self, f"An object named '{o.__name__}' was already registered in '{self._name}' registry!"
            self[o.__name__] = o
            return o
        return deco(obj) if obj is not None else deco

    def get(self, name):
        if name not in self:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return self[name]


def _registries():
    return dict(meta_arch=Registry("META_ARCH"), backbone=Registry("BACKBONE"),
                proposal_generator=Registry("PROPOSAL_GENERATOR"), roi_heads=Registry("ROI_HEADS"))


def test_registers_under_the_reference_names():
    from glass_text_spotting_b200 import d2_adapter as ad
    regs = ad.register_all(**_registries())
    assert set(regs["META_ARCH"]) == {"GlassRCNN"}                      # d2's own GeneralizedRCNN entry is left alone
    assert set(regs["BACKBONE"]) == {"build_resnet_fpn_backbone"}
    assert set(regs["PROPOSAL_GENERATOR"]) == {"RotatedRPN"}
    assert set(regs["ROI_HEADS"]) == {"MaskRotatedRecognizerHybridHead"}
    regs2 = ad.register_all(**_registries(), generalized_rcnn=True)
    assert set(regs2["META_ARCH"]) == {"GlassRCNN", "GeneralizedRCNN"}


def test_duplicate_needs_override():
    from glass_text_spotting_b200 import d2_adapter as ad
    r = _registries()

    class RotatedRPN:      # the reference's own class, registered first by `import glass`
        pass
    r["proposal_generator"].register(RotatedRPN)
    with pytest.raises(KeyError, match="override=True"):
        ad.register_all(**r)
    r = _registries()
    r["proposal_generator"].register(RotatedRPN)
    regs = ad.register_all(**r, override=True)
    assert regs["PROPOSAL_GENERATOR"].get("RotatedRPN") is ad.RotatedRPN


def test_build_by_name_from_pretrain_config_is_two_phase():
    """cls(cfg[, input_shape]) works without a GPU and without weights; running before load_state_dict raises."""
    from glass_text_spotting_b200 import config, d2_adapter as ad
    regs = ad.register_all(**_registries(), generalized_rcnn=True)
    cfg = config.load_config(PRETRAIN_LIKE)
    model = ad.build_model(cfg, regs)                       # META_ARCHITECTURE: GeneralizedRCNN
    assert isinstance(model, ad.GeneralizedRCNN) and isinstance(model, torch.nn.Module)
    assert model._kw["rpn_kwargs"]["pre_nms_topk"] == 1000 and model._kw["max_word_len"] == 26
    with pytest.raises(RuntimeError, match="weights not loaded"):
        model([{"image": torch.zeros(3, 32, 32)}])
    comp = ad.ComponentRCNN(PRETRAIN_LIKE, regs)            # a plain mapping (what a yacs CfgNode is) is accepted too
    assert isinstance(comp.backbone, ad.B200Backbone) and comp.backbone.size_divisibility == 32
    assert comp.backbone.output_shape()["p2"] == {"channels": 256, "stride": 4}
    assert isinstance(comp.proposal_generator, ad.RotatedRPN)
    assert isinstance(comp.roi_heads, ad.MaskRotatedRecognizerHybridHead)
    # stock GeneralizedRCNN hands over normalised images: the component adapters neutralise the fused normalisation
    assert comp.backbone._mean == (0.0, 0.0, 0.0) and comp.roi_heads._kw["pixel_mean"] == (0.0, 0.0, 0.0)
    assert model._kw["pixel_mean"] == (103.53, 116.28, 123.675)
    with pytest.raises(RuntimeError, match="weights not loaded"):
        comp.backbone(torch.zeros(1, 3, 32, 32))


def test_unsupported_architecture_is_refused_by_key():
    from glass_text_spotting_b200 import config, d2_adapter as ad
    import copy
    bad = copy.deepcopy(PRETRAIN_LIKE)
    bad["MODEL"]["RESNETS"]["DEPTH"] = 101
    with pytest.raises(config.UnsupportedConfig, match="RESNETS.DEPTH"):
        ad.GlassRCNN(bad)


def test_state_dict_routing_by_prefix():
    """A component accepts its own keys or a full-model checkpoint; either way the B200 module sees full d2 names."""
    from glass_text_spotting_b200 import d2_adapter as ad
    seen = {}

    class Probe(ad._LazyB200):
        prefix = "proposal_generator."

        def _build(self, sd):
            seen["keys"] = sorted(sd)
            return object()
    p = Probe()
    p.load_state_dict({"rpn_head.conv.weight": torch.zeros(1), "rpn_head.conv.bias": torch.zeros(1)})
    assert seen["keys"] == ["proposal_generator.rpn_head.conv.bias", "proposal_generator.rpn_head.conv.weight"]
    p.load_state_dict({"proposal_generator.rpn_head.conv.weight": torch.zeros(1), "backbone.x": torch.zeros(1)})
    assert seen["keys"] == ["proposal_generator.rpn_head.conv.weight"]
    # DetectionCheckpointer's path: the parent module's load walks the children with their prefixes
    p._load_from_state_dict({"proposal_generator.rpn_head.conv.weight": torch.zeros(1)}, "proposal_generator.", {}, True,
                            [], [], [])
    assert seen["keys"] == ["proposal_generator.rpn_head.conv.weight"]
