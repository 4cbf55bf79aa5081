"""Registry adapter: the B200 modules under the names detectron2 / GLASS configs select them by.

The reference populates detectron2's registries at import (glass/__init__.py:4-9) and every module is then built BY
NAME from the yacs cfg: ``META_ARCH_REGISTRY["GlassRCNN"]`` (glass/modeling/meta_arch/glass_rcnn.py:13),
``PROPOSAL_GENERATOR_REGISTRY["RotatedRPN"]`` (glass/modeling/proposal_generator/rotated_rpn.py:16),
``ROI_HEADS_REGISTRY["MaskRotatedRecognizerHybridHead"]`` (glass/modeling/fusion/recognizers_hybrid_head.py:66, built
at :116-134 with ``cls(cfg, input_shape)``) and detectron2's own ``BACKBONE_REGISTRY["build_resnet_fpn_backbone"]``
(configs/glass_pretrain.yaml:41-42).  ``register_all()`` puts B200 adapters under exactly those names, so

    import glass_text_spotting_b200.d2_adapter as b200; b200.register_all(override=True)   # instead of / after `import glass`
    model = build_model(cfg); DetectionCheckpointer(model).load(path); model.eval()

builds the B200 path from an unchanged config and an unchanged checkpoint.  detectron2 is imported lazily and only
when no registries are handed in (it is not installable offline here; tests/test_d2_adapter_cpu.py drives the same
code through the stub ``Registry`` the golden-vector tools use).

Construction follows detectron2's two-phase convention -- ``cls(cfg[, input_shape])`` first, weights later through
``load_state_dict`` (what ``DetectionCheckpointer`` calls) -- while the B200 modules pack their weights at construction:
each adapter therefore keeps the cfg-derived keyword arguments and builds its B200 module when the state dict arrives.
Running an adapter before its weights were loaded raises (there are no random-init weights to fall back to).

Image convention: detectron2's ``GeneralizedRCNN.preprocess_image`` hands the backbone / ROI heads an
ALREADY NORMALISED, zero-padded ``ImageList``, whereas ``B200GlassRCNN`` feeds its modules raw pixels and fuses
``(x - mean) / std`` into the stem and the image pooler.  The component adapters below (backbone, ROI heads) are the
ones a stock ``GeneralizedRCNN`` drives, so they build their B200 modules with mean 0 / std 1 (``normalized_input``):
the fused normalisation becomes the identity and zero padding stays exactly zero.  The meta-architecture adapter owns
its preprocessing and keeps the raw-pixel path.
"""
from typing import Any, Dict, List, Mapping, Optional

import torch

from . import config as _config
from .structures import ImageList, Instances

# registry name -> the reference / detectron2 registration it stands in for
REGISTRY_NAMES = {
    "META_ARCH": ("GlassRCNN", "GeneralizedRCNN"),
    "BACKBONE": ("build_resnet_fpn_backbone",),
    "PROPOSAL_GENERATOR": ("RotatedRPN",),
    "ROI_HEADS": ("MaskRotatedRecognizerHybridHead",),
}


def _as_cfg(cfg) -> _config.CfgNode:
    """yacs CfgNode (a dict subclass), a plain mapping, or our CfgNode -> our CfgNode over the reference's defaults."""
    if isinstance(cfg, _config.CfgNode):
        return cfg
    if isinstance(cfg, Mapping):
        return _config.load_config(cfg)
    raise TypeError(f"cfg must be a mapping (yacs CfgNode / dict), got {type(cfg).__name__}")


def _strip(state_dict: Mapping[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    return {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}


class _LazyB200(torch.nn.Module):
    """Two-phase construction: keeps the constructor arguments until ``load_state_dict`` supplies the weights."""

    #: checkpoint prefix of this component inside a full GLASS state_dict (d2 names)
    prefix = ""

    def __init__(self):
        super().__init__()
        self._impl = None
        self._device = "cuda"

    def _build(self, state_dict: Dict[str, torch.Tensor]):
        raise NotImplementedError

    def load_state_dict(self, state_dict: Mapping[str, torch.Tensor], strict: bool = True):
        """Accepts the component's own keys or a full-model state_dict (keys carrying ``self.prefix``)."""
        sd = dict(state_dict)
        if self.prefix and any(k.startswith(self.prefix) for k in sd):
            sd = {k: v for k, v in sd.items() if k.startswith(self.prefix)}
        elif self.prefix:
            sd = {self.prefix + k: v for k, v in sd.items()}
        self._impl = self._build(sd)
        return torch.nn.modules.module._IncompatibleKeys([], [])

    # DetectionCheckpointer walks named submodules and calls _load_from_state_dict on each: route it here as well
    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        own = {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}
        if own:
            self.load_state_dict(own, strict)

    def to(self, device=None, *a, **k):
        if device is not None:
            self._device = str(device)
        return self

    @property
    def impl(self):
        if self._impl is None:
            raise RuntimeError(f"{type(self).__name__}: weights not loaded yet -- call load_state_dict (e.g. through "
                               "DetectionCheckpointer) before running the model; the B200 modules pack their weights at "
                               "construction and have no random-init path")
        return self._impl


class B200Backbone(_LazyB200):
    """``BACKBONE_REGISTRY["build_resnet_fpn_backbone"]`` (configs/glass_pretrain.yaml:41-54): Backbone.forward(Tensor
    [N,3,H,W]) -> {"p2".."p6"}, ``output_shape()``, ``size_divisibility``."""
    prefix = "backbone."
    size_divisibility = 32

    def __init__(self, cfg, input_shape=None, normalized_input: bool = True):
        super().__init__()
        cfg = _as_cfg(cfg)
        _config.check_supported(cfg)
        self._mean = (0.0, 0.0, 0.0) if normalized_input else tuple(cfg.MODEL.PIXEL_MEAN)
        self._std = (1.0, 1.0, 1.0) if normalized_input else tuple(cfg.MODEL.PIXEL_STD)

    def _build(self, sd):
        from .modeling.backbone import B200ResNetFPN
        return B200ResNetFPN(sd, device=self._device, pixel_mean=self._mean, pixel_std=self._std)

    def forward(self, x: torch.Tensor):
        return self.impl(x.contiguous().float())

    def output_shape(self):
        from .modeling.backbone import B200ResNetFPN
        return {k: {"channels": 256, "stride": s} for k, s in B200ResNetFPN.strides.items()}


def build_resnet_fpn_backbone(cfg, input_shape=None):
    """detectron2 registers a builder FUNCTION under this name; so does the adapter."""
    return B200Backbone(cfg, input_shape)


class RotatedRPN(_LazyB200):
    """``PROPOSAL_GENERATOR_REGISTRY["RotatedRPN"]`` (rotated_rpn.py:16-17): forward(images, features, gt_instances=None)
    -> (list[Instances{proposal_boxes, objectness_logits}], {})."""
    prefix = "proposal_generator."

    def __init__(self, cfg, input_shape=None):
        super().__init__()
        self._kw = _config.model_kwargs(_as_cfg(cfg))["rpn_kwargs"]

    def _build(self, sd):
        from .modeling.rpn import B200RotatedRPN
        return B200RotatedRPN(sd, device=self._device, **self._kw)

    def forward(self, images, features, gt_instances=None):
        return self.impl(images, features, gt_instances)


class MaskRotatedRecognizerHybridHead(_LazyB200):
    """``ROI_HEADS_REGISTRY["MaskRotatedRecognizerHybridHead"]`` (recognizers_hybrid_head.py:66): forward(images,
    features, proposals, targets=None) -> (list[Instances], {}) and forward_with_given_boxes(images, features, instances)
    (:571 -- with the extra ``images`` argument of the reference)."""
    prefix = "roi_heads."

    def __init__(self, cfg, input_shape=None, normalized_input: bool = True):
        super().__init__()
        cfg = _as_cfg(cfg)
        kw = _config.model_kwargs(cfg)
        self._mask = kw.pop("mask_inference")
        for k in ("rpn_kwargs", "filter_small_boxes", "inflate_ratio", "drop_overlapping_boxes"):
            kw.pop(k, None)
        if normalized_input:
            kw["pixel_mean"], kw["pixel_std"] = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
        self._kw = kw

    def _build(self, sd):
        from .modeling.mask_head import B200MaskHead
        from .modeling.roi_heads import B200GlassROIHeads
        heads = B200GlassROIHeads(sd, device=self._device, **self._kw)
        if self._mask:
            heads.mask_head = B200MaskHead(sd, device=self._device)
        return heads

    def forward(self, images, features, proposals, targets=None):
        return self.impl(images, features, proposals, targets)

    def forward_with_given_boxes(self, images, features, instances):
        return self.impl.forward_with_given_boxes(images, features, instances)


class GlassRCNN(_LazyB200):
    """``META_ARCH_REGISTRY["GlassRCNN"]`` (glass_rcnn.py:13; d2's build_model calls ``cls(cfg)``): forward(list[dict]) ->
    list[{"instances": Instances}], inference(batched_inputs, detected_instances=None, do_postprocess=True)
    (glass_rcnn.py:57-62).  Owns its preprocessing, hence the fused raw-pixel path."""
    prefix = ""

    def __init__(self, cfg):
        super().__init__()
        self._kw = _config.model_kwargs(_as_cfg(cfg))
        self.training = False

    def _build(self, sd):
        from .modeling.glass_rcnn import B200GlassRCNN
        return B200GlassRCNN(sd, device=self._device, **self._kw)

    def forward(self, batched_inputs: List[dict]):
        return self.impl.inference(batched_inputs)

    def inference(self, batched_inputs, detected_instances: Optional[List[Instances]] = None, do_postprocess: bool = True):
        return self.impl.inference(batched_inputs, detected_instances, do_postprocess)


class GeneralizedRCNN(GlassRCNN):
    """configs/glass_pretrain.yaml:40 selects detectron2's stock meta-architecture; same inference sequence."""


ADAPTERS = {
    "META_ARCH": {"GlassRCNN": GlassRCNN, "GeneralizedRCNN": GeneralizedRCNN},
    "BACKBONE": {"build_resnet_fpn_backbone": build_resnet_fpn_backbone},
    "PROPOSAL_GENERATOR": {"RotatedRPN": RotatedRPN},
    "ROI_HEADS": {"MaskRotatedRecognizerHybridHead": MaskRotatedRecognizerHybridHead},
}


def _put(registry, name: str, obj, override: bool) -> None:
    """detectron2's fvcore Registry refuses duplicates; ``override`` swaps the entry in place (its ``_obj_map``), a plain
    dict-like registry is simply assigned."""
    present = name in registry
    if present and not override:
        raise KeyError(f"'{name}' is already registered in {getattr(registry, '_name', registry)!r}; "
                       "pass override=True to replace the reference's module with the B200 one")
    obj_map = getattr(registry, "_obj_map", None)
    if obj_map is not None:
        obj_map[name] = obj
    else:
        registry[name] = obj


def register_all(meta_arch=None, backbone=None, proposal_generator=None, roi_heads=None, override: bool = False,
                 generalized_rcnn: bool = False) -> Dict[str, Any]:
    """Register the adapters under the reference's names.  Registries default to detectron2's (imported here, lazily).
    ``generalized_rcnn`` additionally replaces d2's stock ``GeneralizedRCNN`` entry (needed only when the fused meta-arch
    is wanted for configs/glass_pretrain.yaml; otherwise d2's own GeneralizedRCNN drives the three component adapters)."""
    if None in (meta_arch, backbone, proposal_generator, roi_heads):
        try:
            from detectron2.modeling import (BACKBONE_REGISTRY, META_ARCH_REGISTRY, PROPOSAL_GENERATOR_REGISTRY,
                                             ROI_HEADS_REGISTRY)
        except ImportError as e:  # not installable offline: the caller must hand the registries in
            raise ImportError("detectron2 is not importable: pass the four registries explicitly") from e
        meta_arch = meta_arch if meta_arch is not None else META_ARCH_REGISTRY
        backbone = backbone if backbone is not None else BACKBONE_REGISTRY
        proposal_generator = proposal_generator if proposal_generator is not None else PROPOSAL_GENERATOR_REGISTRY
        roi_heads = roi_heads if roi_heads is not None else ROI_HEADS_REGISTRY
    regs = {"META_ARCH": meta_arch, "BACKBONE": backbone, "PROPOSAL_GENERATOR": proposal_generator, "ROI_HEADS": roi_heads}
    for kind, entries in ADAPTERS.items():
        for name, obj in entries.items():
            if name == "GeneralizedRCNN" and not generalized_rcnn:
                continue
            _put(regs[kind], name, obj, override)
    return regs


def build_model(cfg, registries: Mapping[str, Any]):
    """What detectron2's ``build_model(cfg)`` does, on the given registries: ``META_ARCH_REGISTRY.get(name)(cfg)``."""
    cfg = _as_cfg(cfg)
    return registries["META_ARCH"].get(cfg.MODEL.META_ARCHITECTURE)(cfg)


class ComponentRCNN:
    """The sequence detectron2's stock ``GeneralizedRCNN.inference`` runs over registry-built components
    (glass_rcnn.py:82-101 is the same sequence): normalise + pad -> backbone -> proposal generator -> roi_heads.
    A stand-in for d2's class where d2 is absent; used by the tests to drive the three component adapters together."""

    def __init__(self, cfg, registries: Mapping[str, Any]):
        cfg = _as_cfg(cfg)
        M = cfg.MODEL
        self.backbone = registries["BACKBONE"].get(M.BACKBONE.NAME)(cfg, None)
        self.proposal_generator = registries["PROPOSAL_GENERATOR"].get(M.PROPOSAL_GENERATOR.NAME)(cfg, self.backbone.output_shape())
        self.roi_heads = registries["ROI_HEADS"].get(M.ROI_HEADS.NAME)(cfg, self.backbone.output_shape())
        self.pixel_mean, self.pixel_std = tuple(M.PIXEL_MEAN), tuple(M.PIXEL_STD)

    def load_state_dict(self, sd, strict: bool = True):
        for m in (self.backbone, self.proposal_generator, self.roi_heads):
            m.load_state_dict(sd, strict)

    @torch.no_grad()
    def inference(self, batched_inputs: List[dict], device="cuda"):
        mean = torch.tensor(self.pixel_mean, device=device).view(3, 1, 1)
        std = torch.tensor(self.pixel_std, device=device).view(3, 1, 1)
        imgs = [(x["image"].to(device).float() - mean) / std for x in batched_inputs]
        images = ImageList.from_tensors(imgs, self.backbone.size_divisibility, pad_value=0.0)
        images.normalized = True
        feats = self.backbone(images.tensor)
        proposals, _ = self.proposal_generator(images, feats, None)
        results, _ = self.roi_heads(images, feats, proposals, None)
        return results
