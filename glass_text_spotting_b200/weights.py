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


def _kaiming_u(g, *shape):
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    b = 1.0 / math.sqrt(fan_in)
    return (torch.rand(*shape, generator=g) * 2 - 1) * b


def _linear(sd, prefix, g, out_f, in_f, std=None, bias=True):
    sd[prefix + ".weight"] = (torch.randn(out_f, in_f, generator=g) * std) if std else _kaiming_u(g, out_f, in_f)
    if bias:
        sd[prefix + ".bias"] = torch.zeros(out_f) if std else _kaiming_u(g, out_f, in_f)[:, 0].clone()


def random_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Full GLASS (glass_pretrain.yaml) state_dict with random weights; every BatchNorm keeps activations O(1)."""
    g = torch.Generator().manual_seed(seed)
    sd = random_backbone_state_dict(seed)
    # RPN head (normal 0.01; a larger std on the predictors gives a non-degenerate ranking)
    p = "proposal_generator.rpn_head."
    sd[p + "conv.weight"] = torch.randn(256, 256, 3, 3, generator=g) * 0.02
    sd[p + "conv.bias"] = torch.zeros(256)
    sd[p + "objectness_logits.weight"] = torch.randn(12, 256, 1, 1, generator=g) * 0.05
    sd[p + "objectness_logits.bias"] = torch.zeros(12)
    sd[p + "anchor_deltas.weight"] = torch.randn(60, 256, 1, 1, generator=g) * 0.02
    sd[p + "anchor_deltas.bias"] = torch.zeros(60)
    r = "roi_heads."
    sd[r + "box_head.fc1.weight"] = _xavier(g, 2048, 12544, 1, 1).view(2048, 12544)
    sd[r + "box_head.fc1.bias"] = torch.zeros(2048)
    sd[r + "box_head.fc2.weight"] = _xavier(g, 2048, 2048, 1, 1).view(2048, 2048)
    sd[r + "box_head.fc2.bias"] = torch.zeros(2048)
    # foreground-biased classifier so that a realistic number of words survives the 0.05 threshold
    sd[r + "box_predictor.cls_score.weight"] = torch.randn(2, 2048, generator=g) * 0.01
    sd[r + "box_predictor.cls_score.bias"] = torch.tensor([0.5, -0.5])
    sd[r + "box_predictor.bbox_pred.weight"] = torch.randn(5, 2048, generator=g) * 0.001
    sd[r + "box_predictor.bbox_pred.bias"] = torch.zeros(5)
    sd[r + "box_predictor.orientation_pred.weight"] = torch.randn(4, 2048, generator=g) * 0.01
    sd[r + "box_predictor.orientation_pred.bias"] = torch.zeros(4)
    sd[r + "recognizer_feature_fusion.conv1.weight"] = _msra(g, 256, 256, 1, 1)
    sd[r + "recognizer_feature_fusion.conv2.weight"] = _msra(g, 256, 256, 1, 1)
    # hybrid_net (ResNetFeatureExtractor, layers [1,2,5,3], 3 -> 256)
    h = r + "hybrid_net.ConvNet."

    def conv_bn(cname, bname, cout, cin, k, gain=1.0):
        sd[h + cname + ".weight"] = _msra(g, cout, cin, *k)
        _bn(sd, h + bname, cout, g, gain)

    conv_bn("conv0_1", "bn0_1", 16, 3, (3, 3))
    sd[h + "conv0_1.weight"] /= 50.0  # raw pixel scale
    conv_bn("conv0_2", "bn0_2", 32, 16, (3, 3))
    inpl = 32
    for li, (planes, nblk) in enumerate(zip([64, 128, 256, 256], [1, 2, 5, 3]), start=1):
        for b in range(nblk):
            q = f"layer{li}.{b}."
            conv_bn(q + "conv1", q + "bn1", planes, inpl, (3, 3))
            conv_bn(q + "conv2", q + "bn2", planes, planes, (3, 3), 0.5)
            if inpl != planes:
                conv_bn(q + "downsample.0", q + "downsample.1", planes, inpl, (1, 1), 0.7)
            inpl = planes
        if li <= 3:
            conv_bn(f"conv{li}", f"bn{li}", planes, planes, (3, 3))
    conv_bn("conv4_1", "bn4_1", 256, 256, (2, 2))
    # fusion_net (MultiAspectGCAttention)
    f = r + "fusion_net."
    sd[f + "conv_mask.weight"] = _kaiming_u(g, 1, 64, 1, 1)
    sd[f + "conv_mask.bias"] = torch.zeros(1)
    sd[f + "channel_add_conv.0.weight"] = _kaiming_u(g, 256, 512, 1, 1)
    sd[f + "channel_add_conv.0.bias"] = torch.zeros(256)
    sd[f + "channel_add_conv.1.weight"] = torch.ones(256, 1, 1)
    sd[f + "channel_add_conv.1.bias"] = torch.zeros(256, 1, 1)
    sd[f + "channel_add_conv.3.weight"] = _kaiming_u(g, 512, 256, 1, 1)
    sd[f + "channel_add_conv.3.bias"] = torch.zeros(512)
    sd[f + "out.weight"] = _kaiming_u(g, 256, 512, 3, 3)
    sd[f + "out.bias"] = torch.zeros(256)
    # recognizer head
    q = r + "recognizer_head."
    sd[q + "backbone.conv1.weight"] = _msra(g, 256, 256, 2, 1)
    _bn(sd, q + "backbone.conv1.norm", 256, g)
    sd[q + "backbone.conv2.weight"] = _msra(g, 256, 256, 3, 3)
    _bn(sd, q + "backbone.conv2.norm", 256, g)
    for l in range(2):
        e = q + f"encoder.bilsm_stack.{l}."
        for suf in ("", "_reverse"):
            sd[e + "rnn.weight_ih_l0" + suf] = _kaiming_u(g, 1024, 256)
            sd[e + "rnn.weight_hh_l0" + suf] = _kaiming_u(g, 1024, 256)
            sd[e + "rnn.bias_ih_l0" + suf] = torch.zeros(1024)
            sd[e + "rnn.bias_hh_l0" + suf] = torch.zeros(1024)
        sd[e + "linear.weight"] = _kaiming_u(g, 256, 512)
        sd[e + "linear.bias"] = torch.zeros(256)
    d = q + "decoder.recognizer.decoder."
    for nm in ("sEmbed", "xEmbed"):
        sd[d + f"attention_unit.{nm}.weight"] = _kaiming_u(g, 256, 256)
        sd[d + f"attention_unit.{nm}.bias"] = torch.zeros(256)
    sd[d + "attention_unit.wEmbed.weight"] = _kaiming_u(g, 1, 256)
    sd[d + "attention_unit.wEmbed.bias"] = torch.zeros(1)
    sd[d + "tgt_embedding.weight"] = torch.randn(97, 256, generator=g)
    sd[d + "gru.weight_ih_l0"] = _kaiming_u(g, 768, 512)
    sd[d + "gru.weight_hh_l0"] = _kaiming_u(g, 768, 256)
    sd[d + "gru.bias_ih_l0"] = torch.zeros(768)
    sd[d + "gru.bias_hh_l0"] = torch.zeros(768)
    sd[d + "fc.weight"] = _kaiming_u(g, 97, 256)
    sd[d + "fc.bias"] = torch.zeros(97)
    sd[d + "temperature"] = torch.ones(1)
    return sd


def random_mask_head_state_dict(seed: int = 0, prefix: str = "roi_heads.mask_head.") -> Dict[str, torch.Tensor]:
    """detectron2 MaskRCNNConvUpsampleHead keys (ROI_MASK_HEAD: NUM_CONV 4, CONV_DIM 256, one class): mask_fcn1..4,
    deconv, predictor.  Own generator, so adding it never changes the weights random_state_dict() draws."""
    g = torch.Generator().manual_seed(seed + 7777)
    sd = {}
    for k in range(1, 5):
        sd[f"{prefix}mask_fcn{k}.weight"] = _msra(g, 256, 256, 3, 3)
        sd[f"{prefix}mask_fcn{k}.bias"] = torch.randn(256, generator=g) * 0.05
    sd[prefix + "deconv.weight"] = torch.randn(256, 256, 2, 2, generator=g) * (2.0 / 1024) ** 0.5
    sd[prefix + "deconv.bias"] = torch.randn(256, generator=g) * 0.05
    sd[prefix + "predictor.weight"] = torch.randn(1, 256, 1, 1, generator=g) * 0.5
    sd[prefix + "predictor.bias"] = torch.zeros(1)
    return sd


# ------------------------------------------------------------------------------------------ released checkpoints
_BASE_KEYS = None


def required_keys(orientation: bool = True, mask: bool = False):
    """Parameter names the inference path reads (the detectron2 / GLASS names of SURVEY.md A.10)."""
    global _BASE_KEYS
    if _BASE_KEYS is None:
        _BASE_KEYS = frozenset(random_state_dict(0).keys())
    keys = set(_BASE_KEYS)
    if not orientation:
        keys = {k for k in keys if "orientation_pred" not in k}
    if mask:
        keys |= set(random_mask_head_state_dict(0).keys())
    return keys


def _numpy_safe_globals():
    """The globals a pickled numpy array needs (array reconstruction, dtypes): data constructors only."""
    import numpy as np
    out = [np.ndarray, np.dtype]
    for mod in ("numpy._core.multiarray", "numpy.core.multiarray"):
        try:
            m = __import__(mod, fromlist=["_reconstruct"])
            out += [m._reconstruct, m.scalar]
        except Exception:
            pass
    out += [type(np.dtype(t)) for t in ("float32", "float64", "float16", "int64", "int32", "uint8", "bool")]
    return list(dict.fromkeys(out))


def load_checkpoint(path: str, mask: bool = False, allow_pickle: bool = False) -> Dict[str, torch.Tensor]:
    """What ``DetectionCheckpointer(model).load(path)`` does for a ``.pth`` file (glass/inference/glass_runner.py:58-60):
    read the ``"model"`` entry (or a bare state_dict), drop a DataParallel ``module.`` prefix, turn numpy arrays into
    tensors, and check that every parameter the inference path reads is there -- naming the missing ones instead of
    failing somewhere inside weight packing.  Keys the inference path does not read (training-only heads, mask head when
    ``mask`` is False, ``num_batches_tracked``) are dropped.  A checkpoint without ``box_predictor.orientation_pred``
    (MODEL.ORIENTATION_ON false) is accepted.

    The file is read with ``weights_only=True`` (tensors, numpy arrays and plain containers only -- what released
    detectron2 ``.pth`` checkpoints hold); a checkpoint that needs full unpickling (arbitrary code execution when the
    file is untrusted) is refused unless ``allow_pickle=True`` is passed explicitly."""
    if not str(path).endswith((".pth", ".pt")):
        raise ValueError(f"{path}: only torch checkpoints (.pth) are supported; convert Caffe2 .pkl weights with detectron2")
    try:
        with torch.serialization.safe_globals(_numpy_safe_globals()):   # numpy arrays only construct data
            ckpt = torch.load(path, map_location="cpu", weights_only=True)
    except Exception as e:
        if not allow_pickle:
            raise RuntimeError(f"{path}: not loadable with weights_only=True ({type(e).__name__}: {str(e)[:200]}); if the file "
                               "is trusted, pass allow_pickle=True") from e
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
    sd = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt and isinstance(ckpt["model"], dict) else ckpt
    if not isinstance(sd, dict):
        raise ValueError(f"{path}: no state_dict found (expected a dict or a dict under 'model')")
    out = {}
    for k, v in sd.items():
        if k.startswith("module."):
            k = k[len("module."):]
        if k.endswith("num_batches_tracked"):
            continue
        out[k] = v if isinstance(v, torch.Tensor) else torch.as_tensor(v)
    orientation = "roi_heads.box_predictor.orientation_pred.weight" in out
    need = required_keys(orientation=orientation, mask=mask)
    missing = sorted(need - set(out))
    if missing:
        raise KeyError(f"{path}: {len(missing)} parameters of the GLASS inference path are missing, e.g. {missing[:5]}")
    return {k: out[k].detach().float() for k in need}
