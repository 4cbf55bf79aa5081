"""Seeds, inputs and the seeded weight fill shared by tools/make_golden.py (which runs
the reference's modules in the authoring container) and tests/test_oracle_golden.py
(which replays the same seeds through oracle/nets.py)."""
import math

import torch
from torch import nn

SEEDS = {"hybrid": 11, "fusion": 12, "p2p3": 13, "encoder": 14, "decoder": 15, "decoder_break": 16}


def seeded_input(case):
    g = torch.Generator().manual_seed(1000 + SEEDS[case])
    if case == "hybrid":
        return torch.randn(2, 3, 128, 128, generator=g) * 50.0
    if case == "fusion":
        return torch.randn(2, 512, 8, 32, generator=g)
    if case == "p2p3":
        return torch.randn(1, 256, 16, 16, generator=g), torch.randn(1, 256, 8, 8, generator=g)
    if case == "encoder":
        return torch.randn(3, 256, 4, 32, generator=g)
    if case == "decoder":
        return torch.randn(3, 32, 256, generator=g)
    if case == "decoder_break":
        return torch.randn(2, 32, 256, generator=g)
    raise KeyError(case)


def seeded_fill(module: nn.Module, seed: int):
    """Deterministic, scale-preserving fill keyed on parameter *names* (sorted), so a
    reference module and its oracle restatement with equal state_dict keys get equal
    values."""
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    for name in sorted(sd.keys()):
        t = sd[name]
        if name.endswith("num_batches_tracked"):
            continue
        if name.endswith("running_var"):
            t.copy_(0.5 + torch.rand(t.shape, generator=g))
        elif name.endswith("running_mean"):
            t.copy_(0.1 * torch.randn(t.shape, generator=g))
        elif name.endswith("temperature"):
            t.fill_(1.0)
        elif t.dim() >= 2:
            fan_in = t[0].numel()
            t.copy_(torch.randn(t.shape, generator=g) * math.sqrt(1.5 / fan_in))
        elif ("bn" in name or "norm" in name or "downsample.1" in name or "channel_add_conv.1" in name) \
                and name.endswith("weight"):
            t.copy_(0.75 + 0.5 * torch.rand(t.shape, generator=g))
        else:
            t.copy_(0.1 * torch.randn(t.shape, generator=g))
    module.load_state_dict(sd)


def force_eos_bias(head):
    """Liven the decoder up (larger embeddings / output weights) and bias class 0 so that the
    reference's early break `dones.min() != 0` fires mid-sequence (after 6 of 26 steps: word 1
    emits class 0 at step 0 and keeps decoding, word 0 first emits it at step 5)."""
    with torch.no_grad():
        head.decoder.tgt_embedding.weight *= 6.0
        head.decoder.fc.weight *= 4.0
        head.decoder.fc.bias[0] += 2.0


from glass_text_spotting_b200.synthetic import make_postprocess_case  # noqa: E402,F401  (also used by bench.py)


def make_eval_inputs(seed, n):
    g = torch.Generator().manual_seed(seed)
    boxes = torch.stack([torch.rand(n, generator=g) * 900 + 50, torch.rand(n, generator=g) * 900 + 50,
                         torch.rand(n, generator=g) * 200 + 10, torch.rand(n, generator=g) * 60 + 8,
                         torch.rand(n, generator=g) * 360 - 180], 1)
    scores = torch.rand(n, generator=g)
    probs = torch.softmax(torch.randn(n, 26, 97, generator=g) * 9.0, dim=2)
    for i in range(n):
        stop = int(torch.randint(0, 9, (1,), generator=g))   # stop == 0: empty word (record dropped)
        probs[i, stop] = 0.001
        probs[i, stop, 1] = 0.904
    return boxes, scores, probs


def make_paste_inputs(seed: int, n: int, h: int, w: int, side: int = 28):
    """Seeded soft masks [n, side, side] in (0, 1) (a blurred blob per mask) and rotated boxes [n, 5] partly hanging
    over the image border, for the rotated mask paste (tools/make_golden_paste.py and the tests)."""
    g = torch.Generator().manual_seed(4000 + seed)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, side), torch.linspace(-1, 1, side), indexing="ij")
    masks = []
    for _ in range(n):
        a, b = 0.5 + 0.6 * torch.rand(1, generator=g).item(), 0.4 + 0.6 * torch.rand(1, generator=g).item()
        r = (xx / a) ** 2 + (yy / b) ** 2
        m = torch.sigmoid(6.0 * (1.0 - r)) * (0.75 + 0.25 * torch.rand(side, side, generator=g))
        masks.append(m)
    masks = torch.stack(masks) if n else torch.zeros(0, side, side)
    boxes = torch.stack([torch.rand(n, generator=g) * (w + 20) - 10, torch.rand(n, generator=g) * (h + 20) - 10,
                         torch.rand(n, generator=g) * (w / 2) + 6, torch.rand(n, generator=g) * (h / 3) + 4,
                         torch.rand(n, generator=g) * 360 - 180], 1) if n else torch.zeros(0, 5)
    return masks.float().contiguous(), boxes.float().contiguous()


def make_box_inference_inputs(seed: int, r: int, hw):
    """Seeded box-head outputs for R proposals of one image: class logits [R,2] (fg, bg), deltas [R,5], orientation
    logits [R,4], proposals [R,5]; a few rows are made non-finite and many proposals overlap so that the finite filter,
    the score threshold, the rotated NMS and the top-k all act."""
    g = torch.Generator().manual_seed(5000 + seed)
    h, w = hw
    n_c = max(r // 4, 1)
    centers = torch.stack((torch.rand(n_c, generator=g) * w, torch.rand(n_c, generator=g) * h), 1)
    idx = torch.randint(0, n_c, (r,), generator=g)
    bw = torch.rand(r, generator=g) * (w / 4) + 8
    proposals = torch.stack((centers[idx, 0] + torch.randn(r, generator=g) * 4, centers[idx, 1] + torch.randn(r, generator=g) * 3,
                             bw, bw * (0.15 + 0.5 * torch.rand(r, generator=g)),
                             torch.rand(r, generator=g) * 360 - 180), 1)
    logits = torch.stack((torch.randn(r, generator=g) * 2.0, torch.randn(r, generator=g) * 2.0), 1)
    deltas = torch.randn(r, 5, generator=g) * torch.tensor([1.0, 1.0, 0.5, 0.5, 2.0])
    orient = torch.randn(r, 4, generator=g) * 2.0
    if r >= 20:
        deltas[3, 2] = float("inf")
        logits[7, 0] = float("nan")
    return logits.float(), deltas.float(), orient.float(), proposals.float().contiguous()


def make_meta_postprocess_inputs(seed: int, n: int, hw):
    """Seeded raw detections for GlassRCNN._postprocess (glass_rcnn.py:103-128): a mix of ordinary words, boxes
    thinner than MIN_BOX_DIMENSION, near-axis-aligned boxes hanging over the border (the only ones
    RotatedBoxes.clip touches) and boxes wholly outside the image (empty after the clip)."""
    g = torch.Generator().manual_seed(7000 + seed)
    h, w = hw
    cx = torch.rand(n, generator=g) * w
    cy = torch.rand(n, generator=g) * h
    bw = torch.exp(torch.rand(n, generator=g) * math.log(40.0)) * 6.0
    bh = bw * (0.15 + 0.6 * torch.rand(n, generator=g))
    ang = (torch.rand(n, generator=g) - 0.5) * 120.0
    for i in range(n):
        r = i % 8
        if r == 1:      # too thin (min side < 2)
            bh[i] = 0.5 + 1.4 * float(torch.rand(1, generator=g))
        elif r == 2:    # near-axis-aligned, over the right / bottom border
            ang[i] = (float(torch.rand(1, generator=g)) - 0.5) * 1.8
            cx[i], cy[i] = w - 3.0, h - 2.0
        elif r == 3:    # axis-aligned and completely outside
            ang[i] = 0.0
            cx[i] = w + 50.0 + bw[i]
        elif r == 4:    # exactly on the small-box limit
            bh[i] = 2.0
        elif r == 5:    # angle outside [-180, 180): normalised by the clip
            ang[i] = 200.0 + 100.0 * float(torch.rand(1, generator=g))
    boxes = torch.stack((cx, cy, bw, bh, ang), 1).float()
    scores = torch.rand(n, generator=g)
    return boxes, scores


def make_recognizer_branch_inputs(seed: int, k: int, hw=(96, 160)):
    """Seeded inputs of MaskRotatedRecognizerHybridHead._forward_recognizer (recognizers_hybrid_head.py:513-569):
    the normalised image [3,H,W], p2 [1,256,H/4,W/4], p3 [1,256,H/8,W/8] and k detected boxes inside the image."""
    g = torch.Generator().manual_seed(8000 + seed)
    h, w = hw
    image = torch.randn(3, h, w, generator=g) * 5.0   # keeps every activation of the seeded branch below ~1e3
    p2 = torch.randn(1, 256, h // 4, w // 4, generator=g)
    p3 = torch.randn(1, 256, h // 8, w // 8, generator=g)
    cx = 20 + torch.rand(k, generator=g) * (w - 40)
    cy = 15 + torch.rand(k, generator=g) * (h - 30)
    bw = 24 + torch.rand(k, generator=g) * 70
    bh = 8 + torch.rand(k, generator=g) * 20
    ang = (torch.rand(k, generator=g) - 0.5) * 90
    return image, p2, p3, torch.stack((cx, cy, bw, bh, ang), 1).float()


def make_runner_case(seed: int, hw):
    """A uint8 HWC image and model-space detections for the GlassRunner flow (glass/inference/glass_runner.py:72-109)."""
    import numpy as np
    rng = np.random.RandomState(9000 + seed)
    image = rng.randint(0, 256, size=tuple(hw) + (3,), dtype=np.uint8)
    g = torch.Generator().manual_seed(9000 + seed)
    n = 7
    boxes = torch.stack((torch.rand(n, generator=g) * hw[1], torch.rand(n, generator=g) * hw[0],
                         4 + torch.rand(n, generator=g) * 30, 3 + torch.rand(n, generator=g) * 10,
                         (torch.rand(n, generator=g) - 0.5) * 60), 1).float()
    scores = torch.rand(n, generator=g)
    return image, boxes, scores


def make_box_branch_inputs(seed: int, r: int, hw=(128, 160)):
    """Seeded FPN maps p2..p6 of one image and r proposals for the box branch (recognizers_hybrid_head.py:291-339)."""
    g = torch.Generator().manual_seed(9500 + seed)
    h, w = hw
    feats = {f"p{k}": torch.randn(1, 256, math.ceil(h / 2 ** k), math.ceil(w / 2 ** k), generator=g) for k in range(2, 7)}
    cx = torch.rand(r, generator=g) * w
    cy = torch.rand(r, generator=g) * h
    bw = torch.exp(torch.rand(r, generator=g) * math.log(10.0)) * 12.0
    bh = bw * (0.15 + 0.6 * torch.rand(r, generator=g))
    ang = (torch.rand(r, generator=g) - 0.5) * 180.0
    props = torch.stack((cx, cy, bw, bh, ang), 1).float()
    k = r // 3                       # near-duplicates so that the rotated NMS has something to suppress
    props[k:2 * k] = props[:k] + torch.randn(k, 5, generator=g) * torch.tensor([1.0, 1.0, 0.8, 0.5, 2.0])
    props[:, 2:4] = props[:, 2:4].clamp_min(2.0)
    return feats, props, hw
