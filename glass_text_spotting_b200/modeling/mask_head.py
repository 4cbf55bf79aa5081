"""Mask branch on B200 kernels: what the reference runs when MODEL.ROI_MASK_HEAD.MASK_INFERENCE is on
(glass/modeling/fusion/recognizers_hybrid_head.py:341-442 ``_forward_mask``, :595-601; head =
glass/modeling/roi_heads/rotated_mask_head.py:409-442 on detectron2's MaskRCNNConvUpsampleHead; paste =
glass/postprocess/post_processor_academic.py:187-335).

mask pooler (14x14, five FPN levels, adaptive sampling grid) -> 4 x (conv3x3 + ReLU) -> deconv 2x2/s2 + ReLU ->
conv1x1 -> sigmoid -> pred_masks [K, 1, 28, 28]; ``paste`` turns them into per-detection image masks.
Everything dense runs on glass_conv_gemm: the deconv is ONE GEMM with N = 4 x 256 (a 256-column block per output
sub-pixel), the predictor a block-diagonal GEMM over those columns, so a pooled pixel's row ends with its 2 x 2 logits.
"""
from typing import Dict, Sequence, Tuple

import torch

from .. import ops, packing
from ..ops import Act
from .backbone import Workspace


class B200MaskHead:
    def __init__(self, state_dict: Dict[str, torch.Tensor], prefix: str = "roi_heads.mask_head.", device="cuda",
                 mode: int = ops.MODE_SPLIT, in_features: Sequence[str] = ("p2", "p3", "p4", "p5", "p6"),
                 strides: Sequence[int] = (4, 8, 16, 32, 64), pooler_resolution: int = 14, sampling_ratio: int = 0):
        sd = {k[len(prefix):]: v.detach().float().cpu() for k, v in state_dict.items() if k.startswith(prefix)}
        assert "deconv.weight" in sd and "predictor.weight" in sd, "no mask head weights under " + prefix
        self.device, self.mode = device, mode
        self.in_features, self.strides = tuple(in_features), tuple(strides)
        self.res, self.sampling = pooler_resolution, sampling_ratio
        self.convs = []
        k = 1
        while f"mask_fcn{k}.weight" in sd:
            self.convs.append(packing.pack_conv(sd[f"mask_fcn{k}.weight"], None, sd[f"mask_fcn{k}.bias"], (1, 1), (1, 1),
                                                device=device))
            k += 1
        wd = sd["deconv.weight"]                       # [cin, cout, 2, 2]
        cin, cout = wd.shape[0], wd.shape[1]
        assert tuple(wd.shape[2:]) == (2, 2) and cin == 256 and cout == 256
        w2 = wd.permute(2, 3, 1, 0).reshape(4 * cout, cin)   # row (dy*2+dx)*cout + co
        self.deconv = packing.pack_linear(w2, sd["deconv.bias"].repeat(4), device=device)
        wp = sd["predictor.weight"]
        assert wp.shape[0] == 1, "one mask class (ROI_HEADS.NUM_CLASSES = 1)"
        w3 = torch.zeros((4, 4 * cout))
        for s in range(4):
            w3[s, s * cout:(s + 1) * cout] = wp.view(-1)
        self.predictor = packing.pack_linear(w3, sd["predictor.bias"].repeat(4), n_align=16, device=device)
        self.ws = Workspace(device)

    @torch.no_grad()
    def forward(self, features: Dict[str, Act], rois: torch.Tensor, cap: int = 0) -> torch.Tensor:
        """rois fp32 [K, 6] (batch, cx, cy, w, h, angle) -> pred_masks fp32 [K, 1, 28, 28] (probabilities)."""
        K, r = rois.shape[0], self.res
        out = torch.empty((K, 1, 2 * r, 2 * r), dtype=torch.float32, device=rois.device)
        if K == 0:
            return out
        ws, cap = self.ws, max(cap, K)
        x = ws.act("mask.pool", K, 256, r, r, cap=cap)
        ops.roi_align_rotated([features[f] for f in self.in_features], rois, (r, r), [1.0 / s for s in self.strides],
                              self.sampling, min_level=2, out_f32=False,
                              out_split=(x.buf, x.hp, x.wp, x.border, 0, x.cp))
        for i, w in enumerate(self.convs):
            x = ops.conv2d(x, w, relu=True, out=ws.act(f"mask.fcn{i}", K, 256, r, r, cap=cap), mode=self.mode)
        rows = x.rows
        d = ws.rows("mask.deconv", rows, 4 * 256, cap * x.hp * x.wp)
        ops.conv_gemm(x.hi, x.lo, rows, x.cp, [0], self.deconv, (x.n, x.hp, x.wp, x.border), out_hi=d[0], out_lo=d[1],
                      out_geom=(x.hp, x.wp, x.border), ld_out=4 * 256, relu_post=True, mode=self.mode)
        lg = ws.raw("mask.logits", (cap * x.hp * x.wp, 16), torch.float32)
        ops.conv_gemm(d[0], d[1], rows, 4 * 256, [0], self.predictor, (1, rows, 1, 0), out_f32=lg, ld_f32=16,
                      out_geom=(rows, 1, 0), mode=self.mode)
        ops.mask_finalize(lg, K, r, r, out)
        return out

    __call__ = forward

    @staticmethod
    def paste(pred_masks: torch.Tensor, boxes: torch.Tensor, image_shape: Tuple[int, int], threshold: float = 0.5,
              want_soft: bool = False):
        """paste_masks_in_image: pred_masks [K, 1, M, M] or [K, M, M], boxes [K, 5] (image coordinates) ->
        bool [K, H, W] (and the sampled soft masks when ``want_soft``)."""
        m = pred_masks[:, 0] if pred_masks.dim() == 4 else pred_masks
        return ops.paste_masks_rotated(m.contiguous(), boxes.contiguous().float(), image_shape, threshold, want_soft)
