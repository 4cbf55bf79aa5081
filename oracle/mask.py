"""oracle/mask.py -- TEST INFRASTRUCTURE, not product code.

CPU restatement of the mask branch, which the reference runs when
MODEL.ROI_MASK_HEAD.MASK_INFERENCE is on (glass/modeling/fusion/recognizers_hybrid_head.py:341-442, :595-601):
  * mask pooler: ROIPooler(14, all five FPN levels, sampling 0, ROIAlignRotated) -- oracle/d2_ops.roi_pooler;
  * mask head: detectron2 v0.6 MaskRCNNConvUpsampleHead as configured by configs/glass_finetune_totaltext.yaml
    (ROI_MASK_HEAD: NUM_CONV 4, CONV_DIM 256, NORM ""), which glass/modeling/roi_heads/rotated_mask_head.py:409-442
    subclasses without changing the layers: 4 x (conv3x3 + ReLU), ConvTranspose2d(2, stride 2) + ReLU, conv1x1 -> 1;
    plain torch layers, so torch itself is the ground truth for them;
  * mask_rcnn_inference (detectron2): sigmoid of the logits, one class;
  * the rotated paste: paste_masks_in_image / _do_paste_mask, glass/postprocess/post_processor_academic.py:187-335
    (5-column branch), pinned by tests/golden/paste_masks.pt which the reference's own function produced
    (tools/make_golden_paste.py).
"""
from typing import Tuple

import torch
import torch.nn.functional as F
from torch import nn


class MaskHead(nn.Module):
    """state_dict keys as in detectron2: mask_fcn1..4, deconv, predictor (under roi_heads.mask_head.)."""

    def __init__(self, in_channels: int = 256, conv_dim: int = 256, num_conv: int = 4, num_classes: int = 1):
        super().__init__()
        cur = in_channels
        for k in range(num_conv):
            setattr(self, f"mask_fcn{k + 1}", nn.Conv2d(cur, conv_dim, 3, 1, 1))
            cur = conv_dim
        self.num_conv = num_conv
        self.deconv = nn.ConvTranspose2d(cur, conv_dim, 2, 2, 0)
        self.predictor = nn.Conv2d(conv_dim, num_classes, 1, 1, 0)

    def layers(self, x: torch.Tensor) -> torch.Tensor:
        for k in range(self.num_conv):
            x = F.relu(getattr(self, f"mask_fcn{k + 1}")(x))
        x = F.relu(self.deconv(x))
        return self.predictor(x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """pooled features [K, 256, 14, 14] -> pred_masks [K, 1, 28, 28] probabilities (mask_rcnn_inference)."""
        return self.layers(x).sigmoid()


def seeded_mask_head(seed: int) -> MaskHead:
    """d2 initialises the convs with c2_msra_fill and the predictor with normal(0.001); scaled up here so that the
    random head produces masks that are not all 0.5."""
    g = torch.Generator().manual_seed(seed)
    m = MaskHead().eval()
    with torch.no_grad():
        for name, p in m.named_parameters():
            if p.dim() > 1:
                fan = p[0].numel() if "deconv" not in name else p.shape[0] * 4
                p.copy_(torch.randn(p.shape, generator=g) * (2.0 / fan) ** 0.5)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
        m.predictor.weight.mul_(6.0)
    return m


def do_paste_mask_rotated(masks: torch.Tensor, boxes: torch.Tensor, img_h: int, img_w: int) -> torch.Tensor:
    """_do_paste_mask, 5-column branch with skip_empty=False (post_processor_academic.py:237-335).
    masks [N,1,M,M], boxes [N,5] -> soft masks [N, img_h, img_w]."""
    n = masks.shape[0]
    cx, cy, w, h, a = torch.split(boxes, 1, dim=1)
    a = torch.deg2rad(a)
    cos_a, sin_a = torch.cos(a), torch.sin(a)
    rot = torch.reshape(torch.stack([cos_a, sin_a, -sin_a, cos_a], 1), (-1, 2, 2))
    x0, x1 = cx - w / 2, cx + w / 2   # sin_t = 0, cos_t = 1 in the reference
    y0, y1 = cy - h / 2, cy + h / 2
    grid = torch.zeros([n, img_h, img_w, 2], dtype=torch.float32)
    for i in range(n):
        img_y = torch.arange(0, img_h, dtype=torch.float32) + 0.5 - cy[i]
        img_x = torch.arange(0, img_w, dtype=torch.float32) + 0.5 - cx[i]
        gx = img_x[None, :].expand(img_y.size(0), img_x.size(0))
        gy = img_y[:, None].expand(img_y.size(0), img_x.size(0))
        igrid = torch.stack([gx, gy], dim=2) @ rot[i]
        igrid[..., 0] += cx[i]
        igrid[..., 1] += cy[i]
        igrid[..., 0] = (igrid[..., 0] - x0[i]) / (x1[i] - x0[i]) * 2 - 1
        igrid[..., 1] = (igrid[..., 1] - y0[i]) / (y1[i] - y0[i]) * 2 - 1
        grid[i] = igrid
    return F.grid_sample(masks.float(), grid, align_corners=False)[:, 0]


def paste_masks_in_image(masks: torch.Tensor, boxes: torch.Tensor, image_shape: Tuple[int, int],
                         threshold: float = 0.5) -> torch.Tensor:
    """post_processor_academic.py:187-234: masks [N,M,M] probabilities -> [N, H, W] bool."""
    n = len(masks)
    if n == 0:
        return masks.new_empty((0,) + tuple(image_shape), dtype=torch.uint8)
    soft = do_paste_mask_rotated(masks[:, None, :, :], boxes, image_shape[0], image_shape[1])
    return soft >= threshold
