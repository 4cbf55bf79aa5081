"""B200GlassRunner: the single-image Python API of the reference (glass/inference/glass_runner.py:72-109) on
the B200 hot path: numpy HWC image -> device resize (bilinear, align_corners=False) -> model -> boxes rescaled
to the original image -> the post-processor (glass_runner.py:106; here the device-side merge loop of
``postprocess.B200PostProcessor``)."""
from typing import Dict

import numpy as np
import torch

from . import ops
from .modeling.glass_rcnn import B200GlassRCNN
from .postprocess import B200PostProcessor
from .structures import Instances
from .text import TextDecoder


class B200GlassRunner:
    def __init__(self, state_dict: Dict[str, torch.Tensor], min_target_size: int = 1200, max_target_size: int = 1600,
                 max_upscale_ratio: float = 2.0, input_format: str = "BGR", device="cuda", post_processor="default",
                 **model_kwargs):
        # defaults: configs/glass_finetune_totaltext.yaml:22-25 (INFERENCE_TH_TEST block)
        self.min_target_size, self.max_target_size = min_target_size, max_target_size
        self.max_upscale_ratio, self.input_format, self.device = max_upscale_ratio, input_format, device
        assert self.input_format in ["RGB", "BGR", "GREY"], self.input_format      # glass_runner.py:66
        self.model = B200GlassRCNN(state_dict, device=device, **model_kwargs)
        self.text_decoder = TextDecoder()
        # build_post_processor(cfg) (glass_runner.py:70); pass None to get the raw detections
        self.post_processor = B200PostProcessor() if post_processor == "default" else post_processor

    def get_inference_scale_ratio(self, image_shape) -> float:
        """glass_runner.py:111-121."""
        m = max(image_shape[:2])
        if m > self.max_target_size:
            return self.max_target_size / m
        if m < self.min_target_size:
            return min(self.max_upscale_ratio, self.min_target_size / m)
        return 1

    def image_to_tensor(self, original_image: np.ndarray):
        """glass_runner.py:83-87 (channel order / greyscale) + :123-148 (resize), on the device: uint8 H2D, then one
        resize/convert kernel (the RGB flip is fused into it; the greyscale conversion is the reference's uint8 numpy
        expression, glass/utils/common_utils.py:38-41, applied before the upload)."""
        if self.input_format == "GREY":
            grey = np.uint8(0.2125 * original_image[:, :, 0] + 0.7154 * original_image[:, :, 1]
                            + 0.0721 * original_image[:, :, 2])
            original_image = np.repeat(grey[:, :, None], 3, axis=2)
        h, w = original_image.shape[:2]
        scale = self.get_inference_scale_ratio(original_image.shape)
        nh, nw = (int(np.round(scale * h)), int(np.round(scale * w))) if scale != 1 else (h, w)
        src = torch.as_tensor(np.ascontiguousarray(original_image)).to(self.device, non_blocking=True)
        return ops.resize_bilinear_u8(src, (nh, nw), flip_channels=(self.input_format == "RGB")), scale

    @torch.no_grad()
    def __call__(self, original_image: np.ndarray) -> Instances:
        h, w = original_image.shape[:2]
        tensor, scale = self.image_to_tensor(original_image)
        preds = self.model([{"image": tensor, "height": tensor.shape[1], "width": tensor.shape[2]}])[0]["instances"]
        if scale != 1:
            preds.pred_boxes.scale(1 / scale, 1 / scale)
        preds._image_size = (h, w)
        if self.post_processor is not None:
            preds = self.post_processor(preds)
        return preds

    def read_text(self, preds: Instances):
        """pred_text_prob -> [{"text", "score", "character_scores"}] (TextEncoder.decode_attention)."""
        return self.text_decoder.decode_probs(preds.pred_text_prob)
