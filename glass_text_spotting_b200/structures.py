"""Minimal stand-ins for the detectron2 structures that cross the hot path's boundary
(SURVEY.md 8a "types crossing the path"): RotatedBoxes, Instances, ImageList.  detectron2 itself is
not installable offline; when it is present the integration layer converts to the
real classes -- the field names and semantics here are the same."""
from typing import Any, Dict, List, Sequence, Tuple

import torch


class RotatedBoxes:
    """[n,5] = (cx, cy, w, h, angle_deg CCW), like detectron2.structures.RotatedBoxes."""

    def __init__(self, tensor: torch.Tensor):
        self.tensor = tensor.reshape(-1, 5).float()

    def __len__(self):
        return self.tensor.shape[0]

    def __getitem__(self, item):
        return RotatedBoxes(self.tensor[item].reshape(-1, 5))

    def area(self):
        return self.tensor[:, 2] * self.tensor[:, 3]

    def normalize_angles(self):
        self.tensor[:, 4] = (self.tensor[:, 4] + 180.0) % 360.0 - 180.0

    def clip(self, box_size: Tuple[int, int], clip_angle_threshold: float = 1.0):
        h, w = box_size
        self.normalize_angles()
        t = self.tensor
        idx = torch.where(torch.abs(t[:, 4]) <= clip_angle_threshold)[0]
        x1 = (t[idx, 0] - t[idx, 2] / 2.0).clamp(min=0, max=w)
        y1 = (t[idx, 1] - t[idx, 3] / 2.0).clamp(min=0, max=h)
        x2 = (t[idx, 0] + t[idx, 2] / 2.0).clamp(min=0, max=w)
        y2 = (t[idx, 1] + t[idx, 3] / 2.0).clamp(min=0, max=h)
        t[idx, 0] = (x1 + x2) / 2.0
        t[idx, 1] = (y1 + y2) / 2.0
        t[idx, 2] = torch.min(t[idx, 2], x2 - x1)
        t[idx, 3] = torch.min(t[idx, 3], y2 - y1)

    def nonempty(self, threshold: float = 0.0):
        return (self.tensor[:, 2] > threshold) & (self.tensor[:, 3] > threshold)

    def scale(self, scale_x: float, scale_y: float):
        """d2 RotatedBoxes.scale: anisotropic scaling of a rotated box."""
        t = self.tensor
        t[:, 0] *= scale_x
        t[:, 1] *= scale_y
        theta = t[:, 4] * 3.141592653589793 / 180.0
        c, s = torch.cos(theta), torch.sin(theta)
        t[:, 2] *= torch.sqrt((scale_x * c) ** 2 + (scale_y * s) ** 2)
        t[:, 3] *= torch.sqrt((scale_x * s) ** 2 + (scale_y * c) ** 2)
        t[:, 4] = torch.atan2(scale_x * s, scale_y * c) * 180.0 / 3.141592653589793

    def to(self, device):
        return RotatedBoxes(self.tensor.to(device))


class Instances:
    """Per-image field bag with len / indexing / .to, like detectron2.structures.Instances."""

    def __init__(self, image_size: Tuple[int, int], **fields: Any):
        self._image_size = tuple(image_size)
        self._fields: Dict[str, Any] = {}
        for k, v in fields.items():
            self.set(k, v)

    @property
    def image_size(self):
        return self._image_size

    def set(self, name: str, value: Any):
        if self._fields:
            assert len(value) == len(self), f"field {name} has length {len(value)}, expected {len(self)}"
        self._fields[name] = value

    def has(self, name: str) -> bool:
        return name in self._fields

    def get(self, name: str):
        return self._fields[name]

    def get_fields(self):
        return self._fields

    def __getattr__(self, name):
        if name.startswith("_") or name not in self._fields:
            raise AttributeError(name)
        return self._fields[name]

    def __setattr__(self, name, value):
        if name.startswith("_"):
            super().__setattr__(name, value)
        else:
            self.set(name, value)

    def __len__(self):
        for v in self._fields.values():
            return len(v)
        return 0

    def __getitem__(self, item):
        out = Instances(self._image_size)
        for k, v in self._fields.items():
            out.set(k, v[item])
        return out

    def to(self, device):
        out = Instances(self._image_size)
        for k, v in self._fields.items():
            out.set(k, v.to(device) if hasattr(v, "to") else v)
        return out


class ImageList:
    """Batch of images padded to a common size divisible by ``size_divisibility`` (d2 ImageList.from_tensors).
    ``pad_value`` lets the caller pad RAW pixels with the pixel mean so the normalised padding is exactly 0."""

    def __init__(self, tensor: torch.Tensor, image_sizes: List[Tuple[int, int]], normalized: bool = False):
        # ``normalized``: the tensor already holds (x - mean) / std (what detectron2's preprocess_image produces); the B200
        # modules built with a non-trivial pixel mean expect RAW pixels and refuse such a list (see d2_adapter.py)
        self.tensor, self.image_sizes, self.normalized = tensor, image_sizes, normalized

    @staticmethod
    def from_tensors(tensors: Sequence[torch.Tensor], size_divisibility: int = 0, pad_value=0.0) -> "ImageList":
        sizes = [(int(t.shape[-2]), int(t.shape[-1])) for t in tensors]
        mh, mw = max(s[0] for s in sizes), max(s[1] for s in sizes)
        if size_divisibility > 1:
            d = size_divisibility
            mh, mw = (mh + d - 1) // d * d, (mw + d - 1) // d * d
        c = tensors[0].shape[0]
        pv = torch.as_tensor(pad_value, dtype=torch.float32, device=tensors[0].device).reshape(-1, 1, 1)
        out = torch.empty((len(tensors), c, mh, mw), dtype=torch.float32, device=tensors[0].device)
        out[:] = pv
        for i, t in enumerate(tensors):
            out[i, :, : t.shape[-2], : t.shape[-1]] = t
        return ImageList(out, sizes)
