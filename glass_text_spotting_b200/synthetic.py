"""Seeded synthetic detections for the post-processing row (tests, golden generation, bench.py)."""
import math

import torch


def make_postprocess_case(seed: int, n_lines: int, n_iso: int, img: float = 1024.0):
    """Seeded detections: text lines broken into overlapping word pieces (merge candidates: similar angle and height,
    IoA above / below the threshold), nested and crossing boxes, tiny boxes, low scores."""
    g = torch.Generator().manual_seed(seed)

    def u(a, b, n=1):
        return (torch.rand(n, generator=g) * (b - a) + a)
    rows = []
    for _ in range(n_lines):
        cx, cy = u(100, img - 100).item(), u(100, img - 100).item()
        ang = u(-180, 180).item()
        h = u(12, 60).item()
        pieces = int(torch.randint(2, 5, (1,), generator=g))
        t = math.radians(-ang)
        x = 0.0
        for _p in range(pieces):
            w = h * u(1.5, 4.0).item()
            overlap = u(-0.1, 0.6).item() * w  # negative: a gap (no merge)
            x += w / 2
            jitter_h = h * u(0.8, 1.25).item() if u(0, 1).item() < 0.8 else h * u(0.2, 0.4).item()
            da = u(-6, 6).item() if u(0, 1).item() < 0.8 else u(20, 60).item()
            if u(0, 1).item() < 0.15:
                da += 180.0  # flipped reading direction: "similar angle" through the 180 - diff branch
            px = cx + x * math.cos(t)
            py = cy + x * math.sin(t)
            rows.append([px, py, w, jitter_h, ((ang + da + 180) % 360) - 180, u(0.05, 1.0).item()])
            x += w / 2 - overlap
    for _ in range(n_iso):
        w = u(1, 200).item()
        rows.append([u(0, img).item(), u(0, img).item(), w, w * u(0.05, 1.0).item(), u(-180, 180).item(), u(0.01, 1.0).item()])
    t = torch.tensor(rows, dtype=torch.float32)
    perm = torch.randperm(len(t), generator=g)
    t = t[perm]
    return t[:, :5].contiguous(), t[:, 5].contiguous()
