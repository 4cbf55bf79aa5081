"""GPU diagnostic: host-side issue time of one full step (Python + ctypes + tensor-map encode) vs GPU time."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glass_text_spotting_b200 import weights, lib
from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN

model = B200GlassRCNN(weights.random_state_dict(0))
g = torch.Generator().manual_seed(1000)
imgs = torch.randint(0, 256, (4, 3, 1024, 1024), generator=g).float().cuda()
hw = torch.tensor([[1024, 1024]] * 4, dtype=torch.float32, device="cuda")
for _ in range(3):
    model.forward_device(imgs, hw)
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter()
    feats, det = model.detect(imgs, hw)
    t1 = time.perf_counter()
    counts = det["count"].cpu().tolist()
    t2 = time.perf_counter()
    probs, starts = model.recognize(imgs, feats, det, counts)
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print(f"detect issue {1e3*(t1-t0):.2f} ms | wait counts {1e3*(t2-t1):.2f} ms | recognize issue {1e3*(t3-t2):.2f} ms | "
          f"drain {1e3*(t4-t3):.2f} ms | total {1e3*(t4-t0):.2f} ms | launches {lib.load().glass_launch_count()}")
