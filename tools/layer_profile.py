"""Per-launch table of the conv GEMM for one step of a bench workload (CUDA events around every launch).
  python tools/layer_profile.py [full_bs4|backbone_bs8]
Columns: M rows, N, K (taps), ms, algorithmic TFLOP/s, issued (x3) TFLOP/s, share of the GEMM time."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from glass_text_spotting_b200 import ops, weights  # noqa: E402
from glass_text_spotting_b200.modeling.backbone import B200ResNetFPN  # noqa: E402
from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "full_bs4"
B = 4 if wl == "full_bs4" else 8
g = torch.Generator().manual_seed(1000)
x = torch.randint(0, 256, (B, 3, 1024, 1024), generator=g, dtype=torch.uint8).cuda().float()
hw = torch.tensor([[1024, 1024]] * B, dtype=torch.float32, device="cuda")
if wl == "full_bs4":
    model = B200GlassRCNN(weights.random_state_dict(0))
    step = lambda: model.forward_device(x, hw)
else:
    model = B200ResNetFPN(weights.random_backbone_state_dict(0))
    step = lambda: model(x)
for _ in range(3):
    step()
torch.cuda.synchronize()
REP = 3
ops.PROFILE = []
for _ in range(REP):
    step()
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
n = len(prof) // REP
rows = []
for i in range(n):
    ms = sum(prof[r * n + i][0].elapsed_time(prof[r * n + i][1]) for r in range(REP)) / REP
    rows.append((i, prof[i][3], ms, ops.profile_flops(prof[i])))
tot = sum(r[2] for r in rows)
print(f"# {wl}: {n} conv_gemm launches, {tot:.3f} ms, {sum(r[3] for r in rows) / tot / 1e9:.1f} TFLOP/s algorithmic")
print("idx      M      N      K taps a_ld res f32     ms   alg_TF  iss_TF  share")
for i, m, ms, fl in rows:
    print(f"{i:3d} {m['m']:8d} {m['n']:5d} {m['k']:6d} {m['taps']:3d} {m['a_ld']:4d} {int(m['res']):3d} {int(m['f32']):3d} "
          f"{ms:7.3f} {fl / ms / 1e9:7.1f} {3 * fl / ms / 1e9:7.1f} {100 * ms / tot:5.1f}%")
