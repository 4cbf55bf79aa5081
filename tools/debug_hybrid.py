"""GPU diagnostic: per-layer error of the local CNN (hybrid_net) vs the oracle, to localise parity gaps."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_common import calibrate_bn  # noqa: E402
from test_gpu_roi_heads import _boxes  # noqa: E402
from glass_text_spotting_b200 import ops  # noqa: E402
from glass_text_spotting_b200.modeling.roi_heads import B200GlassROIHeads  # noqa: E402
from oracle import d2_ops  # noqa: E402
from oracle import model as om  # noqa: E402

o = om.build_oracle(seed=1)
g = torch.Generator().manual_seed(7)
h, w, n = 192, 256, 2
images = torch.randint(0, 256, (n, 3, h, w), generator=g).float()
mean = torch.tensor(o.cfg.pixel_mean).view(1, 3, 1, 1)
norm = images - mean
_ = {f"p{k}": torch.randn(n, 256, -(-h // 2 ** k), -(-w // 2 ** k), generator=g) for k in range(2, 7)}
boxes = [_boxes(g, 3, h, w, 30.0, 150.0), _boxes(g, 2, h, w, 30.0, 150.0)]
net = o.roi_heads.hybrid_net.ConvNet
with torch.no_grad():
    crops = torch.cat([d2_ops.roi_pooler([norm[i:i + 1]], [boxes[i]], (128, 128), [1.0], 2) for i in range(n)])
    calibrate_bn(net, lambda: net(crops))
    # oracle intermediates
    ref = {}
    x = F.relu(net.bn0_1(net.conv0_1(crops))); ref["hyb.c01"] = x
    x = F.relu(net.bn0_2(net.conv0_2(x))); ref["hyb.c02"] = x
    x = F.max_pool2d(x, 2, 2, 0); ref["hyb.pool1"] = x
    for b, blk in enumerate(net.layer1):
        x = blk(x); ref[f"hyb.l1.{b}.out"] = x
    x = F.relu(net.bn1(net.conv1(x))); ref["hyb.c1"] = x
    x = F.max_pool2d(x, 2, 2, 0); ref["hyb.pool2"] = x
    for b, blk in enumerate(net.layer2):
        x = blk(x); ref[f"hyb.l2.{b}.out"] = x
    x = F.relu(net.bn2(net.conv2(x))); ref["hyb.c2"] = x
    x = F.max_pool2d(x, kernel_size=2, stride=(2, 1), padding=(0, 1)); ref["hyb.pool3"] = x
    for b, blk in enumerate(net.layer3):
        x = blk(x); ref[f"hyb.l3.{b}.out"] = x
    x = F.relu(net.bn3(net.conv3(x))); ref["hyb.c3"] = x
    for b, blk in enumerate(net.layer4):
        x = blk(x); ref[f"hyb.l4.{b}.out"] = x
    x = F.relu(net.bn4_1(net.conv4_1(x))); ref["final"] = x

heads = B200GlassROIHeads(o.state_dict())
crops_act = ops.Act.from_nchw(crops.cuda(), cp=8)
heads._word_cap = crops.shape[0]
fused = ops.Act(crops.shape[0], 512, 8, 32)
heads.hybrid_net(crops_act, fused)
torch.cuda.synchronize()
print(f"{'layer':16s} {'relL2':>10s} {'max_err':>10s} {'scale':>9s} {'viol':>6s}   bn scale max (folded)")
for name, r in ref.items():
    got = fused.to_nchw()[:, :256].cpu() if name == "final" else heads.ws._acts[name].to_nchw().cpu()
    e = (got - r).abs()
    tol = 1e-4 * max(1.0, r.abs().max().item()) + 1e-3 * r.abs()
    print(f"{name:16s} {((got - r).norm() / r.norm()).item():10.2e} {e.max().item():10.2e} {r.abs().max().item():9.2e} "
          f"{(e > tol).sum().item():6d}")
for nm in ["bn0_1", "bn0_2", "bn1", "bn2", "bn3", "bn4_1"]:
    bn = getattr(net, nm)
    s = bn.weight / torch.sqrt(bn.running_var + 1e-5)
    print(nm, "folded scale max", s.abs().max().item(), "min var", bn.running_var.min().item())
