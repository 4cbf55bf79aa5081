"""Idle time between the kernels of one graph-replayed full_bs4 step (torch.profiler / CUPTI kernel records).
  python tools/gap_profile.py [out.txt]
Prints the step's span, the summed kernel time per stream-merged timeline, the idle time, and the largest gaps with the
kernels on either side."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from glass_text_spotting_b200 import weights  # noqa: E402
from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN  # noqa: E402

g = torch.Generator().manual_seed(1000)
x = torch.randint(0, 256, (4, 3, 1024, 1024), generator=g, dtype=torch.uint8).cuda().float()
hw = torch.tensor([[1024, 1024]] * 4, dtype=torch.float32, device="cuda")
model = B200GlassRCNN(weights.random_state_dict(0))
for _ in range(5):
    model.graph_step(x, hw)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        model.graph_step(x, hw)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "glass" in e.name or
      (e.device_type == torch.autograd.DeviceType.CUDA and "conv_gemm" in e.name)]
ev = sorted(ev, key=lambda e: e.time_range.start)
# last step: from the last stem_s2d kernel on
starts = [i for i, e in enumerate(ev) if "stem_s2d" in e.name]
ev = ev[starts[-1]:]
t0 = ev[0].time_range.start
t1 = max(e.time_range.end for e in ev)
# merged busy time
busy, cur_end = 0.0, t0
gaps = []
prev = None
for e in ev:
    s, en = e.time_range.start, e.time_range.end
    if s > cur_end:
        gaps.append((s - cur_end, prev.name if prev else "", e.name))
    busy += max(0.0, en - max(s, cur_end))
    if en > cur_end:
        cur_end, prev = en, e
out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
span = t1 - t0
print(f"kernels {len(ev)}  span {span / 1e3:.3f} ms  busy {busy / 1e3:.3f} ms  idle {(span - busy) / 1e3:.3f} ms "
      f"({100 * (span - busy) / span:.1f} %)  gaps {len(gaps)}", file=out)


def short(n):
    n = n.replace("glass::", "").replace("void ", "")
    return n[:44]


by = {}
for gdur, a, b in gaps:
    k = (short(a), short(b))
    c = by.setdefault(k, [0, 0.0])
    c[0] += 1
    c[1] += gdur
print("idle by (previous kernel -> next kernel): count, total us, mean us", file=out)
for k, (c, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{t:8.1f} us  {c:4d} x {t / c:6.2f}  {k[0]} -> {k[1]}", file=out)
