"""GPU diagnostic: accuracy of the tcgen05 GEMM (fp16x3 split / single fp16) against an fp64 reference,
next to torch's fp32 CPU matmul, for growing K -- separates operand-representation error from the
tensor core's own accumulation error (signed mean error on all-positive data reveals a rounding bias)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glass_text_spotting_b200 import ops, packing  # noqa: E402


def run(rows, k, n, positive):
    g = torch.Generator().manual_seed(k + n)
    a = torch.randn(rows, k, generator=g)
    w = torch.randn(n, k, generator=g) / k ** 0.5
    if positive:
        a, w = a.abs(), w.abs()
    ref = a.double() @ w.double().t()
    cpu32 = (a @ w.t()).double()
    pw = packing.pack_linear(w, None)
    a2 = packing.split_act(a).cuda()
    res = {}
    for name, mode in (("split", ops.MODE_SPLIT), ("fast", ops.MODE_FAST)):
        _, of = ops.linear(a2, pw, want_split=False, want_f32=True, mode=mode)
        res[name] = of.cpu().double()[:, :n]
    # operand representation only: fp64 product of the rounded (hi+lo) operands
    ar = (a2[0].cpu().double() + a2[1].cpu().double()) / ops.ACT_SCALE
    def stats(x):
        d = x - ref
        return f"relL2={d.norm() / ref.norm():.2e} max={d.abs().max():.2e} mean_signed={d.mean():.2e}"
    print(f"K={k:6d} positive={positive}: cpu_fp32 [{stats(cpu32)}]  split [{stats(res['split'])}]  "
          f"fast [{stats(res['fast'])}]  A-repr relL2={((ar - a.double()).norm() / a.double().norm()):.2e}  |ref|~{ref.abs().mean():.2f}")


if __name__ == "__main__":
    for kbc in (1000, 4, 2, 1):
        ops.KB_PER_CHUNK = kbc
        print(f"---- kb_per_chunk = {kbc}")
        for positive in (False, True):
            for k in (576, 2304, 12544):
                run(256, k, 256, positive)
