"""A/B timing of the two LSTM kernels at the full-inference size (357 words, T = 32): the streaming fp32 kernel
(default) against the opt-in cluster-resident tensor-core kernel (GLASS_LSTM_CLUSTER=1).  The variant is a per-process
static, so each arm runs in its own child process.  GPU box only:

    python tools/lstm_ab.py            # prints ms per launch of both arms (CUDA events, 50 launches after 5 warm-ups)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, torch
sys.path.insert(0, %r)
from glass_text_spotting_b200 import ops
n, T, H = int(sys.argv[1]), 32, 256
g = torch.Generator().manual_seed(0)
gates = (torch.randn(n * T, 8 * H, generator=g) * 0.5).cuda()
whh_t = (torch.randn(2, H, 4 * H, generator=g) * 0.06).cuda()
out = torch.zeros((2, n * T, 2 * H), dtype=torch.float16, device="cuda")
for _ in range(5):
    ops.lstm_bidir(gates, whh_t, n, T, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    ops.lstm_bidir(gates, whh_t, n, T, out)
e1.record()
torch.cuda.synchronize()
print(e0.elapsed_time(e1) / 50)
""" % ROOT


def main():
    n = sys.argv[1] if len(sys.argv) > 1 else "357"
    for name, env in (("streaming fp32 (default)", {}), ("cluster mma.sync (GLASS_LSTM_CLUSTER=1)", {"GLASS_LSTM_CLUSTER": "1"})):
        r = subprocess.run([sys.executable, "-c", CHILD, n], env=dict(os.environ, **env), capture_output=True, text=True)
        print(f"{name}: {r.stdout.strip() or r.stderr[-300:]} ms per launch ({n} words)")


if __name__ == "__main__":
    main()
