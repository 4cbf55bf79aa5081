"""Numerical feasibility probe for the round-2 plan: do the recurrent products of the BiLSTM
(recognizer_encoder.py:136-144) and of the decoder's GRU (prediction_aster.py:291-302) keep fp32-grade accuracy when they
move to tcgen05 with split-fp16 operands (hi + lo, 22 bits; 3 MMAs hi*hi + hi*lo + lo*hi, fp32 accumulation), the way
the conv GEMM already does?  CPU emulation: operands are quantised exactly like csrc/common.cuh's split16 (power-of-two
pre-scale, round-to-nearest fp16 hi, fp16 residual lo), the three partial products are formed in fp64 and rounded to fp32
once per 64-wide k-block (the chunked accumulation), the cell non-linearities stay fp32.  Reference: torch.nn.LSTM / GRUCell
in fp32 and in fp64.  Prints the error after 32 (LSTM) / 26 (GRU) recurrent steps.

    python tools/lstm_split_probe.py
"""
import torch


def split16(x: torch.Tensor, scale: float):
    hi = (x * scale).to(torch.float16)
    lo = (x * scale - hi.float()).to(torch.float16)
    return hi.double(), lo.double()


def split_matmul(a: torch.Tensor, w: torch.Tensor, a_scale=16.0):
    """a [m,k] @ w[n,k]^T with both operands split; w rows pre-scaled by a power of two into [256, 512)."""
    w_scale = torch.exp2(torch.floor(torch.log2(256.0 / w.abs().amax(1).clamp_min(1e-30))) + 1.0).clamp(max=2.0 ** 24)
    ah, al = split16(a, a_scale)
    wh, wl = split16(w * w_scale[:, None], 1.0)
    acc = torch.zeros(a.shape[0], w.shape[0], dtype=torch.float32)
    for k0 in range(0, a.shape[1], 64):      # one fp32 rounding per 64-wide k-block
        s = slice(k0, k0 + 64)
        part = ah[:, s] @ wh[:, s].T + ah[:, s] @ wl[:, s].T + al[:, s] @ wh[:, s].T
        acc = acc + part.float()
    return acc / (a_scale * w_scale[None, :].float())


def lstm_probe(seed=0, k=64, t=32, h=256):
    g = torch.Generator().manual_seed(seed)
    lstm = torch.nn.LSTM(h, h, batch_first=True)
    for p in lstm.parameters():
        torch.nn.init.uniform_(p, -1 / 16, 1 / 16, generator=g)
    x = torch.randn(k, t, h, generator=g)
    with torch.no_grad():
        ref32, _ = lstm(x)
        ref64, _ = lstm.double()(x.double())
        lstm.float()
        wih, whh = lstm.weight_ih_l0, lstm.weight_hh_l0
        b = lstm.bias_ih_l0 + lstm.bias_hh_l0
        gates_in = x.reshape(-1, h) @ wih.T + b           # the input projection already runs on the conv GEMM
        gates_in = gates_in.view(k, t, 4 * h)
        hs, c, hcur = [], torch.zeros(k, h), torch.zeros(k, h)
        for s in range(t):
            gt = gates_in[:, s] + split_matmul(hcur, whh)
            i, f, gg, o = gt.chunk(4, 1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            hcur = torch.sigmoid(o) * torch.tanh(c)
            hs.append(hcur)
        got = torch.stack(hs, 1)
    e_split = (got.double() - ref64).abs().max().item()
    e_fp32 = (ref32.double() - ref64).abs().max().item()
    return e_split, e_fp32, ref64.abs().max().item()


def gru_probe(seed=0, k=64, steps=26, h=256):
    g = torch.Generator().manual_seed(seed)
    cell = torch.nn.GRUCell(2 * h, h)
    for p in cell.parameters():
        torch.nn.init.uniform_(p, -1 / 16, 1 / 16, generator=g)
    xs = torch.randn(steps, k, 2 * h, generator=g)
    with torch.no_grad():
        h32 = torch.zeros(k, h)
        h64 = torch.zeros(k, h, dtype=torch.float64)
        hsp = torch.zeros(k, h)
        c64 = torch.nn.GRUCell(2 * h, h).double()
        c64.load_state_dict({n: p.double() for n, p in cell.state_dict().items()})
        for s in range(steps):
            h32 = cell(xs[s], h32)
            h64 = c64(xs[s].double(), h64)
            gi = split_matmul(xs[s], cell.weight_ih) + cell.bias_ih
            gh = split_matmul(hsp, cell.weight_hh) + cell.bias_hh
            ir, iz, inn = gi.chunk(3, 1)
            hr, hz, hn = gh.chunk(3, 1)
            r, z = torch.sigmoid(ir + hr), torch.sigmoid(iz + hz)
            n = torch.tanh(inn + r * hn)
            hsp = (1 - z) * n + z * hsp
    return (hsp.double() - h64).abs().max().item(), (h32.double() - h64).abs().max().item(), h64.abs().max().item()


if __name__ == "__main__":
    torch.set_num_threads(8)
    for name, fn in (("BiLSTM recurrence, 32 steps", lstm_probe), ("decoder GRU, 26 steps", gru_probe)):
        rows = [fn(seed) for seed in range(3)]
        print(f"{name}: max |err| vs fp64 -- split-fp16 x3: {max(r[0] for r in rows):.2e}, torch fp32: "
              f"{max(r[1] for r in rows):.2e} (|h| up to {max(r[2] for r in rows):.2f}; tolerance atol 1e-4)")
