"""GPU diagnostic for the tcgen05 GEMM: exact small-integer problems whose result is known bit-for-bit,
printed as error summaries (not asserts) so one gpurun call tells where a layout/descriptor bug is."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from glass_text_spotting_b200 import ops, packing  # noqa: E402


def probe(rows, k, n, mode, taps=1, tag=""):
    g = torch.Generator().manual_seed(rows * 7 + k + n)
    a = torch.randint(-4, 5, (rows, k), generator=g).float()
    w = torch.randint(-4, 5, (n, taps * k), generator=g).float()
    pw = packing.pack_linear(w, None, k_p=taps * k, n_align=16)
    pw.cin_p = k
    a2 = packing.split_act(a).cuda()
    out = torch.full((rows, pw.n_p), float("nan"), device="cuda")
    shifts = list(range(taps))
    ops.conv_gemm(a2[0], a2[1], rows, k, shifts, pw, (1, rows, 1, 0), out_f32=out, ld_f32=pw.n_p,
                  out_geom=(rows, 1, 0), mode=mode)
    torch.cuda.synchronize()
    ref = torch.zeros(rows, n)
    for t in range(taps):
        sh = torch.zeros(rows, k)
        sh[: rows - t] = a[t:]
        ref += sh @ w[:, t * k:(t + 1) * k].t()
    got = out.cpu()[:, :n]
    err = (got - ref).abs()
    nan = torch.isnan(got).sum().item()
    print(f"[probe {tag}] rows={rows} k={k} n={n} taps={taps} mode={mode}: max_err={err.nan_to_num(1e9).max().item():.3g} "
          f"nan={nan} bad_rows={(err.nan_to_num(1e9).amax(1) > 0).sum().item()} bad_cols={(err.nan_to_num(1e9).amax(0) > 0).sum().item()}")
    if err.nan_to_num(1e9).max().item() > 0:
        bad = torch.nonzero(err.nan_to_num(1e9) > 0)
        print("   first bad (row,col):", bad[:8].tolist())
        print("   got:", got[bad[0, 0], :8].tolist(), "\n   ref:", ref[bad[0, 0], :8].tolist())


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for mode in (1, 0):
        probe(128, 64, 64, mode, tag="1tile-1kblock")
        probe(128, 256, 64, mode, tag="4kblocks")
        probe(128, 64, 256, mode, tag="n256")
        probe(100, 64, 16, mode, tag="partial-m,n16")
        probe(1000, 128, 80, mode, tag="multi-tile,n80")
        probe(128 * 300, 128, 512, mode, tag="persistent,2 n tiles")
        probe(512, 64, 64, mode, taps=3, tag="taps")
        probe(4096, 1024, 2048, mode, tag="long pipeline")
