#!/usr/bin/env python
"""bench.py -- images/sec @1024x1024 of GLASS's dense forward path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle) on host cores

One "step" = one pass of the hot path over one synthetic batch.  The workload is named in
config.workload (see WORKLOADS).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[3]: the configuration the images/sec metric is quoted on
    "full_bs4": dict(batch=4, h=1024, w=1024, full=True,
                     desc="full GLASS inference (backbone+RPN+rotated RoI+global-local recognizer), bs=4/GPU, "
                          "1024x1024 synthetic (BASELINE.json configs[3])"),
    # BASELINE.json configs[1]
    "backbone_bs8": dict(batch=8, h=1024, w=1024, full=False,
                         desc="ResNet-50+FPN backbone only, bs=8/GPU, 1024x1024 synthetic (BASELINE.json configs[1])"),
}
WORKLOADS["roialign_512"] = dict(batch=1, h=1024, w=1024, full=False, roialign=True,
                                 desc="RotatedROIAlign microbench: 512 rotated RoIs over 5 FPN levels, 7x7, sampling 2 "
                                      "(BASELINE.json configs[2])")
WORKLOADS["postprocess_bs4"] = dict(batch=4, h=1024, w=1024, full=False, postprocess=True,
                                   desc="word post-processor (merge loop + text-score filter, SURVEY.md 8f #1) on the "
                                        "detections of 4 images x 100 words (synthetic broken text lines)")
WORKLOADS["totaltext_loop"] = dict(batch=4, h=1024, w=1024, full=True, totaltext=True,
                                   desc="TotalText-shape eval loop (BASELINE.json configs[4]): uint8 1024x1024 images -> "
                                        "device resize to 1200x1200 (glass_finetune_totaltext.yaml MIN_SIZE_TEST) -> pad "
                                        "1216 -> full GLASS inference -> word post-processor -> 24 KB/image records -> "
                                        "one NCCL all-gather, bs=4/GPU")
WORKLOADS["mask_bs4"] = dict(batch=4, h=1024, w=1024, full=False, mask=True,
                             desc="mask branch (SURVEY.md 8f #3: 14x14 rotated mask pooler over 5 FPN levels -> 4 conv3x3 + "
                                  "deconv + predictor -> sigmoid -> rotated paste at 1024x1024) for 4 images x 89 detections")
MASK_CPU_DETECTIONS = 16  # mask_bs4's CPU leg pastes this many detections per image (bounded sample; stated in its line)


def _config(name, n_gpus):
    """The `config` object of the JSON line -- built by ONE function for both arms (`--impl b200` and `--impl reference`)
    so that the two lines describe the same workload key for key; run-dependent facts (words found, chunk policy,
    clocks) live outside it."""
    wl = WORKLOADS[name]
    return {"workload": wl["desc"], "name": name, "images_per_step_per_gpu": wl["batch"],
            "global_batch": n_gpus * wl["batch"], "image_hw": [wl["h"], wl["w"]],
            "parallelism": f"image-sharded x{n_gpus}",
            "model": "configs/glass_pretrain.yaml geometry: R=100 proposals, <=100 detections/image, every detection "
                     "recognised (26 steps x 97 classes)" if wl["full"] else "ResNet-50 + FPN (p2..p6)",
            "weights": "random init (glass_text_spotting_b200.weights.random_state_dict(0)), BatchNorm folded",
            "images": "uint8 synthetic, torch.Generator seed 1000 + rank",
            "l2": "inputs rotated between 2 batches; per-step working set (GBs of activations) >> 126 MB L2"}


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ============================================================================================ CPU (reference) arm
def _cpu_runner(wl):
    """Returns (callable running ONE image through the reference's CPU path, sample description)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    if wl.get("mask"):
        from oracle import d2_ops, mask as omask  # test infrastructure; allowed here only as the timed CPU baseline
        feats, rois = _mask_inputs(1, MASK_CPU_DETECTIONS)
        head = omask.seeded_mask_head(0)

        def run():
            with torch.no_grad():
                pooled = d2_ops.roi_pooler(feats, [rois[:, 1:]], (14, 14), [1 / 4, 1 / 8, 1 / 16, 1 / 32, 1 / 64], 0)
                omask.paste_masks_in_image(head(pooled)[:, 0], rois[:, 1:].contiguous(), (wl["h"], wl["w"]), 0.5)
        return run, (f"1 image, {MASK_CPU_DETECTIONS} detections through the oracle's mask branch (torch fp32 CPU: rotated pooler, "
                     f"mask head, reference paste at 1024x1024); the GPU arm does 89 per image")
    if wl.get("postprocess"):
        from oracle import postprocess as opp  # test infrastructure; allowed here only as the timed CPU baseline
        from glass_text_spotting_b200.text import TextDecoder
        boxes, scores, probs = _postprocess_inputs(wl["batch"])
        dec = TextDecoder()

        def run():
            for i in range(1):  # ONE image per call, like every other CPU leg
                ts = torch.tensor([w["score"] for w in dec.decode_probs(probs[i])], dtype=torch.float32)
                opp.post_process(boxes[i], scores[i], ts)
        return run, ("1 image x 100 detections through the oracle's post-processor (reference algorithm: torch fp32 + "
                     "cv2.minAreaRect + C rotated IoU/NMS)")
    if not wl["full"]:
        from oracle import nets  # test infrastructure; allowed here only as the timed CPU baseline
        net = nets.ResNetFPN().eval()
        x = torch.randint(0, 256, (1, 3, wl["h"], wl["w"]), generator=g).float() - 110.0

        def run():
            with torch.no_grad():
                net(x)
        return run, "1 image 1024x1024 through the oracle's ResNet-50+FPN (fp32, torch CPU)"
    from glass_text_spotting_b200 import weights
    from oracle import model as om
    o = om.GlassOracle(om.HotPathConfig())   # same geometry as the GPU arm: every detection (<= 100) is recognised
    o.load_state_dict(weights.random_state_dict(0), strict=False)
    # the GPU arm's rank-0 batch (seed 1000): this leg runs its images one per forward, like the reference does
    g = torch.Generator().manual_seed(1000)
    batch = torch.randint(0, 256, (wl["batch"], 3, wl["h"], wl["w"]), generator=g, dtype=torch.uint8)
    last = {"i": 0, "words": []}

    def run():
        img = batch[last["i"] % wl["batch"]].float()
        last["i"] += 1
        with torch.no_grad():
            r = o.inference([{"image": img}])
        last["words"].append(int(r[0]["instances"]["pred_boxes"].shape[0]))
    run.state = last
    return run, ("1 image 1024x1024 (of the GPU arm's rank-0 batch, same weights) through the oracle's full GLASS inference "
                 "(CPU restatement of the detectron2 path, fp32 torch CPU), every detection recognised like the GPU arm")


def _torch_backbone_on_gpu(B, H, W):
    """BASELINE.json configs[1] says "tcgen05 conv kernels vs torch.conv2d": the same ResNet-50+FPN in plain PyTorch
    (cuDNN) on this GPU, INFORMATIONAL only -- library kernels at their own precisions (strict fp32, TF32, bf16), none of
    which is the fp32-grade split arithmetic of the product path."""
    import torch
    from oracle import nets  # test infrastructure; a comparison leg like cpu_baseline, never the measured product
    res = {}
    try:
        net = nets.ResNetFPN().eval().cuda()
        x = torch.randn(B, 3, H, W, device="cuda")
        for name, tf32, dtype in (("fp32_strict", False, torch.float32), ("tf32", True, torch.float32),
                                  ("bf16_channels_last", True, torch.bfloat16)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            m = net.to(dtype)
            xi = x.to(dtype)
            if dtype == torch.bfloat16:
                m, xi = m.to(memory_format=torch.channels_last), xi.contiguous(memory_format=torch.channels_last)
            with torch.no_grad():
                for _ in range(2):
                    m(xi)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    m(xi)
                e1.record()
                torch.cuda.synchronize()
            res[name + "_images_per_s"] = B * 5 / (e0.elapsed_time(e1) / 1e3)
            net = net.float()
    except Exception as e:  # informational: never fail the bench line over it
        res["error"] = repr(e)[:200]
    finally:
        torch.backends.cudnn.allow_tf32 = True
    return res


def cpu_sample(wl, repeats: int):
    import torch
    run, desc = _cpu_runner(wl)
    run()  # warm-up
    t0 = time.perf_counter()
    for _ in range(repeats):
        run()
    dt = time.perf_counter() - t0
    return repeats / dt, torch.get_num_threads(), f"{repeats} x {desc}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    wl = WORKLOADS[args.workload]
    run, desc = _cpu_runner(wl)
    for _ in range(max(1, min(args.warmup, 1))):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    ips = args.steps / dt
    words = getattr(run, "state", {}).get("words")
    print(json.dumps({
        "impl": "reference", "metric": "images/sec @1024x1024", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": _config(args.workload, args.gpus),
        "run": {"words_per_image": (sum(words) / len(words)) if words else None,
                "note": "each timed step = ONE image of the configured batch (bounded sample; the reference runs one image "
                        "per forward anyway, SURVEY.md section 0 fact 4); value = images/s"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "each step = " + desc},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ============================================================================================ RoIAlign microbench
def _roialign_inputs(seed=0):
    """SURVEY.md 8d cfg 3: p2..p6 [1,256,{256..16}^2] N(0,1); 512 RoIs, centres U(0,1024)^2, w log-uniform
    [16,512], h = w*U(0.1,1), angle U(-180,180)."""
    import math
    import torch
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(1, 256, s, s, generator=g) for s in (256, 128, 64, 32, 16)]
    n = 512
    cx, cy = torch.rand(n, generator=g) * 1024, torch.rand(n, generator=g) * 1024
    w = torch.exp(torch.rand(n, generator=g) * (math.log(512) - math.log(16)) + math.log(16))
    h = w * (0.1 + 0.9 * torch.rand(n, generator=g))
    a = torch.rand(n, generator=g) * 360 - 180
    return feats, torch.stack((torch.zeros(n), cx, cy, w, h, a), 1).contiguous()


def _roialign_algorithmic_bytes(rois, sizes=(256, 128, 64, 32, 16), strides=(4, 8, 16, 32, 64), res=7, sr=2, c=256):
    """Sum over RoIs of (distinct feature cells touched by the 4 taps of the res*sr x res*sr samples) * C * 4 B
    + output + rois (SURVEY.md 8d cfg 3)."""
    import math
    import torch
    total = 0
    for r in rois.tolist():
        _, cx, cy, w, h, ang = r
        lvl = int(min(max(math.floor(4 + math.log2(math.sqrt(w * h) / 224 + 1e-8)), 2), 6)) - 2
        s, H = 1.0 / strides[lvl], sizes[lvl]
        th = ang * math.pi / 180
        cs, sn = math.cos(th), math.sin(th)
        g = res * sr
        ii = (torch.arange(g, dtype=torch.float64) + 0.5) / g
        yy = (-h * s / 2 + ii * h * s).view(-1, 1).expand(g, g)
        xx = (-w * s / 2 + ii * w * s).view(1, -1).expand(g, g)
        y = yy * cs - xx * sn + cy * s - 0.5
        x = yy * sn + xx * cs + cx * s - 0.5
        ok = (y >= -1) & (y <= H) & (x >= -1) & (x <= H)
        y0 = y.clamp(min=0).floor().clamp(max=H - 1).long()
        x0 = x.clamp(min=0).floor().clamp(max=H - 1).long()
        y1, x1 = (y0 + 1).clamp(max=H - 1), (x0 + 1).clamp(max=H - 1)
        cells = set()
        for yy_, xx_ in ((y0, x0), (y0, x1), (y1, x0), (y1, x1)):
            cells.update((yy_[ok] * H + xx_[ok]).tolist())
        total += len(cells) * c * 4
    return total + rois.shape[0] * c * res * res * 4 + rois.numel() * 4


def _postprocess_inputs(batch, m=100):
    import torch
    from glass_text_spotting_b200.synthetic import make_postprocess_case
    boxes, scores, probs = [], [], []
    for i in range(batch):
        b, s = make_postprocess_case(100 + i, 24, 40)
        assert len(b) >= m
        boxes.append(b[:m]); scores.append(s[:m])
        g = torch.Generator().manual_seed(200 + i)
        pr = torch.softmax(torch.randn(m, 26, 97, generator=g) * 12.0, dim=2)
        stops = torch.randint(1, 12, (m,), generator=g)
        for k in range(m):
            pr[k, stops[k]] = 0.002
            pr[k, stops[k], 1] = 0.808
        probs.append(pr)
    return torch.stack(boxes), torch.stack(scores), torch.stack(probs)


def _mask_inputs(n_img, per_img, seed=0):
    import math
    import torch
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(n_img, 256, s, s, generator=g) for s in (256, 128, 64, 32, 16)]
    n = n_img * per_img
    cx, cy = torch.rand(n, generator=g) * 1024, torch.rand(n, generator=g) * 1024
    w = torch.exp(torch.rand(n, generator=g) * (math.log(400) - math.log(24)) + math.log(24))
    h = w * (0.15 + 0.6 * torch.rand(n, generator=g))
    a = torch.rand(n, generator=g) * 360 - 180
    b = torch.arange(n_img).repeat_interleave(per_img).float()
    return feats, torch.stack((b, cx, cy, w, h, a), 1).contiguous()


def run_mask(args):
    """SURVEY.md 8f #3 measured like the hot path (device-resident pyramid; e2e adds the D2H of the pasted masks)."""
    import torch
    from glass_text_spotting_b200 import lib, ops, weights
    from glass_text_spotting_b200.modeling.mask_head import B200MaskHead
    wl = WORKLOADS["mask_bs4"]
    B, per = wl["batch"], 89
    feats, rois = _mask_inputs(B, per)
    acts = {f"p{i + 2}": ops.Act.from_nchw(f.cuda()) for i, f in enumerate(feats)}
    rois_d = rois.cuda()
    head = B200MaskHead(weights.random_mask_head_state_dict(0))
    L = lib.load()

    def step():
        m = head(acts, rois_d, cap=B * 100)
        return B200MaskHead.paste(m, rois_d[:, 1:].contiguous(), (wl["h"], wl["w"]), 0.5)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        out = step()
    torch.cuda.synchronize()
    launches0 = L.glass_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) / 1e3
    launches = L.glass_launch_count() - launches0
    host = torch.empty(out.shape, dtype=torch.bool).pin_memory()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        host.copy_(step(), non_blocking=True)
    f1.record()
    torch.cuda.synchronize()
    t_e2e = f0.elapsed_time(f1) / 1e3
    res = {
        "metric": "images/sec through the mask branch (89 detections each, masks pasted at 1024x1024)",
        "value": B * args.steps / t_dev, "unit": "images/s", "n_gpus": 1, "steps": args.steps, "warmup": warm,
        "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "fp16x3 split GEMMs, fp32 paste", "data": "synthetic",
        "config": {"workload": wl["desc"], "pixels_set_fraction": float(out.float().mean()),
                   "l2": "every step writes 373 MB of pasted masks (> 126 MB L2)"},
        "e2e": {"value": B * args.steps / t_e2e, "unit": "images/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": host.numel(), "note": "device-resident pyramid and boxes -> D2H of the bool masks"},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": None, "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                     "kernel": "paste_masks_rotated_kernel + conv_gemm_kernel (not profiled separately this round)"},
    }
    if not args.no_cpu:
        v, cores, sample = cpu_sample(wl, repeats=2)
        res["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(res))


def run_postprocess(args):
    """SURVEY.md 8f #1 measured like the hot path: device-resident inputs for `value`; `e2e` = detections and text
    probabilities from pinned host memory, survivors back to the host."""
    import torch
    from glass_text_spotting_b200 import lib, ops
    wl = WORKLOADS["postprocess_bs4"]
    B, m = wl["batch"], 100
    boxes, scores, probs = _postprocess_inputs(B, m)
    host = [t.contiguous().pin_memory() for t in (boxes, scores, probs)]
    dev = [t.cuda() for t in host]
    L = lib.load()

    def step(b, s, p):
        ts = ops.text_scores(p.view(B * m, 26, 97), 1).view(B, m)
        return ops.postprocess_merge(b, s, None, ts)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        r = step(*dev)
    torch.cuda.synchronize()
    kept = r["count"].cpu().tolist()
    iters = r["iters"].cpu().tolist()
    sampler = ClockSampler(0)
    if not args.no_clocks:
        sampler.start()
    launches0 = L.glass_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(*dev)
    e1.record()
    torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) / 1e3
    launches = L.glass_launch_count() - launches0
    clocks = sampler.stop() if not args.no_clocks else None
    dbuf = [torch.empty_like(t) for t in dev]
    out_host = None
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        for d, h in zip(dbuf, host):
            d.copy_(h, non_blocking=True)
        r = step(*dbuf)
        packed = torch.cat((r["boxes"].view(B, -1), r["index"].float(), r["count"].float().view(B, 1)), 1)
        if out_host is None:
            out_host = torch.empty(packed.shape, dtype=packed.dtype).pin_memory()
        out_host.copy_(packed, non_blocking=True)
    f1.record()
    torch.cuda.synchronize()
    t_e2e = f0.elapsed_time(f1) / 1e3
    # the merge kernel alone
    ts = ops.text_scores(dev[2].view(B * m, 26, 97), 1).view(B, m)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(args.steps):
        ops.postprocess_merge(dev[0], dev[1], None, ts)
    k1.record()
    torch.cuda.synchronize()
    merge_ms = k0.elapsed_time(k1) / args.steps
    out = {
        "metric": "images/sec post-processed (100 detections each)", "value": B * args.steps / t_dev, "unit": "images/s",
        "n_gpus": 1, "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 * t_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "survivors_per_image": kept, "merge_rounds_per_image": iters,
                   "l2": "working set is 4 MB of text probabilities + 10 KB of boxes per step: L2-resident by nature "
                         "(latency-bound, one CTA per image)"},
        "e2e": {"value": B * args.steps / t_e2e, "unit": "images/s",
                "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host),
                "d2h_bytes_per_step": out_host.numel() * out_host.element_size(),
                "note": "pinned host detections + text probabilities -> H2D -> text scores + merge loop -> D2H of survivors"},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": None, "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                     "kernel": "postprocess_merge_kernel (one CTA per image; latency-bound: sequential merge rounds x "
                               "greedy NMS, no roofline applies)", "kernel_ms_per_launch": merge_ms},
    }
    if not args.no_cpu:
        v, cores, sample = cpu_sample(wl, repeats=5)
        out["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(out))


def run_totaltext(args):
    """BASELINE.json configs[4]: the evaluation loop of tools/eval_glass.py / GlassRunner on synthetic images, image-sharded
    over the ranks: uint8 images -> device resize to 1200 (glass_finetune_totaltext.yaml MIN_SIZE_TEST) -> pad 1216 ->
    full GLASS inference (one CUDA graph, no host sync) -> word scores -> device post-processor -> 24 KB/image records kept
    in a device-side loop buffer -> ONE all-gather at the end of the loop (the reference's evaluator gathers
    once, glass/evaluation/text_evaluator.py:246-249).  Reports end-to-end images/s and the all-gather time (CUDA events)."""
    import torch
    import torch.distributed as dist
    from glass_text_spotting_b200 import lib, ops, weights
    from glass_text_spotting_b200.modeling.backbone import PIXEL_MEAN
    from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN
    from glass_text_spotting_b200.postprocess import B200PostProcessor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = WORKLOADS["totaltext_loop"]
    B, H, W = wl["batch"], wl["h"], wl["w"]
    scale = min(2.0, 1200 / max(H, W))                      # GlassRunner.get_inference_scale_ratio
    nh, nw = int(round(scale * H)), int(round(scale * W))   # 1200 x 1200
    ph, pw = (nh + 31) // 32 * 32, (nw + 31) // 32 * 32     # 1216 x 1216 (size_divisibility 32)
    L = lib.load()
    stream = torch.cuda.current_stream()
    model = B200GlassRCNN(weights.random_state_dict(0))
    post = B200PostProcessor()
    m = model.roi_heads.max_det
    steps_t, nc = model.roi_heads.steps, model.roi_heads.num_classes
    steps, warm = args.steps, max(args.warmup, 3)
    g = torch.Generator().manual_seed(2000 + rank)
    host = [torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(2)]
    dev = [h.cuda() for h in host]
    dbuf = [torch.empty_like(d) for d in dev]
    padded = torch.empty((B, 3, ph, pw), dtype=torch.float32, device="cuda")
    padded[:] = torch.tensor(PIXEL_MEAN, device="cuda").view(1, 3, 1, 1)   # mean padding normalises to exactly 0
    img_hw = torch.tensor([[nh, nw]] * B, dtype=torch.float32, device="cuda")
    slot = torch.arange(m, device="cuda", dtype=torch.int64).view(1, m)
    rec = torch.zeros((B, m, 8 + 2 * steps_t), dtype=torch.float32, device="cuda")
    use_graph = not args.no_graph

    def step(images_u8):
        """One loop iteration, no host synchronisation anywhere: -> rec [B, m, 8 + 2*steps] = (box 5, score, original
        index, valid) + per-step argmax and its probability of the words that survive the post-processor."""
        for i in range(B):
            padded[i, :, :nh, :nw] = ops.resize_bilinear_u8(images_u8[i], (nh, nw), flip_channels=False)
        if use_graph:
            model.graph_step(padded, img_hw)
            gs = model.last_graph
            det, probs, word_start = gs["det"], gs["probs"], gs["word_start"]
        else:
            _, det, probs, word_start = model.forward_packed(padded, img_hw)
        total = word_start[B:B + 1]
        ts, tidx, tmax = ops.text_scores(probs, 1, want_steps=True, n_dev=total)
        # word w of image i sits at row word_start[i] + w of the compacted word list
        widx = (word_start[:B].long().view(B, 1) + slot).clamp_(max=probs.shape[0] - 1)
        live = slot < det["count"].long().view(B, 1)
        ts_pad = torch.where(live, ts[widx], torch.zeros((), device="cuda"))
        boxes = det["pred_boxes"].clone()
        boxes[:, :, :4] *= 1.0 / scale                       # back to the original image (glass_runner.py:100-101)
        r = post.batch(boxes.contiguous(), det["scores"].contiguous(), det["count"], ts_pad.contiguous())
        kept = r["index"] >= 0
        sel = word_start[:B].long().view(B, 1) + r["index"].clamp(min=0).long()
        rec[:, :, 0:5], rec[:, :, 5], rec[:, :, 6] = r["boxes"], r["scores"], r["index"].float()
        rec[:, :, 7] = kept.float()
        rec[:, :, 8:8 + steps_t] = torch.where(kept.unsqueeze(-1), tidx[sel].float(), torch.zeros((), device="cuda"))
        rec[:, :, 8 + steps_t:] = torch.where(kept.unsqueeze(-1), tmax[sel], torch.zeros((), device="cuda"))
        return rec, r["count"], total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0 and not args.no_clocks:
        sampler.start()
    for i in range(warm):
        step(dev[i % 2])
    torch.cuda.synchronize()
    loop_buf = torch.empty((steps,) + tuple(rec.shape), dtype=rec.dtype, device="cuda")
    stats = torch.zeros((steps, 2), dtype=torch.float32, device="cuda")     # (survivors, words) per step
    gathered = torch.empty((world,) + tuple(loop_buf.shape), dtype=rec.dtype, device="cuda") if world > 1 else None
    if gathered is not None:
        dist.all_gather_into_tensor(gathered.view(world, -1), loop_buf.view(-1))   # NCCL warm-up
    ag = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    barrier()

    def loop(src, e2e):
        for i in range(steps):
            if e2e:
                dbuf[i % 2].copy_(host[i % 2], non_blocking=True)
            r, cnt, total = step(dbuf[i % 2] if e2e else src[i % 2])
            loop_buf[i].copy_(r, non_blocking=True)
            stats[i, 0] = cnt.float().mean()
            stats[i, 1] = total.float()[0]
        if gathered is not None:
            ag[2 if e2e else 0].record(stream)
            dist.all_gather_into_tensor(gathered.view(world, -1), loop_buf.view(-1))
            ag[3 if e2e else 1].record(stream)

    launches0 = L.glass_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    loop(dev, False)
    e1.record(stream)
    barrier()
    t_dev = e0.elapsed_time(e1) / 1e3
    launches = L.glass_launch_count() - launches0
    if use_graph:
        launches += model.last_graph["launches"] * steps
    clocks = sampler.stop() if rank == 0 else None
    kept_mean, words_per_step = stats.mean(0).tolist()
    # e2e: pinned host uint8 batch -> H2D -> loop body -> (end of loop) all-gather -> D2H of the records on rank 0
    out_host = torch.empty(tuple((gathered if gathered is not None else loop_buf).shape), dtype=rec.dtype).pin_memory()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    f0.record(stream)
    loop(None, True)
    if rank == 0:
        out_host.copy_(gathered if gathered is not None else loop_buf, non_blocking=True)
    f1.record(stream)
    barrier()
    t_e2e = f0.elapsed_time(f1) / 1e3
    ag_ms = ag[0].elapsed_time(ag[1]) if gathered is not None else 0.0
    ag_e2e_ms = ag[2].elapsed_time(ag[3]) if gathered is not None else 0.0
    times = torch.tensor([t_dev, t_e2e, ag_ms, ag_e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_dev, t_e2e, ag_ms, ag_e2e_ms = times.tolist()
    if rank == 0:
        cfg = _config("totaltext_loop", world)
        cfg.update(model_input=[ph, pw], scale_ratio=scale, config_file="configs/glass_finetune_totaltext.yaml geometry")
        print(json.dumps({
            "metric": "images/sec @1024x1024 (TotalText-shape eval loop)", "value": world * B * steps / t_dev,
            "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": 1e3 * t_dev / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16x3 split (22-bit operands, 3 tcgen05 MMAs per product, chunked fp32 RN accumulation)",
            "data": "synthetic", "config": cfg,
            "run": {"words_recognized_per_step": words_per_step, "words_after_postprocess_per_image": kept_mean,
                    "note": "random weights give word scores far below TEXT_THRESHOLD 0.25, so the (faithful) "
                            "post-processor drops nearly everything; its merge loop is exercised by the "
                            "postprocess_bs4 workload",
                    "step": "resize x4 -> one CUDA graph replay -> text scores -> device post-processor -> record; no "
                            "host synchronisation inside the loop" if use_graph else "eager launches",
                    "collective": "ONE NCCL all-gather of the loop's 24 KB/image records at the end of the loop, inside the "
                                  "timed region" if world > 1 else "none (N=1)",
                    "allgather_ms_per_loop": ag_ms, "allgather_ms_per_loop_e2e": ag_e2e_ms,
                    "allgather_ms_per_step": ag_ms / steps,
                    "allgather_bytes_per_rank": loop_buf.numel() * 4 if world > 1 else 0},
            "e2e": {"value": world * B * steps / t_e2e, "unit": "images/s", "h2d_bytes_per_step": B * H * W * 3,
                    "d2h_bytes_per_step": out_host.numel() * 4 / steps,
                    "note": "pinned host uint8 HWC batch -> H2D -> resize/pad -> hot path -> post-processor -> records; the "
                            "loop ends with the single all-gather and rank 0's D2H of all records, inside the timed region"},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None,
                         "traffic": None, "kernel": "conv_gemm_kernel (see the full_bs4 workload for its roofline line)"}}))
    if world > 1:
        dist.destroy_process_group()


def _roialign_glass_shapes(rois_d, flush, reps: int = 10):
    """SURVEY.md 8d cfg 3, second half: the two GLASS-specific poolers on the first 100 RoIs of the cfg-3 set --
    the recognizer pooler ([100,256,8,32], ADAPTIVE grid ceil(roi/bins), on the 256^2 P2P3 map; a10) and the image pooler
    ([100,3,128,128], sampling 2, normalisation fused, on the 1024^2 image; a11).  Single launches after an L2 flush.
    Bytes = output written (in the kernel's own storage format) + the RoIs' footprint on the input, clipped to it."""
    import torch
    from glass_text_spotting_b200 import ops
    from glass_text_spotting_b200.ops import Act
    k = rois_d.shape[0]
    g = torch.Generator().manual_seed(1)
    gmap = Act.from_nchw(torch.randn(1, 256, 256, 256, generator=g).cuda())
    image = torch.randint(0, 256, (1, 3, 1024, 1024), generator=g).float().cuda()
    fused = Act(k, 512, 8, 32)          # [local | global]: the pooler writes channels 256..511 (roi_heads.forward_recognizer)
    crops = Act(k, 3, 128, 128, 1, 8)

    def recog():
        ops.roi_align_rotated([gmap], rois_d, (8, 32), [0.25], 0, out_f32=False,
                              out_split=(fused.buf, fused.hp, fused.wp, fused.border, 256, fused.cp))

    img4 = torch.empty((1, 1024, 1024, 4), dtype=torch.float32, device="cuda")

    def img():
        ops.image_roi_align_rotated(image, (1024, 1024), (103.530, 116.280, 123.675), (1.0, 1.0, 1.0), rois_d, (128, 128), 2,
                                    out_act=crops, workspace=img4)

    r = rois_d.cpu()
    area = (r[:, 3].clamp(max=1024) * r[:, 4].clamp(max=1024)).sum().item()      # px^2 on the image
    out = {}
    for name, fn, nbytes in (
            ("recognizer_pooler_100x256x8x32_adaptive", recog, k * 256 * 8 * 32 * 4 + area / 16 * 256 * 4),
            ("image_pooler_100x3x128x128", img, k * 128 * 128 * 8 * 4 + area * 3 * 4)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        out[name] = {"ms": ms, "bytes": int(nbytes), "gbs": nbytes / (ms / 1e3) / 1e9}
    return out


def _ncu(name):
    """Numbers that only a profiler can give (DRAM bytes, tensor-pipe activity) come from the committed ncu passes:
    profiles/ncu_metrics.json (written by tools/gpu/ncu_metrics.sh + tools/ncu_to_json.py; says which round)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_metrics.json"))).get(name) or {}
    except Exception:
        return {}


def measure_roialign(steps, no_clocks, min_seconds=1.2, extras=True):
    """BASELINE.json configs[2].  Timed region = back-to-back launches (CUDA events on the launching stream), each
    on the NEXT of 4 copies of the feature pyramid (4 x 91 MB of split-fp16 maps + 4 x 26 MB of outputs >> 126 MB L2), so
    every launch finds its inputs in HBM, not in L2; the replay loop lasts >= ``min_seconds`` so that the 100 ms clock
    sampler sees >= 5 samples under load.  The single-launch, L2-flushed time is reported beside it."""
    import torch
    from glass_text_spotting_b200 import ops
    from glass_text_spotting_b200.ops import Act
    feats, rois = _roialign_inputs()
    ncopy = 4
    acts = [[Act.from_nchw(f.cuda()) for f in feats] for _ in range(ncopy)]
    rois_d = rois.cuda()
    outs = [torch.empty((2, 512, 49 * 256), dtype=torch.float16, device="cuda") for _ in range(ncopy)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    scales = [1 / 4, 1 / 8, 1 / 16, 1 / 32, 1 / 64]

    def step(i):
        ops.roi_align_rotated(acts[i % ncopy], rois_d, (7, 7), scales, 2, out_f32=False,
                              out_split=(outs[i % ncopy], 7, 7, 0, 0, 256))

    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    # (1) single launch, L2 flushed by a 256 MB write before it
    ts = []
    for i in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(i)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms_flushed = sum(ts) / len(ts)
    # (2) back-to-back launches over rotating inputs, replayed from a CUDA graph of 2 rounds over the copies (the
    # Python/ctypes call costs as much host time as the kernel runs, so eager launches would time the host)
    per_graph = 2 * ncopy
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for i in range(per_graph):
                step(i)
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(20):
        graph.replay()
    c1.record()
    torch.cuda.synchronize()
    est = c0.elapsed_time(c1) / 20 / 1e3                      # seconds per replay
    replays = max((steps + per_graph - 1) // per_graph, int(min_seconds / max(est, 1e-6)) + 1)
    sampler = ClockSampler(torch.cuda.current_device())
    if not no_clocks:
        sampler.start()
        time.sleep(0.15)
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(replays):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    launches = replays * per_graph
    ms = e0.elapsed_time(e1) / launches
    clocks = sampler.stop() if not no_clocks else None
    nbytes = _roialign_algorithmic_bytes(rois)
    peaks, src = _peaks()
    gbs = nbytes / (ms / 1e3) / 1e9
    dram = _ncu("roialign_512").get("dram_bytes_per_launch")
    res = {"ms_per_launch": ms, "launches": launches, "timed_s": ms * launches / 1e3, "algorithmic_bytes": nbytes,
           "gbs_algorithmic": gbs, "frac_algorithmic": gbs / peaks["hbm_gbs"],
           "dram_bytes_per_launch_ncu": dram,
           "gbs_dram": (dram / (ms / 1e3) / 1e9) if dram else None,
           "frac_dram": (dram / (ms / 1e3) / 1e9 / peaks["hbm_gbs"]) if dram else None,
           "peak_gbs": peaks["hbm_gbs"], "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({src})",
           "single_launch_l2_flushed_ms": ms_flushed, "single_launch_l2_flushed_gbs": nbytes / (ms_flushed / 1e3) / 1e9,
           "rois_per_s": 512 / (ms / 1e3), "clocks": clocks,
           "kernel": "roi_align_rotated_split8_kernel<2>",
           "l2": f"inputs larger than L2: launches rotate over {ncopy} copies of the pyramid and output "
                 f"({ncopy} x 117 MB > 126 MB L2)"}
    if extras:
        res["glass_shapes"] = _roialign_glass_shapes(rois_d[:100].contiguous(), flush)
    return res


def run_roialign(args):
    wl = WORKLOADS["roialign_512"]
    r = measure_roialign(args.steps, args.no_clocks)
    cfg = {"workload": wl["desc"], "name": "roialign_512", "l2": r["l2"], "rois_per_s": r["rois_per_s"],
           "single_launch_l2_flushed_ms": r["single_launch_l2_flushed_ms"],
           "single_launch_l2_flushed_gbs": r["single_launch_l2_flushed_gbs"], "glass_shapes": r.get("glass_shapes"),
           "timed_s": r["timed_s"]}
    print(json.dumps({
        "metric": "RotatedROIAlign GB/s (algorithmic bytes)", "value": r["gbs_algorithmic"], "unit": "GB/s", "n_gpus": 1,
        "steps": r["launches"], "warmup": 4, "ms_per_step": r["ms_per_launch"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (split-fp16 storage)", "data": "synthetic", "config": cfg,
        "gpu_launches": r["launches"], "clocks": r["clocks"],
        "roofline": {"bound": "hbm", "achieved": r["gbs_algorithmic"], "peak": r["peak_gbs"], "unit": "GB/s",
                     "frac": r["frac_algorithmic"], "traffic": r["dram_bytes_per_launch_ncu"],
                     "achieved_dram_gbs": r["gbs_dram"], "frac_dram": r["frac_dram"], "kernel": r["kernel"],
                     "algorithmic_bytes": r["algorithmic_bytes"], "peak_source": r["peak_source"]}}))


def measure_backbone(min_seconds=1.2, no_clocks=False, mode=0, graph=True):
    """BASELINE.json configs[1]: ResNet-50 + FPN, bs = 8, 1024x1024, inputs resident in HBM, rotating between 2 batches;
    the timed loop lasts >= ``min_seconds`` with the clock sampler running."""
    import torch
    from glass_text_spotting_b200 import ops, weights
    from glass_text_spotting_b200.modeling.backbone import B200ResNetFPN
    B, H, W = 8, 1024, 1024
    g = torch.Generator().manual_seed(1000)
    dev = [torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8).cuda().float() for _ in range(2)]
    model = B200ResNetFPN(weights.random_backbone_state_dict(0), mode=mode)
    for i in range(3):
        model(dev[i % 2])
    torch.cuda.synchronize()
    eager = model
    if graph:
        # one CUDA graph per input batch (the forward has no host work between its ~75 launches; eager, the ctypes launch
        # path makes the small layers host-bound: 8.7-9.1 ms instead of 7.9)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        graphs = []
        for d in dev:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                model(d)
            graphs.append(gr)
        torch.cuda.synchronize()

        class _Replay:
            def __call__(self, x):
                graphs[0 if x is dev[0] else 1].replay()
        model = _Replay()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for i in range(4):
        model(dev[i % 2])
    c1.record()
    torch.cuda.synchronize()
    steps = int(min_seconds / (c0.elapsed_time(c1) / 4 / 1e3)) + 1
    sampler = ClockSampler(torch.cuda.current_device())
    if not no_clocks:
        sampler.start()
        time.sleep(0.15)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        model(dev[i % 2])
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 1e3
    clocks = sampler.stop() if not no_clocks else None
    ops.PROFILE = []
    for i in range(2):
        eager(dev[i % 2])
    torch.cuda.synchronize()
    gemm_ms = sum(r[0].elapsed_time(r[1]) for r in ops.PROFILE) / 2
    gemm_flops = sum(ops.profile_flops(r) for r in ops.PROFILE) / 2
    ops.PROFILE = None
    peaks, src = _peaks()
    peak = peaks["bf16_tflops_sustained"]
    ach = gemm_flops / (gemm_ms / 1e3) / 1e12
    mul = 1 if mode else 3
    nc = _ncu("backbone_bs8")
    del model, eager
    return {"images_s": B * steps / t, "ms_per_step": 1e3 * t / steps, "steps": steps, "timed_s": t,
            "step": "one CUDA graph replay per batch" if graph else "eager launches",
            "conv_gemm_ms_per_step": gemm_ms, "algorithmic_tflops": ach, "algorithmic_frac": ach / peak,
            "issued_tflops": ach * mul, "issued_frac": ach * mul / peak,
            "tensor_pipe_active_pct": nc.get("time_weighted_tensor_pipe_active_pct"),
            "tensor_pipe_active_pct_mma_bound_layers": nc.get("tensor_pipe_active_pct_mma_bound_layers"),
            "tensor_pipe_source": nc.get("source"), "peak_tflops": peak,
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({src})", "clocks": clocks}


# ============================================================================================ B200 arm
def run_b200(args):
    """full_bs4 / backbone_bs8.  One step = one pass of the hot path over one batch of 4 (8) images.

    full_bs4: the step is ONE CUDA graph (B200GlassRCNN.graph_step: no host sync, no per-kernel host work); its result --
    the fixed-size packed detection records of the batch -- is appended to a device-side loop buffer, and the loop ends
    with the SINGLE all-gather of SURVEY.md 8e / the north star ("a single NCCL all-gather of detections at the end",
    what the reference's evaluator does with comm.gather in text_evaluator.py:246-249).  The gather (and, for e2e, rank 0's
    read-back of the gathered records) is INSIDE the timed region."""
    import torch
    import torch.distributed as dist
    from glass_text_spotting_b200 import lib, ops, weights
    from glass_text_spotting_b200.modeling.backbone import B200ResNetFPN
    from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = WORKLOADS[args.workload]
    B, H, W = wl["batch"], wl["h"], wl["w"]
    full = wl["full"]
    L = lib.load()
    mode = ops.MODE_FAST if args.fast else ops.MODE_SPLIT
    stream = torch.cuda.current_stream()
    steps, warm = args.steps, max(args.warmup, 3)
    use_graph = full and not args.no_graph

    g = torch.Generator().manual_seed(1000 + rank)
    # host images are uint8 like the reference's dataset mapper output; they become fp32 on the device
    host = [torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8).pin_memory() for _ in range(2)]
    dev = [h.cuda().float() for h in host]
    dbuf = [torch.empty((B, 3, H, W), dtype=torch.uint8, device="cuda") for _ in range(2)]
    fbuf = torch.empty((B, 3, H, W), dtype=torch.float32, device="cuda")
    img_hw = torch.tensor([[H, W]] * B, dtype=torch.float32, device="cuda")

    if full:
        model = B200GlassRCNN(weights.random_state_dict(0), mode=mode)

        def step(images):
            """-> the packed per-image detection records [B, 100, 10 + 26*97] of this batch (a persistent device buffer)"""
            if use_graph:
                return model.graph_step(images, img_hw)
            return model.forward_packed(images, img_hw)[0]
    else:
        model = B200ResNetFPN(weights.random_backbone_state_dict(0), mode=mode)

        def step(images):
            out = model(images)
            return torch.cat([out[k].hi.float().abs().mean().view(1) for k in ["p2", "p3", "p4", "p5", "p6"]] +
                             [out["p6"].buf.float().view(-1)])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (allocates every workspace buffer, captures the graph); the clock sampler runs from here to the end of
    # the timed region so that it collects enough 100 ms samples under load
    sampler = ClockSampler(local_rank)
    if rank == 0 and not args.no_clocks:
        sampler.start()
    for i in range(warm):
        res = step(dev[i % 2])
    torch.cuda.synchronize()
    # loop buffer: every step's result stays on the device until the single end-of-loop gather
    loop_buf = torch.empty((steps,) + tuple(res.shape), dtype=res.dtype, device="cuda")
    gathered = torch.empty((world,) + tuple(loop_buf.shape), dtype=res.dtype, device="cuda") if (full and world > 1) else None
    ag = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    if gathered is not None:   # NCCL warm-up (communicator set-up is not part of the loop)
        dist.all_gather_into_tensor(gathered.view(world, -1), loop_buf.view(-1))
    barrier()

    # ---- timed: inputs resident in HBM
    launches0 = L.glass_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(steps):
        loop_buf[i].copy_(step(dev[i % 2]), non_blocking=True)
    if gathered is not None:
        ag[0].record(stream)
        dist.all_gather_into_tensor(gathered.view(world, -1), loop_buf.view(-1))
        ag[1].record(stream)
    e1.record(stream)
    barrier()
    t_dev = e0.elapsed_time(e1) / 1e3
    launches = L.glass_launch_count() - launches0
    if use_graph:
        launches = model.last_graph["launches"] * steps     # kernels inside the replayed graph
    clocks = sampler.stop() if rank == 0 else None
    ag_ms = ag[0].elapsed_time(ag[1]) if gathered is not None else 0.0
    words_per_step = float(loop_buf[..., 0].sum().item()) / steps if full else None

    # ---- e2e: host buffers; H2D of every step's batch and D2H of its result are inside the timed region.  The H2D of
    # step i+1 and the D2H of step i run on a copy stream while the next step computes (double-buffered staging, events
    # both ways), the way a serving loop feeds the model; every step still waits for ITS OWN batch and ships ITS OWN
    # result.  The loop ends with the single all-gather and rank 0's read-back of the other ranks' records.
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]     # H2D of buffer b finished
    consumed = [torch.cuda.Event(), torch.cuda.Event()]  # the step reading buffer b has finished with it
    done = [torch.cuda.Event() for _ in range(steps)]    # step i's record is in the loop buffer
    res_host = torch.empty(tuple(loop_buf.shape), dtype=res.dtype).pin_memory()
    gathered_host = torch.empty(tuple(gathered.shape), dtype=res.dtype).pin_memory() if (gathered is not None and rank == 0) \
        else None

    def upload(b):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])
            dbuf[b].copy_(host[b], non_blocking=True)
            ready[b].record(copy_stream)

    def e2e_loop(n):
        upload(0)
        for i in range(n):
            b = i % 2
            if i + 1 < n:
                upload((i + 1) % 2)           # next step's batch, overlapped with this step's compute
            stream.wait_event(ready[b])
            if full and use_graph:
                # uint8 -> fp32 happens in the copy into the graph's static input (one pass over the batch)
                loop_buf[i].copy_(step(dbuf[b]), non_blocking=True)
                consumed[b].record(stream)
            else:
                fbuf.copy_(dbuf[b])           # uint8 -> fp32 on the device
                consumed[b].record(stream)
                loop_buf[i].copy_(step(fbuf), non_blocking=True)
            done[i].record(stream)
            with torch.cuda.stream(copy_stream):   # this step's own result to the host, off the compute stream
                copy_stream.wait_event(done[i])
                res_host[i].copy_(loop_buf[i], non_blocking=True)
        if gathered is not None:
            ag[2].record(stream)
            dist.all_gather_into_tensor(gathered.view(world, -1), loop_buf.view(-1))
            ag[3].record(stream)
            if rank == 0:                      # the gathered records are consumed on rank 0 (comm.gather(dst=0))
                gathered_host.copy_(gathered, non_blocking=True)
        stream.wait_stream(copy_stream)

    for b in range(2):
        consumed[b].record(stream)
    e2e_loop(min(2, steps))                    # untimed warm-up of the e2e loop
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    f0.record(stream)
    copy_stream.wait_event(f0)                 # no upload starts before the timed region does
    e2e_loop(steps)
    f1.record(stream)
    barrier()
    t_e2e = f0.elapsed_time(f1) / 1e3
    d2h_bytes = res_host[0].numel() * res_host.element_size()
    if gathered_host is not None:
        d2h_bytes += gathered_host.numel() * gathered_host.element_size() / steps
    ag_e2e_ms = ag[2].elapsed_time(ag[3]) if gathered is not None else 0.0

    # ---- per-kernel profile pass: CUDA events around every launch of the dominant kernel (conv GEMM), eager launches
    ops.PROFILE = []
    nprof = 2
    for i in range(nprof):
        if full:
            model.forward_packed(dev[i % 2], img_hw)
        else:
            step(dev[i % 2])
    torch.cuda.synchronize()
    gemm_ms = sum(r[0].elapsed_time(r[1]) for r in ops.PROFILE) / nprof
    gemm_flops = sum(ops.profile_flops(r) for r in ops.PROFILE) / nprof
    n_gemm = len(ops.PROFILE) // nprof
    ops.PROFILE = None

    times = torch.tensor([t_dev, t_e2e, ag_ms, ag_e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_dev, t_e2e, ag_ms, ag_e2e_ms = times.tolist()

    if rank == 0:
        peaks, peak_src = _peaks()
        peak = peaks["bf16_tflops_sustained"]
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12
        cfg = _config(args.workload, world)
        run = {"kb_per_chunk": ops.KB_PER_CHUNK or "2 (bottom-up ResNet body), 4 (FPN, RPN, box head, P2P3), 6 (per-word recognizer convs)",
               "step": "one CUDA graph replay (B200GlassRCNN.graph_step), no host synchronisation inside the loop"
                       if use_graph else "eager launches"}
        if full:
            # NOTE the GEMM's M space follows the live word count on the device: the profile pass' FLOPs are the
            # algorithmic FLOPs of the words actually detected, not of the capacity
            run["words_per_step"] = words_per_step
            run["collective"] = ("ONE NCCL all-gather of the loop's packed detection records at the end of the loop, inside "
                                 "the timed region") if world > 1 else "none (N=1)"
            run["allgather_ms_per_loop"] = ag_ms
            run["allgather_ms_per_loop_e2e"] = ag_e2e_ms
            run["allgather_bytes_per_rank"] = loop_buf.numel() * 4 if world > 1 else 0
        nc = _ncu(args.workload)
        out = {
            "metric": "images/sec @1024x1024", "value": world * B * steps / t_dev, "unit": "images/s",
            "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": 1e3 * t_dev / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "fp16 (single tcgen05 pass)" if args.fast else
                     "fp16x3 split (22-bit operands, 3 tcgen05 MMAs per product, chunked fp32 RN accumulation)",
            "data": "synthetic", "config": cfg, "run": run,
            "e2e": {"value": world * B * steps / t_e2e, "unit": "images/s",
                    "h2d_bytes_per_step": B * 3 * H * W, "d2h_bytes_per_step": d2h_bytes,
                    "note": "pinned host uint8 batch -> H2D (copy stream, overlapped with the previous step) -> hot path -> "
                            "D2H of the step's result on the copy stream "
                            + ("(packed detection records); at N > 1 the loop ends with the single all-gather and rank 0's "
                               "read-back of the gathered records, both inside the timed region"
                               if full else "(p6 + per-level checksums)")},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": nc.get("dram_bytes_per_step"),
                         "traffic_note": "dram__bytes_read+write summed over the step's conv_gemm launches (ncu, "
                                         "profiles/ncu_metrics.json); null when no capture exists for this workload",
                         "kernel": "conv_gemm_kernel (tcgen05 implicit GEMM), all launches of one step",
                         "launches_per_step": n_gemm, "kernel_ms_per_step": gemm_ms,
                         "kernel_share_of_step": gemm_ms / (1e3 * t_dev / steps),
                         "algorithmic_flops_per_step": gemm_flops,
                         "issued_mma_flops_per_step": gemm_flops * (1 if args.fast else 3),
                         "issued_frac": achieved * (1 if args.fast else 3) / peak,
                         "tensor_pipe_active_pct_ncu": nc.get("time_weighted_tensor_pipe_active_pct"),
                         # SURVEY.md 8d cfg 4: the WHOLE step against t_min = algorithmic FLOPs / tensor peak (the gather
                         # term bytes / HBM peak is < 0.1 ms and is left out)
                         "whole_step_frac_of_tensor_roofline": (gemm_flops / (peak * 1e12)) / (t_dev / steps),
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peak_src})"},
        }
        if world == 1 and full and not args.no_submetrics:
            # BASELINE.json's metric names three quantities; the other two are measured here, in the same process on the
            # same GPU right after the timed region, each in a >= 1.2 s loop with its own clock record
            del model
            torch.cuda.empty_cache()
            out["submetrics"] = {
                "backbone": dict(measure_backbone(no_clocks=args.no_clocks, mode=mode),
                                 workload=WORKLOADS["backbone_bs8"]["desc"]),
                "roialign": dict(measure_roialign(64, args.no_clocks, extras=False),
                                 workload=WORKLOADS["roialign_512"]["desc"])}
        if world == 1 and not args.no_cpu:
            v, cores, sample = cpu_sample(wl, repeats=2 if full else 3)
            out["cpu_baseline"] = {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample}
            if not full:
                out["config"]["torch_cudnn_informational"] = _torch_backbone_on_gpu(B, H, W)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="full_bs4", choices=sorted(WORKLOADS))
    ap.add_argument("--fast", action="store_true", help="single-pass fp16 (NOT the parity precision)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="full_bs4: eager launches instead of the CUDA-graph step")
    ap.add_argument("--no-submetrics", action="store_true", help="skip the backbone / RoIAlign sub-benches of full_bs4")
    ap.add_argument("--no-clocks", action="store_true", help="do not spawn the nvidia-smi clock sampler (ncu runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif WORKLOADS[args.workload].get("roialign"):
        run_roialign(args)
    elif WORKLOADS[args.workload].get("postprocess"):
        run_postprocess(args)
    elif WORKLOADS[args.workload].get("totaltext"):
        run_totaltext(args)
    elif WORKLOADS[args.workload].get("mask"):
        run_mask(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
