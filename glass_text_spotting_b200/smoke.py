"""One small invocation of the WHOLE hot path on cuda:0 (backbone -> rotated RPN -> box branch -> recognizer, i.e. every
kernel the model launches), checked against the oracle (test infrastructure: only this smoke check, tests/ and bench.py's
CPU legs may import ``oracle``).  Mirrors tests/test_gpu_e2e.py on one 224 x 288 image with at most 6 detections."""
import torch


def _tol(got, ref, name, rtol=1e-3, atol=1e-4):
    got, ref = got.detach().cpu().float(), ref.detach().cpu().float()
    err = (got - ref).abs()
    bad = int((err > atol + rtol * ref.abs()).sum())
    assert bad == 0, f"smoke: {name}: {bad}/{err.numel()} out of rtol {rtol:g} / atol {atol:g}, max err {err.max().item():.3e}"
    return err.max().item()


def run() -> None:
    from . import lib, ops
    from .modeling.glass_rcnn import B200GlassRCNN
    from oracle import model as om
    L = lib.load()
    torch.cuda.set_device(0)
    K = 6
    cfg = om.HotPathConfig(max_detections_override=K)
    img = om.synthetic_image(5, 224, 288)
    o = om.build_oracle(seed=0, calib_images=[img], cfg=cfg)
    taps = {}
    with torch.no_grad():
        want = o.inference([{"image": img, "height": 448, "width": 576}], taps=taps)[0]["instances"]
    t = taps["per_image"][0]
    model = B200GlassRCNN(o.state_dict(), detections_per_image=K)
    launches0 = L.glass_launch_count()
    mt = {}
    got = model.inference([{"image": img, "height": 448, "width": 576}], taps=mt)[0]["instances"]
    torch.cuda.synchronize()
    launches = L.glass_launch_count() - launches0

    # dense stages, free-running from the image: the shallow taps at the literal tolerance, the deep ones (a random-weight
    # ResNet amplifies rounding noise ~4x per stage) by relative L2
    worst = _tol(mt["features"]["res2"].to_nchw(), t["res2"], "res2")
    for k in ("p2", "p3", "p4", "p5", "p6"):
        rel = ((mt["features"][k].to_nchw().cpu() - t[k]).norm() / t[k].norm()).item()
        assert rel < 1e-3, f"smoke: {k} relL2 {rel:.2e}"
    # detections: same count, same boxes / scores in the same order
    assert len(got) == len(want["pred_boxes"]) == K, (len(got), len(want["pred_boxes"]))
    _tol(got.pred_boxes.tensor, want["pred_boxes"], "pred_boxes", rtol=5e-3, atol=5e-3)
    _tol(got.scores, want["scores"], "scores", rtol=5e-3, atol=1e-4)
    rows = got.pred_text_prob.sum(-1).cpu()
    assert ((rows - 1).abs() < 1e-4).logical_or(rows == 0).all(), "smoke: text rows are softmax rows or zero"
    # recognizer teacher-forced with the oracle's pyramid and detections: per-character probabilities at the literal tolerance
    det = t["det_boxes"]
    rois = torch.cat((torch.zeros(len(det), 1), det), 1).cuda().contiguous()
    ws = torch.tensor([0, len(det)], dtype=torch.int32).cuda()
    feats = {k: ops.Act.from_nchw(t[k].cuda()) for k in ("p2", "p3")}
    il = model.preprocess_image([{"image": img}])
    probs = model.roi_heads.forward_recognizer(il.tensor, tuple(il.tensor.shape[-2:]), feats, rois, ws, 1)
    torch.cuda.synchronize()
    perr = _tol(probs, t["pred_text_prob"], "pred_text_prob | oracle pyramid + detections", atol=1e-5)
    print(f"smoke ok: full model on a 224x288 image, {K} detections == oracle; res2 max err {worst:.2e}, p2..p6 relL2 < 1e-3, "
          f"text probabilities max err {perr:.2e}; {launches} kernel launches in model.inference()")
