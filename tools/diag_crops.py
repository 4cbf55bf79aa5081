import os, sys, math, ctypes
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from glass_text_spotting_b200 import ops
from glass_text_spotting_b200.modeling.glass_rcnn import B200GlassRCNN
from oracle import model as om
seed = int(sys.argv[1])
H = W = 1024
img = om.synthetic_image(seed, H, W)
o = om.build_oracle(seed=seed, calib_images=[img])
taps = {}
with torch.no_grad():
    want = o.inference([{"image": img}], taps=taps, do_postprocess=False)[0]["instances"]
t = taps["per_image"][0]
model = B200GlassRCNN(o.state_dict())
det = t["det_boxes"]; k = det.shape[0]
rois = torch.cat((torch.zeros(k, 1), det), 1).contiguous().cuda()
mean = o.cfg.pixel_mean; std = o.cfg.pixel_std
act = ops.Act(k, 3, 128, 128, 1, 8)
ops.image_roi_align_rotated(img[None].cuda().contiguous(), (H, W), mean, std, rois, (128, 128), 2, out_act=act)
got = act.to_nchw().cpu(); ref = t["local_crops"]
err = (got - ref).abs(); bad = err > 1e-4 + 1e-3 * ref.abs()
libm = ctypes.CDLL("libm.so.6"); libm.cosf.restype = ctypes.c_float; libm.cosf.argtypes = [ctypes.c_float]; libm.sinf.restype = ctypes.c_float; libm.sinf.argtypes = [ctypes.c_float]
print("misses", int(bad.sum()), "max err", err.max().item())
for w in range(k):
    nb = int(bad[w].sum())
    a = np.float32(det[w, 4].item())
    th = np.float32(a * np.float32(math.pi) / np.float32(180.0))
    dc = np.float32(libm.cosf(float(th))) != np.float32(np.cos(np.float64(th)))
    ds = np.float32(libm.sinf(float(th))) != np.float32(np.sin(np.float64(th)))
    if nb or dc or ds:
        print("word", w, "misses", nb, "max err %.2e" % err[w].max().item(), "box", [round(v, 3) for v in det[w].tolist()], "cosf differs", bool(dc), "sinf differs", bool(ds))
