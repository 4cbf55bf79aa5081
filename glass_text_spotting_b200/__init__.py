"""glass_text_spotting_b200 -- B200-native (sm_100a) implementation of GLASS's per-image dense
forward path behind the detectron2 GeneralizedRCNN / ROIHeads plugin surface.

The compute lives in ``_lib/libglass_b200.so`` (hand-written CUDA, C ABI declared in
``include/glass_b200.h``); this package is the thin Python host: a ctypes binding (``lib``),
tensor-level operator wrappers (``ops``), weight pre-packing (``packing``) and the module mirror of
the reference's plugin interface (``modeling``).  There is no CPU fallback: importing ``lib`` and
calling an op without the built extension or without a GPU raises.
"""
__version__ = "0.1.0"
