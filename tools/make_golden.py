"""Generate tests/golden/*.pt by running the REFERENCE's own pure-torch modules.

Runs only in the authoring container (needs /root/reference).  detectron2 / fvcore are
not installable offline, so the four reference files that are pure torch are imported
under a ~30-line stub of ``detectron2.utils.registry.Registry``,
``detectron2.config.configurable`` and ``fvcore.nn.weight_init``.
The reference modules get the *oracle's* seeded weights (same parameter names), a
seeded input, and their outputs are stored (strided sub-samples + full-tensor sums so
the fixtures stay small).  tests/test_oracle_golden.py replays the same seeds through
oracle/nets.py and compares.

    python tools/make_golden.py
"""
import importlib.util
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def install_stubs():
    class Registry(dict):
        def __init__(self, name):
            super().__init__()
            self._name = name

        def register(self, obj=None):
            def deco(o):
                self[o.__name__] = o
                return o
            return deco(obj) if obj is not None else deco

        def get(self, name):
            return self[name]

    def configurable(fn=None, **kw):
        return fn  # classes are constructed with explicit kwargs below

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("detectron2")
    mod("detectron2.utils")
    mod("detectron2.utils.registry", Registry=Registry)
    mod("detectron2.config", configurable=configurable)
    mod("fvcore")
    mod("fvcore.nn")
    wi = mod("fvcore.nn.weight_init", c2_msra_fill=lambda m: None, c2_xavier_fill=lambda m: None)
    sys.modules["fvcore.nn"].weight_init = wi


def load_ref(relpath, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def seeded_state(module, seed, bn_stats=True):
    """Scale-preserving seeded weights for a golden case (identical code in the test)."""
    from tests.golden_common import seeded_fill
    seeded_fill(module, seed)


def summarize(t, stride):
    sl = t[tuple(slice(None, None, s) for s in stride)].contiguous().clone()
    return {"shape": list(t.shape), "stride": list(stride), "sample": sl,
            "sum": t.double().sum().item(), "abssum": t.double().abs().sum().item()}


def main():
    install_stubs()
    from tests import golden_common as gc
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)

    lfe = load_ref("glass/modeling/fusion/local_feature_extraction.py", "ref_lfe")
    fus = load_ref("glass/modeling/fusion/fusion_modules.py", "ref_fus")
    enc = load_ref("glass/modeling/recognition/recognizer_encoder.py", "ref_enc")
    ast = load_ref("glass/modeling/recognition/prediction_aster.py", "ref_ast")
    ast.device = torch.device("cpu")

    with torch.no_grad():
        # a12 ResNetFeatureExtractor
        m = lfe.ResNetFeatureExtractor(3, 256).eval()
        gc.seeded_fill(m, gc.SEEDS["hybrid"])
        x = gc.seeded_input("hybrid")
        torch.save(summarize(m(x), (1, 8, 1, 4)), os.path.join(OUT, "hybrid_net.pt"))

        # a13 MultiAspectGCAttention
        m = fus.MultiAspectGCAttention(inplanes=512, ratio=0.5, headers=8, outplane=256,
                                       fusion_type="channel_add").eval()
        gc.seeded_fill(m, gc.SEEDS["fusion"])
        x = gc.seeded_input("fusion")
        torch.save(summarize(m(x), (1, 8, 1, 4)), os.path.join(OUT, "fusion_net.pt"))

        # a9 P2P3Fusion
        m = fus.P2P3Fusion(256).eval()
        gc.seeded_fill(m, gc.SEEDS["p2p3"])
        p2, p3 = gc.seeded_input("p2p3")
        torch.save(summarize(m(p2, p3), (1, 4, 2, 2)), os.path.join(OUT, "p2p3.pt"))

        # a15 BiLSTMBlockV2
        m = enc.BiLSTMBlockV2(input_size=256, hidden_size=256, output_size=256, num_of_layers=2).eval()
        gc.seeded_fill(m, gc.SEEDS["encoder"])
        x = gc.seeded_input("encoder")
        torch.save(summarize(m(x), (1, 1, 4)), os.path.join(OUT, "encoder.pt"))

        # a16 AttentionRecognitionHead.sample (two cases: normal, and one tuned to trigger the early break)
        for case in ["decoder", "decoder_break"]:
            m = ast.AttentionRecognitionHead(num_classes=97, in_planes=256, sDim=256, attDim=256,
                                             max_len_labels=26).eval()
            gc.seeded_fill(m, gc.SEEDS[case])
            if case == "decoder_break":
                gc.force_eos_bias(m)
            x = gc.seeded_input(case)
            probs, alphas = m.sample(x, None, 26, 0)
            d = summarize(probs, (1, 1, 1))
            d["num_steps"] = len(alphas)
            torch.save(d, os.path.join(OUT, f"{case}.pt"))
            print(case, "steps", len(alphas))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
