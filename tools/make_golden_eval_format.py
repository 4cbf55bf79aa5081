"""Generate tests/golden/eval_format.pt with the REFERENCE's own evaluator code (glass/evaluation/text_evaluator.py):
``instances_to_coco_json`` (:351-415, the per-detection record), ``get_instances_text`` (:323-348),
``rotated_boxes_to_polygons`` (:434-462) and ``TextEvaluator.to_eval_format`` (:156-239, the
``x,y,...,####text`` line files).  Authoring container only: the module's heavy imports (rasterio, shapely,
Levenshtein, detectron2, fvcore) are absent offline and are stubbed -- none of them is reached by these functions.

    python tools/make_golden_eval_format.py
"""
import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"


class Boxes5:
    def __init__(self, t):
        self.tensor = t


class Instances:
    def __init__(self, **fields):
        self._fields = fields

    def has(self, k):
        return k in self._fields

    def __getattr__(self, k):
        if k.startswith("_") or k not in self._fields:
            raise AttributeError(k)
        return self._fields[k]

    def __len__(self):
        return len(self._fields["scores"])


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference():
    stub("rasterio", features=None, Affine=None)
    stub("rasterio.features")
    stub("shapely")
    stub("shapely.geometry", Polygon=object, LinearRing=object)
    stub("fvcore")
    stub("fvcore.common")
    stub("fvcore.common.file_io", PathManager=object)
    stub("detectron2")
    stub("detectron2.data", MetadataCatalog=object)
    stub("detectron2.evaluation")
    stub("detectron2.evaluation.evaluator", DatasetEvaluator=object)
    stub("detectron2.utils", comm=None)
    for name in ("glass", "glass.evaluation", "glass.modeling", "glass.modeling.recognition"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, *name.split("."))]
        sys.modules[name] = m
    stub("glass.evaluation.text_eval_script")
    stub("glass.evaluation.lexicon_utils", find_match_word=None, get_lexicon=None)
    sys.modules["glass.evaluation"].text_eval_script = sys.modules["glass.evaluation.text_eval_script"]

    def load(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[modname] = m
        spec.loader.exec_module(m)
        return m
    te = load("glass.modeling.recognition.text_encoder", "glass/modeling/recognition/text_encoder.py")
    ev = load("glass.evaluation.text_evaluator", "glass/evaluation/text_evaluator.py")
    return te, ev


from tests.golden_common import make_eval_inputs as make_inputs  # noqa: E402  (shared with the tests)


def main():
    te, ev = load_reference()
    y = yaml.safe_load(open(os.path.join(REF, "configs/glass_pretrain.yaml")))
    rh = y["MODEL"]["ROI_RECOGNIZER_HEAD"]
    head = types.SimpleNamespace(NAME="RecognizerRCNNHeadV3", MAX_WORD_LENGTH=rh["MAX_WORD_LENGTH"], CHARACTER_SET=rh["CHARACTER_SET"],
                                 UNK_SYMBOL_PRED=rh.get("UNK_SYMBOL_PRED", False), LABELS_TYPE="attention", IGNORE_TEXT=[],
                                 IGNORE_EMPTY_TEXT=True)
    enc = te.TextEncoder(types.SimpleNamespace(MODEL=types.SimpleNamespace(ROI_RECOGNIZER_HEAD=head)))
    cases = []
    all_records = []
    for seed, n, image_id in [(0, 12, 17), (1, 30, 204), (2, 0, 5), (3, 7, 1500)]:
        boxes, scores, probs = make_inputs(seed, n)
        inst = Instances(pred_boxes=Boxes5(boxes), scores=scores, pred_text_prob=probs)
        recs = {flag: ev.instances_to_coco_json(inst, image_id, enc, flag) for flag in (True, False)}
        texts, text_scores, _ = ev.get_instances_text(probs, enc, True) if n else ([], [], [])
        # character_probs is np.float64(probs[i]).tolist() (26 x 97 numbers per record): store the detection index
        # it came from instead, the test rebuilds and compares the full lists
        full = [np.float64(probs[i].numpy()).tolist() for i in range(n)]
        for rr in recs.values():
            for r in rr:
                r["character_probs"] = {"det_index": full.index(r["character_probs"])}
        cases.append({"seed": seed, "n": n, "image_id": image_id, "boxes": boxes, "scores": scores,  # probs: regenerated from the seed
                     
                      "records": recs[True], "records_keep_specials": recs[False],
                      "texts": list(texts), "text_scores": [float(s) for s in text_scores]})
        all_records += recs[True]
    polys = ev.rotated_boxes_to_polygons(cases[1]["boxes"].numpy())
    # TextEvaluator.to_eval_format on the gathered records (no lexicon, end-to-end mode), both file-name conventions
    files = {}
    for dataset in ("totaltext", "icdar15"):
        evaluator = object.__new__(ev.TextEvaluator)
        evaluator.lexicon, evaluator._word_spotting, evaluator.dataset = None, False, dataset
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            try:
                with open("text_results.json", "w") as f:
                    json.dump([dict(r, character_probs=[]) for r in all_records], f)  # unused without a lexicon
                evaluator.to_eval_format("text_results.json", "out", text_cf_th=0.3, detection_cf_th=0.2)
                files[dataset] = {fn: open(os.path.join("out", fn)).read() for fn in sorted(os.listdir("out"))}
            finally:
                os.chdir(cwd)
    torch.save({"cases": cases, "polygons_case1": torch.from_numpy(polys), "eval_files": files,
                "text_cf_th": 0.3, "detection_cf_th": 0.2},
               os.path.join(ROOT, "tests", "golden", "eval_format.pt"))
    print("records per case:", [len(c["records"]) for c in cases], "files:", {k: len(v) for k, v in files.items()})
    print(list(files["totaltext"].items())[0])


if __name__ == "__main__":
    main()
