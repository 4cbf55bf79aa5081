"""Generate tests/golden/text_decode.pt with the REFERENCE's own TextEncoder.decode_attention
(glass/modeling/recognition/text_encoder.py is pure numpy/torch, importable here).  Authoring container only."""
import importlib.util
import os
import types

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

spec = importlib.util.spec_from_file_location("ref_text_encoder", os.path.join(REF, "glass/modeling/recognition/text_encoder.py"))
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

y = yaml.safe_load(open(os.path.join(REF, "configs/glass_pretrain.yaml")))
rh = y["MODEL"]["ROI_RECOGNIZER_HEAD"]
head = types.SimpleNamespace(NAME="RecognizerRCNNHeadV3", MAX_WORD_LENGTH=rh["MAX_WORD_LENGTH"], CHARACTER_SET=rh["CHARACTER_SET"],
                             UNK_SYMBOL_PRED=rh.get("UNK_SYMBOL_PRED", False), LABELS_TYPE="attention", IGNORE_TEXT=[],
                             IGNORE_EMPTY_TEXT=True)
cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(ROI_RECOGNIZER_HEAD=head))
enc = mod.TextEncoder(cfg)

rng = np.random.RandomState(0)
n, t = 64, 26
idx = rng.randint(2, 97, size=(n, t))
for i in range(n):   # stop symbols at assorted places: none, first step, last step, repeated
    if i % 4 != 0:
        idx[i, rng.randint(0, t)] = 1
    if i % 7 == 0:
        idx[i, rng.randint(0, t)] = 1
idx[1, 0] = 1
idx[2, t - 1] = 1
probs = rng.uniform(0.2, 1.0, size=(n, t)).astype(np.float32)
out = enc.decode_attention(idx.copy(), probs.copy(), include_stop_symbol_conf=True)
out_nostop = enc.decode_attention(idx.copy(), probs.copy(), include_stop_symbol_conf=False)
torch.save({"charset": rh["CHARACTER_SET"], "characters": enc.character, "idx": torch.from_numpy(idx), "probs": torch.from_numpy(probs),
            "text": [o["text"] for o in out], "score": [float(o["score"]) for o in out],
            "nchar": [len(o["character_scores"]) for o in out],
            "text_nostop": [o["text"] for o in out_nostop], "score_nostop": [float(o["score"]) for o in out_nostop]},
           os.path.join(ROOT, "tests", "golden", "text_decode.pt"))
print("ok", len(enc.character), repr(out[0]["text"]), out[0]["score"])
