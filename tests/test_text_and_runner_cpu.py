"""The host-side text decoder against golden vectors produced by the reference's own
TextEncoder.decode_attention (tools/make_golden_text.py), and the runner's scale logic."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "text_decode.pt")


def test_text_decoder_matches_reference_golden():
    from glass_text_spotting_b200.text import DEFAULT_CHARSET, TextDecoder
    g = torch.load(GOLD)
    assert g["charset"] == DEFAULT_CHARSET
    dec = TextDecoder()
    assert dec.character == g["characters"] and len(dec.character) == 97
    out = dec.decode_attention(g["idx"].numpy(), g["probs"].numpy(), include_stop_symbol_conf=True)
    assert [o["text"] for o in out] == g["text"]
    assert np.allclose([o["score"] for o in out], g["score"], rtol=1e-6)
    assert [len(o["character_scores"]) for o in out] == g["nchar"]
    out2 = dec.decode_attention(g["idx"].numpy(), g["probs"].numpy(), include_stop_symbol_conf=False)
    assert [o["text"] for o in out2] == g["text_nostop"]
    assert np.allclose([o["score"] for o in out2], g["score_nostop"], rtol=1e-6)


def test_decode_probs_uses_greedy_argmax():
    from glass_text_spotting_b200.text import TextDecoder
    dec = TextDecoder()
    p = torch.zeros(2, 26, 97)
    word = [dec.character.index(c) for c in "Hi!"] + [1]
    for t, c in enumerate(word):
        p[0, t, c] = 0.9
    p[0, len(word):, 5] = 1.0
    p[1, :, 1] = 1.0   # immediate stop -> empty word
    out = dec.decode_probs(p)
    assert out[0]["text"] == "Hi!" and abs(out[0]["score"] - 0.9 ** 4) < 1e-6 and out[1]["text"] == ""
    assert dec.decode_probs(torch.zeros(0, 26, 97)) == []


def test_runner_scale_ratio_matches_reference_rule():
    """glass/inference/glass_runner.py:111-121 with the totaltext thresholds (1200 / 1600 / 2x)."""
    from glass_text_spotting_b200.runner import B200GlassRunner
    r = B200GlassRunner.__new__(B200GlassRunner)
    r.min_target_size, r.max_target_size, r.max_upscale_ratio = 1200, 1600, 2.0
    assert r.get_inference_scale_ratio((1024, 1024, 3)) == 1200 / 1024      # the 1.171875 of SURVEY 8d cfg 5
    assert r.get_inference_scale_ratio((400, 300, 3)) == 2.0                 # capped up-scaling
    assert r.get_inference_scale_ratio((3200, 1000, 3)) == 0.5
    assert r.get_inference_scale_ratio((1300, 1500, 3)) == 1
