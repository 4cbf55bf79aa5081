"""Host-side text decoding of the recognizer output -- the step right after the hot path in every caller
(glass/modeling/recognition/text_encoder.py:113-151 ``decode_attention``; used by
glass/evaluation/text_evaluator.py:325-326 on the per-step argmax / max of ``pred_text_prob``)."""
from typing import Dict, List

import numpy as np
import torch

# configs/glass_pretrain.yaml:7-8 (95 printable characters; YAML '' un-escapes to ')
DEFAULT_CHARSET = "0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ!\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~ "


class TextDecoder:
    """index -> character table of TextEncoder in 'attention' mode: ['[GO]', '[s]'] + character set
    (text_encoder.py:31-46; 2 + 95 = 97 classes)."""

    def __init__(self, character_set: str = DEFAULT_CHARSET, unk_symbol: bool = False):
        self.character = ["[GO]", "[s]"] + (["[UNK]"] if unk_symbol else []) + list(character_set)
        self.stop_index = self.character.index("[s]")

    def decode_attention(self, pred_indices: np.ndarray, pred_probs: np.ndarray = None,
                         include_stop_symbol_conf: bool = True) -> List[Dict]:
        """Same contract as TextEncoder.decode_attention: text up to (excluding) the first stop symbol; word
        score = product of the character probabilities (incl. the stop symbol's when requested)."""
        pred_indices = np.asarray(pred_indices)
        n, t = pred_indices.shape
        is_stop = pred_indices == self.stop_index
        first_stop = np.where(is_stop.any(axis=1), is_stop.argmax(axis=1), t)
        steps = np.arange(t)[None, :]
        mask = steps < first_stop[:, None]
        if include_stop_symbol_conf:
            mask = steps <= first_stop[:, None]
        out = []
        for i in range(n):
            idx = pred_indices[i, mask[i]]
            if include_stop_symbol_conf and len(idx) and idx[-1] == self.stop_index:
                chars = idx[:-1]
            else:
                chars = idx
            text = "".join(self.character[c] for c in chars)
            if pred_probs is not None:
                conf = np.asarray(pred_probs)[i, mask[i]]
                out.append({"text": text, "score": float(np.prod(conf)), "character_scores": conf})
            else:
                out.append({"text": text, "score": 1, "character_scores": [1] * len(text)})
        return out

    def decode_probs(self, pred_text_prob: torch.Tensor) -> List[Dict]:
        """[K, steps, classes] probabilities -> decoded words (greedy argmax per step, as text_evaluator.py:326)."""
        if pred_text_prob.shape[0] == 0:
            return []
        probs, idx = pred_text_prob.max(dim=2)
        return self.decode_attention(idx.cpu().numpy(), probs.cpu().numpy())
