"""Evaluator wire format: what the reference's TextEvaluator makes of the model output, so that the
packed detection records gathered at the end of the loop can feed the reference's evaluation scripts unchanged.

* ``get_instances_text``        glass/evaluation/text_evaluator.py:323-348
* ``instances_to_coco_json``    :351-415 (one dict per word: polys / boxes / rboxes / rec / score_text /
                                character_probs / score_detection)
* ``rotated_boxes_to_polygons`` :434-462, ``boxes_to_polygons`` :418-431
* ``to_eval_lines``             TextEvaluator.to_eval_format :156-239 without a lexicon: per-image
                                ``x1,y1,...,x4,y4,####text`` lines, thresholded on the rounded text / detection scores
* ``instances_from_packed``     the inverse of B200GlassRCNN.pack_detections (the all-gathered record)
Host-side formatting only (numpy / python); masks (``pred_masks`` -> rasterio polygons) are out of scope with the
mask branch."""
from typing import Dict, List, Tuple

import numpy as np
import torch

from .structures import Instances, RotatedBoxes
from .text import TextDecoder

# text_evaluator.py:333 -- a non-raw python string: "\#" keeps its backslash, so '\\' is one of the characters
SPECIAL_CHARACTERS = "'!?.:,*+\"()·[]/\\#$%;<=>@^_`{|}~"


def boxes_to_polygons(boxes: np.ndarray) -> np.ndarray:
    n = len(boxes)
    if n == 0:
        return np.array([]).reshape((0, 4, 2))
    p = np.zeros((n, 4, 2))
    p[:, 0, 0], p[:, 0, 1] = boxes[:, 0], boxes[:, 1]
    p[:, 1, 0], p[:, 1, 1] = boxes[:, 2], boxes[:, 1]
    p[:, 2, 0], p[:, 2, 1] = boxes[:, 2], boxes[:, 3]
    p[:, 3, 0], p[:, 3, 1] = boxes[:, 0], boxes[:, 3]
    return p


def rotated_boxes_to_polygons(boxes: np.ndarray) -> np.ndarray:
    n = len(boxes)
    if n == 0:
        return np.array([]).reshape((0, 4, 2))
    assert boxes.shape[-1] == 5, "The last dimension of input shape must be 5 for XYWHA format"
    cx, cy, w, h, a = (boxes[:, i] for i in range(5))
    t = np.deg2rad(-a)
    p = np.zeros((n, 4, 2))
    sin_t, cos_t = np.sin(t), np.cos(t)
    p[:, 0, 0] = cx + (h * sin_t - w * cos_t) / 2
    p[:, 1, 0] = cx + (h * sin_t + w * cos_t) / 2
    p[:, 2, 0] = cx - (h * sin_t - w * cos_t) / 2
    p[:, 3, 0] = cx - (h * sin_t + w * cos_t) / 2
    p[:, 0, 1] = cy - (h * cos_t + w * sin_t) / 2
    p[:, 1, 1] = cy - (h * cos_t - w * sin_t) / 2
    p[:, 2, 1] = cy + (h * cos_t + w * sin_t) / 2
    p[:, 3, 1] = cy + (h * cos_t - w * sin_t) / 2
    return p


def get_instances_text(text_probs: torch.Tensor, text_decoder: TextDecoder, only_remove_first_last_character: bool = True):
    """-> (texts, word scores, probabilities as numpy)."""
    if len(text_probs) == 0:
        return [], [], []
    text_probs = text_probs.detach().cpu()
    pred_probs, pred_indices = text_probs.max(dim=2)
    words = text_decoder.decode_attention(pred_indices.numpy(), pred_probs.numpy())
    texts = [w["text"] for w in words]
    scores = [w["score"] for w in words]
    if only_remove_first_last_character:
        for i in range(len(texts)):
            if len(texts[i]) > 0 and SPECIAL_CHARACTERS.find(texts[i][0]) > -1:
                texts[i] = texts[i][1:]
            if len(texts[i]) > 0 and SPECIAL_CHARACTERS.find(texts[i][-1]) > -1:
                texts[i] = texts[i][:-1]
    return texts, scores, text_probs.numpy()


def instances_to_coco_json(instances: Instances, file_name, text_decoder: TextDecoder,
                           only_remove_first_last_character: bool = True) -> List[Dict]:
    if len(instances) == 0:
        return []
    assert not instances.has("pred_masks"), "mask polygons are not supported (the mask branch is out of scope)"
    boxes = instances.pred_boxes.tensor.detach().cpu().numpy()
    polygons = (boxes_to_polygons(boxes) if boxes.shape[1] == 4 else rotated_boxes_to_polygons(boxes)).tolist()
    rboxes = (rotated_boxes_to_polygons(instances.pred_rboxes.tensor.detach().cpu().numpy()).tolist()
              if instances.has("pred_rboxes") else [[]] * len(polygons))
    # the reference feeds pred_boxes (5 columns when rotated) to the axis-aligned helper: columns 0..3 are used
    bxs = boxes_to_polygons(boxes).tolist()
    texts, scores_text, probs = get_instances_text(instances.pred_text_prob, text_decoder, only_remove_first_last_character)
    scores_detection = instances.scores.detach().cpu().tolist()
    out = []
    for poly, rec, score_text, cprobs, box, rbox, score_det in zip(polygons, texts, scores_text, probs, bxs, rboxes,
                                                                   scores_detection):
        if len(rec) > 0 and len(poly) >= 3:
            out.append({"image_id": file_name, "category_id": 1, "polys": poly, "boxes": box, "rboxes": rbox, "rec": rec,
                        "score_text": np.float64(score_text).tolist(), "character_probs": np.float64(cprobs).tolist(),
                        "score_detection": np.float64(score_det).tolist()})
    return out


def eval_file_name(image_id, dataset: str) -> str:
    """to_eval_format :218-225."""
    if dataset in ("totaltext", "textocr"):
        return "{:07d}.txt".format(int(image_id))
    if dataset.startswith("icdar"):
        return "{}.txt".format(int(image_id))
    raise ValueError(dataset)


def to_eval_lines(records: List[Dict], dataset: str = "totaltext", text_cf_th: float = 0.5,
                  detection_cf_th: float = 0.0) -> Dict[str, str]:
    """file name -> file content of the per-image detection files the official scripts read (no lexicon, end-to-end
    mode).  Scores go through the same text round trip as in the reference: rounded to 3 digits, printed, parsed."""
    files: Dict[str, str] = {}
    for r in records:
        if not r["score_text"] > 0.001:
            continue
        cors = "".join(str(int(pt[0])) + "," + str(int(pt[1])) + "," for pt in r.get("polys", []))
        rec = "".join(c for c in r["rec"] if ord(c) < 128)
        score_text, score_det = str(round(r["score_text"], 3)), str(round(r["score_detection"], 3))
        # the reference writes "<id>: <cors><st>|<sd>,####<rec>", strips the line and splits it again
        line = (cors + score_text + "|" + score_det + ",####" + rec).strip()
        ptr = line.split(",####")
        if float(score_text) < text_cf_th or float(score_det) < detection_cf_th:
            fn = eval_file_name(r["image_id"], dataset)
            files.setdefault(fn, "")   # the reference opens the file in append mode before it tests the thresholds
            continue
        fn = eval_file_name(r["image_id"], dataset)
        files[fn] = files.get(fn, "") + ",".join(ptr[0].split(",")[:-1]) + ",####" + ptr[1] + "\n"
    return files


def instances_from_packed(rec: torch.Tensor, image_sizes: List[Tuple[int, int]], steps: int = 26,
                          num_classes: int = 97) -> List[Instances]:
    """[n_img, max_det, 10 + steps*classes] records (B200GlassRCNN.pack_detections, possibly all-gathered) ->
    per-image Instances with the reference's field names."""
    rec = rec.detach().cpu()
    out = []
    for i in range(rec.shape[0]):
        k = int((rec[i, :, 0] > 0.5).sum())
        r = rec[i, :k]
        out.append(Instances(image_sizes[i], pred_boxes=RotatedBoxes(r[:, 1:6].clone()), scores=r[:, 6].clone(),
                             pred_classes=r[:, 7].long(), orientations=r[:, 8:10].clone(),
                             pred_text_prob=r[:, 10:].reshape(k, steps, num_classes).clone()))
    return out
