"""oracle/ -- TEST INFRASTRUCTURE, not product code.

CPU restatement (PyTorch fp32 + a small C library) of the reference's algorithm for
GLASS's per-image dense forward path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import anything from here; the product package
(``glass_text_spotting_b200``) never does.

Parity pinning: the reference ships no tests or golden vectors, and its
arithmetic below the GLASS modules lives in detectron2 v0.6, which is not vendored and
not installable offline.  The oracle is therefore pinned by
  * detectron2's upstream known-answer tests for the restated operators
    (tests/test_oracle_d2_ops.py), torchvision angle-0 equivalences, and
  * golden vectors generated in the authoring container by importing the reference's
    own pure-torch modules from /root/reference (tools/make_golden.py ->
    tests/golden/*.pt; checked by tests/test_oracle_golden.py), including the box branch's inference
    (tools/make_golden_box_inference.py), the recognizer branch end to end through the reference's own
    ``_forward_recognizer`` + ``RecognizerRCNNHeadV3`` (tools/make_golden_recognizer_branch.py), GlassRCNN's
    ``_postprocess`` / ``detector_postprocess`` (tools/make_golden_meta_postprocess.py), the word post-processor,
    the evaluator formats, the GlassRunner flow and the mask paste.
    The box branch end to end likewise (tools/make_golden_box_branch.py).
The detectron2-recalled parts (backbone wiring, RPN, poolers) have no reference-run
pin: for them parity is "unpinned beyond the upstream KATs"  -- plus, for
the ResNet-50 / FPN wiring, a cross-check against torchvision's independent
implementations of the same architectures (tests/test_oracle_backbone_torchvision.py), and for the
proposal selection / box decode a cross-check against torchvision's RPN ``filter_proposals`` and ``BoxCoder``
at angle 0 (tests/test_oracle_rpn_torchvision.py), for the rotated IoU OpenCV's rotated-rectangle
intersection and for RoIAlignRotated a torch grid_sample formulation (tests/test_oracle_rotated_independent.py).
"""
