"""Pins the GLASS-module restatements in oracle/nets.py against golden vectors produced
by the reference's own modules (tools/make_golden.py, run in the authoring container
from /root/reference; fixtures in tests/golden/).  CPU only."""
import os

import pytest
import torch

from oracle import nets
from tests import golden_common as gc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _check(out, name, rtol=1e-4, atol=1e-5):
    g = torch.load(os.path.join(GOLD, name + ".pt"))
    assert list(out.shape) == g["shape"]
    sl = out[tuple(slice(None, None, s) for s in g["stride"])]
    assert torch.allclose(sl, g["sample"], rtol=rtol, atol=atol), (sl - g["sample"]).abs().max()
    assert abs(out.double().sum().item() - g["sum"]) <= 1e-4 * max(1.0, g["abssum"])
    assert abs(out.double().abs().sum().item() - g["abssum"]) <= 1e-4 * max(1.0, g["abssum"])
    return g


@torch.no_grad()
def test_hybrid_net_golden():
    m = nets.ResNetFeatureExtractor(3, 256).eval()
    gc.seeded_fill(m, gc.SEEDS["hybrid"])
    _check(m(gc.seeded_input("hybrid")), "hybrid_net")


@torch.no_grad()
def test_fusion_net_golden():
    m = nets.MultiAspectGCAttention(512, 0.5, 8, 256).eval()
    gc.seeded_fill(m, gc.SEEDS["fusion"])
    _check(m(gc.seeded_input("fusion")), "fusion_net")


@torch.no_grad()
def test_p2p3_golden():
    m = nets.P2P3Fusion(256).eval()
    gc.seeded_fill(m, gc.SEEDS["p2p3"])
    p2, p3 = gc.seeded_input("p2p3")
    _check(m(p2, p3), "p2p3")


@torch.no_grad()
def test_encoder_golden():
    m = nets.BiLSTMBlockV2(256, 2).eval()
    gc.seeded_fill(m, gc.SEEDS["encoder"])
    _check(m(gc.seeded_input("encoder")), "encoder")


@pytest.mark.parametrize("case", ["decoder", "decoder_break"])
@torch.no_grad()
def test_decoder_golden(case):
    m = nets.AttentionRecognitionHead(97, 256, 256, 256, 26).eval()
    gc.seeded_fill(m, gc.SEEDS[case])
    if case == "decoder_break":
        gc.force_eos_bias(m)
    taps = {}
    probs = m.sample(gc.seeded_input(case), 26, 0, taps=taps)
    g = _check(probs, case, rtol=1e-4, atol=1e-6)
    assert taps["decoder_steps"] == g["num_steps"]
    # rows after the break stay zero (prediction_aster.py:73,98)
    assert probs[:, g["num_steps"]:].abs().sum().item() == 0.0
    if case == "decoder_break":
        assert 1 < g["num_steps"] < 26


@torch.no_grad()
def test_cnn_v1_1_golden():
    """a14: the recognizer's CNN_V1_1 against the reference's own class (tools/make_golden_cnn_v1_1.py)."""
    m = nets.CNN_V1_1(256).eval()
    g = torch.load(os.path.join(GOLD, "cnn_v1_1.pt"))
    assert sorted(m.state_dict().keys()) == g["keys"]   # same parameter names as the reference module
    gc.seeded_fill(m, 17)
    x = torch.randn(3, 256, 8, 32, generator=torch.Generator().manual_seed(1017))
    _check(m(x), "cnn_v1_1")
