"""The product's random weight factory produces exactly the oracle's state_dict names and shapes
(= the detectron2 / reference names of SURVEY.md A.10), so released checkpoints load on both sides."""
import torch


def test_random_state_dict_matches_oracle_names_and_shapes():
    from glass_text_spotting_b200 import weights
    from oracle import model as om
    sd = weights.random_state_dict(0)
    ref = om.GlassOracle().state_dict()
    ref = {k: v for k, v in ref.items() if not k.endswith("num_batches_tracked")}
    missing = sorted(set(ref) - set(sd))
    extra = sorted(set(sd) - set(ref))
    assert not missing, missing[:10]
    assert not extra, extra[:10]
    for k, v in ref.items():
        assert tuple(sd[k].shape) == tuple(v.shape), (k, tuple(sd[k].shape), tuple(v.shape))
        assert torch.isfinite(sd[k]).all()


def test_oracle_loads_product_weights():
    from glass_text_spotting_b200 import weights
    from oracle import model as om
    o = om.GlassOracle()
    missing, unexpected = o.load_state_dict(weights.random_state_dict(1), strict=False)
    assert not unexpected and all(k.endswith("num_batches_tracked") for k in missing)
