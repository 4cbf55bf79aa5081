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


def test_load_checkpoint_d2_format(tmp_path):
    """A DetectionCheckpointer-style .pth ({"model": state_dict, ...}) loads by name; problems are named."""
    import numpy as np
    import pytest
    from glass_text_spotting_b200 import weights
    sd = weights.random_state_dict(2)
    ck = {"model": {("module." + k): (v.numpy() if i % 2 else v) for i, (k, v) in enumerate(sd.items())},
          "iteration": 599999, "optimizer": {}}
    ck["model"]["module.backbone.bottom_up.stem.conv1.norm.num_batches_tracked"] = torch.tensor(7)
    ck["model"]["module.roi_heads.mask_head.deconv.weight"] = np.zeros((2, 2), dtype=np.float32)     # not read without masks
    p = tmp_path / "model_0599999.pth"
    torch.save(ck, p)
    got = weights.load_checkpoint(str(p))
    assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    # a bare state_dict without the orientation head (MODEL.ORIENTATION_ON false) is accepted
    bare = {k: v for k, v in sd.items() if "orientation_pred" not in k}
    torch.save(bare, tmp_path / "bare.pth")
    assert set(weights.load_checkpoint(str(tmp_path / "bare.pth"))) == set(bare)
    # missing parameters are reported by name
    del bare["roi_heads.hybrid_net.ConvNet.conv0_1.weight"]
    torch.save(bare, tmp_path / "broken.pth")
    with pytest.raises(KeyError, match="conv0_1.weight"):
        weights.load_checkpoint(str(tmp_path / "broken.pth"))
    with pytest.raises(KeyError, match="mask_head"):
        weights.load_checkpoint(str(p), mask=True)
    with pytest.raises(ValueError, match="pkl"):
        weights.load_checkpoint("R-50.pkl")


def test_load_checkpoint_refuses_arbitrary_pickles(tmp_path):
    """weights_only=True first; a file that needs full unpickling is refused unless opted in."""
    import pytest
    import torch
    from glass_text_spotting_b200 import weights

    class Evil:
        def __reduce__(self):
            return (print, ("arbitrary code ran",))
    p = tmp_path / "evil.pth"
    torch.save({"model": {"x": Evil()}}, str(p))
    with pytest.raises(RuntimeError, match="allow_pickle=True"):
        weights.load_checkpoint(str(p))
