"""``ViTModel.from_pretrained`` (ref:ssp_vit2spn_tiny.py:112): resolution, formats, and loud failure."""
import os

import pytest
import torch


def test_from_pretrained_raises_without_checkpoint(monkeypatch):
    import vit2spn
    monkeypatch.setenv("V2S_ALLOW_RANDOM_INIT", "0")
    with pytest.raises(OSError, match="V2S_ALLOW_RANDOM_INIT"):
        vit2spn.ViTModel.from_pretrained("WinKawaks/vit-tiny-patch16-224", output_hidden_states=True)
    with pytest.raises(OSError):
        vit2spn.ViTBackbone()


def test_from_pretrained_random_init_is_announced(monkeypatch, capfd):
    import vit2spn
    from vit2spn import modules
    monkeypatch.setenv("V2S_ALLOW_RANDOM_INIT", "1")
    monkeypatch.setattr(modules, "_warned_random_init", False)
    vit2spn.ViTBackbone()
    assert "RANDOM INIT" in capfd.readouterr().err


@pytest.mark.parametrize("fmt", ["safetensors", "bin"])
def test_from_pretrained_loads_hf_classification_checkpoint(tmp_path, fmt, capfd):
    """A ViTForImageClassification-style checkpoint directory (``vit.`` prefix, classifier, no pooler — the layout of
    WinKawaks/vit-tiny-patch16-224) in both weight formats; a checkpoint of another geometry is rejected."""
    import vit2spn
    src = vit2spn.ViTModel(vit2spn.ViTConfig())
    sd = {"vit." + k: v.detach().clone().contiguous() for k, v in src.state_dict().items() if not k.startswith("pooler.")}
    sd["classifier.weight"] = torch.zeros(1000, 192)
    sd["classifier.bias"] = torch.zeros(1000)
    if fmt == "safetensors":
        from safetensors.torch import save_file
        save_file(sd, str(tmp_path / "model.safetensors"))
    else:
        torch.save(sd, str(tmp_path / "pytorch_model.bin"))
    got = vit2spn.ViTModel.from_pretrained(str(tmp_path), output_hidden_states=True)
    for k, v in src.state_dict().items():
        if not k.startswith("pooler."):
            assert torch.equal(v, got.state_dict()[k]), k
    assert "missing ['pooler.dense.weight', 'pooler.dense.bias']" in capfd.readouterr().err
    bad = dict(sd)
    del bad["vit.encoder.layer.3.output.dense.weight"]
    bad["vit.encoder.layer.12.output.dense.bias"] = torch.zeros(192)
    other = tmp_path / "other"
    os.makedirs(other)
    torch.save(bad, str(other / "pytorch_model.bin"))
    with pytest.raises(RuntimeError, match="does not match"):
        vit2spn.ViTModel.from_pretrained(str(other))
