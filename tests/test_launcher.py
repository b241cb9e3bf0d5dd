"""The launcher / compat layer of SURVEY §8(b): unmodified reference-style scripts on the CUDA backbone."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_mod():
    import vit2spn  # noqa: F401
    import importlib
    return importlib.import_module("vit2spn.run")


def test_shims_only_fill_gaps(monkeypatch):
    run = _run_mod()
    before = {n: sys.modules.get(n) for n in run.SHIMS}
    done = run.install_shims()
    try:
        for name in run.SHIMS:
            assert name in sys.modules or name not in done
        for name in done:                       # a shim is registered only when the real module is absent
            assert "compat" in sys.modules[name].__file__
        assert "numpy" not in run.install_shims(("numpy",))
    finally:
        for n in done:
            for k in [k for k in sys.modules if k == n or k.startswith(n + ".")]:
                if before.get(n) is None:
                    del sys.modules[k]


def test_medmnist_shim_contract(monkeypatch):
    """Item / labels contract the reference relies on (ref:octmnist_ft_vit2spn.py:47-50,177)."""
    monkeypatch.setenv("V2S_SHIM_DATASET_SIZE", "12")
    monkeypatch.setenv("V2S_SYNTHETIC_DATA", "1")
    import importlib.util
    d = os.path.join(ROOT, "vit-2spn_b200", "compat", "medmnist")
    spec = importlib.util.spec_from_file_location("_v2s_medmnist", os.path.join(d, "__init__.py"),
                                                  submodule_search_locations=[d])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["_v2s_medmnist"] = mod
    try:
        spec.loader.exec_module(mod)
        ds = mod.OCTMNIST(split="train", transform=None, download=True)
        assert len(ds) == 12 and ds.labels.shape == (12, 1)
        img, target = ds[3]
        assert img.mode == "L" and img.size == (28, 28)
        assert isinstance(target, np.ndarray) and target.shape == (1,) and 0 <= int(target[0]) < 4
        assert np.array_equal(np.asarray(ds[3][0]), np.asarray(img))          # deterministic
        assert len(mod.INFO["octmnist"]["label"]) == 4
        seen = []
        ds2 = mod.OCTMNIST(split="val", transform=lambda im: seen.append(im.size) or 1.5)
        assert ds2[0][0] == 1.5 and seen == [(28, 28)]
        monkeypatch.setenv("V2S_SYNTHETIC_DATA", "0")          # fabricated data is opt-in
        with pytest.raises(ImportError, match="V2S_SYNTHETIC_DATA"):
            mod.OCTMNIST(split="train")
    finally:
        for k in [k for k in sys.modules if k.startswith("_v2s_medmnist")]:
            del sys.modules[k]


def test_flop_count_shim():
    import importlib.util
    f = os.path.join(ROOT, "vit-2spn_b200", "compat", "fvcore", "nn", "__init__.py")
    spec = importlib.util.spec_from_file_location("_v2s_fvnn", f)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    head = nn.Sequential(nn.Linear(384, 1024), nn.ReLU(), nn.Dropout(0.3), nn.Linear(1024, 128))
    assert mod.FlopCountAnalysis(head, torch.randn(1, 384)).total() == 384 * 1024 + 1024 * 128
    assert mod.FlopCountAnalysis(head, (torch.randn(3, 384),)).total() == 3 * (384 * 1024 + 1024 * 128)
    assert mod.VIT_TINY_MAC_PER_IMAGE == 1253491200


def test_launcher_usage():
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "vit2spn.run"], capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 2 and "vit2spn.run" in r.stdout
    r = subprocess.run([sys.executable, "-m", "vit2spn.run", "/nonexistent.py"], capture_output=True, text=True,
                       env=env, cwd=ROOT)
    assert r.returncode == 2 and "no such script" in r.stderr


@pytest.mark.gpu
def test_launcher_runs_reference_style_script(tmp_path):
    """A script written against transformers / medmnist / fvcore / matplotlib, with its own dual-stream model,
    autocast + GradScaler loop and ``.data`` EMA, runs unchanged; the saved backbone has HF key names and loads
    into the stock ``transformers.ViTModel``."""
    env = dict(os.environ, PYTHONPATH=ROOT, V2S_SHIM_DATASET_SIZE="12", V2S_SYNTHETIC_DATA="1")
    out = tmp_path / "bb.pth"
    r = subprocess.run([sys.executable, "-m", "vit2spn.run", os.path.join(ROOT, "tests", "data", "mini_user_script.py"),
                        str(out)], capture_output=True, text=True, env=env, cwd=tmp_path, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "MINI_OK 6" in r.stdout and "SYNTHETIC DATA" in r.stdout
    sd = torch.load(out, map_location="cpu")
    assert len(sd) == 200 and "vit.embeddings.cls_token" in sd
    from transformers import ViTConfig, ViTModel
    hf = ViTModel(ViTConfig(hidden_size=192, num_hidden_layers=12, num_attention_heads=3, intermediate_size=768))
    hf.load_state_dict({k[4:]: v for k, v in sd.items()}, strict=True)
