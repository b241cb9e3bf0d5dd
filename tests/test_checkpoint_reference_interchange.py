"""Checkpoint interchange with the REAL reference code.  Needs /root/reference (present in the build
container only; skipped elsewhere): the reference's own `DualStreamNetwork`, `save_checkpoint` and `load_checkpoint`
(AST-extracted from ref:ssp_vit2spn_tiny.py:53-72,109-166, executed unmodified) write and read the files."""
import os
import sys

import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
REF_SCRIPT = "/root/reference/ssp_vit2spn_tiny.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF_SCRIPT), reason="the reference is not mounted on this machine")


def _reference():
    import make_golden as mg
    ns = mg.reference_namespace("ssp_vit2spn_tiny.py")
    ns["os"] = os
    mg.extract_classes(REF_SCRIPT, {"save_checkpoint", "load_checkpoint"}, ns)
    return ns


def test_reference_checkpoint_loads_here_and_ours_loads_there(tmp_path):
    import vit2spn
    from oracle import vit2spn_oracle as orc
    ns = _reference()
    state = orc.init_state(31, 0.01)
    ref_model = ns["DualStreamNetwork"]()
    ref_model.load_state_dict(state, strict=True)
    ref_opt = torch.optim.Adam(ref_model.parameters(), lr=1e-4)
    # give the reference optimizer real state: one step on synthetic gradients for every trainable tensor
    g = torch.Generator().manual_seed(0)
    for p in ref_model.parameters():
        if p.requires_grad:
            p.grad = torch.randn(p.shape, generator=g) * 1e-3
    ref_opt.step()
    path = os.path.join(tmp_path, "ref_ckpt.pth")
    ns["save_checkpoint"](ref_model, ref_opt, 10, 0.25, path)                 # the reference writes ...
    ours = vit2spn.DualStreamNetwork()
    ours_opt = vit2spn.FusedAdam(ours.parameters(), lr=1e-4)
    _, _, epoch, loss = vit2spn.load_checkpoint(ours, ours_opt, path)         # ... this library reads
    assert epoch == 10 and loss == 0.25
    ref_sd, our_sd = ref_model.state_dict(), ours.state_dict()
    assert list(ref_sd) == list(our_sd)
    assert all(torch.equal(ref_sd[k], our_sd[k]) for k in ref_sd)
    r_state, o_state = ref_opt.state_dict()["state"], ours_opt.state_dict()["state"]
    assert set(r_state) == set(o_state) and len(r_state) == 408               # Adam state of the 400 + 8 trainable tensors
    assert len(ref_opt.state_dict()["param_groups"][0]["params"]) == 808 == len(ours_opt.state_dict()["param_groups"][0]["params"])
    for k in r_state:
        for f in ("exp_avg", "exp_avg_sq"):
            assert torch.equal(r_state[k][f], o_state[k][f].cpu()), (k, f)
        assert float(r_state[k]["step"]) == float(o_state[k]["step"])
    # and back: this library writes, the reference's load_checkpoint reads
    back = os.path.join(tmp_path, "our_ckpt.pth")
    vit2spn.save_checkpoint(ours, ours_opt, 11, 0.5, back)
    ref2 = ns["DualStreamNetwork"]()
    ref2_opt = torch.optim.Adam(ref2.parameters(), lr=1e-4)
    ns["device"] = torch.device("cpu")
    _, _, epoch2, loss2 = ns["load_checkpoint"](ref2, ref2_opt, back)
    assert epoch2 == 11 and loss2 == 0.5
    assert all(torch.equal(ref2.state_dict()[k], ref_sd[k]) for k in ref_sd)
    assert len(ref2_opt.state_dict()["state"]) == 408
