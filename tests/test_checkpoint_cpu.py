"""CPU: checkpoint format compatibility with the reference (ref:ssp_vit2spn_tiny.py:53-72, 246;
ref:octmnist_ft_vit2spn.py:190) — same dictionary keys, same 808 state-dict keys, strict hand-off of
``online_network_1.state_dict()`` into a fine-tune backbone.  No kernels involved."""
import os

import torch

import vit2spn
from oracle import vit2spn_oracle as orc


def test_checkpoint_roundtrip_and_reference_key_layout(tmp_path):
    model = vit2spn.DualStreamNetwork()
    state = orc.init_state(21, 0.01)
    model.load_state_dict(state, strict=True)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)           # the reference's optimizer (ref:173)
    path = os.path.join(tmp_path, "ckpt.pth")
    vit2spn.save_checkpoint(model, opt, 7, 0.123, path)
    ck = torch.load(path)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "loss"}
    assert list(ck["model_state_dict"]) == orc.model_param_names()
    assert len(ck["optimizer_state_dict"]["param_groups"][0]["params"]) == 808
    model2 = vit2spn.DualStreamNetwork()
    opt2 = vit2spn.FusedAdam(model2.parameters(), lr=1e-4)        # drop-in optimizer loads the same state
    m, o, epoch, loss = vit2spn.load_checkpoint(model2, opt2, path)
    assert epoch == 7 and loss == 0.123
    sd = model2.state_dict()
    assert all(torch.equal(sd[k], state[k]) for k in state)
    # missing file → (model, optimizer, 0, inf), as the reference
    _, _, e0, l0 = vit2spn.load_checkpoint(model2, opt2, os.path.join(tmp_path, "nope.pth"))
    assert e0 == 0 and l0 == float("inf")
    # final export of the reference script (ref:246) → strict load into the fine-tune model's backbone
    export = os.path.join(tmp_path, "pretrained.pth")
    torch.save(model.online_network_1.state_dict(), export)
    ft = vit2spn.FineTunedModel(num_classes=4)
    ft.backbone.load_state_dict(torch.load(export), strict=True)
    assert torch.equal(ft.backbone.vit.embeddings.cls_token, state["online_network_1.vit.embeddings.cls_token"])


def test_single_stream_variant_structure():
    m = vit2spn.SingleStreamNetwork()
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 5_561_472 + (192 * 1024 + 1024 + 1024 * 128 + 128) + 33_024
    assert all(not p.requires_grad for p in m.target_network.parameters())
