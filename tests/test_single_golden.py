"""Single-stream variant against vectors produced by the REFERENCE's own `SingleStreamNetwork`
(tests/golden/make_single_golden.py → single_golden.npz): the oracle on CPU, the CUDA backbones on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_single_golden as mk  # noqa: E402  (helpers only; the reference is not touched at import)
from oracle import vit2spn_oracle as orc  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "single_golden.npz"))


def _check_grads(named_grads, rtol_norm):
    names = list(G["grad_names"])
    assert set(named_grads) == set(names)
    floor = 1e-6 * float(G["grad_norms"].max())          # analytically-zero key-bias gradients hold rounding noise
    for i, n in enumerate(names):
        g = named_grads[n].detach().cpu()
        ref_norm = float(G["grad_norms"][i])
        assert abs(g.double().norm().item() - ref_norm) <= rtol_norm * ref_norm + floor, n
        k = min(mk.SLICE, g.numel())
        np.testing.assert_allclose(g.flatten()[:k].numpy(), G["grad_slices"][i][:k], rtol=0, atol=rtol_norm * ref_norm + floor)


def _heads():
    ph = nn.Sequential(nn.Linear(192, 1024), nn.ReLU(), nn.Dropout(0.0), nn.Linear(1024, 128))
    qh = nn.Sequential(nn.Linear(128, 128), nn.ReLU(), nn.Linear(128, 128))
    hs = mk.heads_state()
    ph.load_state_dict({k[len("projection_head."):]: v for k, v in hs.items() if k.startswith("projection_head.")})
    qh.load_state_dict({k[len("prediction_head."):]: v for k, v in hs.items() if k.startswith("prediction_head.")})
    return ph, qh


def test_oracle_matches_reference_single_stream():
    base = orc.init_state(13, 0.01)
    v1, v2 = mk.inputs()
    leaves = {k: v.clone().requires_grad_(True) for k, v in orc.sub_state(base, "online_network_1").items()}
    ph, qh = _heads()
    fo = orc.backbone_features(leaves, v1)
    with torch.no_grad():
        ft = orc.backbone_features(orc.sub_state(base, "target_network_1"), v2)
    pred, tgt = qh(ph(fo)), ph(ft).detach()
    loss = -torch.mean(nn.CosineSimilarity(dim=1)(pred, tgt)) / mk.ACCUM
    loss.backward()
    np.testing.assert_allclose(pred.detach().numpy(), G["pred"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(tgt.numpy(), G["tgt"], rtol=0, atol=2e-6)
    assert abs(loss.item() - float(G["loss"])) <= 1e-5 * abs(float(G["loss"]))
    grads = {"online_network.vit." + k: v.grad for k, v in leaves.items() if v.grad is not None}
    grads.update({"projection_head." + n: p.grad for n, p in ph.named_parameters()})
    grads.update({"prediction_head." + n: p.grad for n, p in qh.named_parameters()})
    _check_grads(grads, 1e-4)


@pytest.mark.gpu
def test_cuda_single_stream_matches_reference_golden():
    """`vit2spn.SingleStreamNetwork` takes the reference's state_dict (strict, same key order) and in fp32 check mode
    reproduces one micro-step of ref:dsn_ssn/ssp_single.py:195-206: projections, loss, every gradient, Adam and the
    momentum-0.99 target update."""
    import vit2spn
    dev = torch.device("cuda", 0)
    model = vit2spn.SingleStreamNetwork()
    assert list(model.state_dict().keys()) == list(G["keys"])
    res = model.load_state_dict(mk.full_state(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    model.to(dev).train()
    model.projection_head[2].p = 0.0
    for net in (model.online_network, model.target_network):
        net.vit.compute_mode = "fp32"
    v1, v2 = mk.inputs()
    opt = vit2spn.FusedAdam(model.parameters(), lr=1e-4)
    opt.zero_grad()
    pred, tgt = model(v1.to(dev), v2.to(dev))
    loss = -torch.mean(nn.CosineSimilarity(dim=1)(pred, tgt)) / mk.ACCUM
    loss.backward()
    np.testing.assert_allclose(pred.detach().cpu().numpy(), G["pred"], rtol=0, atol=5e-6)
    np.testing.assert_allclose(tgt.detach().cpu().numpy(), G["tgt"], rtol=0, atol=5e-6)
    assert abs(loss.item() - float(G["loss"])) <= 1e-5 * abs(float(G["loss"]))
    _check_grads({n: p.grad for n, p in model.named_parameters() if p.grad is not None}, 1e-4)
    opt.step()
    model.update_target_network()                       # the reference's default momentum (0.99)
    post = model.state_dict()
    norms = np.array([post[k].double().norm().item() for k in post])
    np.testing.assert_allclose(norms, G["post_norms"], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(post["target_network.vit.embeddings.position_embeddings"].flatten()[:mk.SLICE].cpu().numpy(),
                               G["post_target_pos_slice"], rtol=0, atol=1e-6)
