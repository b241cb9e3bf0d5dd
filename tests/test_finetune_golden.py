"""Fine-tune / evaluation path against vectors produced by the REFERENCE's own `FineTunedModel`
(tests/golden/make_finetune_golden.py → finetune_golden.npz): the oracle on CPU, the CUDA backbone on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_finetune_golden as mk  # noqa: E402  (helpers only; the reference is not touched at import)
from oracle import vit2spn_oracle as orc  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "finetune_golden.npz"))


def _check_grads(named_grads, rtol_norm):
    """per-tensor norm and leading slice against the reference.  Some gradients are analytically zero (key biases;
    the last block's output bias, whose uniform shift of the features BatchNorm removes): both sides then hold
    rounding noise, hence the floor relative to the largest gradient norm of the model."""
    names = list(G["grad_names"])
    assert set(named_grads) == set(names)
    floor = 1e-6 * float(G["grad_norms"].max())
    for i, n in enumerate(names):
        g = named_grads[n].detach().cpu()
        ref_norm = float(G["grad_norms"][i])
        assert abs(g.double().norm().item() - ref_norm) <= rtol_norm * ref_norm + floor, n
        k = min(mk.SLICE, g.numel())
        np.testing.assert_allclose(g.flatten()[:k].numpy(), G["grad_slices"][i][:k], rtol=0, atol=rtol_norm * ref_norm + floor)


def test_oracle_backbone_matches_reference_finetune_model():
    x, y = mk.inputs()
    leaves = {k: v.clone().requires_grad_(True) for k, v in orc.sub_state(orc.init_state(9, 0.02), "online_network_1").items()}
    head = nn.Sequential(nn.Linear(192, 128), nn.BatchNorm1d(128), nn.ReLU(), nn.Dropout(0.0), nn.Linear(128, mk.NUM_CLASSES))
    head.load_state_dict({k[3:]: v for k, v in mk.head_state().items()})
    head.eval()
    with torch.no_grad():
        probs = torch.softmax(head(orc.backbone_features(leaves, x)), dim=1)
    np.testing.assert_allclose(probs.numpy(), G["eval_probs"], rtol=0, atol=2e-6)
    head.train()
    logits = head(orc.backbone_features(leaves, x))
    loss = nn.CrossEntropyLoss(weight=torch.tensor(mk.CLASS_WEIGHTS))(logits, y)
    loss.backward()
    np.testing.assert_allclose(logits.detach().numpy(), G["logits"], rtol=0, atol=5e-6)
    assert abs(loss.item() - float(G["loss"])) <= 1e-6 * abs(float(G["loss"]))
    grads = {"backbone.vit." + k: v.grad for k, v in leaves.items() if v.grad is not None}
    grads.update({"fc." + n: p.grad for n, p in head.named_parameters()})
    _check_grads(grads, 1e-4)      # fp32 summation order differs between the functional oracle and the HF modules


@pytest.mark.gpu
def test_cuda_finetune_model_matches_reference_golden():
    """`vit2spn.FineTunedModel` takes the reference's state_dict (strict), and in fp32 check mode reproduces its
    evaluation probabilities (ref:129-137) and one training step (ref:95-104,187-192: weighted CE, Adam + L2) —
    logits, loss, every gradient, the updated head and the BatchNorm running statistics."""
    import vit2spn
    dev = torch.device("cuda", 0)
    model = vit2spn.FineTunedModel(mk.NUM_CLASSES)
    assert list(model.state_dict().keys()) == list(G["keys"])                 # same keys, same order as the reference
    res = model.load_state_dict(mk.full_state(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    model.to(dev)
    model.backbone.vit.compute_mode = "fp32"
    x, y = mk.inputs()
    x, y = x.to(dev), y.to(dev)
    model.eval()
    with torch.no_grad():
        probs = torch.softmax(model(x), dim=1)
    np.testing.assert_allclose(probs.cpu().numpy(), G["eval_probs"], rtol=0, atol=5e-6)
    model.train()
    model.fc[3].p = 0.0
    crit = nn.CrossEntropyLoss(weight=torch.tensor(mk.CLASS_WEIGHTS, device=dev))
    opt = vit2spn.FusedAdam(model.parameters(), lr=1e-4, weight_decay=1e-4)
    opt.zero_grad()
    logits = model(x)
    loss = crit(logits, y)
    loss.backward()
    np.testing.assert_allclose(logits.detach().cpu().numpy(), G["logits"], rtol=0, atol=1e-5)
    assert abs(loss.item() - float(G["loss"])) <= 1e-5 * abs(float(G["loss"]))
    _check_grads({n: p.grad for n, p in model.named_parameters() if p.grad is not None}, 1e-4)
    opt.step()
    post = model.state_dict()
    norms = np.array([post[k].double().norm().item() for k in post])
    # Tensors whose gradient is analytically zero (key biases, the last block's output bias) carry pure rounding
    # noise, and Adam's first step moves every element by +-lr whatever the gradient's magnitude: their post-step
    # norm is only defined up to lr * sqrt(numel).  Every other tensor must match the reference to 2e-6.
    gnorm = dict(zip(list(G["grad_names"]), G["grad_norms"]))
    noise_floor = 1e-5 * float(G["grad_norms"].max())
    for k, got, want in zip(post, norms, G["post_norms"]):
        noisy = k in gnorm and float(gnorm[k]) <= noise_floor
        atol = 2.0 * 1e-4 * post[k].numel() ** 0.5 if noisy else 1e-7
        assert abs(got - want) <= 2e-6 * abs(want) + atol, (k, got, want, noisy)
    np.testing.assert_allclose(post["fc.0.weight"].flatten()[:mk.SLICE].cpu().numpy(), G["post_fc0_weight_slice"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(post["fc.1.running_mean"].cpu().numpy(), G["post_bn_running_mean"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(post["fc.1.running_var"].cpu().numpy(), G["post_bn_running_var"], rtol=0, atol=2e-6)
