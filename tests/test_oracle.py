"""CPU: the oracle restatement vs the golden vectors produced by the REFERENCE classes
(tests/golden/make_golden.py).  This is what pins ``oracle/`` to the reference."""
import numpy as np
import pytest
import torch

from oracle import vit2spn_oracle as orc

SLICE = 64


def _case(golden, tag):
    seed, perturb, B, accum = golden[f"{tag}/meta"]
    state = orc.init_state(int(seed), float(perturb))
    x1, x2 = orc.synthetic_views(int(B), seed=int(seed))
    return state, x1, x2, int(accum)


def test_structure_counts():
    # README "11.68 M" / SURVEY §6: 11 681 408 trainable (requires_grad) parameters
    shapes = orc.backbone_param_shapes()
    assert len(shapes) == 200
    per_backbone = sum(int(np.prod(s)) for s in shapes.values())
    assert per_backbone == 5_561_472
    heads = sum(int(np.prod(s)) for s in orc.head_param_shapes().values())
    assert 2 * per_backbone + heads == 11_681_408
    assert len(orc.model_param_names()) == 808
    st = orc.init_state(1)
    assert sum(st[k].numel() for k in orc.trainable_names()) == 11_606_528
    assert len(orc.trainable_names()) == 400


@pytest.mark.parametrize("tag", ["init", "perturbed"])
def test_oracle_matches_reference_golden(golden, tag):
    state, x1, x2, accum = _case(golden, tag)
    loss, pred, tgt, grads = orc.loss_and_grads(dict(state), x1, x2, accum)
    ref_loss = float(golden[f"{tag}/loss"])
    assert abs(loss.item() - ref_loss) <= 1e-5 * abs(ref_loss)          # fp32-check tolerance
    np.testing.assert_allclose(pred.numpy(), golden[f"{tag}/pred"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(tgt.numpy(), golden[f"{tag}/tgt"], rtol=1e-4, atol=2e-6)
    with torch.no_grad():
        hid = orc.backbone_hidden(orc.sub_state(state, "online_network_1"), x1)
    np.testing.assert_allclose(hid.mean(dim=1).numpy(), golden[f"{tag}/feat1"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(hid[:, ::49, :].numpy(), golden[f"{tag}/hidden1_slice"], rtol=1e-4, atol=2e-5)
    names = orc.trainable_names()
    gn = np.array([grads[k].double().norm().item() for k in names])
    np.testing.assert_allclose(gn, golden[f"{tag}/grad_norms"], rtol=2e-4, atol=1e-9)
    gs = golden[f"{tag}/grad_slices"]
    num = den = 0.0
    for i, k in enumerate(names):
        a = grads[k].flatten()[:SLICE].numpy()
        b = gs[i][: len(a)]
        num += float(((a - b) ** 2).sum()); den += float((b ** 2).sum())
    assert np.sqrt(num / den) < 1e-5

    # Adam (lr 1e-4) + EMA (0.999) on top of the same grads → post-step weights
    state2, _ = orc.adam_step(dict(state), grads, {}, lr=1e-4)
    state2 = orc.ema_update(state2, 0.999)
    ps = golden[f"{tag}/post_slices"]
    for i, k in enumerate(orc.model_param_names()):
        a = state2[k].flatten()[:SLICE].numpy()
        # Adam's first step moves each weight by ~lr*sign(g): compare to 1e-6 abs (EMA gate)
        np.testing.assert_allclose(a, ps[i][: len(a)], rtol=0, atol=1e-6, err_msg=k)


def test_loss_sharded_mean_equals_global_mean():
    # the reference loss has no negatives → DP with equal shards is exact
    g = torch.Generator().manual_seed(0)
    p, z = torch.randn(8, 128, generator=g), torch.randn(8, 128, generator=g)
    full = orc.ssp_loss(p, z)
    halves = 0.5 * (orc.ssp_loss(p[:4], z[:4]) + orc.ssp_loss(p[4:], z[4:]))
    assert abs(full.item() - halves.item()) < 1e-7


def test_preprocess_shapes():
    x1, x2 = orc.synthetic_views(2, 0)
    assert x1.shape == (2, 3, 224, 224) and x1.dtype == torch.float32
    assert not torch.equal(x1, x2)
