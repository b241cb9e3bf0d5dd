"""197-token attention operator (tcgen05 kernels, `v2s_test_attention`) against a plain torch fp32 reference of
the same op on the same bf16 inputs (HF:modeling_vit.py:220-251 → SDPA, scale 64^-0.5, no mask, dropout 0)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(qkv, dctx):
    B = qkv.shape[0]
    x = qkv.float().requires_grad_(True)
    q, k, v = [t.view(B, 197, 3, 64).transpose(1, 2) for t in x.split(192, dim=-1)]
    s = (q @ k.transpose(-1, -2)) * 0.125
    lse = torch.logsumexp(s, dim=-1)
    o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, 197, 192)
    o.backward(dctx.float())
    return o.detach(), lse.detach(), x.grad


@pytest.mark.parametrize("batch,scale", [(1, 1.0), (3, 1.0), (5, 4.0), (64, 0.5)])
@pytest.mark.parametrize("variant", [0, 1, 2, 3])    # bit 0: SIMT reference kernels instead of tensor-core; bit 1: fp16 tensors
def test_attention_forward_backward(batch, scale, variant):
    import vit2spn  # noqa: F401
    from vit2spn import _lib
    dev = torch.device("cuda", 0)
    _lib.init_device(0)
    g = torch.Generator(device=dev).manual_seed(batch)
    lp = torch.float16 if variant & 2 else torch.bfloat16
    qkv = (torch.randn(batch, 197, 576, device=dev, generator=g) * scale).to(lp)
    dctx = torch.randn(batch, 197, 192, device=dev, generator=g).to(lp)
    o_ref, lse_ref, dqkv_ref = _ref(qkv, dctx)
    ctx = torch.full((batch, 197, 192), float("nan"), device=dev, dtype=lp)
    lse = torch.full((batch, 3, 197), float("nan"), device=dev)
    _lib.check(_lib.lib.v2s_test_attention(0, _lib.ptr(qkv), _lib.ptr(ctx), _lib.ptr(lse), None, None, batch, variant,
                                           _lib.stream_ptr()), "attention fwd")
    torch.cuda.synchronize()
    assert _lib.lib.v2s_debug_flag() == 0
    # bf16 P and bf16 output: 2^-8 relative on values of O(1)
    assert torch.isfinite(ctx.float()).all()
    assert (ctx.float() - o_ref).abs().max().item() <= 0.02 * max(1.0, o_ref.abs().max().item())
    assert (lse - lse_ref).abs().max().item() <= 2e-3 * max(1.0, lse_ref.abs().max().item())
    dqkv = torch.full((batch, 197, 576), float("nan"), device=dev, dtype=lp)
    _lib.check(_lib.lib.v2s_test_attention(1, _lib.ptr(qkv), _lib.ptr(ctx), _lib.ptr(lse), _lib.ptr(dctx), _lib.ptr(dqkv),
                                           batch, variant, _lib.stream_ptr()), "attention bwd")
    torch.cuda.synchronize()
    assert _lib.lib.v2s_debug_flag() == 0
    assert torch.isfinite(dqkv.float()).all()
    for i, name in enumerate(("dq", "dk", "dv")):
        got, ref = dqkv[..., i * 192:(i + 1) * 192].float(), dqkv_ref[..., i * 192:(i + 1) * 192]
        rel = (got - ref).norm().item() / max(ref.norm().item(), 1e-12)
        assert rel <= 2e-2, (name, rel)
        assert (got - ref).abs().max().item() <= 0.03 * max(1.0, ref.abs().max().item()), name
