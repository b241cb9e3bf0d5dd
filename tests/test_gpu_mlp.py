"""GPU (B200): the chained MLP kernel (mlp_tc.cu: fc1 -> GELU -> fc2 + residual + LayerNorm forward, and the
dgrad chain of the same two layers backward) against a plain PyTorch fp32 reference of the same op on the same
16-bit-rounded operands (HF:modeling_vit.py:296-312,340-346)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

LP = {0: torch.bfloat16, 1: torch.float16}
# relative-to-max tolerances of a 16-bit result / an fp32 result computed from 16-bit operands
TOL = {0: 1.0e-2, 1: 2.0e-3}


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _lib():
    from vit2spn import _lib
    _lib.init_device(0)
    return _lib


@pytest.mark.parametrize("rows", [128 * 3 + 37, 128 * 300, 5])
@pytest.mark.parametrize("lp", [0, 1])
@pytest.mark.parametrize("save", [True, False])
def test_mlp_forward_matches_torch(rows, lp, save):
    L = _lib()
    dev, dt = torch.device("cuda:0"), LP[lp]
    g = torch.Generator(device="cpu").manual_seed(rows + lp)
    xn2 = torch.randn(rows, 192, generator=g).to(dev).to(dt)
    w1 = (torch.randn(768, 192, generator=g) * 0.05).to(dev).to(dt)
    w2 = (torch.randn(192, 768, generator=g) * 0.05).to(dev).to(dt)
    b1 = (torch.randn(768, generator=g) * 0.1).to(dev)
    b2 = (torch.randn(192, generator=g) * 0.1).to(dev)
    xmid = torch.randn(rows, 192, generator=g).to(dev)
    gamma = (1.0 + 0.1 * torch.randn(192, generator=g)).to(dev)
    beta = (0.1 * torch.randn(192, generator=g)).to(dev)
    u = torch.full((rows, 768), float("nan"), device=dev, dtype=dt) if save else None
    h = torch.full((rows, 768), float("nan"), device=dev, dtype=dt) if save else None
    out = torch.full((rows, 192), float("nan"), device=dev)
    xn = torch.full((rows, 192), float("nan"), device=dev, dtype=dt)
    mean = torch.empty(rows, device=dev)
    rstd = torch.empty(rows, device=dev)
    L.check(L.lib.v2s_test_mlp(0, L.ptr(xn2), L.ptr(w1), L.ptr(w2), L.ptr(b1), L.ptr(b2), L.ptr(u), L.ptr(h), L.ptr(xmid),
                               L.ptr(out), L.ptr(xn), L.ptr(gamma), L.ptr(beta), L.ptr(mean), L.ptr(rstd), rows, lp,
                               L.stream_ptr()), "test_mlp")
    torch.cuda.synchronize()
    assert L.lib.v2s_debug_flag() == 0
    u_ref = xn2.float() @ w1.float().t() + b1
    h_ref = torch.nn.functional.gelu(u_ref).to(dt)                       # the kernel feeds the 16-bit h to fc2
    out_ref = xmid + b2 + h_ref.float() @ w2.float().t()
    if save:
        assert _rel(u.float(), u_ref) < TOL[lp]
        assert _rel(h.float(), h_ref.float()) < TOL[lp]
    assert _rel(out, out_ref) < (2e-3 if lp == 0 else 5e-4)
    mu = out.mean(dim=1)
    assert float((mean - mu).abs().max()) < 1e-5
    assert float((rstd - 1.0 / torch.sqrt(out.var(dim=1, unbiased=False) + 1e-12)).abs().max() / rstd.abs().max()) < 1e-4
    xn_ref = torch.nn.functional.layer_norm(out, (192,), gamma, beta, 1e-12)
    assert _rel(xn.float(), xn_ref) < TOL[lp]


@pytest.mark.parametrize("rows", [128 * 2 + 1, 128 * 300])
@pytest.mark.parametrize("lp", [0, 1])
def test_mlp_backward_matches_torch(rows, lp):
    L = _lib()
    dev, dt = torch.device("cuda:0"), LP[lp]
    g = torch.Generator(device="cpu").manual_seed(7 * rows + lp)
    dx = torch.randn(rows, 192, generator=g).to(dev).to(dt)
    w1 = (torch.randn(768, 192, generator=g) * 0.05).to(dev).to(dt)
    w2 = (torch.randn(192, 768, generator=g) * 0.05).to(dev).to(dt)
    u = torch.randn(rows, 768, generator=g).to(dev).to(dt)
    du = torch.full((rows, 768), float("nan"), device=dev, dtype=dt)
    dxn = torch.full((rows, 192), float("nan"), device=dev, dtype=dt)
    runs = []
    for _ in range(3):      # warm launches too: a release/acquire hole in the u-tile ring showed only from the 2nd launch on
        du.fill_(float("nan")); dxn.fill_(float("nan"))
        L.check(L.lib.v2s_test_mlp(1, L.ptr(dx), L.ptr(w1), L.ptr(w2), None, None, L.ptr(u), L.ptr(du), None, L.ptr(dxn),
                                   None, None, None, None, None, rows, lp, L.stream_ptr()), "test_mlp")
        torch.cuda.synchronize()
        runs.append((du.clone(), dxn.clone()))
    assert L.lib.v2s_debug_flag() == 0
    for a, b in runs[1:]:
        assert torch.equal(a, runs[0][0]) and torch.equal(b, runs[0][1]), "the kernel has no atomics: runs must be bit-identical"
    uf = u.float().requires_grad_(True)
    torch.nn.functional.gelu(uf).sum().backward()
    du_ref = ((dx.float() @ w2.float()) * uf.grad).to(dt)
    dxn_ref = du_ref.float() @ w1.float()
    assert _rel(du.float(), du_ref.float()) < TOL[lp]
    assert _rel(dxn.float(), dxn_ref) < TOL[lp]
    # no element may be far off (a stale operand tile corrupts a handful of elements, which a max-relative check over
    # a 30 M-element tensor can miss only if the bound is loose: bound each element by a few 16-bit ulps of its value)
    err = (du.float() - du_ref.float()).abs()
    assert int((err > 0.02 * du_ref.float().abs() + 0.02).sum()) == 0
