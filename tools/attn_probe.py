"""Diagnostic probe for the tcgen05 attention kernels vs the SIMT reference kernels and torch."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vit2spn  # noqa: E402
from vit2spn import _lib  # noqa: E402

dev = torch.device("cuda:0")
_lib.init_device(0)


def torch_ref(qkv, dctx=None):
    B = qkv.shape[0]
    x = qkv.float().requires_grad_(dctx is not None)
    q, k, v = [t.view(B, 197, 3, 64).transpose(1, 2) for t in x.split(192, dim=-1)]
    s = (q @ k.transpose(-1, -2)) * 0.125
    lse = torch.logsumexp(s, dim=-1)
    o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, 197, 192)
    if dctx is None:
        return o, lse, None
    o.backward(dctx.float())
    return o.detach(), lse.detach(), x.grad


def stats(name, got, ref):
    err = (got.float() - ref.float()).abs()
    print(f"  {name}: max_err {err.max().item():.4g} mean_err {err.mean().item():.4g} ref_max {ref.abs().max().item():.3g} "
          f"nan {int(torch.isnan(got.float()).sum())}")
    return err.max().item()


def run(B, scale=1.0, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    qkv = (torch.randn(B, 197, 576, device=dev, generator=g) * scale).bfloat16()
    dctx = torch.randn(B, 197, 192, device=dev, generator=g).bfloat16()
    o_ref, lse_ref, dqkv_ref = torch_ref(qkv, dctx)
    ok = True
    for variant in (1, 0):
        ctx = torch.full((B, 197, 192), float("nan"), device=dev, dtype=torch.bfloat16)
        lse = torch.full((B, 3, 197), float("nan"), device=dev)
        rc = _lib.lib.v2s_test_attention(0, _lib.ptr(qkv), _lib.ptr(ctx), _lib.ptr(lse), None, None, B, variant, _lib.stream_ptr())
        if rc:
            print("fwd error:", _lib.lib.v2s_last_error().decode()); ok = False; continue
        torch.cuda.synchronize()
        flag = _lib.lib.v2s_debug_flag()
        print(f"B={B} scale={scale} fwd variant={variant} flag={flag}")
        e1 = stats("ctx", ctx, o_ref); e2 = stats("lse", lse, lse_ref)
        ok &= flag == 0 and e1 < 0.03 * max(1.0, o_ref.abs().max().item()) and e2 < 2e-2
        if e1 >= 0.03 and variant == 0:
            bad = ((ctx.float() - o_ref).abs() > 0.03) | torch.isnan(ctx.float())
            print("   bad by head:", [int(bad[:, :, h * 64:(h + 1) * 64].sum()) for h in range(3)],
                  " bad rows<128:", int(bad[:, :128].sum()), " rows>=128:", int(bad[:, 128:].sum()),
                  " by image:", [int(bad[i].sum()) for i in range(min(B, 4))])
        dq = torch.full((B, 197, 576), float("nan"), device=dev, dtype=torch.bfloat16)
        rc = _lib.lib.v2s_test_attention(1, _lib.ptr(qkv), _lib.ptr(ctx), _lib.ptr(lse), _lib.ptr(dctx), _lib.ptr(dq), B, variant, _lib.stream_ptr())
        if rc:
            print("  bwd:", _lib.lib.v2s_last_error().decode()); continue
        torch.cuda.synchronize()
        flag = _lib.lib.v2s_debug_flag()
        print(f"B={B} scale={scale} bwd variant={variant} flag={flag}")
        names = ["dq", "dk", "dv"]
        for i in range(3):
            e = stats(names[i], dq[..., i * 192:(i + 1) * 192], dqkv_ref[..., i * 192:(i + 1) * 192])
            ok &= flag == 0 and e < 0.03 * max(1.0, dqkv_ref.abs().max().item())
    return ok


if __name__ == "__main__":
    allok = True
    for B, sc in [(1, 1.0), (3, 1.0), (8, 3.0), (128, 1.0)]:
        allok &= run(B, sc)
    print("ALL OK" if allok else "SOME FAILED")
