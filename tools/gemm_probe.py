"""Diagnostic probe for the tcgen05 GEMM (run on the B200): compares the tensor-core path with
torch fp32 matmul of the same bf16 operands for the three operand-major combinations and prints
where errors sit (rows / column blocks), so that descriptor mistakes can be localised in one run."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vit2spn  # noqa: E402
from vit2spn import _lib  # noqa: E402


def run(which, m, n, k, variant=0, seed=0):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(seed)
    if which == 0:
        a = torch.randn(m, k, device=dev, generator=g).bfloat16(); b = torch.randn(n, k, device=dev, generator=g).bfloat16()
        ref = a.float() @ b.float().t()
        c = torch.full((m, n), float("nan"), device=dev, dtype=torch.bfloat16)
    elif which == 1:
        a = torch.randn(m, k, device=dev, generator=g).bfloat16(); b = torch.randn(k, n, device=dev, generator=g).bfloat16()
        ref = a.float() @ b.float()
        c = torch.full((m, n), float("nan"), device=dev, dtype=torch.bfloat16)
    else:
        a = torch.randn(k, m, device=dev, generator=g).bfloat16(); b = torch.randn(k, n, device=dev, generator=g).bfloat16()
        ref = a.float().t() @ b.float()
        c = torch.zeros(m, n, device=dev, dtype=torch.float32)
    rc = _lib.lib.v2s_test_gemm(which, _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), m, n, k, variant, _lib.stream_ptr())
    if rc:
        print(f"which={which} {m}x{n}x{k} variant={variant}: ERROR {_lib.lib.v2s_last_error().decode()}")
        return False
    torch.cuda.synchronize()
    flag = _lib.lib.v2s_debug_flag()
    out = c.float()
    err = (out - ref).abs()
    tol = 0.02 * ref.abs().max().item() + 0.05
    bad = ~(err <= tol)      # catches NaN
    ok = not bool(bad.any()) and flag == 0
    print(f"which={which} {m}x{n}x{k} variant={variant}: flag={flag} max_err={err.nan_to_num(1e9).max().item():.4g} "
          f"ref_max={ref.abs().max().item():.3g} bad={int(bad.sum())}/{bad.numel()} {'OK' if ok else 'FAIL'}")
    if not ok and bad.any():
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        print(f"   bad rows: n={len(rows)} first={rows[:8].tolist()} last={rows[-4:].tolist()}  "
              f"bad cols: n={len(cols)} first={cols[:8].tolist()} last={cols[-4:].tolist()}")
        print("   bad fraction per 32-col block:", [round(float(bad[:, j:j + 32].float().mean()), 2) for j in range(0, n, 32)][:24])
        print("   bad fraction per 8-row block (first 16):", [round(float(bad[i:i + 8].float().mean()), 2) for i in range(0, min(m, 128), 8)])
        print("   nan count:", int(torch.isnan(out).sum()), " sample out/ref:", out[0, :4].tolist(), ref[0, :4].tolist())
    return ok


if __name__ == "__main__":
    _lib.init_device(0)
    allok = True
    # SIMT reference path first (sanity of the hook), then tensor-core path
    allok &= run(0, 256, 192, 192, variant=1)
    for which, m, n, k in [(0, 128, 192, 64), (0, 128, 192, 192), (0, 256, 576, 192), (0, 1576, 768, 192),
                           (0, 300, 192, 768), (0, 25216, 576, 192),
                           (1, 128, 192, 64), (1, 256, 192, 576), (1, 1576, 768, 192), (1, 1576, 192, 768),
                           (2, 128, 192, 64), (2, 192, 192, 256), (2, 576, 192, 1576), (2, 192, 768, 1576),
                           (2, 768, 192, 25216)]:
        allok &= run(which, m, n, k)
    print("ALL OK" if allok else "SOME FAILED")
