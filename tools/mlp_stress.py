"""Repeats the chained MLP kernel (forward and backward, through the test hook) many times on fixed inputs and
reports every run whose output differs from the first run: any difference is a synchronisation bug (the kernel has no
atomics, so results must be bit-identical run to run).  Prints where the differing elements sit.

    python tools/mlp_stress.py [rows] [iterations]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit2spn import _lib as L  # noqa: E402


def describe(name, a, b):
    bad = (a != b) & ~(torch.isnan(a) & torch.isnan(b))
    if a.dtype.is_floating_point:
        bad |= torch.isnan(a) != torch.isnan(b)
    idx = bad.nonzero()
    if idx.numel() == 0:
        return False
    rows = idx[:, 0]
    cols = idx[:, 1] if idx.shape[1] > 1 else torch.zeros_like(rows)
    print(f"   {name}: {idx.shape[0]} elements differ; rows {int(rows.min())}..{int(rows.max())} (tiles "
          f"{sorted(set((rows // 128).tolist()))[:12]}), cols {int(cols.min())}..{int(cols.max())}; "
          f"nan {int(torch.isnan(a[bad]).sum())}; max |diff| {float((a - b)[bad].abs().max()):.4g} (max |ref| {float(b.abs().max()):.4g})")
    r0 = int(rows[0])
    cs = cols[rows == r0][:8].tolist()
    print(f"      row {r0} (row in tile {r0 % 128}): cols {cs} got {[round(float(a[r0, c]), 4) for c in cs]} first-run {[round(float(b[r0, c]), 4) for c in cs]}")
    print(f"      rows-in-tile affected: {sorted(set((rows % 128).tolist()))[:40]}")
    return True


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 128 * 300
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    dev = torch.device("cuda:0")
    L.init_device(0)
    dt = torch.bfloat16
    g = torch.Generator(device="cpu").manual_seed(1)
    xa = torch.randn(rows, 192, generator=g).to(dev).to(dt)
    w1 = (torch.randn(768, 192, generator=g) * 0.05).to(dev).to(dt)
    w2 = (torch.randn(192, 768, generator=g) * 0.05).to(dev).to(dt)
    b1 = (torch.randn(768, generator=g) * 0.1).to(dev)
    b2 = (torch.randn(192, generator=g) * 0.1).to(dev)
    xmid = torch.randn(rows, 192, generator=g).to(dev)
    gamma, beta = torch.ones(192, device=dev), torch.zeros(192, device=dev)
    uin = torch.randn(rows, 768, generator=g).to(dev).to(dt)
    junk = torch.empty(64 << 20, dtype=torch.uint8, device=dev)

    def run_fwd(save):
        u = torch.full((rows, 768), float("nan"), device=dev, dtype=dt)
        h = torch.full((rows, 768), float("nan"), device=dev, dtype=dt)
        out = torch.full((rows, 192), float("nan"), device=dev)
        xn = torch.full((rows, 192), float("nan"), device=dev, dtype=dt)
        mean, rstd = torch.zeros(rows, device=dev), torch.zeros(rows, device=dev)
        L.check(L.lib.v2s_test_mlp(0, L.ptr(xa), L.ptr(w1), L.ptr(w2), L.ptr(b1), L.ptr(b2), L.ptr(u if save else None),
                                   L.ptr(h if save else None), L.ptr(xmid), L.ptr(out), L.ptr(xn), L.ptr(gamma), L.ptr(beta),
                                   L.ptr(mean), L.ptr(rstd), rows, 0, L.stream_ptr()))
        return {"u": u, "h": h, "out": out, "xn": xn, "mean": mean[:, None], "rstd": rstd[:, None]} if save else \
               {"out": out, "xn": xn, "mean": mean[:, None], "rstd": rstd[:, None]}

    def run_bwd():
        du = torch.full((rows, 768), float("nan"), device=dev, dtype=dt)
        dxn = torch.full((rows, 192), float("nan"), device=dev, dtype=dt)
        L.check(L.lib.v2s_test_mlp(1, L.ptr(xa), L.ptr(w1), L.ptr(w2), None, None, L.ptr(uin), L.ptr(du), None, L.ptr(dxn),
                                   None, None, None, None, None, rows, 0, L.stream_ptr()))
        return {"du": du, "dxn": dxn}

    def explain_bwd(first, later):
        """Which u would reproduce the wrong du values of the first run?"""
        bad = (first["du"] != later["du"]).nonzero()
        if bad.numel() == 0:
            return
        r, c = int(bad[0, 0]), int(bad[0, 1])
        cs = [int(x) for x in bad[bad[:, 0] == r][:, 1][:6]]
        acc = (xa[r].float() @ w2.float())                     # (dx W2)[r, :]
        def gp(x):
            xf = x.float().clone().requires_grad_(True)
            torch.nn.functional.gelu(xf).sum().backward()
            return xf.grad
        print(f"      first-run du[{r}, {cs}] = {[round(float(first['du'][r, k]), 4) for k in cs]}; "
              f"later = {[round(float(later['du'][r, k]), 4) for k in cs]}")
        for name, rr, shift in [("same row, previous chunk", r, -128), ("same row, next chunk", r, 128),
                                ("same row, chunk 0", r, -(cs[0] // 128) * 128), ("next tile of the CTA, chunk 0", r + 148 * 128, -(cs[0] // 128) * 128),
                                ("previous tile of the CTA, last chunk", r - 148 * 128, 640 - (cs[0] // 128) * 128),
                                ("same row, same chunk", r, 0)]:
            if 0 <= rr < rows and all(0 <= k + shift < 768 for k in cs):
                cand = [float((acc[k] * gp(uin[rr, k + shift])).to(dt)) for k in cs]
                print(f"        u from {name:38s}: {[round(x, 4) for x in cand]}")

    for name, fn in (("bwd", run_bwd), ("fwd online", lambda: run_fwd(True)), ("fwd target", lambda: run_fwd(False))):
        ref = fn()
        torch.cuda.synchronize()
        if name == "bwd":
            second = fn()
            torch.cuda.synchronize()
            explain_bwd(ref, second)
            uf = uin.float().requires_grad_(True)
            torch.nn.functional.gelu(uf).sum().backward()
            du_t = ((xa.float() @ w2.float()) * uf.grad).to(dt)
            for tag, res in (("first", ref), ("second", second)):
                err = (res["du"].float() - du_t.float()).abs()
                print(f"   {tag} run vs torch: max |du - ref| = {float(err.max()):.4g} (elements off by > 0.05: {int((err > 0.05).sum())})")
        for k, v in ref.items():
            assert not torch.isnan(v.float()).any(), (name, k, "NaN in the first run")
        fails = 0
        for it in range(iters):
            if it % 3 == 0:
                junk.random_(0, 255)          # vary timing / cache state between runs
            cur = fn()
            torch.cuda.synchronize()
            bad = False
            for k in ref:
                if not torch.equal(ref[k], cur[k]):
                    if not bad:
                        print(f"{name}: run {it} differs")
                    bad = describe(k, cur[k].float(), ref[k].float()) or bad
            fails += bad
        print(f"{name}: {fails} of {iters} runs differ from the first; flag {L.lib.v2s_debug_flag()}")


if __name__ == "__main__":
    main()
