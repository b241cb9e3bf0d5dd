"""Times the chained MLP kernel (mlp_tc.cu) in isolation through its test hook: one launch over as many m-tiles
as a B=128 step has (4 backbones x 197 tiles), forward with / without the u,h stores and backward.

    python tools/mlp_timing.py [rows]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit2spn import _lib as L  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 25216
    dev = torch.device("cuda:0")
    L.init_device(0)
    dt = torch.bfloat16
    xn2 = torch.randn(rows, 192, device=dev).to(dt)
    w1 = (torch.randn(768, 192, device=dev) * 0.05).to(dt)
    w2 = (torch.randn(192, 768, device=dev) * 0.05).to(dt)
    b1 = torch.randn(768, device=dev) * 0.1
    b2 = torch.randn(192, device=dev) * 0.1
    xmid = torch.randn(rows, 192, device=dev)
    gamma, beta = torch.ones(192, device=dev), torch.zeros(192, device=dev)
    u = torch.randn(rows, 768, device=dev).to(dt)
    h = torch.empty(rows, 768, device=dev, dtype=dt)
    out = torch.empty(rows, 192, device=dev)
    xn = torch.empty(rows, 192, device=dev, dtype=dt)
    dxn = torch.empty(rows, 192, device=dev, dtype=dt)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def fwd(save, ln=True):
        L.check(L.lib.v2s_test_mlp(0, L.ptr(xn2), L.ptr(w1), L.ptr(w2), L.ptr(b1), L.ptr(b2), L.ptr(u if save else None),
                                   L.ptr(h if save else None), L.ptr(xmid), L.ptr(out), L.ptr(xn if ln else None),
                                   L.ptr(gamma), L.ptr(beta), L.ptr(mean), L.ptr(rstd), rows, 0, L.stream_ptr()))

    def bwd():
        L.check(L.lib.v2s_test_mlp(1, L.ptr(xn2), L.ptr(w1), L.ptr(w2), None, None, L.ptr(u), L.ptr(h), None, L.ptr(dxn),
                                   None, None, None, None, None, rows, 0, L.stream_ptr()))

    def timeit(fn, n=10):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    tiles = (rows + 127) // 128
    waves = -(-tiles // 148)
    for name, fn in (("fwd, target groups (no u/h stores)", lambda: fwd(False)),
                     ("fwd, online groups (u and h stored)", lambda: fwd(True)),
                     ("fwd, no stores, no LayerNorm", lambda: fwd(False, False)),
                     ("bwd", bwd)):
        us = timeit(fn)
        print(f"{name:40s} {us:8.1f} us   {us / waves:6.2f} us per m-tile round ({tiles} tiles, {waves} rounds)  "
              f"{4.0 * rows * 192 * 768 / us / 1e6:7.1f} TFLOP/s")
    print("debug flag", L.lib.v2s_debug_flag())
    if os.environ.get("V2S_GEMM_DEBUG"):
        import ctypes as C
        names_mma = ["acc1_empty", "a1_full", "w_full(S1)", "acc2_empty", "a2_full", "w_full(S2)", "total", "tiles"]
        names_gelu = ["u_full(bwd)", "acc1_full", "a2_free", "u_free(fwd)", "total"]
        names_drain = ["acc2_full", "rs_full", "dbuf_free", "busy (acc2 complete -> released)", "total"]
        for name, fn in (("fwd target", lambda: fwd(False)), ("fwd online", lambda: fwd(True)), ("bwd", bwd)):
            buf = (C.c_int64 * 32)()
            L.lib.v2s_debug_counters(buf)            # clear
            fn(); torch.cuda.synchronize()
            L.check(L.lib.v2s_debug_counters(buf))
            v = list(buf)
            nt = max(v[7], 1)
            print(f"-- {name}: CTA 0, {v[7]} tiles; cycles per tile")
            print("   MMA warp waits:     " + "  ".join(f"{n} {v[i] / nt:.0f}" for i, n in enumerate(names_mma[:7])))
            print("   GELU thread waits:  " + "  ".join(f"{n} {v[8 + i] / nt:.0f}" for i, n in enumerate(names_gelu)))
            print("   drain thread waits: " + "  ".join(f"{n} {v[13 + i] / nt:.0f}" for i, n in enumerate(names_drain)))


if __name__ == "__main__":
    main()
