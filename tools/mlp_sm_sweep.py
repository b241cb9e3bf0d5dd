"""How the chained MLP kernel scales with the number of resident CTAs (v2s_set_sm_limit): per-tile time per CTA from 16 to 148 CTAs,
and SM clock / board power while the kernel loops (pynvml).  Shows how much of a full-chip launch is contention for chip-wide resources
(L2 / memory system, the 1000 W power cap) rather than per-SM work.

    python tools/mlp_sm_sweep.py
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit2spn import _lib as L
L.init_device(0)
dev = torch.device("cuda:0"); dt = torch.bfloat16
rows = 4 * 25216
xn2 = torch.randn(rows, 192, device=dev).to(dt); w1 = (torch.randn(768, 192, device=dev) * 0.05).to(dt); w2 = (torch.randn(192, 768, device=dev) * 0.05).to(dt)
b1 = torch.randn(768, device=dev) * 0.1; b2 = torch.randn(192, device=dev) * 0.1; xmid = torch.randn(rows, 192, device=dev)
gamma, beta = torch.ones(192, device=dev), torch.zeros(192, device=dev)
u = torch.randn(rows, 768, device=dev).to(dt); h = torch.empty(rows, 768, device=dev, dtype=dt)
out = torch.empty(rows, 192, device=dev); xn = torch.empty(rows, 192, device=dev, dtype=dt); mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def fwd(save):
    L.check(L.lib.v2s_test_mlp(0, L.ptr(xn2), L.ptr(w1), L.ptr(w2), L.ptr(b1), L.ptr(b2), L.ptr(u if save else None), L.ptr(h if save else None), L.ptr(xmid), L.ptr(out), L.ptr(xn), L.ptr(gamma), L.ptr(beta), L.ptr(mean), L.ptr(rstd), rows, 0, L.stream_ptr()))
def timeit(fn, n=7):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
tiles = rows // 128
for sms in (148, 111, 74, 37, 16):
    L.check(L.lib.v2s_set_sm_limit(sms if sms < 148 else 0))
    for name, save in (("target", False), ("online", True)):
        us = timeit(lambda: fwd(save))
        per_tile = us / (tiles / sms)      # us per tile per CTA (ignoring the partial last round)
        print(f"{sms:4d} CTAs  fwd {name:6s} {us:8.1f} us   {per_tile:6.2f} us per tile per CTA")
L.lib.v2s_set_sm_limit(0)

# ---- clocks and power while the kernel loops (pynvml), 148 vs 74 CTAs
import threading, time
import pynvml
pynvml.nvmlInit(); hdl = pynvml.nvmlDeviceGetHandleByIndex(0)
for sms in (148, 74):
    L.check(L.lib.v2s_set_sm_limit(sms if sms < 148 else 0))
    for name, save in (("target", False), ("online", True)):
        samples, stop = [], False
        def samp():
            while not stop:
                samples.append((pynvml.nvmlDeviceGetClockInfo(hdl, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(hdl) / 1000.0))
                time.sleep(0.02)
        th = threading.Thread(target=samp); th.start()
        t0 = time.time(); n = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        while time.time() - t0 < 2.0:
            for _ in range(50): fwd(save)
            n += 50; torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        stop = True; th.join()
        s2 = samples[len(samples) // 2:]
        clk = sorted(c for c, _ in s2)[len(s2) // 2]; pw = sorted(p_ for _, p_ in s2)[len(s2) // 2]
        print(f"{sms:4d} CTAs fwd {name:6s}: {e0.elapsed_time(e1) * 1e3 / n:7.1f} us per launch back to back (L2-warm), median SM clock {clk} MHz, power {pw:.0f} W")
L.lib.v2s_set_sm_limit(0)
