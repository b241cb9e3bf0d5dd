"""Per-phase cycle breakdown of the persistent attention forward kernel (CTA 0). V2S_GEMM_DEBUG=1."""
import ctypes as C
import os
import sys
os.environ.setdefault("V2S_GEMM_DEBUG", "1")
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vit2spn  # noqa
from vit2spn import _lib
_lib.init_device(0)
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
qkv = torch.randn(B, 197, 576, device=dev).bfloat16()
ctx = torch.empty(B, 197, 192, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, 3, 197, device=dev)
for _ in range(3):
    _lib.check(_lib.lib.v2s_test_attention(0, _lib.ptr(qkv), _lib.ptr(ctx), _lib.ptr(lse), None, None, B, 0, _lib.stream_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    _lib.lib.v2s_test_attention(0, _lib.ptr(qkv), _lib.ptr(ctx), _lib.ptr(lse), None, None, B, 0, _lib.stream_ptr())
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
buf = (C.c_int64 * 32)()
_lib.check(_lib.lib.v2s_debug_counters(buf))
d = list(buf)
print(f"B={B}: {us:.1f} us per launch; UMMA warp: wait_load {d[0]} wait_tfree {d[1]} wait_p {d[2]} total {d[3]} jobs {d[4]}")
names = ["wait_s", "pass1", "bar_max", "pass2", "bar_sum", "wait_o", "o_read", "stage"]
print("softmax thread (slot 0): " + "  ".join(f"{n} {v}" for n, v in zip(names, d[8:16])))

# ---- backward (one CTA per (b,h)) ----
dctx = torch.randn(B, 197, 192, device=dev).bfloat16()
dqkv = torch.empty(B, 197, 576, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    _lib.check(_lib.lib.v2s_test_attention(1, _lib.ptr(qkv), _lib.ptr(ctx), _lib.ptr(lse), _lib.ptr(dctx), _lib.ptr(dqkv), B, 0, _lib.stream_ptr()))
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    _lib.lib.v2s_test_attention(1, _lib.ptr(qkv), _lib.ptr(ctx), _lib.ptr(lse), _lib.ptr(dctx), _lib.ptr(dqkv), B, 0, _lib.stream_ptr())
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
_lib.check(_lib.lib.v2s_debug_counters(buf))
d = list(buf)
names = ["wait_s", "P", "wait_load", "wait_dp", "dS", "wait_dq", "dQ_stage_store", "wait_kv", "dKdV_drain"]
nj = max(d[27], 1)
print(f"bwd B={B}: {us:.1f} us per launch; thread 32 of CTA 0, cycles per job (both tiles; {nj} jobs): " + "  ".join(f"{n} {v // nj}" for n, v in zip(names, d[16:25])) + f"  D {d[26] // nj}  total {d[25] // nj}")
