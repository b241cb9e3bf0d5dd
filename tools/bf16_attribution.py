"""Attributes the 16-bit loss error of the SSP step to the stages that round.

The fp32 oracle is re-run with ONE rounding stage of oracle.rounding enabled at a time (and with all of them), at
BASELINE config 2's batch (128) and seed 42; the table reports the relative loss error each stage causes on its own.
Runs on the CPU (about a minute) or on a GPU (`--device cuda`); with a GPU and libvit2spn present it also reports
this build and stock torch autocast against the same fp32 value.

    python tools/bf16_attribution.py [--batch 128] [--seed 42] [--dtype bf16|fp16] [--device cpu|cuda] [--out file.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vit2spn_oracle as orc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dt = torch.bfloat16 if a.dtype == "bf16" else torch.float16
    dev = torch.device(a.device)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_num_threads(os.cpu_count() or 8)
    state = {k: v.to(dev) for k, v in orc.init_state(a.seed, 0.0).items()}
    x1, x2 = (t.to(dev) for t in orc.synthetic_views(a.batch, seed=a.seed))

    def loss_with(stages):
        with torch.no_grad(), orc.rounding(stages, dt):
            p, t = orc.dual_stream_forward(state, x1, x2)
            return float(orc.ssp_loss(p, t))

    ref = loss_with(())
    rows = {}
    for st in orc.ALL_ROUNDING_STAGES:
        rows[st] = loss_with((st,))
    rows["all"] = loss_with("all")
    rows["all but w"] = loss_with(tuple(s for s in orc.ALL_ROUNDING_STAGES if s != "w"))
    out = {"batch": a.batch, "seed": a.seed, "dtype": a.dtype, "fp32_loss": ref,
           "stages": {k: {"loss": v, "rel_err": abs(v - ref) / abs(ref), "signed_rel": (v - ref) / abs(ref)} for k, v in rows.items()}}
    if dev.type == "cuda":
        with torch.no_grad(), torch.autocast("cuda", dtype=dt):
            p, t = orc.dual_stream_forward(state, x1, x2)
            ac = float(orc.ssp_loss(p.float(), t.float()))
        out["torch_autocast"] = {"loss": ac, "rel_err": abs(ac - ref) / abs(ref), "signed_rel": (ac - ref) / abs(ref)}
        try:
            os.environ.setdefault("V2S_ALLOW_RANDOM_INIT", "1")
            import vit2spn
            model = vit2spn.DualStreamNetwork()
            model.load_state_dict({k: v.cpu() for k, v in state.items()}, strict=True)
            model.to(dev).train()
            model.projection_head[2].p = 0.0
            model.compute_mode = a.dtype
            lv = float(model.ssp_step(x1, x2, accumulation_steps=1, with_backward=False))
            out["libvit2spn"] = {"loss": lv, "rel_err": abs(lv - ref) / abs(ref), "signed_rel": (lv - ref) / abs(ref),
                                 "rel_err_vs_rounding_model": abs(lv - rows["all"]) / abs(rows["all"])}
        except Exception as e:  # noqa: BLE001
            out["libvit2spn"] = {"error": str(e)}
    print(f"fp32 loss {ref:.8f}  (B={a.batch}, seed {a.seed}, {a.dtype})")
    for k, v in out["stages"].items():
        print(f"  round {k:10s}: loss {v['loss']:.8f}  rel err {v['rel_err']:.2e}  ({v['signed_rel']:+.2e})")
    for k in ("torch_autocast", "libvit2spn"):
        if k in out:
            print(f"  {k:16s}: {out[k]}")
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
