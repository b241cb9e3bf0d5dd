"""torch-eager baseline on the same B200 (the practical bar) + bf16-autocast
numerics of stock PyTorch against the fp32 oracle (calibrates the bf16 tolerances).

Builds the reference's model the way the reference does (4x transformers.ViTModel + nn.Sequential
heads, ref:ssp_vit2spn_tiny.py:121-160) from the installed transformers — /root/reference is not
needed (it does not exist on the GPU box).  Writes gpurun_out/eager_baseline.json.
"""
import json
import os
import sys
import time

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vit2spn_oracle as orc  # noqa: E402


def build_eager():
    from transformers import ViTConfig, ViTModel

    class Backbone(nn.Module):
        def __init__(self):
            super().__init__()
            self.vit = ViTModel(ViTConfig(hidden_size=192, num_hidden_layers=12, num_attention_heads=3,
                                          intermediate_size=768, patch_size=16, image_size=224,
                                          output_hidden_states=True))

        def forward(self, x):
            return self.vit(x).hidden_states[-1].mean(dim=1)

    class Dual(nn.Module):
        def __init__(self):
            super().__init__()
            self.online_network_1, self.online_network_2 = Backbone(), Backbone()
            self.target_network_1, self.target_network_2 = Backbone(), Backbone()
            for p in list(self.target_network_1.parameters()) + list(self.target_network_2.parameters()):
                p.requires_grad = False
            self.projection_head = nn.Sequential(nn.Linear(384, 1024), nn.ReLU(), nn.Dropout(0.3), nn.Linear(1024, 128))
            self.prediction_head = nn.Sequential(nn.Linear(128, 128), nn.ReLU(), nn.Linear(128, 128))

        def forward(self, x1, x2):
            f1, f2 = self.online_network_1(x1), self.online_network_2(x2)
            with torch.no_grad():
                t1, t2 = self.target_network_1(x1), self.target_network_2(x2)
            p = self.prediction_head(self.projection_head(torch.cat([f1, f2], 1)))
            z = self.projection_head(torch.cat([t1, t2], 1)).detach()
            return p, z

        def update_target_network(self, momentum=0.999):
            for a, b in ((self.online_network_1, self.target_network_1), (self.online_network_2, self.target_network_2)):
                for p, t in zip(a.parameters(), b.parameters()):
                    t.data = momentum * t.data + (1 - momentum) * p.data

    return Dual()


def main():
    dev = torch.device("cuda:0")
    out = {"gpu": torch.cuda.get_device_name(0)}
    crit = nn.CosineSimilarity(dim=1)
    # ---- numerics of stock autocast vs the fp32 oracle (B=4, golden 'init' case) ----
    state = orc.init_state(42, 0.0)
    x1, x2 = orc.synthetic_views(4, seed=42)
    o_loss, _, _, o_grads = orc.loss_and_grads(dict(state), x1, x2, 1)
    model = build_eager()
    model.load_state_dict(state, strict=True)
    model.to(dev).train()
    model.projection_head[2].p = 0.0
    for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)),
                      ("bf16", torch.autocast("cuda", dtype=torch.bfloat16)),
                      ("fp16", torch.autocast("cuda", dtype=torch.float16))):
        model.zero_grad(set_to_none=True)
        with ctx:
            p, z = model(x1.to(dev), x2.to(dev))
            loss = -torch.mean(crit(p, z))
        loss.backward()
        num = den = 0.0
        for n, prm in model.named_parameters():
            if n in o_grads:
                g = prm.grad.detach().float().cpu().double()
                num += float(((g - o_grads[n].double()) ** 2).sum()); den += float((o_grads[n].double() ** 2).sum())
        out[f"numerics_{name}"] = dict(loss=loss.item(), oracle_loss=o_loss.item(),
                                       loss_rel=abs(loss.item() - o_loss.item()) / abs(o_loss.item()),
                                       grad_rel_l2=(num / den) ** 0.5)
        print(name, out[f"numerics_{name}"], flush=True)
    # ---- timing, B=128, full step (fwd+bwd+Adam+EMA), accumulation 1 ----
    B = 128
    xa, xb = [t.to(dev) for t in orc.synthetic_views(B, seed=0)]
    for name, dtype in (("bf16", torch.bfloat16), ("fp16", torch.float16), ("fp32", None)):
        model = build_eager().to(dev).train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)
        scaler = torch.amp.GradScaler("cuda", enabled=(dtype == torch.float16))

        def step():
            opt.zero_grad()
            with torch.autocast("cuda", dtype=dtype or torch.bfloat16, enabled=dtype is not None):
                p, z = model(xa, xb)
                loss = -torch.mean(crit(p, z))
            scaler.scale(loss).backward()
            scaler.step(opt); scaler.update()
            model.update_target_network()
            return loss
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10 if dtype is not None else 5
        e0.record()
        for _ in range(n):
            step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        out[f"eager_{name}_b128"] = dict(ms_per_step=ms, pairs_per_s=B / ms * 1e3)
        print(name, out[f"eager_{name}_b128"], flush=True)
        del model, opt
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "eager_baseline.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
