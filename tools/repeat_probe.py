"""Run-to-run differences of the bf16 SSP step on fixed inputs (diagnostic)."""
import os
os.environ.setdefault("V2S_ALLOW_RANDOM_INIT", "1")   # random-init weights by specification (no checkpoint offline)
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vit2spn
from oracle import vit2spn_oracle as orc
dev = torch.device("cuda", 0)
state = orc.init_state(3, 0.02)
x1, x2 = orc.synthetic_views(48, seed=5)
x1, x2 = x1.to(dev), x2.to(dev)
model = vit2spn.DualStreamNetwork(); model.load_state_dict(state, strict=True); model.to(dev).train()
model.projection_head[2].p = 0.0
vit2spn.set_compute_mode("bf16")
runs, losses = [], []
for it in range(6):
    for p in model.parameters():
        p.grad = None
    loss = model.ssp_step(x1, x2, accumulation_steps=1)
    losses.append(loss.item())
    runs.append({n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
print("losses", set(losses))
ref = runs[0]
tot_ref = sum(float(v.pow(2).sum()) for v in ref.values()) ** 0.5
for k in range(1, 6):
    d = {n: (runs[k][n] - ref[n]) for n in ref}
    tot = sum(float(v.pow(2).sum()) for v in d.values()) ** 0.5
    worst = sorted(((float(v.norm()) / (float(ref[n].norm()) + 1e-30), float(v.norm()), n) for n, v in d.items()), reverse=True)[:4]
    big = sorted(((float(v.norm()), n) for n, v in d.items()), reverse=True)[:4]
    print(f"run {k}: rel {tot / tot_ref:.3e}; largest abs-diff tensors {[(f'{a:.2e}', n.split('vit.')[-1]) for a, n in big]}")
    print(f"        worst relative {[(f'{r:.2e}', n.split('vit.')[-1]) for r, a, n in worst]}")
