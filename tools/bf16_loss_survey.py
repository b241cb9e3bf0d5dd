#!/usr/bin/env python
"""How far is a bf16 SSP step from the fp32 reference algorithm, for this build and for stock PyTorch bf16 autocast?

north_star asks for loss rel err <= 1e-3 in bf16.  The loss is a mean of near-zero cosines (|loss| ~ 0.05 at
init), so a relative gate is an absolute 5e-5 on a cosine.  This tool measures, over several weight / data seeds at
B=128, the loss error and the gradient rel-L2 error against the fp32 oracle (run on the GPU with TF32 off) for
 (a) this build's bf16 path and (b) the oracle restatement under torch.autocast(bfloat16) = what the reference's own
PyTorch path does in bf16.   python tools/bf16_loss_survey.py > profiles/bf16_loss_survey_rNN.json"""
import json
import os

os.environ.setdefault("V2S_ALLOW_RANDOM_INIT", "1")   # random-init weights by specification (no checkpoint offline)
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
import vit2spn
from oracle import vit2spn_oracle as orc

dev = torch.device("cuda", 0)
B = 128


def rel_l2(a, b):
    num = sum(float((a[k].float().to(dev) - b[k].float().to(dev)).pow(2).sum()) for k in b if k in a)
    den = sum(float(b[k].float().pow(2).sum()) for k in b if k in a)
    return (num / den) ** 0.5


rows = []
for seed in range(8):
    perturb = 0.02 if seed % 2 else 0.0
    state = orc.init_state(seed, perturb)
    x1, x2 = orc.synthetic_views(B, seed=100 + seed)
    x1, x2 = x1.to(dev), x2.to(dev)
    st = {k: v.to(dev) for k, v in state.items()}
    o_loss, _, _, o_grads = orc.loss_and_grads(dict(st), x1, x2, 1)                    # fp32 on the GPU
    with torch.autocast("cuda", dtype=torch.bfloat16):
        t_loss, _, _, t_grads = orc.loss_and_grads(dict(st), x1, x2, 1)                # stock PyTorch bf16
    model = vit2spn.DualStreamNetwork()
    model.load_state_dict(state, strict=True)
    model.to(dev).train()
    model.projection_head[2].p = 0.0
    vit2spn.set_compute_mode("bf16")
    loss = model.ssp_step(x1, x2, accumulation_steps=1)
    grads = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    o = float(o_loss)
    rows.append(dict(seed=seed, perturb=perturb, fp32_loss=o,
                     ours_abs=abs(float(loss) - o), ours_rel=abs(float(loss) - o) / abs(o),
                     torch_bf16_abs=abs(float(t_loss) - o), torch_bf16_rel=abs(float(t_loss) - o) / abs(o),
                     ours_grad_rel_l2=rel_l2(grads, o_grads), torch_bf16_grad_rel_l2=rel_l2(t_grads, o_grads)))
    del model


def mean(k):
    return sum(r[k] for r in rows) / len(rows)


out = dict(batch=B, rows=rows,
           mean=dict(ours_abs=mean("ours_abs"), torch_bf16_abs=mean("torch_bf16_abs"), ours_rel=mean("ours_rel"),
                     torch_bf16_rel=mean("torch_bf16_rel"), ours_grad_rel_l2=mean("ours_grad_rel_l2"),
                     torch_bf16_grad_rel_l2=mean("torch_bf16_grad_rel_l2")))
print(json.dumps(out, indent=1))
