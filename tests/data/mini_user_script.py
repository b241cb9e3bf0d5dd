"""A small user script in the style of the reference's pretraining scripts (own wording, test fixture):
it imports what they import, builds its *own* two-online / two-target model around ``transformers.ViTModel``,
trains a few micro-steps with autocast + GradScaler + gradient accumulation, updates the targets by rebinding
``.data``, counts FLOPs, plots, and saves the first online backbone.  Run through ``python -m vit2spn.run``."""
import os
import sys

import torch
import torch.nn as nn
from torch.utils.data import DataLoader
from torchvision import transforms
from medmnist.dataset import OCTMNIST
import matplotlib.pyplot as plt
from transformers import ViTModel
from fvcore.nn import FlopCountAnalysis

dev = torch.device("cuda")
torch.manual_seed(7)
ACCUM, EMA_M = 2, 0.9


class TwoViews:
    def __init__(self, t):
        self.t = t

    def __call__(self, img):
        return self.t(img), self.t(img)


aug = transforms.Compose([
    transforms.Grayscale(num_output_channels=3),
    transforms.RandomHorizontalFlip(),
    transforms.Resize((224, 224)),
    transforms.ToTensor(),
    transforms.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225]),
])
loader = DataLoader(OCTMNIST(split="train", transform=TwoViews(aug), download=True), batch_size=4, shuffle=True,
                    num_workers=0)


class Trunk(nn.Module):
    def __init__(self):
        super().__init__()
        self.vit = ViTModel.from_pretrained("WinKawaks/vit-tiny-patch16-224", output_hidden_states=True)

    def forward(self, x):
        return self.vit(x).hidden_states[-1].mean(dim=1)


class TwoStream(nn.Module):
    def __init__(self):
        super().__init__()
        self.on_a, self.on_b, self.tg_a, self.tg_b = Trunk(), Trunk(), Trunk(), Trunk()
        for p in list(self.tg_a.parameters()) + list(self.tg_b.parameters()):
            p.requires_grad = False
        self.proj = nn.Sequential(nn.Linear(384, 256), nn.ReLU(), nn.Linear(256, 64))
        self.pred = nn.Sequential(nn.Linear(64, 64), nn.ReLU(), nn.Linear(64, 64))

    def forward(self, a, b):
        f = torch.cat([self.on_a(a), self.on_b(b)], dim=1)
        with torch.no_grad():
            t = torch.cat([self.tg_a(a), self.tg_b(b)], dim=1)
        return self.pred(self.proj(f)), self.proj(t).detach()

    def ema(self):
        for on, tg in ((self.on_a, self.tg_a), (self.on_b, self.tg_b)):
            for po, pt in zip(on.parameters(), tg.parameters()):
                pt.data = EMA_M * pt.data + (1 - EMA_M) * po.data


net = TwoStream().to(dev)
opt = torch.optim.Adam(net.parameters(), lr=1e-4)
cos = nn.CosineSimilarity(dim=1)
scaler = torch.amp.GradScaler("cuda")
gflops = FlopCountAnalysis(net, (torch.randn(1, 3, 224, 224, device=dev),) * 2).total() / 1e9
print(f"FLOPs per pair: {gflops:.3f} G")
assert 4.9 < gflops < 5.2, gflops        # 4 backbone forwards of 1.2535 GMAC

tg_before = net.tg_a.vit.embeddings.cls_token.detach().clone()
on_before = net.on_a.vit.encoder.layer[3].intermediate.dense.weight.detach().clone()
history = []
net.train()
for epoch in range(2):
    opt.zero_grad()
    for i, (views, _) in enumerate(loader):
        a, b = views[0].to(dev), views[1].to(dev)
        with torch.autocast("cuda"):
            p, z = net(a, b)
            loss = -cos(p, z).mean() / ACCUM
        scaler.scale(loss).backward()
        if (i + 1) % ACCUM == 0 or i + 1 == len(loader):
            scaler.step(opt)
            scaler.update()
            opt.zero_grad()
            net.ema()
        history.append(loss.item() * ACCUM)
        if (i + 1) % 2 == 0:
            torch.cuda.empty_cache()
assert all(map(lambda v: v == v and abs(v) <= 1.0, history)), history
assert not torch.equal(on_before, net.on_a.vit.encoder.layer[3].intermediate.dense.weight), "online did not train"
assert not torch.equal(tg_before, net.tg_a.vit.embeddings.cls_token), "target did not move"
plt.figure(); plt.plot(history); plt.xlabel("it"); plt.savefig("loss.png"); plt.close()
out = sys.argv[1] if len(sys.argv) > 1 else "mini_backbone.pth"
torch.save(net.on_a.state_dict(), out)
print("MINI_OK", len(history), history[-1])
