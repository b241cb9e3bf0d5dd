#!/usr/bin/env python
"""Secondary timings named by SURVEY.md §8(d) / §8(f) that bench.py's single line does not carry.

    python tools/bench_variants.py [--batch 128] [--iters 30] > profiles/variants_rNN.json

* ssp_micro_step      — forward + backward only, gradients accumulated (no optimizer / EMA)
* ssp_accum8          — the reference recipe: 8 micro-steps + 1 Adam + 1 EMA (ref:ssp_vit2spn_tiny.py:205-219)
* ssp_accum1          — accumulation_steps = 1 (bench.py's step)
* adam / ema          — the two optimiser-side kernels alone, with their algorithmic HBM bytes
* single_stream_step  — SingleStreamNetwork (ref:dsn_ssn/ssp_single.py:103-138) full step
* finetune_step       — FineTunedModel(4) fwd + weighted CE + bwd + Adam(L2 1e-4), fp32 and bf16 backbone
                        (ref:octmnist_ft_vit2spn.py:95-104,187-192)
* eval_forward_b1024  — no_grad eval forward at batch 1024 (ref:octmnist_ft_vit2spn.py:129-137)
* augment_*           — input pipeline split (ref:ssp_vit2spn_tiny.py:84-96): GPU finish kernel for 2x128 views, host half per
                        image, and the reference's full CPU transform per image for comparison

All device-timed with CUDA events on the current stream after warm-up, inputs resident in HBM.
"""
import argparse
import json
import os

os.environ.setdefault("V2S_ALLOW_RANDOM_INIT", "1")   # random-init weights by specification (no checkpoint offline)
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.nn as nn

import vit2spn


def timed(fn, iters, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def timed_queued(fn, iters, dev):
    """Device time of a short kernel sequence whose host enqueue cost exceeds its run time: queue the calls
    behind a long blocker kernel so the GPU never waits for the host."""
    blocker = torch.randn(12288, 12288, device=dev)
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        torch.mm(blocker, blocker)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--iters", type=int, default=30)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    B = args.batch
    out = {"batch": B, "device": torch.cuda.get_device_name(0)}
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(2, B, 3, 224, 224, generator=g).to(dev)

    vit2spn.set_compute_mode("bf16")
    torch.manual_seed(42)
    model = vit2spn.DualStreamNetwork().to(dev).train()
    opt = vit2spn.FusedAdam(model.parameters(), lr=1e-4)
    opt.zero_grad()

    def micro(accum):
        return model.ssp_step(x[0], x[1], accumulation_steps=accum)

    def full(accum):
        for _ in range(accum):
            micro(accum)
        opt.step()
        opt.zero_grad()
        model.update_target_network()

    ms = timed(lambda: micro(8), args.iters)
    out["ssp_micro_step"] = {"ms": ms, "pairs_per_s": B / ms * 1e3}
    opt.zero_grad()
    ms = timed(lambda: full(8), max(args.iters // 4, 3), warmup=2)
    out["ssp_accum8"] = {"ms_per_optimizer_step": ms, "pairs_per_s": 8 * B / ms * 1e3}
    ms = timed(lambda: full(1), args.iters)
    out["ssp_accum1"] = {"ms": ms, "pairs_per_s": B / ms * 1e3}

    micro(1)
    ms = timed_queued(opt.step, args.iters, dev)
    nbytes = 11606528 * 28                                # p,g,m,v read + p,m,v written
    out["adam"] = {"ms": ms, "algorithmic_bytes": nbytes, "GBps": nbytes / ms / 1e6}
    ms = timed_queued(model.update_target_network, args.iters, dev)
    nbytes = 11122944 * 12
    out["ema"] = {"ms": ms, "algorithmic_bytes": nbytes, "GBps": nbytes / ms / 1e6}
    del model, opt

    torch.manual_seed(42)
    single = vit2spn.SingleStreamNetwork().to(dev).train()
    sopt = vit2spn.FusedAdam(single.parameters(), lr=1e-4)
    crit = nn.CosineSimilarity(dim=1)

    def single_step():
        p, z = single(x[0], x[1])
        loss = -torch.mean(crit(p, z))
        loss.backward()
        sopt.step()
        sopt.zero_grad()
        single.update_target_network()
    ms = timed(single_step, args.iters)
    out["single_stream_step"] = {"ms": ms, "pairs_per_s": B / ms * 1e3}
    del single, sopt

    labels = torch.randint(0, 4, (B,), generator=g).to(dev)
    weights = torch.tensor([1.0, 2.0, 3.0, 0.5], device=dev)
    for mode in ("fp32", "bf16"):
        vit2spn.set_compute_mode(mode)
        torch.manual_seed(42)
        ft = vit2spn.FineTunedModel(4).to(dev).train()
        fopt = vit2spn.FusedAdam(ft.parameters(), lr=1e-4, weight_decay=1e-4)
        ce = nn.CrossEntropyLoss(weight=weights)

        def ft_step():
            loss = ce(ft(x[0]), labels)
            loss.backward()
            fopt.step()
            fopt.zero_grad()
        ms = timed(ft_step, args.iters)
        out[f"finetune_step_{mode}"] = {"ms": ms, "images_per_s": B / ms * 1e3,
                                        "tflops": B * 7.463e9 / ms / 1e9}
        if mode == "bf16":
            ft.eval()
            xe = torch.randn(1024, 3, 224, 224, generator=g).to(dev)

            def ev():
                with torch.no_grad():
                    return torch.softmax(ft(xe), dim=1)
            ms = timed(ev, 10, warmup=3)
            out["eval_forward_b1024"] = {"ms": ms, "images_per_s": 1024 / ms * 1e3,
                                         "tflops": 1024 * 2.507e9 / ms / 1e9}
        del ft, fopt
    # ---- input pipeline ----
    try:
        import time
        import numpy as np
        from PIL import Image
        from torchvision import transforms
        from vit2spn import augment
        compose = transforms.Compose([
            transforms.Grayscale(num_output_channels=3), transforms.RandomHorizontalFlip(p=0.5),
            transforms.RandomVerticalFlip(p=0.3), transforms.RandomRotation(degrees=30),
            transforms.RandomAffine(degrees=15, translate=(0.1, 0.1), scale=(0.8, 1.2), shear=10),
            transforms.ColorJitter(brightness=0.3, contrast=0.3, saturation=0.3, hue=0.1),
            transforms.Resize((224, 224)), transforms.ToTensor(),
            transforms.GaussianBlur(kernel_size=3, sigma=(0.1, 2.0)),
            transforms.RandomErasing(p=0.5, scale=(0.02, 0.2), ratio=(0.3, 3.3)),
            transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        split = augment.SplitAugment(compose)
        rng = np.random.default_rng(0)
        pil = [Image.fromarray(rng.integers(0, 256, size=(28, 28), dtype=np.uint8), mode="L") for _ in range(64)]
        t0 = time.perf_counter()
        recs = [split(im) for im in pil for _ in range(2)]
        host_ms = (time.perf_counter() - t0) / len(recs) * 1e3
        t0 = time.perf_counter()
        for im in pil[:16]:
            compose(im); compose(im)
        ref_ms = (time.perf_counter() - t0) / 32 * 1e3
        n = 2 * B
        u8 = torch.stack([recs[i % len(recs)][0] for i in range(n)]).to(dev)
        k1d = torch.stack([recs[i % len(recs)][1] for i in range(n)]).to(dev)
        er = torch.stack([recs[i % len(recs)][2] for i in range(n)]).to(dev)
        buf = torch.empty(n, 3, 224, 224, device=dev)
        ms = timed_queued(lambda: augment.finish_views(u8, k1d, er, split.mean, split.std, dev, out=buf), args.iters, dev)
        nbytes = n * 3 * 224 * 224 * 4
        out["augment_finish_gpu"] = {"views": n, "ms": ms, "algorithmic_bytes": nbytes, "GBps": nbytes / ms / 1e6}
        out["augment_host_half_per_view_ms"] = host_ms
        out["augment_reference_cpu_per_view_ms"] = ref_ms
        out["augment_h2d_bytes_per_step"] = {"split": int(n * (784 + 12 + 16)), "reference": int(nbytes)}
        # ---- the real input path end to end: synthetic OCTMNIST-shaped dataset -> DataLoader workers -> SSP step ----
        import importlib.util
        from torch.utils.data import DataLoader
        d = os.path.join(ROOT, "vit-2spn_b200", "compat", "medmnist")
        spec = importlib.util.spec_from_file_location("_v2s_medmnist", os.path.join(d, "__init__.py"),
                                                      submodule_search_locations=[d])
        mm = importlib.util.module_from_spec(spec); sys.modules["_v2s_medmnist"] = mm; spec.loader.exec_module(mm)
        os.environ["V2S_SHIM_DATASET_SIZE"] = str(B * 24)
        workers = min(16, os.cpu_count() or 4)
        vit2spn.set_compute_mode("bf16")
        torch.manual_seed(42)
        model = vit2spn.DualStreamNetwork().to(dev).train()
        opt = vit2spn.FusedAdam(model.parameters(), lr=1e-4)
        opt.zero_grad()

        def run(loader, to_dev):
            n, t0 = 0, None
            for i, (views, _) in enumerate(loader):
                if i == 4:                      # workers warmed up
                    torch.cuda.synchronize(); t0 = time.perf_counter(); n = 0
                v1, v2 = views
                if to_dev:
                    v1, v2 = v1.to(dev, non_blocking=True), v2.to(dev, non_blocking=True)
                model.ssp_step(v1, v2, accumulation_steps=1)
                opt.step(); opt.zero_grad(); model.update_target_network()
                n += v1.shape[0]
            torch.cuda.synchronize()
            return n / (time.perf_counter() - t0)

        class Dual:
            def __init__(self, t): self.t = t
            def __call__(self, x): return self.t(x), self.t(x)

        ours = augment.gpu_dual_view_loader(mm.OCTMNIST(split="train", download=True), compose, batch_size=B, device=dev,
                                            shuffle=True, num_workers=workers, pin_memory=True, persistent_workers=True,
                                            drop_last=True)
        out["real_input_pipeline_gpu_split"] = {"pairs_per_s": run(ours, False), "workers": workers}
        del ours
        ref = DataLoader(mm.OCTMNIST(split="train", transform=Dual(compose), download=True), batch_size=B, shuffle=True,
                         num_workers=workers, pin_memory=True, persistent_workers=True, drop_last=True)
        out["real_input_pipeline_reference_loader"] = {"pairs_per_s": run(ref, True), "workers": workers}
    except Exception as e:      # torchvision / PIL missing
        out["augment"] = f"skipped: {type(e).__name__}: {e}"
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
