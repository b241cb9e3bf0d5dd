"""Training-loop API of the reference (ref:ssp_vit2spn_tiny.py:53-72,197-232) on the fused step."""
from __future__ import annotations

import os

import torch
import torch.nn as nn


def _default_mode():
    from .modules import get_compute_mode
    return get_compute_mode()

# reference constants (ref:ssp_vit2spn_tiny.py:35-39)
batch_size = 128
epochs = 100
learning_rate = 1e-4
momentum = 0.999
accumulation_steps = 8


def save_checkpoint(model, optimizer, epoch, loss, path="checkpoint.pth"):
    """Same dictionary keys as ref:ssp_vit2spn_tiny.py:53-61, so checkpoints interchange."""
    checkpoint = {
        "epoch": epoch,
        "model_state_dict": model.state_dict(),
        "optimizer_state_dict": optimizer.state_dict(),
        "loss": loss,
    }
    torch.save(checkpoint, path)
    print(f"Checkpoint saved at epoch {epoch}")


def load_checkpoint(model, optimizer, path="checkpoint.pth", device=None):
    """ref:ssp_vit2spn_tiny.py:63-72 (``strict=False`` model load, returns epoch 0 / inf if absent)."""
    if os.path.exists(path):
        checkpoint = torch.load(path, map_location=device)
        model.load_state_dict(checkpoint["model_state_dict"], strict=False)
        optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
        epoch = checkpoint["epoch"]
        loss = checkpoint["loss"]
        print(f"Checkpoint loaded from epoch {epoch}, loss: {loss}")
        return model, optimizer, epoch, loss
    return model, optimizer, 0, float("inf")


def train_self_supervised(model, dataloader, epochs, optimizer, criterion, checkpoint_path="checkpoint.pth",
                          accumulation_steps=accumulation_steps, device=None, log=print, scaler=None):
    """ref:ssp_vit2spn_tiny.py:197-232.  With the reference's criterion (``nn.CosineSimilarity(dim=1)``)
    the whole micro-step (4 backbones, heads, loss, backward) runs in the CUDA library via
    ``model.ssp_step``; any other criterion goes through the autograd-compatible ``model(x1, x2)``.
    In the fp16 compute mode (the reference's CUDA precision) a ``torch.amp.GradScaler`` is used exactly as ref:175,
    213,216-217 do (``scaler`` argument, created here if omitted): the loss is scaled on the device inside the fused
    step, ``scaler.step(optimizer)`` skips the update on overflow and ``scaler.update()`` adapts the scale; bf16 and the
    fp32 check mode need no scaling.  The per-micro-step ``loss.item()`` host sync of the reference (ref:220) is
    replaced by a device-side accumulation read once per epoch."""
    model, optimizer, start_epoch, _ = load_checkpoint(model, optimizer, checkpoint_path, device)
    device = device or next(model.parameters()).device
    fused = isinstance(criterion, nn.CosineSimilarity) and criterion.dim == 1 and hasattr(model, "ssp_step")
    from .modules import InfoNCELoss
    if isinstance(criterion, InfoNCELoss) and hasattr(model, "ssp_step"):      # opt-in, never the reference recipe
        model.loss_mode, model.temperature, model.nce_group, fused = "infonce", criterion.temperature, criterion.group, True
    elif hasattr(model, "loss_mode"):
        model.loss_mode = "cosine"
    if scaler is None and getattr(model, "compute_mode", None) == "fp16" or \
            (scaler is None and getattr(model, "compute_mode", None) is None and _default_mode() == "fp16"):
        scaler = torch.amp.GradScaler("cuda")                      # ref:175
    use_scaler = scaler is not None and scaler.is_enabled()
    model.train()
    loss_history = []
    for epoch in range(start_epoch, epochs):
        epoch_loss = torch.zeros((), dtype=torch.float32, device=device)
        optimizer.zero_grad()
        n = len(dataloader)
        for i, (views, _) in enumerate(dataloader):
            view1, view2 = views
            view1, view2 = view1.to(device, non_blocking=True), view2.to(device, non_blocking=True)
            if fused:
                loss = model.ssp_step(view1, view2, accumulation_steps, grad_scale=scaler if use_scaler else 1.0)
            else:
                pred, tgt = model(view1, view2)
                loss = (criterion(pred, tgt) if isinstance(criterion, InfoNCELoss) else -torch.mean(criterion(pred, tgt))) / accumulation_steps
                (scaler.scale(loss) if use_scaler else loss).backward()
            if (i + 1) % accumulation_steps == 0 or (i + 1) == n:
                if use_scaler:
                    scaler.step(optimizer)
                    scaler.update()
                else:
                    optimizer.step()
                optimizer.zero_grad()
                model.update_target_network()
            epoch_loss += loss.detach() * accumulation_steps
        avg_epoch_loss = float(epoch_loss.item()) / max(n, 1)
        loss_history.append(avg_epoch_loss)
        log(f"Epoch {epoch + 1}/{epochs}, Loss: {avg_epoch_loss}")
        if (epoch + 1) % 10 == 0:
            save_checkpoint(model, optimizer, epoch + 1, avg_epoch_loss, checkpoint_path)
    return loss_history
