"""Fused Adam with ``torch.optim.Adam``'s constructor, semantics and ``state_dict`` layout
(ref:ssp_vit2spn_tiny.py:173,216; fine-tune: ref:octmnist_ft_vit2spn.py:192 with L2 weight decay).

Parameters that are views of a vit2spn flat buffer (every parameter of ``DualStreamNetwork`` /
``ViTModel``) are updated by ONE flat-range kernel launch per optimizer step (v2s_adam_step);
any other parameter is updated by the same kernel, one range per tensor.  Tensors whose ``.grad``
is None are skipped, exactly as torch does (final LayerNorm + pooler).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import Range, lib, check, stream_ptr
from .modules import _STORES


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameter")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self._flat_state = {}      # id(store) -> (store, exp_avg_flat, exp_avg_sq_flat)
        self.grad_scale = 1.0      # multiplies every gradient inside the kernel (1/world_size, 1/loss_scale)

    def _state_for(self, p, m_view=None, v_view=None):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = m_view if m_view is not None else torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = v_view if v_view is not None else torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _store_ranges(self, group_params):
        """Split the group's parameters into whole-store flat ranges and leftovers."""
        pset = {id(p) for p in group_params}
        ranges, covered = [], set()
        for store in list(_STORES):
            active = [(p, o) for p, o in zip(store.params, store.offsets)
                      if o < store.active_numel and p.requires_grad]
            if not active or any(id(p) not in pset for p, _ in active):
                continue
            if store.flat is None or not store.flat.is_cuda or not store.grads_attached():
                continue
            if store.active_numel % 4:
                continue
            key = id(store)
            ent = self._flat_state.get(key)
            if ent is None or ent[1].device != store.flat.device:
                m = torch.zeros(store.numel, dtype=torch.float32, device=store.flat.device)
                v = torch.zeros(store.numel, dtype=torch.float32, device=store.flat.device)
                ent = (store, m, v)
                self._flat_state[key] = ent
            _, m, v = ent
            steps = set()
            for p, o in active:
                mv, vv = m[o:o + p.numel()].view(p.shape), v[o:o + p.numel()].view(p.shape)
                st = self._state_for(p, mv, vv)
                if st["exp_avg"].data_ptr() != mv.data_ptr():      # e.g. after load_state_dict
                    mv.copy_(st["exp_avg"]); st["exp_avg"] = mv
                if st["exp_avg_sq"].data_ptr() != vv.data_ptr():
                    vv.copy_(st["exp_avg_sq"]); st["exp_avg_sq"] = vv
                steps.add(float(st["step"]))
            if len(steps) != 1:
                continue                                            # inconsistent history: per-tensor path
            ranges.append((store, m, v, active, int(steps.pop())))
            covered.update(id(p) for p, _ in active)
        return ranges, covered

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            lr, (b1, b2), eps, wd = group["lr"], group["betas"], group["eps"], group["weight_decay"]
            if group.get("amsgrad") or group.get("maximize"):
                raise NotImplementedError("FusedAdam: amsgrad/maximize are not used by the reference")
            with_grad = [p for p in group["params"] if p.grad is not None]
            store_ranges, covered = self._store_ranges(with_grad)
            # group launches by step count (the kernel takes one bias correction per launch)
            launches = {}
            for store, m, v, active, step in store_ranges:
                lp = store.flat_lp
                r = Range(store.flat.data_ptr(), store.flat_grad.data_ptr(), m.data_ptr(), v.data_ptr(),
                          lp.data_ptr() if lp is not None else None, store.active_numel)
                launches.setdefault(step + 1, []).append(r)
                for p, _ in active:
                    self.state[p]["step"] += 1
                if lp is not None and store.active_numel == store.numel or lp is not None and getattr(store, "_lp_tail_ok", False):
                    store.mark_lp_fresh()
                elif lp is not None:
                    # the kernel refreshes only the trained prefix; the never-trained tail (final LN, pooler)
                    # is constant, so one full cast makes every later refresh complete
                    store.lp(refresh=True)
                    store._lp_tail_ok = True
                    store.mark_lp_fresh()
            for p in with_grad:
                if id(p) in covered:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("FusedAdam: parameter on CPU — vit2spn has no CPU fallback")
                if p.grad.is_sparse or p.dtype != torch.float32:
                    raise RuntimeError("FusedAdam supports dense fp32 parameters only")
                st = self._state_for(p)
                g = p.grad.contiguous()
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdam: non-contiguous parameter")
                step = int(float(st["step"]))
                st["step"] += 1
                launches.setdefault(step + 1, []).append(
                    Range(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), None,
                          p.numel()))
                launches.setdefault(("keep", step + 1), []).append(g)
            for key, rs in launches.items():
                if isinstance(key, tuple):
                    continue
                for i in range(0, len(rs), 4):
                    chunk = rs[i:i + 4]
                    arr = (Range * len(chunk))(*chunk)
                    check(lib.v2s_adam_step(arr, len(chunk), int(key), float(lr), float(b1), float(b2), float(eps),
                                            float(wd), float(self.grad_scale), stream_ptr()), "adam_step")
        return loss
