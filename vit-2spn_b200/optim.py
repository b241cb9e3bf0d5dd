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
        self._flat_state = {}      # id(store) -> [store, exp_avg_flat, exp_avg_sq_flat, step]
        self._psets = {}
        self._all_ids = None
        self.grad_scale = 1.0      # multiplies every gradient inside the kernel (1/world_size, 1/loss_scale)

    def _state_for(self, p, m_view=None, v_view=None):
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0, dtype=torch.float32)
            st["exp_avg"] = m_view if m_view is not None else torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = v_view if v_view is not None else torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _store_plan(self, store, pset):
        """(active params, m_flat, v_flat) if the whole store can be updated as one flat range."""
        active = [(p, o) for p, o in zip(store.params, store.offsets) if o < store.active_numel and p.requires_grad]
        if not active or any(id(p) not in pset for p, _ in active):
            return None
        if store.flat is None or not store.flat.is_cuda or store.active_numel % 4:
            return None
        key = id(store)
        ent = self._flat_state.get(key)
        if ent is None or ent[1].device != store.flat.device:
            m = torch.zeros(store.numel, dtype=torch.float32, device=store.flat.device)
            v = torch.zeros(store.numel, dtype=torch.float32, device=store.flat.device)
            ent = [store, m, v, None]          # [.., step count of the store (python int, lazily synced to state)]
            self._flat_state[key] = ent
        _, m, v, step = ent
        if step is None:
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
                return None                                          # inconsistent history: per-tensor path
            ent[3] = int(steps.pop())
        return active, ent

    def _all_param_ids(self):
        n = sum(len(g["params"]) for g in self.param_groups)
        if self._all_ids is None or len(self._all_ids) != n:
            self._all_ids = {id(p) for g in self.param_groups for p in g["params"]}
        return self._all_ids

    def _sync_steps(self):
        """Write the per-store step counters back into the per-parameter state (state_dict layout of torch)."""
        for store, m, v, step in self._flat_state.values():
            if step is None:
                continue
            for p, o in zip(store.params, store.offsets):
                if p in self.state and o < store.active_numel:
                    self.state[p]["step"] = torch.tensor(float(step), dtype=torch.float32)

    def state_dict(self):
        self._sync_steps()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for ent in self._flat_state.values():
            ent[3] = None                       # re-derive the flat views and step counters from the loaded state

    def zero_grad(self, set_to_none: bool = True):
        """Flat-buffer fast path: ONE memset per store, ``.grad`` views stay attached (they read as zeros
        instead of None).  Parameters outside vit2spn stores follow torch's semantics."""
        handled = set()
        mine = self._all_param_ids()
        for store in list(_STORES):
            # only stores whose trainable parameters all belong to this optimizer
            if store.flat_grad is None or not all(id(p) in mine for p in store.params if p.requires_grad):
                continue
            if store.zero_grads_fast():
                handled.update(id(p) for p in store.params)
        for group in self.param_groups:
            for p in group["params"]:
                if id(p) in handled or p.grad is None:
                    continue
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.detach_(); p.grad.zero_()

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
            params = group["params"]
            pset = self._psets.get(id(group))
            if pset is None or len(pset) != len(params):
                pset = {id(p) for p in params}
                self._psets[id(group)] = pset
            launches, covered = {}, set()
            for store in list(_STORES):
                if store.flat_grad is None or not store.grads_attached():
                    continue
                plan = self._store_plan(store, pset)
                if plan is None:
                    continue
                active, ent = plan
                _, m, v, step = ent
                lp = store.flat_lp
                if lp is not None and store.active_numel != store.numel and not getattr(store, "_lp_tail_ok", False):
                    # the kernel refreshes only the trained prefix of the bf16 shadow; the never-trained tail
                    # (final LN, pooler) is constant, so one full cast makes every later refresh complete
                    store.lp(refresh=True)
                    store._lp_tail_ok = True
                launches.setdefault(step + 1, []).append(
                    Range(store.flat.data_ptr(), store.flat_grad.data_ptr(), m.data_ptr(), v.data_ptr(),
                          lp.data_ptr() if lp is not None else None, store.active_numel))
                ent[3] = step + 1
                if lp is not None:
                    store.mark_lp_fresh()
                covered.update(id(p) for p, _ in active)
            keep = []
            if len(covered) != len(params):
                for p in params:
                    if id(p) in covered or p.grad is None:
                        continue
                    if not p.is_cuda:
                        raise RuntimeError("FusedAdam: parameter on CPU — vit2spn has no CPU fallback")
                    if p.grad.is_sparse or p.dtype != torch.float32 or not p.is_contiguous():
                        raise RuntimeError("FusedAdam supports dense contiguous fp32 parameters only")
                    st = self._state_for(p)
                    g = p.grad.contiguous()
                    keep.append(g)
                    step = int(float(st["step"]))
                    st["step"] += 1
                    launches.setdefault(step + 1, []).append(
                        Range(p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), None,
                              p.numel()))
            for key, rs in launches.items():
                for i in range(0, len(rs), 4):
                    chunk = rs[i:i + 4]
                    arr = (Range * len(chunk))(*chunk)
                    check(lib.v2s_adam_step(arr, len(chunk), int(key), float(lr), float(b1), float(b2), float(eps),
                                            float(wd), float(self.grad_scale), stream_ptr()), "adam_step")
        return loss
