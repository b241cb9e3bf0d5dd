"""CPU oracle for the ViT-2SPN dual-stream self-supervised pretraining (SSP) step.

TEST INFRASTRUCTURE — not a product path (see ``oracle/__init__.py``).

A functional, plain-PyTorch fp32 restatement of the arithmetic the reference executes, written
from the reference's call sites and the (un-vendored, un-pinned) HuggingFace ``transformers``
ViT implementation it calls.  Citations:

* ``ref:``  = /root/reference (mrsaraei/ViT-2SPN)
* ``HF:``   = transformers==5.5.0, ``models/vit/modeling_vit.py``

Parity pinning: the reference ships NO tests, golden vectors or fixtures for this path
(SURVEY.md §4, §8c).  The oracle is therefore pinned against outputs of the reference's own
class definitions (AST-extracted from ``ref:ssp_ssl/ssl_vit2spn_scratch.py`` and
``ref:ssp_vit2spn_tiny.py`` and executed against the installed transformers/torch) by
``tests/golden/make_golden.py``; the resulting vectors are committed under ``tests/golden/`` and
re-checked by ``tests/test_oracle.py`` on every run.

Everything is a pure function of explicit parameter dictionaries keyed by the HuggingFace
``state_dict`` names, so the same dictionaries can be loaded (strict) into the reference modules.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------
# Architecture constants: ViT-Tiny/16 @224  (ref:ssp_ssl/ssl_vit2spn_scratch.py:100-108)
# ---------------------------------------------------------------------------------------------
HIDDEN = 192
LAYERS = 12
HEADS = 3
HEAD_DIM = 64
MLP = 768
PATCH = 16
IMAGE = 224
GRID = IMAGE // PATCH            # 14
TOKENS = GRID * GRID + 1         # 197 (CLS + 196 patches)
LN_EPS = 1e-12                   # HF ViTConfig.layer_norm_eps
PROJ_HIDDEN = 1024               # ref:ssp_vit2spn_tiny.py:133-138
PROJ_OUT = 128
DROPOUT_P = 0.3
COS_EPS = 1e-8                   # torch.nn.CosineSimilarity default (ref:ssp_vit2spn_tiny.py:174)
MOMENTUM = 0.999                 # ref:ssp_vit2spn_tiny.py:38
IMAGENET_MEAN = (0.485, 0.456, 0.406)   # ref:ssp_vit2spn_tiny.py:95
IMAGENET_STD = (0.229, 0.224, 0.225)


def backbone_param_shapes() -> "OrderedDict[str, tuple]":
    """HF ``ViTModel`` parameter names/shapes in registration order (200 tensors).

    Order verified against ``ViTModel(ViTConfig(192,12,3,768,16,224)).named_parameters()``.
    """
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["embeddings.cls_token"] = (1, 1, HIDDEN)
    s["embeddings.position_embeddings"] = (1, TOKENS, HIDDEN)
    s["embeddings.patch_embeddings.projection.weight"] = (HIDDEN, 3, PATCH, PATCH)
    s["embeddings.patch_embeddings.projection.bias"] = (HIDDEN,)
    for l in range(LAYERS):
        p = f"encoder.layer.{l}."
        for n in ("query", "key", "value"):
            s[p + f"attention.attention.{n}.weight"] = (HIDDEN, HIDDEN)
            s[p + f"attention.attention.{n}.bias"] = (HIDDEN,)
        s[p + "attention.output.dense.weight"] = (HIDDEN, HIDDEN)
        s[p + "attention.output.dense.bias"] = (HIDDEN,)
        s[p + "intermediate.dense.weight"] = (MLP, HIDDEN)
        s[p + "intermediate.dense.bias"] = (MLP,)
        s[p + "output.dense.weight"] = (HIDDEN, MLP)
        s[p + "output.dense.bias"] = (HIDDEN,)
        s[p + "layernorm_before.weight"] = (HIDDEN,)
        s[p + "layernorm_before.bias"] = (HIDDEN,)
        s[p + "layernorm_after.weight"] = (HIDDEN,)
        s[p + "layernorm_after.bias"] = (HIDDEN,)
    s["layernorm.weight"] = (HIDDEN,)
    s["layernorm.bias"] = (HIDDEN,)
    s["pooler.dense.weight"] = (HIDDEN, HIDDEN)
    s["pooler.dense.bias"] = (HIDDEN,)
    return s


def head_param_shapes() -> "OrderedDict[str, tuple]":
    """``projection_head`` / ``prediction_head`` (ref:ssp_vit2spn_tiny.py:133-143)."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["projection_head.0.weight"] = (PROJ_HIDDEN, 2 * HIDDEN)
    s["projection_head.0.bias"] = (PROJ_HIDDEN,)
    s["projection_head.3.weight"] = (PROJ_OUT, PROJ_HIDDEN)
    s["projection_head.3.bias"] = (PROJ_OUT,)
    s["prediction_head.0.weight"] = (PROJ_OUT, PROJ_OUT)
    s["prediction_head.0.bias"] = (PROJ_OUT,)
    s["prediction_head.2.weight"] = (PROJ_OUT, PROJ_OUT)
    s["prediction_head.2.bias"] = (PROJ_OUT,)
    return s


BACKBONES = ("online_network_1", "online_network_2", "target_network_1", "target_network_2")


def model_param_names() -> list:
    """The 808 ``DualStreamNetwork.state_dict()`` keys in the reference's registration order
    (ref:ssp_vit2spn_tiny.py:124-143): 4 backbones (prefix ``<net>.vit.``) then the two heads."""
    names = []
    for net in BACKBONES:
        names += [f"{net}.vit.{k}" for k in backbone_param_shapes()]
    names += list(head_param_shapes())
    return names


# ---------------------------------------------------------------------------------------------
# Deterministic, platform-independent weights and inputs (numpy PCG64 → torch)
# ---------------------------------------------------------------------------------------------
def _trunc_normal(rng, shape, std):
    # HF:_init_weights (modeling_vit.py:384-398): trunc_normal_(std=0.02) cut at +-2 (absolute).
    a = rng.standard_normal(size=shape) * std
    return np.clip(a, -2.0, 2.0)


def init_state(seed: int = 42, perturb: float = 0.0) -> "OrderedDict[str, torch.Tensor]":
    """Random-init state for the whole ``DualStreamNetwork`` (808 tensors, fp32).

    Distributionally the same as the reference under random init (the four backbones
    are independently initialised; HF init for the ViT, ``nn.Linear`` default init for the heads).
    ``perturb > 0`` adds N(0, perturb) noise to every tensor (biases, LayerNorm affine, ...) so that
    parity runs also exercise non-trivial biases / LN weights, as a partially trained model has.
    """
    rng = np.random.default_rng(seed)
    st: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for net in BACKBONES:
        for k, shp in backbone_param_shapes().items():
            if k.endswith("layernorm.weight") or k.endswith("layernorm_before.weight") or k.endswith(
                "layernorm_after.weight"
            ):
                a = np.ones(shp)
            elif k.endswith(".bias"):
                a = np.zeros(shp)
            else:
                a = _trunc_normal(rng, shp, 0.02)
            if perturb > 0:
                a = a + rng.standard_normal(size=shp) * perturb
            st[f"{net}.vit.{k}"] = torch.from_numpy(a.astype(np.float32))
    for k, shp in head_param_shapes().items():
        fan_in = head_param_shapes()[k.replace(".bias", ".weight")][1]
        bound = 1.0 / math.sqrt(fan_in)          # nn.Linear default (kaiming_uniform a=sqrt(5))
        a = rng.uniform(-bound, bound, size=shp)
        if perturb > 0:
            a = a + rng.standard_normal(size=shp) * perturb
        st[k] = torch.from_numpy(a.astype(np.float32))
    return st


def synthetic_octmnist_u8(batch: int, seed: int = 0) -> np.ndarray:
    """Seeded raw OCTMNIST-shaped source images: uint8 [B,1,28,28]."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(batch, 1, 28, 28), dtype=np.uint8)


def preprocess_u8(img_u8) -> torch.Tensor:
    """uint8 [B,1,28,28] → fp32 [B,3,224,224]: bilinear 224, replicate 3 ch, ImageNet-normalise.

    The deterministic core of ``strong_augment_transform`` (ref:ssp_vit2spn_tiny.py:84-96:
    Grayscale(3) → Resize(224) → ToTensor → Normalize); the random augmentations are off the
    measured path.
    """
    x = torch.as_tensor(np.asarray(img_u8)).to(torch.float32) / 255.0
    x = F.interpolate(x, size=(IMAGE, IMAGE), mode="bilinear", align_corners=False)
    x = x.expand(-1, 3, -1, -1)
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    return ((x - mean) / std).contiguous()


def synthetic_views(batch: int, seed: int = 0):
    """Two independent synthetic views (x1, x2), fp32 [B,3,224,224]."""
    return (preprocess_u8(synthetic_octmnist_u8(batch, seed)),
            preprocess_u8(synthetic_octmnist_u8(batch, seed + 1)))


# ---------------------------------------------------------------------------------------------
# Optional rounding model of the 16-bit compute modes
# ---------------------------------------------------------------------------------------------
# The fp32 restatement above is the oracle proper.  `rounding(...)` turns it into a MODEL of a 16-bit tensor-core
# implementation: the named intermediate tensors are rounded to bf16 / fp16 (round-to-nearest-even, straight-through
# gradient) exactly where libvit2spn stores them in 16 bits, everything else stays fp32:
#   "w"        GEMM weight operands (the bf16 shadow of the fp32 master weights)
#   "patches"  the im2col patch matrix
#   "xn"       LayerNorm outputs (the A operands of the QKV and fc1 GEMMs)
#   "qkv"      the fused QKV projection output
#   "p"        the un-normalised softmax numerators exp(s - max) fed to P V (the row sum stays fp32)
#   "ctx"      the attention output
#   "h"        gelu(u) (u itself stays fp32 in the forward pass)
# Used (a) to attribute the bf16 loss error to stages (tools/bf16_attribution.py) and (b) as the parity oracle of the
# bf16 / fp16 modes: the CUDA path must match the fully rounded model to the north_star loss tolerance, which
# separates "the format's rounding noise" from "a kernel bug".
ALL_ROUNDING_STAGES = ("w", "patches", "xn", "qkv", "p", "ctx", "h")
_ROUND = {"stages": frozenset(), "dtype": torch.bfloat16}


class rounding:
    """Context manager: ``with rounding(("w", "xn"), torch.bfloat16): ...``; ``rounding("all")`` enables every stage."""

    def __init__(self, stages="all", dtype=torch.bfloat16):
        self.new = {"stages": frozenset(ALL_ROUNDING_STAGES if stages == "all" else stages), "dtype": dtype}
        unknown = self.new["stages"] - set(ALL_ROUNDING_STAGES)
        if unknown:
            raise ValueError(f"unknown rounding stages {sorted(unknown)}")

    def __enter__(self):
        self.old = dict(_ROUND)
        _ROUND.update(self.new)
        return self

    def __exit__(self, *exc):
        _ROUND.update(self.old)
        return False


def _rnd(x, stage):
    """Round `x` to the 16-bit format if `stage` is enabled (identity gradient)."""
    if stage not in _ROUND["stages"]:
        return x
    return x + (x.to(_ROUND["dtype"]).to(x.dtype) - x).detach()


# ---------------------------------------------------------------------------------------------
# Backbone forward  (HF:modeling_vit.py)
# ---------------------------------------------------------------------------------------------
def patch_embed(p, x):
    """HF:153-167 Conv2d(3,192,k16,s16) → flatten(2).transpose(1,2), restated as im2col + GEMM
    (K index = c*256 + ky*16 + kx), then CLS prepend + position add (HF:117-124)."""
    B = x.shape[0]
    cols = x.reshape(B, 3, GRID, PATCH, GRID, PATCH).permute(0, 2, 4, 1, 3, 5).reshape(B, GRID * GRID, 3 * PATCH * PATCH)
    w = _rnd(p["embeddings.patch_embeddings.projection.weight"].reshape(HIDDEN, -1), "w")
    tok = _rnd(cols, "patches") @ w.t() + p["embeddings.patch_embeddings.projection.bias"]
    cls = p["embeddings.cls_token"].expand(B, -1, -1)
    return torch.cat([cls, tok], dim=1) + p["embeddings.position_embeddings"]


def layer_norm(x, w, b):
    """nn.LayerNorm(192, eps=1e-12) (HF:325-326): biased variance.  Stated through F.layer_norm so
    that the oracle can also be run under torch.autocast (bf16 parity oracle)."""
    return F.layer_norm(x, (HIDDEN,), w, b, LN_EPS)


def gelu_erf(x):
    """HF ``hidden_act='gelu'`` → exact erf GELU 0.5 x (1 + erf(x / sqrt 2)) (HF:297-298)."""
    return F.gelu(x)


def attention(q, k, v):
    """softmax(q k^T / sqrt(64)) v, no mask, dropout 0 (HF:220-251; SDPA ≡ eager math)."""
    B, N, _ = q.shape
    q = q.view(B, N, HEADS, HEAD_DIM).transpose(1, 2)
    k = k.view(B, N, HEADS, HEAD_DIM).transpose(1, 2)
    v = v.view(B, N, HEADS, HEAD_DIM).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (HEAD_DIM ** -0.5)
    if "p" in _ROUND["stages"]:
        # the tensor-core kernels feed the rounded numerators to P V and divide by the fp32 row sum afterwards
        e = torch.exp(s - s.max(dim=-1, keepdim=True).values)
        return ((_rnd(e, "p") @ v) / e.sum(dim=-1, keepdim=True)).transpose(1, 2).reshape(B, N, HIDDEN)
    a = torch.softmax(s, dim=-1)
    return (a @ v).transpose(1, 2).reshape(B, N, HIDDEN)


def vit_layer(p, l, h):
    """HF ViTLayer.forward :328-346 (pre-LN block)."""
    pre = f"encoder.layer.{l}."
    x = _rnd(layer_norm(h, p[pre + "layernorm_before.weight"], p[pre + "layernorm_before.bias"]), "xn")
    a = pre + "attention.attention."
    q = _rnd(x @ _rnd(p[a + "query.weight"], "w").t() + p[a + "query.bias"], "qkv")
    k = _rnd(x @ _rnd(p[a + "key.weight"], "w").t() + p[a + "key.bias"], "qkv")
    v = _rnd(x @ _rnd(p[a + "value.weight"], "w").t() + p[a + "value.bias"], "qkv")
    ctx = _rnd(attention(q, k, v), "ctx")
    h = h + ctx @ _rnd(p[pre + "attention.output.dense.weight"], "w").t() + p[pre + "attention.output.dense.bias"]
    x = _rnd(layer_norm(h, p[pre + "layernorm_after.weight"], p[pre + "layernorm_after.bias"]), "xn")
    u = x @ _rnd(p[pre + "intermediate.dense.weight"], "w").t() + p[pre + "intermediate.dense.bias"]
    h = h + _rnd(gelu_erf(u), "h") @ _rnd(p[pre + "output.dense.weight"], "w").t() + p[pre + "output.dense.bias"]
    return h


def backbone_hidden(p, x):
    """``ViTModel(x).hidden_states[-1]``: block-12 output BEFORE the final LayerNorm
    (HF:426 ``tie_last_hidden_states=False``).  Final LN + pooler are discarded by the
    reference and are not computed here."""
    h = patch_embed(p, x)
    for l in range(LAYERS):
        h = vit_layer(p, l, h)
    return h


def backbone_features(p, x):
    """``ViTBackbone.forward`` (ref:ssp_vit2spn_tiny.py:114-118): mean over all 197 tokens."""
    return backbone_hidden(p, x).mean(dim=1)


def sub_state(state, net):
    pre = f"{net}.vit."
    return {k[len(pre):]: v for k, v in state.items() if k.startswith(pre)}


# ---------------------------------------------------------------------------------------------
# Heads, loss, optimiser, EMA
# ---------------------------------------------------------------------------------------------
def projection_head(state, f, mask=None):
    """Linear(384,1024)-ReLU-Dropout(0.3)-Linear(1024,128) (ref:133-138).  ``mask`` is the
    already-scaled dropout multiplier (keep/(1-p)) or None for eval / neutralised dropout (D11)."""
    y = torch.relu(f @ state["projection_head.0.weight"].t() + state["projection_head.0.bias"])
    if mask is not None:
        y = y * mask
    return y @ state["projection_head.3.weight"].t() + state["projection_head.3.bias"]


def prediction_head(state, z):
    """Linear(128,128)-ReLU-Linear(128,128) (ref:139-143)."""
    y = torch.relu(z @ state["prediction_head.0.weight"].t() + state["prediction_head.0.bias"])
    return y @ state["prediction_head.2.weight"].t() + state["prediction_head.2.bias"]


def dual_stream_forward(state, x1, x2, mask_online=None, mask_target=None):
    """``DualStreamNetwork.forward`` (ref:ssp_vit2spn_tiny.py:145-160) → (pred, target_proj)."""
    f1 = backbone_features(sub_state(state, "online_network_1"), x1)
    f2 = backbone_features(sub_state(state, "online_network_2"), x2)
    with torch.no_grad():
        t1 = backbone_features(sub_state(state, "target_network_1"), x1)
        t2 = backbone_features(sub_state(state, "target_network_2"), x2)
    pred = prediction_head(state, projection_head(state, torch.cat([f1, f2], dim=1), mask_online))
    with torch.no_grad():
        tgt = projection_head(state, torch.cat([t1, t2], dim=1), mask_target)
    return pred, tgt.detach()


def ssp_loss(pred, tgt, accumulation_steps: int = 1):
    """``-mean(CosineSimilarity(dim=1, eps=1e-8)(p, z)) / accumulation_steps``
    (ref:ssp_vit2spn_tiny.py:174,211).  torch clamps each norm separately."""
    pn = pred.norm(dim=1).clamp_min(COS_EPS)
    zn = tgt.norm(dim=1).clamp_min(COS_EPS)
    cos = (pred * tgt).sum(dim=1) / (pn * zn)
    return -cos.mean() / accumulation_steps


def infonce_loss(pred, keys, label_offset: int = 0, temperature: float = 0.2, accumulation_steps: int = 1):
    """InfoNCE over gathered target projections — BASELINE north_star (3) / config 3.  **No reference counterpart**
    (the reference loss is the negative-free cosine of ref:174,211), so this restatement IS the
    definition the CUDA kernel (kernels.cu infonce_row_kernel) is checked against: parity unpinned by the reference.
    ``logits[i][j] = cos(pred_i, keys_j) / temperature`` with each norm clamped at 1e-8 like
    ``nn.CosineSimilarity`` (ref:174); the positive of local row i is key ``label_offset + i``; keys are detached
    (ref:158)."""
    eps = 1e-8
    ph = pred / pred.norm(dim=1, keepdim=True).clamp_min(eps)
    kh = keys.detach() / keys.detach().norm(dim=1, keepdim=True).clamp_min(eps)
    logits = ph @ kh.t() / temperature
    labels = torch.arange(pred.shape[0], device=pred.device) + label_offset
    return torch.nn.functional.cross_entropy(logits, labels) / accumulation_steps


def trainable_names():
    """Parameters that receive a gradient: both online backbones minus final LN + pooler
    (never used, grad None — SURVEY D6), plus the two heads.  400 tensors, 11 606 528 elements."""
    names = []
    for net in BACKBONES[:2]:
        for k in backbone_param_shapes():
            if k.startswith("layernorm.") or k.startswith("pooler."):
                continue
            names.append(f"{net}.vit.{k}")
    return names + list(head_param_shapes())


def loss_and_grads(state, x1, x2, accumulation_steps=1, mask_online=None, mask_target=None):
    """One micro-step (ref:209-213): returns (loss, pred, tgt, {name: grad})."""
    names = trainable_names()
    leaves = {k: state[k].detach().clone().requires_grad_(True) for k in names}
    st = dict(state)
    st.update(leaves)
    pred, tgt = dual_stream_forward(st, x1, x2, mask_online, mask_target)
    loss = ssp_loss(pred, tgt, accumulation_steps)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names])
    return loss.detach(), pred.detach(), tgt.detach(), dict(zip(names, grads))


def adam_step(state, grads, opt, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
    """``torch.optim.Adam(lr=1e-4)`` defaults (ref:173,216): no weight decay, no amsgrad,
    bias-corrected; tensors without a gradient are skipped.  ``opt`` = {name: (step, m, v)}."""
    b1, b2 = betas
    for k, g in grads.items():
        step, m, v = opt.get(k, (0, torch.zeros_like(g), torch.zeros_like(g)))
        step += 1
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        bc1 = 1 - b1 ** step
        bc2 = 1 - b2 ** step
        denom = v.sqrt() / math.sqrt(bc2) + eps
        state[k] = state[k] - (lr / bc1) * (m / denom)
        opt[k] = (step, m, v)
    return state, opt


def ema_update(state, momentum=MOMENTUM):
    """``update_target_network`` (ref:162-166): every backbone tensor, incl. final LN/pooler."""
    for o, t in (("online_network_1", "target_network_1"), ("online_network_2", "target_network_2")):
        for k in backbone_param_shapes():
            state[f"{t}.vit.{k}"] = momentum * state[f"{t}.vit.{k}"] + (1 - momentum) * state[f"{o}.vit.{k}"]
    return state


def ssp_step(state, opt, x1, x2, lr=1e-4, momentum=MOMENTUM, mask_online=None, mask_target=None):
    """Full step with accumulation_steps=1: fwd → loss → bwd → Adam → EMA."""
    loss, pred, tgt, grads = loss_and_grads(state, x1, x2, 1, mask_online, mask_target)
    state, opt = adam_step(state, grads, opt, lr=lr)
    state = ema_update(state, momentum)
    return loss, grads, state, opt


def train_loop(state, batches, epochs, accumulation_steps=8, lr=1e-4, momentum=MOMENTUM):
    """``train_self_supervised`` (ref:ssp_vit2spn_tiny.py:197-232) without the AMP scaler (disabled on CPU, ref:175) and
    the checkpoint I/O: per epoch ``zero_grad``; per micro-batch ``loss = -mean(cos)/accumulation_steps`` accumulated
    into the gradients (ref:211-213); on every ``accumulation_steps``-th micro-batch AND on the last one of the epoch
    (ref:215) Adam step, ``zero_grad``, EMA update; ``epoch_loss += loss.item() * accumulation_steps`` (ref:220) and the
    logged value is ``epoch_loss / len(dataloader)`` (ref:226).  ``batches`` = list of (view1, view2).
    Returns (loss_history, final_state, optimizer_state)."""
    opt, history = {}, []
    names = trainable_names()
    for _ in range(epochs):
        acc = {k: torch.zeros_like(state[k]) for k in names}
        epoch_loss = 0.0
        for i, (v1, v2) in enumerate(batches):
            loss, _, _, grads = loss_and_grads(state, v1, v2, accumulation_steps)
            for k in names:
                acc[k] += grads[k]
            if (i + 1) % accumulation_steps == 0 or (i + 1) == len(batches):
                state, opt = adam_step(state, acc, opt, lr=lr)
                acc = {k: torch.zeros_like(state[k]) for k in names}
                state = ema_update(state, momentum)
            epoch_loss += float(loss) * accumulation_steps
        history.append(epoch_loss / len(batches))
    return history, state, opt
