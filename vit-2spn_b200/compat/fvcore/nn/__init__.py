"""``FlopCountAnalysis(model, inputs).total()`` (ref:ssp_vit2spn_tiny.py:185-194).  fvcore jit-traces the
model and counts one FLOP per multiply-accumulate of conv / linear / matmul ops; the fused CUDA backbone
cannot be traced, so the count is analytic: 1 253 491 200 MAC per image per ViT-Tiny backbone forward
 plus in_features x out_features per ``nn.Linear`` outside the backbones."""
import torch.nn as nn

VIT_TINY_MAC_PER_IMAGE = 28901376 + 12 * 102049152
_announced = False


def _is_backbone(m):
    return hasattr(m, "_store") and hasattr(m, "config")


def _count(module, batch):
    if _is_backbone(module):
        return VIT_TINY_MAC_PER_IMAGE * batch
    if isinstance(module, nn.Linear):
        return module.in_features * module.out_features * batch
    return sum(_count(c, batch) for c in module.children())


class FlopCountAnalysis:
    def __init__(self, model, inputs):
        self.model = model
        first = inputs[0] if isinstance(inputs, (tuple, list)) else inputs
        self.batch = int(first.shape[0]) if hasattr(first, "shape") and len(first.shape) > 0 else 1

    def total(self):
        global _announced
        if not _announced:
            print("vit2spn: fvcore is not installed; FlopCountAnalysis returns an ANALYTIC count (1 FLOP per MAC of the "
                  "ViT-Tiny backbones and nn.Linear layers), not a traced measurement", flush=True)
            _announced = True
        return int(_count(self.model, self.batch))

    def by_module(self):
        return {name: int(_count(m, self.batch)) for name, m in self.model.named_modules()}
