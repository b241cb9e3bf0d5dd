"""GPU-side input pipeline for the reference's augmentation (ref:ssp_vit2spn_tiny.py:74-107).

The reference builds every 224x224x3 fp32 view on CPU workers and ships 154 MB per step to the GPU.  Here the
`transforms.Compose` is split at its `Resize`:

* everything before it (Grayscale, flips, rotation, affine, colour jitter) runs unchanged - the very same
  torchvision / PIL objects, drawing from the same torch RNG stream - on the 28x28 PIL image (784 pixels);
* the parameters of the transforms after it are drawn on the host in the reference's order (GaussianBlur sigma,
  RandomErasing rectangle) and travel with the 784-byte view;
* `Resize -> ToTensor -> GaussianBlur -> RandomErasing -> Normalize` runs on the GPU (`v2s_augment_finish_u8`):
  Pillow-exact bilinear resize, so with the same seed the views equal the reference's to fp32 rounding.

    loader = gpu_dual_view_loader(dataset, strong_augment_transform, batch_size=128, device="cuda")
    for (view1, view2), labels in loader: ...            # same contract as the reference's DataLoader (ref:205-207)
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib

OUT = 224
_PRECISION_BITS = 32 - 8 - 2


def pil_bilinear_tables(in_size: int, out_size: int = OUT):
    """Coefficient tables of Pillow's BILINEAR resize for 8-bit images (Resample.c precompute_coeffs /
    normalize_coeffs_8bpc): bounds [out,2] (first source index, taps), coefs [out,ksize] in 22-bit fixed point."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    coefs = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [max(1.0 - abs((x + xmin - center + 0.5) / filterscale), 0.0) for x in range(xmax)]
        ww = sum(w)
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            coefs[xx, x] = int(0.5 + k * (1 << _PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, coefs


class SplitAugment:
    """Host half of one view.  ``__call__(pil_image)`` -> (uint8 [s,s], k1d fp32 [3], erase int32 [4])."""

    def __init__(self, compose):
        from torchvision import transforms as T
        ts = list(compose.transforms)
        idx = [i for i, t in enumerate(ts) if isinstance(t, T.Resize)]
        if len(idx) != 1:
            raise NotImplementedError("SplitAugment: the Compose must contain exactly one Resize")
        self.head, tail = ts[:idx[0]], ts[idx[0]:]
        size = tail[0].size
        size = (size, size) if isinstance(size, int) else tuple(size)
        if size != (OUT, OUT):
            raise NotImplementedError(f"SplitAugment: Resize{size} (the CUDA backbone takes 224x224)")
        if tail[0].interpolation not in (T.InterpolationMode.BILINEAR,):
            raise NotImplementedError("SplitAugment: only BILINEAR Resize")
        self.blur = self.erasing = self.normalize = None
        seen_tensor = False
        for t in tail[1:]:
            if isinstance(t, T.ToTensor):
                seen_tensor = True
            elif isinstance(t, T.GaussianBlur) and self.erasing is None and self.normalize is None:
                if tuple(t.kernel_size) != (3, 3):
                    raise NotImplementedError("SplitAugment: GaussianBlur kernel_size must be 3")
                self.blur = t
            elif isinstance(t, T.RandomErasing) and self.normalize is None:
                if t.value != 0 or t.inplace:
                    raise NotImplementedError("SplitAugment: RandomErasing(value=0, inplace=False) only")
                self.erasing = t
            elif isinstance(t, T.Normalize):
                self.normalize = t
            else:
                raise NotImplementedError(f"SplitAugment: unsupported transform after Resize: {t}")
        if not seen_tensor:
            raise NotImplementedError("SplitAugment: expected ToTensor after Resize")
        self.tail_order = [type(t).__name__ for t in tail]
        self.mean = tuple(float(v) for v in (self.normalize.mean if self.normalize else (0.0, 0.0, 0.0)))
        self.std = tuple(float(v) for v in (self.normalize.std if self.normalize else (1.0, 1.0, 1.0)))
        self._meta = torch.empty(3, OUT, OUT, device="meta")

    def __call__(self, img):
        from torchvision.transforms import _functional_tensor as FT
        for t in self.head:                       # the reference's own objects on the small PIL image
            img = t(img)
        a = np.asarray(img)
        if a.ndim == 3:
            if not (np.array_equal(a[..., 0], a[..., 1]) and np.array_equal(a[..., 0], a[..., 2])):
                raise NotImplementedError("SplitAugment: the image is not grey after the PIL-side transforms")
            a = a[..., 0]
        a = np.ascontiguousarray(a, dtype=np.uint8)
        k1d = np.array([0.0, 1.0, 0.0], np.float32)
        erase = np.zeros(4, np.int32)
        if self.blur is not None:                  # same draw as GaussianBlur.forward
            sigma = self.blur.get_params(self.blur.sigma[0], self.blur.sigma[1])
            k1d = FT._get_gaussian_kernel1d(3, sigma, torch.float32, torch.device("cpu")).numpy()
        if self.erasing is not None:               # same draws as RandomErasing.forward
            t = self.erasing
            if torch.rand(1) < t.p:
                i, j, h, w, _ = t.get_params(self._meta, scale=t.scale, ratio=t.ratio, value=[float(t.value)])
                if (h, w) != (OUT, OUT) or (i, j) != (0, 0):      # get_params' "no rectangle found" return
                    erase = np.array([i, j, h, w], np.int32)
        return torch.from_numpy(a), torch.from_numpy(k1d), torch.from_numpy(erase)


class DualViewSplit:
    """ref:ssp_vit2spn_tiny.py:74-82 — two independent augmentations of one image, host half only."""

    def __init__(self, compose):
        self.split = compose if isinstance(compose, SplitAugment) else SplitAugment(compose)

    def __call__(self, x):
        view1 = self.split(x)
        view2 = self.split(x)
        return view1, view2


_TABLES = {}


def finish_views(u8, k1d, erase, mean, std, device=None, out=None):
    """GPU half: uint8 [n,s,s] + per-view blur taps [n,3] + erase rectangles [n,4] -> fp32 [n,3,224,224] on `device`."""
    device = torch.device(device or "cuda")
    if device.type != "cuda":
        raise RuntimeError("vit2spn.augment.finish_views needs a CUDA device (no CPU fallback)")
    _lib.init_device(device.index if device.index is not None else torch.cuda.current_device())
    n, s = int(u8.shape[0]), int(u8.shape[-1])
    key = (s, device)
    if key not in _TABLES:
        b, c = pil_bilinear_tables(s)
        _TABLES[key] = (torch.from_numpy(b).to(device), torch.from_numpy(c).to(device), int(c.shape[1]))
    bounds, coefs, ksize = _TABLES[key]
    u8 = u8.to(device, non_blocking=True).contiguous()
    k1d = k1d.to(device, torch.float32, non_blocking=True).contiguous()
    erase = erase.to(device, torch.int32, non_blocking=True).contiguous()
    if out is None:
        out = torch.empty(n, 3, OUT, OUT, device=device, dtype=torch.float32)
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    _lib.check(_lib.lib.v2s_augment_finish_u8(_lib.ptr(u8), n, s, _lib.ptr(bounds), _lib.ptr(coefs), ksize, _lib.ptr(k1d),
                                              _lib.ptr(erase), m3, s3, _lib.ptr(out), _lib.stream_ptr()), "augment_finish")
    return out


class _GpuDualViewLoader:
    def __init__(self, inner, split, device):
        self.inner, self.split, self.device = inner, split, torch.device(device)

    def __len__(self):
        return len(self.inner)

    def __iter__(self):
        for (v1, v2), labels in self.inner:
            views = [finish_views(v[0], v[1], v[2], self.split.mean, self.split.std, self.device) for v in (v1, v2)]
            yield views, labels


def gpu_dual_view_loader(dataset, compose, batch_size, device="cuda", **loader_kwargs):
    """DataLoader with the reference's contract (``for (views, _) in loader: view1, view2 = views``,
    ref:ssp_vit2spn_tiny.py:100-107,205-207) whose views are finished on the GPU.  `dataset.transform` is replaced
    by the host half of `compose`; workers ship 2 x 784 bytes per image instead of 2 x 602 KB."""
    from torch.utils.data import DataLoader
    dual = DualViewSplit(compose)
    dataset.transform = dual
    inner = DataLoader(dataset, batch_size=batch_size, **loader_kwargs)
    return _GpuDualViewLoader(inner, dual.split, device)
