"""TEST INFRASTRUCTURE ONLY — CPU restatement of the part of the reference's augmentation pipeline that
`v2s_augment_finish_u8` runs on the GPU: everything from `transforms.Resize((224, 224))` on
(ref:ssp_vit2spn_tiny.py:90-95):

    Resize((224,224)) on the 28x28 PIL image  ->  ToTensor  ->  GaussianBlur(3, sigma)  ->  RandomErasing(value 0)
    ->  Normalize(ImageNet mean / std)

The arithmetic lives in un-vendored third-party code, restated here from its published algorithm and pinned by
tests/test_augment.py against the installed libraries themselves:
  * Pillow 12.2.0 `Image.resize(..., BILINEAR)` on 8-bit images (src/libImaging/Resample.c): separable, horizontal
    pass first, coefficients normalised in double precision and rounded to 22-bit fixed point
    (`(int)(0.5 + w * (1 << 22))`), accumulation from `1 << 21`, result `>> 22` clipped to 0..255 after EACH pass.
  * torchvision 0.26.0 `to_tensor` (uint8 / 255 in fp32), `gaussian_blur` (reflect padding, outer-product 3x3
    kernel, fp32), `erase` (rectangle := 0), `normalize` ((x - mean) / std in fp32).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
OUT = 224
MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def pil_bilinear_coeffs(in_size, out_size=OUT):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (support 1.0)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    coefs = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = []
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w.append(1.0 - a if a < 1.0 else 0.0)
        ww = sum(w)
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            coefs[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, coefs


def _pass(img, bounds, coefs):
    """one horizontal resampling pass over the last axis of a uint8 array"""
    out = np.empty(img.shape[:-1] + (bounds.shape[0],), np.uint8)
    src = img.astype(np.int64)
    for xx in range(bounds.shape[0]):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(img.shape[:-1], 1 << (PRECISION_BITS - 1), np.int64)
        for k in range(n):
            acc += src[..., x0 + k] * int(coefs[xx, k])
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_resize_bilinear_u8(img, out_size=OUT):
    """uint8 [H, W] -> uint8 [out, out], bit-exact with PIL's BILINEAR resize of an 8-bit image."""
    bh, ch = pil_bilinear_coeffs(img.shape[1], out_size)
    bv, cv = pil_bilinear_coeffs(img.shape[0], out_size)
    tmp = _pass(img, bh, ch)                       # horizontal first
    return _pass(tmp.T.copy(), bv, cv).T.copy()     # then vertical


def finish_view(img_u8, k1d=None, erase=None, mean=MEAN, std=STD):
    """uint8 [28, 28] (the image after the PIL-side augmentations) -> fp32 [3, 224, 224]"""
    r = pil_resize_bilinear_u8(np.asarray(img_u8, np.uint8))
    x = r.astype(np.float32) / np.float32(255.0)
    if k1d is not None:
        k = np.asarray(k1d, np.float32)
        k2 = (k[:, None] * k[None, :]).astype(np.float32)
        p = np.pad(x, 1, mode="reflect")
        y = np.zeros_like(x)
        for i in range(3):
            for j in range(3):
                y += k2[i, j] * p[i:i + OUT, j:j + OUT]
        x = y
    if erase is not None and erase[2] > 0 and erase[3] > 0:
        i, j, h, w = [int(v) for v in erase]
        x = x.copy()
        x[i:i + h, j:j + w] = 0.0
    out = np.empty((3, OUT, OUT), np.float32)
    for c in range(3):
        out[c] = (x - np.float32(mean[c])) / np.float32(std[c])
    return out
