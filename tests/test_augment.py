"""Input pipeline split: PIL-side augmentations on the host at 28x28, Resize -> ToTensor ->
GaussianBlur -> RandomErasing -> Normalize on the GPU.  The oracle restatement is pinned against Pillow /
torchvision themselves; the CUDA kernel is compared with the oracle, and the whole split pipeline with the reference's
own Compose under the same RNG seed."""
import os
import sys

import numpy as np
import pytest
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import augment_oracle as ao  # noqa: E402


def reference_compose():
    """The transform of ref:ssp_vit2spn_tiny.py:84-96 (parameters are the reference's)."""
    from torchvision import transforms
    return transforms.Compose([
        transforms.Grayscale(num_output_channels=3),
        transforms.RandomHorizontalFlip(p=0.5),
        transforms.RandomVerticalFlip(p=0.3),
        transforms.RandomRotation(degrees=30),
        transforms.RandomAffine(degrees=15, translate=(0.1, 0.1), scale=(0.8, 1.2), shear=10),
        transforms.ColorJitter(brightness=0.3, contrast=0.3, saturation=0.3, hue=0.1),
        transforms.Resize((224, 224)),
        transforms.ToTensor(),
        transforms.GaussianBlur(kernel_size=3, sigma=(0.1, 2.0)),
        transforms.RandomErasing(p=0.5, scale=(0.02, 0.2), ratio=(0.3, 3.3)),
        transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]),
    ])


def _images(n, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for t in range(n):
        if t % 3 == 0:
            yy, xx = np.mgrid[0:28, 0:28]
            a = (96 + 64 * np.sin(0.4 * yy + t) + 30 * np.cos(0.3 * xx)).clip(0, 255).astype(np.uint8)
        else:
            a = rng.integers(0, 256, size=(28, 28), dtype=np.uint8)
        out.append(a)
    return out


def test_oracle_resize_is_bit_exact_with_pillow():
    for a in _images(40) + [np.zeros((28, 28), np.uint8), np.full((28, 28), 255, np.uint8)]:
        ref = np.asarray(Image.fromarray(a, mode="L").resize((224, 224), Image.BILINEAR))
        assert np.array_equal(ao.pil_resize_bilinear_u8(a), ref)
    a = _images(2, 5)[1]
    rgb = np.asarray(Image.fromarray(a, mode="L").convert("RGB").resize((224, 224), Image.BILINEAR))
    assert np.array_equal(rgb[..., 1], ao.pil_resize_bilinear_u8(a))       # the RGB image the reference resizes


def test_oracle_finish_matches_torchvision():
    import torchvision.transforms.functional as F
    import torchvision.transforms._functional_tensor as FT
    for t, a in enumerate(_images(6, 1)):
        pil = Image.fromarray(a, mode="L").convert("RGB")
        x = F.to_tensor(F.resize(pil, [224, 224]))
        sigma = 0.1 + 0.35 * t
        x = F.gaussian_blur(x, [3, 3], [sigma, sigma])
        rect = (10 * t, 5 + 7 * t, 20 + 9 * t, 100 - 11 * t)
        x = F.erase(x, *rect, torch.tensor(0.0))
        x = F.normalize(x, ao.MEAN, ao.STD)
        k1d = FT._get_gaussian_kernel1d(3, sigma, torch.float32, torch.device("cpu")).numpy()
        got = ao.finish_view(a, k1d, rect)
        assert np.abs(got - x.numpy()).max() <= 2e-6


def test_product_tables_equal_oracle_tables():
    import vit2spn  # noqa: F401
    from vit2spn import augment
    for s in (28, 32, 64):
        b, c = augment.pil_bilinear_tables(s)
        bo, co = ao.pil_bilinear_coeffs(s)
        assert np.array_equal(b, bo) and np.array_equal(c, co)


def test_split_pipeline_equals_reference_compose_on_cpu():
    """Same torch seed -> the host half + the (oracle) finish reproduce the reference Compose: the split draws every
    random parameter in the reference's order."""
    import vit2spn  # noqa: F401
    from vit2spn import augment
    compose = reference_compose()
    split = augment.SplitAugment(compose)
    assert split.tail_order == ["Resize", "ToTensor", "GaussianBlur", "RandomErasing", "Normalize"]
    for seed, a in enumerate(_images(12, 2)):
        pil = Image.fromarray(a, mode="L")
        torch.manual_seed(seed)
        ref = [compose(pil), compose(pil)]                    # DualViewTransform: two independent draws
        torch.manual_seed(seed)
        for r in ref:
            u8, k1d, erase = split(pil)
            got = ao.finish_view(u8.numpy(), k1d.numpy(), erase.numpy(), split.mean, split.std)
            assert np.abs(got - r.numpy()).max() <= 2e-6, seed


def test_split_rejects_what_it_cannot_reproduce():
    import vit2spn  # noqa: F401
    from vit2spn import augment
    from torchvision import transforms
    with pytest.raises(NotImplementedError):
        augment.SplitAugment(transforms.Compose([transforms.ToTensor()]))
    with pytest.raises(NotImplementedError):
        augment.SplitAugment(transforms.Compose([transforms.Resize((224, 224)), transforms.ToTensor(),
                                                 transforms.RandomErasing(value=1.0)]))
    with pytest.raises(NotImplementedError):
        augment.SplitAugment(transforms.Compose([transforms.Resize((128, 128)), transforms.ToTensor()]))


@pytest.mark.gpu
def test_cuda_finish_matches_oracle_and_pillow():
    import vit2spn  # noqa: F401
    from vit2spn import augment
    dev = torch.device("cuda", 0)
    imgs = _images(9, 3)
    rng = np.random.default_rng(7)
    k1d = np.zeros((9, 3), np.float32); k1d[:, 1] = 1.0
    erase = np.zeros((9, 4), np.int32)
    for t in range(9):
        if t % 3 != 0:
            e = np.exp(-0.5 / rng.uniform(0.1, 2.0) ** 2)
            k1d[t] = np.array([e, 1.0, e], np.float32) / np.float32(1 + 2 * e)
        if t % 2 == 1:
            erase[t] = (rng.integers(0, 100), rng.integers(0, 100), rng.integers(1, 120), rng.integers(1, 120))
    out = augment.finish_views(torch.from_numpy(np.stack(imgs)), torch.from_numpy(k1d), torch.from_numpy(erase),
                               ao.MEAN, ao.STD, dev).cpu().numpy()
    for t in range(9):
        ref = ao.finish_view(imgs[t], None if t % 3 == 0 else k1d[t], erase[t])
        assert np.abs(out[t] - ref).max() <= 2e-6, t
    # no blur, no erasing: undoing Normalize / ToTensor gives Pillow's resize bit for bit
    t = 0
    back = np.rint((out[t][1] * np.float32(ao.STD[1]) + np.float32(ao.MEAN[1])) * 255.0).astype(np.uint8)
    assert np.array_equal(back, np.asarray(Image.fromarray(imgs[t], mode="L").resize((224, 224), Image.BILINEAR)))


@pytest.mark.gpu
def test_gpu_loader_equals_reference_dataloader():
    """The reference's DataLoader contract (views, labels) with GPU-finished views == the reference Compose run on the
    CPU with the same seed (num_workers=0 so that one RNG stream serves both)."""
    import vit2spn  # noqa: F401
    from vit2spn import augment

    class Tiny(torch.utils.data.Dataset):
        def __init__(self, transform):
            self.imgs, self.transform = _images(10, 4), transform

        def __len__(self):
            return len(self.imgs)

        def __getitem__(self, i):
            return self.transform(Image.fromarray(self.imgs[i], mode="L")), np.array([i % 4])

    compose = reference_compose()

    class Dual:
        def __call__(self, x):
            return compose(x), compose(x)

    torch.manual_seed(123)
    ref = list(torch.utils.data.DataLoader(Tiny(Dual()), batch_size=4, shuffle=False))
    torch.manual_seed(123)
    ours = list(augment.gpu_dual_view_loader(Tiny(None), compose, batch_size=4, device="cuda", shuffle=False))
    assert len(ref) == len(ours) == 3
    for (rv, rl), (ov, ol) in zip(ref, ours):
        assert torch.equal(rl, ol)
        for a, b in zip(rv, ov):
            assert b.is_cuda and b.shape == a.shape and b.dtype == torch.float32
            assert (a - b.cpu()).abs().max().item() <= 2e-6
