import os

import numpy as np
from PIL import Image

from .info import INFO


class OCTMNIST:
    """Seeded synthetic OCTMNIST: items are ``(PIL 'L' 28x28, np.ndarray[1] int label)``; ``labels`` is [N,1].
    ``V2S_SHIM_DATASET_SIZE`` bounds N (default: the real split sizes)."""
    flag = "octmnist"

    def __init__(self, split, transform=None, target_transform=None, download=False, as_rgb=False, root=None,
                 size=None, mmap_mode=None):
        if os.environ.get("V2S_SYNTHETIC_DATA", "0") != "1":
            raise ImportError("vit2spn's medmnist stand-in fabricates data; set V2S_SYNTHETIC_DATA=1 to opt in")
        print(f"*** SYNTHETIC DATA *** OCTMNIST('{split}') is a seeded stand-in (stripes + noise, random labels), not medmnist",
              flush=True)
        self.info = INFO[self.flag]
        self.synthetic = True
        self.split, self.transform, self.target_transform, self.as_rgb = split, transform, target_transform, as_rgb
        n = self.info["n_samples"][split]
        cap = os.environ.get("V2S_SHIM_DATASET_SIZE")
        if cap:
            n = min(n, int(cap))
        rng = np.random.default_rng({"train": 0, "val": 1, "test": 2}[split])
        self.labels = rng.integers(0, len(self.info["label"]), size=(n, 1)).astype(np.uint8)
        self._seed = int(rng.integers(0, 2 ** 31))
        self._n = n

    def __len__(self):
        return self._n

    def _pixels(self, index):
        # smooth blob + label-dependent stripe frequency, so a classifier has something to learn
        rng = np.random.default_rng(self._seed + index)
        yy, xx = np.mgrid[0:28, 0:28].astype(np.float32)
        lab = int(self.labels[index, 0])
        img = 96 + 64 * np.sin((lab + 1) * 0.45 * yy + rng.uniform(0, 6.28)) + 24 * rng.standard_normal((28, 28))
        return np.clip(img, 0, 255).astype(np.uint8)

    def __getitem__(self, index):
        img = Image.fromarray(self._pixels(index), mode="L")
        target = self.labels[index].astype(int)
        if self.as_rgb:
            img = img.convert("RGB")
        if self.transform is not None:
            img = self.transform(img)
        if self.target_transform is not None:
            target = self.target_transform(target)
        return img, target
