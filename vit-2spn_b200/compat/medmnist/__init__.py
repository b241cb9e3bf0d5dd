"""Synthetic stand-in for ``medmnist`` (absent from this image, and there is no network for its
``download=True``): same constructor, item and ``labels`` contract as the classes the reference scripts use
(ref:ssp_vit2spn_tiny.py:100-104, ref:octmnist_ft_vit2spn.py:47-50,177)."""
from .info import INFO  # noqa: F401
from .dataset import OCTMNIST  # noqa: F401
