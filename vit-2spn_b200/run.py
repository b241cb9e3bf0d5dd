"""Launcher: run an unmodified ViT-2SPN training script on the B200 backbone.

    python -m vit2spn.run /path/to/ssp_vit2spn_tiny.py [script args...]

What it does before handing control to the script ("what scripts call must keep working unchanged"):

* ``transformers.ViTModel`` / ``transformers.ViTConfig`` are rebound to the accelerated classes, so
  ``ViTModel.from_pretrained("WinKawaks/vit-tiny-patch16-224", output_hidden_states=True)``
  (ref:ssp_vit2spn_tiny.py:112) and ``ViTModel(ViTConfig(...))`` (ref:ssp_ssl/ssl_vit2spn_scratch.py:100-109)
  build the CUDA backbone; the script's own ``ViTBackbone`` / ``DualStreamNetwork`` / training loop,
  ``torch.optim.Adam``, ``GradScaler`` and ``.data`` EMA rebinding run as written.
* modules the image lacks (``fvcore``, ``matplotlib``) are served from ``vit2spn/compat`` when — and only when —
  the real ones cannot be imported.  The ``medmnist`` stand-in FABRICATES data (seeded synthetic OCTMNIST-shaped
  images with random labels), so it is opt-in: ``V2S_SYNTHETIC_DATA=1``; without it a missing ``medmnist`` stays an
  ImportError.  Every run on fabricated data says so on stdout (``SYNTHETIC DATA`` banner), and the analytic FLOP
  count of the ``fvcore`` stand-in is labelled as such.
* the script's output directories are created relative to the current directory, as it assumes.

The script's own classes call the backbone once per network (4 launches of the grouped kernels instead of 1);
the fully fused step is the native API (``vit2spn.DualStreamNetwork.ssp_step``).
"""
import importlib.util
import os
import runpy
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = ("medmnist", "fvcore", "matplotlib")


def install_shims(names=SHIMS):
    """Register compat stand-ins for modules that are not importable; returns the names shimmed."""
    done = []
    for name in names:
        if name in sys.modules:
            continue
        if name == "medmnist" and os.environ.get("V2S_SYNTHETIC_DATA", "0") != "1":
            continue              # fabricated data is opt-in: the script's own `import medmnist` fails loudly instead
        try:
            found = importlib.util.find_spec(name) is not None
        except (ImportError, ValueError):
            found = False
        if found:
            continue
        pkg_dir = os.path.join(_HERE, "compat", name)
        spec = importlib.util.spec_from_file_location(name, os.path.join(pkg_dir, "__init__.py"),
                                                      submodule_search_locations=[pkg_dir])
        module = importlib.util.module_from_spec(spec)
        sys.modules[name] = module
        spec.loader.exec_module(module)
        done.append(name)
    return done


def patch_transformers():
    """``from transformers import ViTModel, ViTConfig`` → the accelerated classes."""
    import transformers
    import vit2spn
    transformers.ViTModel = vit2spn.ViTModel
    transformers.ViTConfig = vit2spn.ViTConfig
    return transformers


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 0 if argv else 2
    script = argv[0]
    if not os.path.isfile(script):
        print(f"vit2spn.run: no such script: {script}", file=sys.stderr)
        return 2
    shimmed = install_shims()
    patch_transformers()
    if shimmed:
        print(f"vit2spn.run: stand-ins active for {', '.join(shimmed)} (not installed in this image)", file=sys.stderr)
    if "medmnist" in shimmed:
        print("vit2spn.run: *** SYNTHETIC DATA *** medmnist is not installed and V2S_SYNTHETIC_DATA=1: every dataset the "
              "script loads is FABRICATED (seeded stripes + noise, random labels); losses, accuracies and AUCs of this run "
              "say nothing about OCT images", flush=True)
    elif "medmnist" not in sys.modules and importlib.util.find_spec("medmnist") is None:
        print("vit2spn.run: medmnist is not installed; set V2S_SYNTHETIC_DATA=1 to run on fabricated OCTMNIST-shaped data",
              file=sys.stderr)
    # the fine-tune scripts savefig into this directory without creating it (ref:octmnist_ft_vit2spn.py:166)
    os.makedirs("./ssp_retinaloct_tbme/vit2spn_tiny/result", exist_ok=True)
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
