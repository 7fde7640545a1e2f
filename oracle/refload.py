"""ORACLE / TEST INFRASTRUCTURE ONLY.

Imports the reference's OWN ``model/unet.py``, ``model/layers.py`` and ``model/loss.py`` verbatim from
/root/reference (build container only -- the directory does not exist on the GPU box) after putting the
``resnest.torch`` / ``monai.losses`` stand-ins on ``sys.path``.  Used to (a) validate
the functional restatement in ``oracle/functional.py`` and (b) generate tests/golden/*.pt
(tools/make_golden.py).  Nothing is copied: the reference files are executed where they lie.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("XV2_REFERENCE_ROOT", "/root/reference")
_STANDINS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "standins")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "unet.py"))


def load_reference():
    """Returns (unet_module, layers_module, loss_module) of the reference, executed verbatim."""
    if not reference_available():
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    import torchvision.models as tvm

    for name in ("resnet50", "resnet101", "resnet152"):  # H7: pretrained=True needs the network
        fn = getattr(tvm, name)
        if getattr(fn, "_xv2_offline", False):
            continue

        def offline(pretrained=False, _fn=fn, **kw):
            return _fn(weights=None, **kw)

        offline._xv2_offline = True
        setattr(tvm, name, offline)
    if _STANDINS not in sys.path:
        sys.path.insert(0, _STANDINS)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "model" or k.startswith("model.")}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        unet = importlib.import_module("model.unet")
        layers = importlib.import_module("model.layers")
        loss = importlib.import_module("model.loss")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            sys.modules["_xv2ref_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    return unet, layers, loss
