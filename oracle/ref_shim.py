"""Import the UNMODIFIED reference `gbm/model.py` on a CPU-only box.  TEST INFRASTRUCTURE ONLY.

Needs the reference's two source files: the checkout of the build container (/root/reference) or the
byte-for-byte copies oracle/make_ref.py puts under oracle/_ref/ (git-ignored; they travel to the GPU box).
Used by tests/golden/make_golden.py to generate the committed golden vectors, by the live
oracle-vs-reference test, and by bench.py's CPU legs (--impl reference, cpu_baseline) -- never by the
product path.

The shim is harness code, not a restatement:
  1. `PyTorchHelpers` (gbm/model.py:7) is not shipped with the reference -> empty stub module;
  2. `.cuda()` (gbm/model.py:135,154,189 - the last one is evaluated at class-definition time)
     -> identity;
  3. `nn.DataParallel(..., device_ids=[0,1,2,3])` (gbm/model.py:132-135) -> pass-through wrapper
     that keeps the `.module` attribute, hence the `cnn.module.*` state-dict keys.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/make_ref.py (git-ignored copies)


def _find_root() -> str:
    env = os.environ.get("MIL_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/gbm/model.py"):
        return "/root/reference"
    return _VENDORED


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "gbm", "model.py"))


def load_reference_model_module():
    """Returns the reference's `model` module (gbm/model.py), imported under the shim."""
    import torch
    from torch import nn

    if "model" in sys.modules and getattr(sys.modules["model"], "__mil_shim__", False):
        return sys.modules["model"]
    if not reference_available():
        raise FileNotFoundError(f"reference checkout not found under {REFERENCE_ROOT}")

    sys.modules.setdefault("PyTorchHelpers", types.ModuleType("PyTorchHelpers"))

    class _PassThroughDP(nn.Module):
        def __init__(self, module, device_ids=None, **_):
            super().__init__()
            self.module = module

        def forward(self, *a, **k):
            return self.module(*a, **k)

    saved = (torch.Tensor.cuda, nn.Module.cuda, nn.DataParallel)
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    nn.DataParallel = _PassThroughDP
    for pth in (os.path.join(REFERENCE_ROOT, "gbm"), REFERENCE_ROOT):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    try:
        import model as ref_model  # noqa: the reference's gbm/model.py
    finally:
        # keep the shims in place: the reference calls .cuda() inside __init__ too
        pass
    ref_model.__mil_shim__ = True
    ref_model.__mil_saved__ = saved
    return ref_model


def build_reference(class_weights=None, seed: int = 0, quiet: bool = True):
    """`Attention(n_classes=3, class_weights)` of the reference with its own init under
    torch.manual_seed(seed)."""
    import torch
    ref_model = load_reference_model_module()
    torch.manual_seed(seed)
    ctx = contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()
    with ctx:
        net = ref_model.Attention(n_classes=3, class_weights=class_weights)
    return net


def alt_resnet_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "alt_resnet.py"))


def load_reference_alt_resnet():
    """The reference's UNMODIFIED `alt_resnet.py`.  It cannot be imported as it stands -- line 3 is the relative import
    `from .utils import load_state_dict_from_url` and the checkout has neither a package nor a `utils` module (SURVEY.md
    section 2) -- so it is loaded as the submodule of a stub package whose `utils` provides that one name (only used
    by `pretrained=True`, which nothing here asks for).  Harness code, not a restatement."""
    import importlib.util
    name = "_mil_altpkg.alt_resnet"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(REFERENCE_ROOT, "alt_resnet.py")
    if not os.path.isfile(path):
        raise FileNotFoundError(f"{path} not found")
    pkg = types.ModuleType("_mil_altpkg")
    pkg.__path__ = []
    utils = types.ModuleType("_mil_altpkg.utils")

    def load_state_dict_from_url(*a, **k):  # pragma: no cover
        raise RuntimeError("no network: pretrained weights are not available")
    utils.load_state_dict_from_url = load_state_dict_from_url
    sys.modules["_mil_altpkg"] = pkg
    sys.modules["_mil_altpkg.utils"] = utils
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "_mil_altpkg"
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def build_reference_wide(layers=(2, 2, 2, 2), class_weights=None, seed: int = 0, quiet: bool = True):
    """The reference's `Attention` with its `cnn` replaced by the reference's own alt ResNet:
    `nn.DataParallel(alt_resnet.ResNet(alt_resnet.BasicBlock, layers, num_classes=80))` (gbm/model.py:132-135 with
    alt_resnet.py:70-145 in place of gbm/model.py:14-61) -- both classes unmodified."""
    from torch import nn
    net = build_reference(class_weights, seed, quiet)
    alt = load_reference_alt_resnet()
    net.cnn = nn.DataParallel(alt.ResNet(alt.BasicBlock, list(layers), num_classes=net.L))
    return net
