"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the *reference's own* Triton kernel modules
(/root/reference/unsloth/kernels/{rms_layernorm,rope_embedding,cross_entropy_loss,
swiglu,geglu,fast_lora,utils}.py) on CPU under TRITON_INTERPRET=1, without running
`unsloth/__init__.py` (which needs unsloth_zoo / peft / trl / bitsandbytes -- all
absent offline).  Recipe follows SURVEY.md section 8c and the reference's own
GPU-free harness (tests/conftest.py:52-207 in the reference).

Only usable inside the build container (where /root/reference exists); it is used by
oracle/make_golden.py to produce the committed fixtures under tests/golden/ and by
tests that are skipped when the reference tree is absent (e.g. on the GPU box).
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types
from contextlib import nullcontext
from types import SimpleNamespace

REFERENCE_ROOT = os.environ.get("UNSLOTH_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "unsloth", "kernels"))


_LOADED = None


def load_reference_kernels():
    """Return a namespace with the reference kernel modules (interpreter mode)."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    os.environ["TRITON_INTERPRET"] = "1"
    import torch  # noqa: F401
    import triton
    import triton.language as tl

    def _mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    # skeleton packages: __path__ points into the reference so submodules resolve,
    # but neither __init__.py executes.
    pkg = _mod("unsloth")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "unsloth")]
    kpkg = _mod("unsloth.kernels")
    kpkg.__path__ = [os.path.join(REFERENCE_ROOT, "unsloth", "kernels")]

    from packaging.version import Version

    def _get_dtype(x):
        return x

    _mod("unsloth_zoo").__path__ = []
    _mod("unsloth_zoo.utils", Version=Version, _get_dtype=_get_dtype,
         get_quant_type=lambda *a, **k: None)
    _mod("unsloth_zoo.loss_utils",
         patch_loss_functions=lambda *a, **k: None,
         post_patch_loss_function=lambda m: m)
    _mod("unsloth_zoo.patching_utils", patch_layernorm=lambda *a, **k: None)
    _mod("unsloth.device_type",
         is_hip=lambda: False, get_device_type=lambda: "cuda",
         DEVICE_TYPE="cuda", DEVICE_TYPE_TORCH="cuda", DEVICE_COUNT=0,
         ALLOW_PREQUANTIZED_MODELS=True, ALLOW_BITSANDBYTES=False)
    _mod("unsloth.kernels.fp8", weight_dequant=None, fp8_linear=None)

    # the real bnb_availability module is dependency-free: exec it as-is
    spec = importlib.util.spec_from_file_location(
        "unsloth.bnb_availability",
        os.path.join(REFERENCE_ROOT, "unsloth", "bnb_availability.py"))
    bnbm = importlib.util.module_from_spec(spec)
    sys.modules["unsloth.bnb_availability"] = bnbm
    spec.loader.exec_module(bnbm)

    mods = {}
    for name in ("utils", "rms_layernorm", "rope_embedding", "cross_entropy_loss",
                 "swiglu", "geglu", "fast_lora"):
        mods[name] = importlib.import_module("unsloth.kernels." + name)

    # interpreter cannot call the import-time aliases of un-patched builtins
    from oracle._interp_aliases import interp_cast as _cast, interp_tanh as _tanh

    for m in mods.values():
        if hasattr(m, "torch_gpu_device"):
            m.torch_gpu_device = lambda d: nullcontext()
        if hasattr(m, "triton_cast"):
            m.triton_cast = _cast
        if hasattr(m, "triton_tanh"):
            m.triton_tanh = _tanh
    mods["rope_embedding"].torch_device_stream = \
        lambda d: SimpleNamespace(synchronize=lambda: None)
    _LOADED = SimpleNamespace(**mods)
    return _LOADED
