"""TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the *reference's own* Triton kernel modules
(unsloth/kernels/{rms_layernorm,rope_embedding,cross_entropy_loss,swiglu,geglu,fast_lora,
utils}.py) without running `unsloth/__init__.py` (which needs unsloth_zoo / peft / trl /
bitsandbytes -- all absent offline).  Recipe follows SURVEY.md section 8c and the reference's own
GPU-free harness (tests/conftest.py:52-207 in the reference).

Two modes:

  * `load_reference_kernels()` -- CPU, TRITON_INTERPRET=1 (fp32 / fp16).  Used inside the build
    container by oracle/make_golden.py to produce the committed fixtures under tests/golden/.
  * `load_reference_kernels_native()` -- on a CUDA device, real Triton compilation: the reference's
    kernels exactly as a user of the reference runs them (bf16 included).  Used by the `-m gpu`
    parity tests (tests/test_gpu_vs_reference.py) and by the GPU reference baseline
    (benchmarks/ref_triton_bench.py, `bench.py --impl gpu-reference`).

Where the reference comes from: `/root/reference` in the build container, else the UNMODIFIED
offline install of that tree under `baseline/_ref` (`pip install --no-deps --target baseline/_ref`,
git-ignored, travels to the GPU box with the gpurun snapshot; recipe in DESIGN.md section 5).
Nothing is copied into the repository's history.

bitsandbytes is not installable offline, so in native mode the five C symbols the reference binds
from it (kernels/utils.py:273-284) are rebound to the same-signature symbols exported by
libunsloth_b200.so -- the swap INTEGRATION.md documents.  Everything else on the path (Triton
kernels, torch.matmul / addmm_ schedule, autograd functions) is the reference's own code.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types
from contextlib import nullcontext
from types import SimpleNamespace

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference_root():
    cands = [os.environ.get("UNSLOTH_REFERENCE_ROOT"), "/root/reference",
             os.path.join(_ROOT, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "unsloth", "kernels")):
            return c
    return cands[1]


REFERENCE_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "unsloth", "kernels"))


_LOADED = None
_MODE = None

KERNEL_MODULES = ("utils", "rms_layernorm", "rope_embedding", "cross_entropy_loss", "swiglu",
                  "geglu", "fast_lora")


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _skeleton(device_count):
    """Skeleton packages: __path__ points into the reference so submodules resolve, but neither
    __init__.py executes; third-party imports of the kernel modules are stubbed."""
    pkg = _mod("unsloth")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "unsloth")]
    kpkg = _mod("unsloth.kernels")
    kpkg.__path__ = [os.path.join(REFERENCE_ROOT, "unsloth", "kernels")]

    from packaging.version import Version

    _mod("unsloth_zoo").__path__ = []
    _mod("unsloth_zoo.utils", Version=Version, _get_dtype=lambda x: x,
         get_quant_type=lambda *a, **k: None)
    _mod("unsloth_zoo.loss_utils",
         patch_loss_functions=lambda *a, **k: None,
         post_patch_loss_function=lambda m: m)
    _mod("unsloth_zoo.patching_utils", patch_layernorm=lambda *a, **k: None)
    _mod("unsloth.device_type",
         is_hip=lambda: False, get_device_type=lambda: "cuda",
         DEVICE_TYPE="cuda", DEVICE_TYPE_TORCH="cuda", DEVICE_COUNT=device_count,
         ALLOW_PREQUANTIZED_MODELS=True, ALLOW_BITSANDBYTES=False)
    _mod("unsloth.kernels.fp8", weight_dequant=None, fp8_linear=None)

    # the real bnb_availability module is dependency-free: exec it as-is
    spec = importlib.util.spec_from_file_location(
        "unsloth.bnb_availability",
        os.path.join(REFERENCE_ROOT, "unsloth", "bnb_availability.py"))
    bnbm = importlib.util.module_from_spec(spec)
    sys.modules["unsloth.bnb_availability"] = bnbm
    spec.loader.exec_module(bnbm)


def load_reference_kernels():
    """Return a namespace with the reference kernel modules (CPU interpreter mode)."""
    global _LOADED, _MODE
    if _LOADED is not None:
        if _MODE != "interpret":
            raise RuntimeError("reference kernels already loaded in %s mode" % _MODE)
        return _LOADED
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    os.environ["TRITON_INTERPRET"] = "1"
    import torch  # noqa: F401
    import triton  # noqa: F401

    _skeleton(device_count=0)
    mods = {name: importlib.import_module("unsloth.kernels." + name) for name in KERNEL_MODULES}

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
    _LOADED, _MODE = SimpleNamespace(**mods), "interpret"
    return _LOADED


def load_reference_kernels_native(bind_bnb_symbols=True):
    """The reference kernel modules compiled by Triton for the visible CUDA device.

    `bind_bnb_symbols`: rebind the bitsandbytes C symbols of kernels/utils.py:273-284 (and
    `get_ptr`, HAS_CUDA_STREAM) to libunsloth_b200.so's same-signature exports so that the
    reference's `fast_dequantize` / `matmul_lora` run on NF4 weights (bitsandbytes itself is not
    installable offline)."""
    global _LOADED, _MODE
    if _LOADED is not None:
        if _MODE != "native":
            raise RuntimeError("reference kernels already loaded in %s mode" % _MODE)
        return _LOADED
    if not reference_available():
        raise RuntimeError("reference not present (looked at %s)" % REFERENCE_ROOT)
    if os.environ.get("TRITON_INTERPRET") == "1":
        raise RuntimeError("TRITON_INTERPRET=1 is set: native mode needs real Triton compilation")
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("native mode needs a CUDA device")

    _skeleton(device_count=1)       # 1: CUDA_STREAMS / WEIGHT_BUFFERS sized for cuda:0, no per-layer
    #                                 stream.synchronize() (rope_embedding.py:278-279 is DEVICE_COUNT > 1)
    if bind_bnb_symbols and "bitsandbytes" not in sys.modules:
        # a stand-in `bitsandbytes` whose native library IS libunsloth_b200.so: the reference's
        # module-scope binds (kernels/utils.py:138-165, 263-284) then take their normal CUDA branch
        # (HAS_CUDA_STREAM, get_ptr, the five ctypes handles) with no edits to its code.
        import ctypes
        so = os.path.join(_ROOT, "unsloth_b200", "_C", "libunsloth_b200.so")
        lib = ctypes.CDLL(so)
        for name in ("cdequantize_blockwise_fp32", "cdequantize_blockwise_fp16_nf4",
                     "cdequantize_blockwise_bf16_nf4", "cgemm_4bit_inference_naive_fp16",
                     "cgemm_4bit_inference_naive_bf16"):
            getattr(lib, name).restype = None
        fn = _mod("bitsandbytes.functional", lib=lib,
                  get_ptr=lambda t: None if t is None else ctypes.c_void_p(t.data_ptr()))
        _mod("bitsandbytes", __version__="0.45.5", functional=fn).__path__ = []
    mods = {name: importlib.import_module("unsloth.kernels." + name) for name in KERNEL_MODULES}
    _LOADED, _MODE = SimpleNamespace(**mods), "native"
    return _LOADED
