"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports
every symbol include/unsloth_b200.h declares; the Python surface mirrors the reference's names."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "unsloth_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(ub200_\w+|cdequantize_\w+|cgemm_4bit_\w+)\s*\(", src))
    return sorted(names)


def _header_prototypes():
    """{name: (restype, [ctype kind per parameter])} parsed from the header; kinds: 'p' pointer /
    cudaStream_t, 'l' int64_t, 'i' int, 'f' float.  The `#else` (plain C) duplicates of the typed
    bitsandbytes prototypes are parsed too and must agree with the CUDA ones."""
    src = open(os.path.join(ROOT, "include", "unsloth_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = {}
    for m in re.finditer(r"\b(int|void)\s+(ub200_\w+|cdequantize_\w+|cgemm_4bit_\w+)\s*\(([^)]*)\)\s*;", src):
        res, name, params = m.group(1), m.group(2), m.group(3).strip()
        kinds = []
        if params and params != "void":
            for prm in params.split(","):
                prm = " ".join(prm.split())
                if "*" in prm or "cudaStream_t" in prm:
                    kinds.append("p")
                elif re.search(r"\bint64_t\b", prm):
                    kinds.append("l")
                elif re.search(r"\bfloat\b", prm):
                    kinds.append("f")
                elif re.search(r"\bint\b", prm):
                    kinds.append("i")
                else:
                    raise AssertionError("unparsed parameter %r of %s" % (prm, name))
        protos.setdefault(name, []).append((res, kinds))
    return protos


def test_ctypes_signatures_match_the_header():
    """Every ctypes signature in unsloth_b200/_lib.py has the header's parameter count, parameter
    kinds (pointer / int64 / int / float) and return type: a mismatch would pass garbage silently."""
    import ctypes
    from unsloth_b200 import _lib
    protos = _header_prototypes()
    assert set(protos) == set(_lib._SIGS), set(protos) ^ set(_lib._SIGS)

    def kind(t):
        if t in (ctypes.c_void_p,) or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
            return "p"
        return {ctypes.c_int64: "l", ctypes.c_int: "i", ctypes.c_float: "f"}[t]
    for name, variants in protos.items():
        args, res = _lib._SIGS[name]
        for hres, hkinds in variants:
            assert [kind(a) for a in args] == hkinds, (name, [kind(a) for a in args], hkinds)
            assert (res is None) == (hres == "void"), name


def test_library_exports_every_declared_symbol():
    from unsloth_b200 import _lib
    import ctypes
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), "missing export %s" % n
        assert n in _lib.EXPORTED_SYMBOLS, "no ctypes signature for %s" % n
    assert lib.ub200_abi_version() == 1


def test_python_surface_matches_reference_names():
    import unsloth_b200.kernels as K
    for name in ("fast_rms_layernorm", "fast_rope_embedding", "fast_cross_entropy_loss",
                 "unsloth_fused_ce_loss", "fast_dequantize", "matmul_lora", "get_lora_parameters",
                 "get_lora_parameters_bias", "apply_lora_mlp_swiglu", "apply_lora_mlp_geglu_approx",
                 "apply_lora_mlp_geglu_exact", "apply_lora_qkv", "apply_lora_o", "LoRA_MLP",
                 "LoRA_QKV", "LoRA_W", "swiglu_fg_kernel", "swiglu_DWf_DW_dfg_kernel",
                 "geglu_approx_forward_kernel", "geglu_approx_backward_kernel",
                 "Fast_RMS_Layernorm", "Fast_RoPE_Embedding", "Fast_RoPE_Embedding_QK",
                 "Fast_CrossEntropyLoss", "patch_rms_layernorm", "patch_loss_functions",
                 "fast_gemv", "fast_linear_forward", "fast_add_rms_layernorm"):
        assert hasattr(K, name), name


def test_no_cpu_fallback():
    """The product path must refuse CPU tensors instead of silently computing on the host."""
    import torch
    import unsloth_b200.kernels as K

    class N:
        weight = torch.ones(64)
        variance_epsilon = 1e-5
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        K.fast_rms_layernorm(N(), torch.randn(2, 3, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        K.swiglu_fg_kernel(torch.randn(1, 2, 64), torch.randn(1, 2, 64))


def test_product_does_not_import_oracle():
    """Nothing under unsloth_b200/ may import oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "unsloth_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dp, f)


def test_get_lora_parameters_contract():
    """Attribute contract of the PEFT scraper (reference: tests/test_fast_gemv_dispatch.py:36-63
    and kernels/utils.py:335-397): disabled / merged adapters yield A = B = s = None."""
    import torch
    from unsloth_b200.kernels import get_lora_parameters, get_lora_parameters_bias
    from unsloth_b200.lora import LoraLinear

    base = torch.nn.Linear(16, 8, bias=False)
    m = LoraLinear(base, r=4, lora_alpha=8)
    W, Wq, A, B, s = get_lora_parameters(m)
    assert W is base.weight and Wq is None and A.shape == (4, 16) and B.shape == (8, 4) and s == 2.0
    m.disable_adapters = True
    assert get_lora_parameters(m)[2:] == (None, None, None)
    m.disable_adapters = False
    m.merged = True
    assert get_lora_parameters(m)[2:] == (None, None, None)
    m.merged = False
    assert len(get_lora_parameters_bias(m)) == 6
    # a bare base layer (no LoRA wrapper) behaves as "adapters disabled"
    assert get_lora_parameters(base)[2:] == (None, None, None)
