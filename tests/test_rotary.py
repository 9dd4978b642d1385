"""RoPE tables of `patch.RotaryCache` (a12: models/llama.py:1775-1914) against transformers' own
rope initialisers, the bar of the reference's tests/utils/test_rope_scaling_drift.py:207-240
(rtol 1e-4, atol 1e-6), plus the table contract the RoPE kernel relies on:
[size, D] with both halves duplicated, size grown in steps of 8192, table dtype per architecture."""
import torch

from unsloth_b200.patch import CONFIGS, RotaryCache, hf_config

HEAD_DIM, ROPE_THETA, MAX_POS = 64, 500000.0, 131072
LLAMA3 = {"rope_type": "llama3", "factor": 32.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
          "original_max_position_embeddings": 8192}


def _hf_inv_freq(rope_scaling, rope_type):
    from transformers import LlamaConfig
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
    kw = dict(hidden_size=256, num_attention_heads=4, num_key_value_heads=2, head_dim=HEAD_DIM,
              max_position_embeddings=MAX_POS)
    try:
        cfg = LlamaConfig(rope_theta=ROPE_THETA, rope_scaling=rope_scaling, **kw)
    except TypeError:
        cfg = LlamaConfig(rope_parameters=dict(rope_scaling or {"rope_type": "default"}, rope_theta=ROPE_THETA), **kw)
    inv, _ = ROPE_INIT_FUNCTIONS[rope_type](cfg, "cpu")
    return inv.float().cpu()


def _vanilla():
    return 1.0 / (ROPE_THETA ** (torch.arange(0, HEAD_DIM, 2, dtype=torch.int64).float() / HEAD_DIM))


import pytest


@pytest.mark.parametrize("factor", [32.0, 8.0])      # Llama-3.2-1B's factor; the reference test's (:41-47)
def test_llama3_scaling_matches_transformers(factor):
    rs = dict(LLAMA3, factor=factor)
    ours = RotaryCache(HEAD_DIM, ROPE_THETA, "cpu", torch.float32, rs).inv_freq()
    expected = _hf_inv_freq(rs, "llama3")
    assert not torch.allclose(expected, _vanilla(), rtol=1e-4)          # the guard of the reference test
    assert torch.allclose(ours, expected, rtol=1e-4, atol=1e-6)


def test_default_rope_is_vanilla():
    ours = RotaryCache(HEAD_DIM, ROPE_THETA, "cpu", torch.float32, None).inv_freq()
    assert torch.allclose(ours, _vanilla(), rtol=1e-4, atol=1e-6)
    ours = RotaryCache(HEAD_DIM, ROPE_THETA, "cpu", torch.float32, {"rope_type": "default"}).inv_freq()
    assert torch.allclose(ours, _vanilla(), rtol=1e-4, atol=1e-6)


def test_table_contract():
    rc = RotaryCache(HEAD_DIM, 10000.0, "cpu", torch.bfloat16, None)
    cos, sin = rc.get(100)
    assert cos.shape == (8192, HEAD_DIM) and cos.dtype == torch.bfloat16          # grown by 8192
    assert torch.equal(cos[:, :HEAD_DIM // 2], cos[:, HEAD_DIM // 2:])            # halves duplicated
    assert torch.equal(sin[:, :HEAD_DIM // 2], sin[:, HEAD_DIM // 2:])
    t = torch.arange(8192, dtype=torch.float32)
    fr = torch.outer(t, rc.inv_freq())
    assert torch.equal(cos[:, :HEAD_DIM // 2], fr.cos().to(torch.bfloat16))       # fp32 math, one rounding
    c2, _ = rc.get(8192)
    assert c2.data_ptr() == cos.data_ptr()                                         # cached
    c3, _ = rc.get(8193)
    assert c3.shape[0] == 16384 and torch.equal(c3[:8192], cos)
    assert RotaryCache(256, 10000.0, "cpu", torch.float32, None).get(8)[0].dtype == torch.float32   # Gemma tables


def test_config_table_matches_survey_section_8():
    """The synthetic model configs as HF configs."""
    c = hf_config("llama-3-8b")
    assert (c.hidden_size, c.num_hidden_layers, c.num_attention_heads, c.num_key_value_heads,
            c.intermediate_size, c.vocab_size) == (4096, 32, 32, 8, 14336, 128256)
    c = hf_config("llama-3.2-1b")
    assert (c.hidden_size, c.num_hidden_layers, c.intermediate_size, c.tie_word_embeddings) == (2048, 16, 8192, True)
    c = hf_config("mistral-7b-v0.3")
    assert (c.hidden_size, c.intermediate_size, c.vocab_size) == (4096, 14336, 32768)
    c = hf_config("gemma-2-9b")
    assert (c.hidden_size, c.num_hidden_layers, c.head_dim, c.vocab_size) == (3584, 42, 256, 256000)
    assert c.attn_logit_softcapping == 50.0 and c.final_logit_softcapping == 30.0 and c.sliding_window == 4096
    assert set(CONFIGS) >= {"llama-3-8b", "llama-3.2-1b", "mistral-7b-v0.3", "gemma-2-9b"}
