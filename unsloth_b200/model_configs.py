"""Model shapes of the BASELINE.json configs (public model cards) and the
HuggingFace config builder.  Importing this module loads nothing else of the package (in particular
not libunsloth_b200.so), so the CPU reference arm of bench.py can use it without loading the product."""
CONFIGS = {
    "llama-3.2-1b": dict(arch="llama", hidden_size=2048, num_hidden_layers=16, num_attention_heads=32,
                         num_key_value_heads=8, head_dim=64, intermediate_size=8192, vocab_size=128256,
                         rms_norm_eps=1e-5, rope_theta=500000.0, tie_word_embeddings=True,
                         max_position_embeddings=131072,
                         rope_scaling=dict(rope_type="llama3", factor=32.0, low_freq_factor=1.0,
                                           high_freq_factor=4.0, original_max_position_embeddings=8192)),
    "llama-3-8b": dict(arch="llama", hidden_size=4096, num_hidden_layers=32, num_attention_heads=32,
                       num_key_value_heads=8, head_dim=128, intermediate_size=14336, vocab_size=128256,
                       rms_norm_eps=1e-5, rope_theta=500000.0),
    "mistral-7b-v0.3": dict(arch="mistral", hidden_size=4096, num_hidden_layers=32, num_attention_heads=32,
                            num_key_value_heads=8, head_dim=128, intermediate_size=14336, vocab_size=32768,
                            rms_norm_eps=1e-5, rope_theta=1000000.0, sliding_window=None),
    "gemma-2-9b": dict(arch="gemma2", hidden_size=3584, num_hidden_layers=42, num_attention_heads=16,
                       num_key_value_heads=8, head_dim=256, intermediate_size=14336, vocab_size=256000,
                       rms_norm_eps=1e-6, rope_theta=10000.0, sliding_window=4096,
                       attn_logit_softcapping=50.0, final_logit_softcapping=30.0,
                       query_pre_attn_scalar=256),
}


def hf_config(name, num_hidden_layers=None, **overrides):
    spec = dict(CONFIGS[name])
    spec.update(overrides)
    arch = spec.pop("arch")
    if num_hidden_layers is not None:
        spec["num_hidden_layers"] = num_hidden_layers
    spec.setdefault("max_position_embeddings", 8192)
    spec["attention_bias"] = False
    if arch == "llama":
        from transformers import LlamaConfig as C
        spec["mlp_bias"] = False
    elif arch == "mistral":
        from transformers import MistralConfig as C
    else:
        from transformers import Gemma2Config as C
        spec["hidden_activation"] = "gelu_pytorch_tanh"
    return C(**spec)
