"""Kernel API of unsloth_b200: the same names `unsloth/kernels/__init__.py:15-62` exports."""
from .cross_entropy_loss import (fast_cross_entropy_loss, Fast_CrossEntropyLoss,
                                 unsloth_fused_ce_loss, MAX_FUSED_SIZE, patch_loss_functions,
                                 unpatch_loss_functions, UnslothForCausalLMLoss)
from .rms_layernorm import (fast_rms_layernorm, Fast_RMS_Layernorm, patch_rms_layernorm,
                            unpatch_rms_layernorm, fast_add_rms_layernorm, Fast_Add_RMS_Layernorm)
from .rope_embedding import fast_rope_embedding, Fast_RoPE_Embedding, Fast_RoPE_Embedding_QK
from .swiglu import swiglu_fg_kernel, swiglu_DWf_DW_dfg_kernel
from .geglu import (geglu_exact_forward_kernel, geglu_exact_backward_kernel,
                    geglu_approx_forward_kernel, geglu_approx_backward_kernel)
from .fast_lora import (get_lora_parameters, get_lora_parameters_bias, apply_lora_mlp_swiglu,
                        apply_lora_mlp_geglu_exact, apply_lora_mlp_geglu_approx, apply_lora_qkv,
                        apply_lora_o, LoRA_MLP, LoRA_QKV, LoRA_W)
from .utils import (fast_dequantize, matmul_lora, QUANT_STATE, gemm, fast_gemv,
                    fast_linear_forward)
from .attention import fast_attention, Fast_Attention, attention_forward, attention_backward
