"""unsloth_b200 -- Blackwell-native (sm_100a) fused-kernel QLoRA fine-tuning hot path that drops
in behind Unsloth's FastLanguageModel monkey-patch surface.  See DESIGN.md / INTEGRATION.md."""
__version__ = "0.1.0"
