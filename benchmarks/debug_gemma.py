import sys, os, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_model import _reference_from, TINY
from unsloth_b200.patch import build_qlora_model, hf_config
from unsloth_b200.kernels import get_lora_parameters
name = sys.argv[1] if len(sys.argv) > 1 else "gemma-2-9b"
kw = dict(TINY, query_pre_attn_scalar=64) if "gemma" in name else dict(TINY)
for variant in ("full", "nosoftcap"):
    ex = dict(kw)
    if variant == "nosoftcap" and "gemma" in name:
        ex.update(attn_logit_softcapping=None, final_logit_softcapping=None)
    model = build_qlora_model(name, r=8, lora_alpha=16, device="cuda", num_hidden_layers=2, init_b_std=0.05, **ex)
    cfg = hf_config(name, 2, **ex)
    ref = _reference_from(model, cfg)
    torch.manual_seed(1)
    ids = torch.randint(0, kw["vocab_size"], (2, 72))
    out = model(input_ids=ids.cuda(), labels=ids.cuda()); out.loss.backward()
    ro = ref(input_ids=ids, labels=ids); ro.loss.backward()
    print(variant, "loss", out.loss.item(), ro.loss.item())
    for li, (lo, lr) in enumerate(zip(model.model.layers, ref.model.layers)):
        for po, pr in ((lo.self_attn, lr.self_attn), (lo.mlp, lr.mlp)):
            for pn in ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj"):
                if not hasattr(po, pn): continue
                _, _, A, B, _ = get_lora_parameters(getattr(po, pn))
                r = []
                for ours, theirs in ((A.grad, getattr(pr, pn).A.grad), (B.grad, getattr(pr, pn).B.grad)):
                    a, b = ours.float().cpu().flatten(), theirs.flatten()
                    r.append(round((torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item(), 4))
                    r.append(round((a.norm() / (b.norm() + 1e-30)).item(), 3))
                print(" L%d %-9s cosA %.4f ratioA %.3f cosB %.4f ratioB %.3f" % (li, pn, *r))
