import torch
from flash_attn import flash_attn_func
torch.manual_seed(0)
B, S, H, Hk, D, cap = 1, 72, 4, 2, 64, 50.0
for hk in (H, Hk):
    q = torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    k = torch.randn(B, S, hk, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    v = torch.randn(B, S, hk, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    g = torch.randn(B, S, H, D, device="cuda", dtype=torch.bfloat16)
    for sc in (0.0, cap):
        o = flash_attn_func(q, k, v, causal=True, softmax_scale=D ** -0.5, softcap=sc)
        dq, dk, dv = torch.autograd.grad(o, (q, k, v), g)
        qf, kf, vf = q.float(), k.float().repeat_interleave(H // hk, 2), v.float().repeat_interleave(H // hk, 2)
        s = torch.einsum("bihd,bjhd->bhij", qf, kf) * D ** -0.5
        if sc:
            s = sc * torch.tanh(s / sc)
        i = torch.arange(S, device="cuda")[:, None]; j = torch.arange(S, device="cuda")[None, :]
        s = s.masked_fill(~(j <= i), float("-inf"))
        ref = torch.einsum("bhij,bjhd->bihd", torch.softmax(s, -1), vf)
        rq, rk, rv = torch.autograd.grad(ref, (q, k, v), g.float())
        def cs(a, b):
            a, b = a.float().flatten(), b.float().flatten()
            return (torch.dot(a, b) / (a.norm() * b.norm())).item()
        print("hk", hk, "softcap", sc, "out err", (o.float() - ref).abs().max().item(), "cos dq dk dv", cs(dq, rq), cs(dk, rk), cs(dv, rv))
