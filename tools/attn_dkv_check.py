"""dK/dV kernel variants against each other and against fp32 torch on one (batch, head) slice.
TAVK_DKV_AUG=0 python tools/attn_dkv_check.py save; TAVK_DKV_AUG=1 python tools/attn_dkv_check.py compare"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_emotion_b200 import _lib as L  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "save"
path = "/tmp/attn_dkv_check.pt"
out = {}
for B, S in ((16, 1464), (16, 323), (2, 185), (3, 130)):
    nh, H = 12, 768
    g = torch.Generator().manual_seed(S)
    qkv = (torch.randn(B, S, 3 * H, generator=g) * 1.5).cuda().bfloat16()     # |scores| up to ~40: lse well away from 0
    do = torch.randn(B, S, H, generator=g).cuda().bfloat16()
    q, k, v = qkv[..., :H], qkv[..., H:2 * H], qkv[..., 2 * H:]
    o = torch.empty(B, S, H, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, nh, S, device="cuda")
    delta = torch.empty_like(lse)
    dqkv = torch.full((B, S, 3 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.attn_fwd(q, k, v, o, lse, B=B, S=S, nh=nh, ld_qkv=3 * H, ld_o=H)
    L.attn_bwd(q, k, v, o, do, lse, delta, dqkv[..., :H], dqkv[..., H:2 * H], dqkv[..., 2 * H:], B=B, S=S, nh=nh,
               ld_qkv=3 * H, ld_o=H, ld_dqkv=3 * H)
    torch.cuda.synchronize()
    b, h = B - 1, nh - 1
    sl = slice(h * 64, (h + 1) * 64)
    qf, kf, vf = (t[b, :, sl].float().detach().requires_grad_(True) for t in (q, k, v))
    ref = torch.softmax(qf @ kf.t() * 0.125, dim=-1) @ vf
    ref.backward(do[b, :, sl].float())
    rel = lambda a, r: ((a.double() - r.double()).norm() / r.double().norm()).item()  # noqa: E731
    print("B=%d S=%d |lse| max %.1f: dK vs fp32 %.3e, dV vs fp32 %.3e, finite %s" % (
        B, S, lse.abs().max().item(), rel(dqkv[b, :, H:2 * H][:, sl], kf.grad), rel(dqkv[b, :, 2 * H:][:, sl], vf.grad),
        bool(torch.isfinite(dqkv.float()).all())))
    out[(B, S)] = dqkv[..., H:].cpu()
if mode == "save":
    torch.save(out, path)
else:
    old = torch.load(path)
    for key, t in out.items():
        a, r = t.double(), old[key].double()
        print("%s: this run vs saved run rel-L2 %.3e, max abs %.3e (|x| max %.2f)" % (
            key, ((a - r).norm() / r.norm()).item(), (a - r).abs().max().item(), r.abs().max().item()))
print("ok")
