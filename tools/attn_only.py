"""Attention kernels alone (for ncu / per-kernel timing).  python tools/attn_only.py [B] [S] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_emotion_b200 import _lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1464
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
nh, H = 12, 768
qkv = torch.randn(B, S, 3 * H, device="cuda").bfloat16()
q, k, v = qkv[..., :H], qkv[..., H:2 * H], qkv[..., 2 * H:]
o = torch.empty(B, S, H, device="cuda", dtype=torch.bfloat16)
do = torch.randn(B, S, H, device="cuda").bfloat16()
lse = torch.empty(B, nh, S, device="cuda")
delta = torch.empty_like(lse)
dqkv = torch.empty_like(qkv)
for _ in range(iters):
    L.attn_fwd(q, k, v, o, lse, B=B, S=S, nh=nh, ld_qkv=3 * H, ld_o=H)
    L.attn_bwd(q, k, v, o, do, lse, delta, dqkv[..., :H], dqkv[..., H:2 * H], dqkv[..., 2 * H:], B=B, S=S, nh=nh,
               ld_qkv=3 * H, ld_o=H, ld_dqkv=3 * H)
torch.cuda.synchronize()
if os.environ.get("TAVK_NO_KINETO"):      # under ncu: CUPTI has one subscriber
    for _ in range(2):
        L.attn_fwd(q, k, v, o, lse, B=B, S=S, nh=nh, ld_qkv=3 * H, ld_o=H)
        L.attn_bwd(q, k, v, o, do, lse, delta, dqkv[..., :H], dqkv[..., H:2 * H], dqkv[..., 2 * H:], B=B, S=S, nh=nh,
                   ld_qkv=3 * H, ld_o=H, ld_dqkv=3 * H)
    torch.cuda.synchronize()
    print("ok")
    sys.exit(0)
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        L.attn_fwd(q, k, v, o, lse, B=B, S=S, nh=nh, ld_qkv=3 * H, ld_o=H)
        L.attn_bwd(q, k, v, o, do, lse, delta, dqkv[..., :H], dqkv[..., H:2 * H], dqkv[..., 2 * H:], B=B, S=S, nh=nh,
                   ld_qkv=3 * H, ld_o=H, ld_dqkv=3 * H)
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total):
    if e.self_device_time_total > 0:
        print("%-60s n=%d avg %.1f us" % (e.key[:60], e.count, e.self_device_time_total / e.count))
print("ok")
