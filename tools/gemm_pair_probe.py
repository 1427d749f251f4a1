"""Single-CTA tiles against CTA pairs (tcgen05 cta_group::2) on the step's GEMM signatures: device-side duration of each
(20 launches in a CUDA graph, CUDA events around 5 replays) and a bitwise comparison of the two outputs.
python tools/gemm_pair_probe.py [name ...]   (names: tools/gemm_probe.py SHAPES; default = all)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from multi_modal_emotion_b200 import _lib as L  # noqa: E402
from gemm_probe import SHAPES  # noqa: E402

SHAPES = dict(SHAPES)
SHAPES.update({   # the fusion block at 128 samples per GPU (M = 128 * 323)
    "fusion128_qkv": (41344, 2304, 768, 0, 0, 0, 1, 1, 0, 1),
    "fusion128_out_proj": (41344, 768, 768, 0, 0, 0, 0, 1, 1, 1),
    "fusion128_ffn_up_gelu_grad": (41344, 3072, 768, 0, 0, L.EPI_GELU_GRAD, 1, 1, 0, 1),
    "fusion128_ffn_down": (41344, 768, 3072, 0, 0, 0, 0, 1, 1, 1),
    "fusion_wgrad_ffn_up": (3072, 768, 5168, 1, 1, 0, 0, 0, 0, 4),
    "fusion_qkv_dgrad": (5168, 768, 2304, 0, 1, 0, 0, 0, 0, 1),
    "fusion_wgrad_out": (768, 768, 5168, 1, 1, 0, 0, 0, 0, 8),
    "fusion_wgrad_qkv": (2304, 768, 5168, 1, 1, 0, 0, 0, 0, 2),
    "fusion_wgrad_ffn_down": (768, 3072, 5168, 1, 1, 0, 0, 0, 0, 2),
    "roberta_wgrad_ffn_up": (3072, 768, 1120, 1, 1, 0, 0, 0, 0, 2),
    "roberta_wgrad_out": (768, 768, 1120, 1, 1, 0, 0, 0, 0, 3),
    "w2v_wgrad_ffn_up": (3072, 768, 2384, 1, 1, 0, 0, 0, 0, 2),
    "w2v_wgrad_out": (768, 768, 2384, 1, 1, 0, 0, 0, 0, 5),
})


def timed(fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3


def run(name):
    m, n, k, a_mn, b_mn, epi, obf, hb, hr, ks = SHAPES[name]
    gen = torch.Generator(device="cuda").manual_seed(1)
    A = (torch.randn((k, m) if a_mn else (m, k), device="cuda", generator=gen) * 0.5).bfloat16()
    B = (torch.randn((k, n) if b_mn else (n, k), device="cuda", generator=gen) * 0.5).bfloat16()
    kw = dict(M=m, N=n, K=k, a_mn=bool(a_mn), b_mn=bool(b_mn), epilogue=epi, accumulate=ks > 1 or bool(a_mn and b_mn), k_splits=ks)
    if hb:
        kw["bias"] = torch.randn(n, device="cuda", generator=gen)
    if hr:
        kw["resid"] = torch.randn((m, n), device="cuda", generator=gen)
    if epi in (L.EPI_GELU, L.EPI_GELU_GRAD):
        kw["out2"] = torch.empty((m, n), device="cuda", dtype=torch.bfloat16)
    if epi in (L.EPI_GELU_BWD, L.EPI_MUL):
        kw["aux"] = torch.randn((m, n), device="cuda", generator=gen).bfloat16()
    res, outs = [], []
    for pair in (1, 2):
        for bn in (128, 256):
            out = torch.zeros((m, n), device="cuda", dtype=torch.bfloat16 if obf else torch.float32)
            us = timed(lambda: L.gemm(A, B, out, block_n=bn, cta_pair=pair, **kw))
            out.zero_()
            L.gemm(A, B, out, block_n=bn, cta_pair=pair, **kw)
            torch.cuda.synchronize()
            res.append((pair, bn, us))
            outs.append(out)
    same = all(torch.equal(outs[0], o) for o in outs[1:]) if not kw["accumulate"] else all(
        (outs[0] - o).abs().max().item() <= 1e-3 * (outs[0].abs().max().item() + 1e-6) for o in outs[1:])
    best = min(res, key=lambda r: r[2])
    fl = 2.0 * m * n * k
    print("%-28s M=%5d N=%4d K=%5d  " % (name, m, n, k) + "  ".join(
        "%s/bn%d %6.1f us %6.0f TF" % ("pair" if p == 2 else "one ", bn, us, fl / us / 1e6) for p, bn, us in res) +
        "  best=%s/bn%d  outputs %s" % ("pair" if best[0] == 2 else "one", best[1], "identical" if same else "DIFFER"), flush=True)
    return same


if __name__ == "__main__":
    L.require_device()
    ok = True
    for nm in (sys.argv[1:] or list(SHAPES)):
        ok = run(nm) and ok
    sys.exit(0 if ok else 1)
