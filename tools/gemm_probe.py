"""Times (CUDA events) the step's dominant GEMM signatures one by one; the small driver used under ncu.
python tools/gemm_probe.py [name ...]     names: see SHAPES; default = all.  TAVK_PROBE_REPS=n (default 5)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_emotion_b200 import _lib as L  # noqa: E402

M = 23424
SHAPES = {  # name: (M, N, K, a_mn, b_mn, epilogue, out_bf16, bias, resid, k_splits)
    "ffn_up_gelu": (M, 3072, 768, 0, 0, L.EPI_GELU, 1, 1, 0, 1),
    "ffn_up_dgrad_gelubwd": (M, 3072, 768, 0, 1, L.EPI_GELU_BWD, 1, 0, 0, 1),
    "ffn_up_gelu_grad": (M, 3072, 768, 0, 0, L.EPI_GELU_GRAD, 1, 1, 0, 1),
    "ffn_up_dgrad_mul": (M, 3072, 768, 0, 1, L.EPI_MUL, 1, 0, 0, 1),
    "ffn_down_resid": (M, 768, 3072, 0, 0, 0, 0, 1, 1, 1),
    "ffn_down_dgrad": (M, 768, 3072, 0, 1, 0, 0, 0, 0, 1),
    "qkv": (M, 2304, 768, 0, 0, 0, 1, 1, 0, 1),
    "qkv_dgrad": (M, 768, 2304, 0, 1, 0, 0, 0, 0, 1),
    "out_proj_resid": (M, 768, 768, 0, 0, 0, 0, 1, 1, 1),
    "out_proj_dgrad_bf16": (M, 768, 768, 0, 1, 0, 1, 0, 0, 1),
    "wgrad_ffn_up": (3072, 768, M, 1, 1, 0, 0, 0, 0, 2),
    "wgrad_ffn_down": (768, 3072, M, 1, 1, 0, 0, 0, 0, 2),
    "wgrad_qkv": (2304, 768, M, 1, 1, 0, 0, 0, 0, 2),
    "wgrad_out": (768, 768, M, 1, 1, 0, 0, 0, 0, 8),
    "fusion_ffn_up_gelu": (5168, 3072, 768, 0, 0, L.EPI_GELU, 1, 1, 0, 1),
    "fusion_out_proj": (5168, 768, 768, 0, 0, 0, 0, 1, 1, 1),
    "roberta_out_proj": (1120, 768, 768, 0, 0, 0, 0, 1, 1, 1),
    "posconv": (2384, 48, 6144, 0, 0, 0, 0, 1, 0, 1),
    "roberta_ffn_down": (1120, 768, 3072, 0, 0, 0, 0, 1, 1, 1),
    "roberta_ffn_up_gelu": (1120, 3072, 768, 0, 0, L.EPI_GELU, 1, 1, 0, 1),
    "roberta_qkv": (1120, 2304, 768, 0, 0, 0, 1, 1, 0, 1),
    "w2v_ffn_down": (2384, 768, 3072, 0, 0, 0, 0, 1, 1, 1),
    "w2v_ffn_up_gelu": (2384, 3072, 768, 0, 0, L.EPI_GELU, 1, 1, 0, 1),
    "w2v_out_proj": (2384, 768, 768, 0, 0, 0, 0, 1, 1, 1),
    "fusion_ffn_down": (5168, 768, 3072, 0, 0, 0, 0, 1, 1, 1),
    "fusion_qkv": (5168, 2304, 768, 0, 0, 0, 1, 1, 0, 1),
}
BLOCK_N = int(os.environ.get("TAVK_PROBE_BLOCK_N", "0"))


def run(name, reps):
    m, n, k, a_mn, b_mn, epi, obf, hb, hr, ks = SHAPES[name]
    A = torch.randn((k, m) if a_mn else (m, k), device="cuda").bfloat16()
    B = torch.randn((k, n) if b_mn else (n, k), device="cuda").bfloat16()
    out = torch.zeros((m, n), device="cuda", dtype=torch.bfloat16 if obf else torch.float32)
    kw = dict(M=m, N=n, K=k, a_mn=bool(a_mn), b_mn=bool(b_mn), epilogue=epi, accumulate=ks > 1 or (a_mn and b_mn), k_splits=ks,
              block_n=BLOCK_N)
    if hb:
        kw["bias"] = torch.zeros(n, device="cuda")
    if hr:
        kw["resid"] = torch.zeros((m, n), device="cuda")
    if epi in (L.EPI_GELU, L.EPI_GELU_GRAD):
        kw["out2"] = torch.empty((m, n), device="cuda", dtype=torch.bfloat16)
    if epi in (L.EPI_GELU_BWD, L.EPI_MUL):
        kw["aux"] = torch.randn((m, n), device="cuda").bfloat16()
    for _ in range(2):
        L.gemm(A, B, out, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        L.gemm(A, B, out, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-24s M=%d N=%d K=%d bn=%d  %.4f ms  %.1f TFLOP/s" % (name, m, n, k, BLOCK_N, ms, 2.0 * m * n * k / ms / 1e9))


if __name__ == "__main__":
    L.require_device()
    reps = int(os.environ.get("TAVK_PROBE_REPS", "5"))
    for nm in (sys.argv[1:] or list(SHAPES)):
        run(nm, reps)
