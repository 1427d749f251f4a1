"""GPU-side duration of small GEMMs (CUDA-graph of 20 launches, so CPU submission does not pace them)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_emotion_b200 import _lib as L  # noqa: E402

L.require_device()


def timed(M, N, K, bn=0, reps=20, **kw):
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            L.gemm(A, B, out, M=M, N=N, K=K, block_n=bn, **kw)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            L.gemm(A, B, out, M=M, N=N, K=K, block_n=bn, **kw)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3


for (M, N, K) in ((128, 128, 64), (128, 128, 768), (128, 128, 3072), (1120, 768, 768), (1120, 768, 3072), (1120, 3072, 768),
                  (2384, 768, 768), (5168, 768, 768), (5168, 2304, 768), (23424, 768, 768)):
    print("M=%5d N=%4d K=%4d:" % (M, N, K), "  ".join("bn%d %.1f us" % (bn, timed(M, N, K, bn)) for bn in (64, 128, 256)))
