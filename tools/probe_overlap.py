"""Checks that a TMA tensor map with overlapping rows (row pitch < row length) is accepted and loads what the strided
conv-as-GEMM needs: Conv1d(C, N, k, stride s) over channels-last x[T, C] == GEMM with A[t, :] = x_flat[t*s*C : t*s*C + k*C]."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_modal_emotion_b200 import _lib as L  # noqa: E402

L.require_device()
torch.manual_seed(0)
for (C, N, k, s, T_in) in ((64, 128, 3, 2, 1001), (512, 512, 3, 2, 9600), (512, 512, 2, 2, 600)):
    x = torch.randn(T_in, C, device="cuda").bfloat16()
    w = (torch.randn(N, C, k, device="cuda") * 0.05).bfloat16()
    T_out = (T_in - k) // s + 1
    wk = w.permute(0, 2, 1).contiguous().view(N, k * C)          # [N, (tap, c)]
    out = torch.empty(T_out, N, device="cuda")
    L.gemm(x, wk, out, M=T_out, N=N, K=k * C, lda=s * C)
    ref = torch.nn.functional.conv1d(x.float().t()[None], w.float(), stride=s)[0].t()
    err = ((out - ref).norm() / ref.norm()).item()
    print("C=%d N=%d k=%d s=%d T_in=%d -> T_out=%d rel err %.3e" % (C, N, k, s, T_in, T_out, err))
    assert err < 1e-3
print("ok")
