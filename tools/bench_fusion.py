"""Times the 12-layer fusion encoder (forward + backward) alone at a BASELINE shape and prints the fraction of the
measured bf16 tensor peak (MEASURED_PEAKS.json).  Usage: python tools/bench_fusion.py [B] [S] [regime: R|none]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transformers import VideoMAEConfig  # noqa: E402

from multi_modal_emotion_b200 import _lib as L, synthetic as syn  # noqa: E402
from multi_modal_emotion_b200.tavformer import VideoMAEEncoder  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 323
    regime = sys.argv[3] if len(sys.argv) > 3 else "R"
    peak = 1651.1
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"]
    except Exception:
        pass
    enc = VideoMAEEncoder(VideoMAEConfig(), 12)
    enc.load_state_dict(syn.synth_state_dict(enc, seed=3))
    enc = enc.cuda()
    x = torch.randn(B, S, 768, device="cuda", requires_grad=True)
    mask = None
    if regime == "R":
        T, K = 70, 104
        Ta = S - T - K
        mask = syn.reference_masks(B, T, Ta, K, torch.full((B,), T), torch.full((B,), Ta)).cuda()
    flops = 3 * 12 * (14155776 * S + 3072 * S * S) * B

    def step():
        y = enc(x, mask)
        y.backward(torch.ones_like(y) / y.numel())
        enc.zero_grad(set_to_none=True)
        x.grad = None

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n0 = L.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("fusion fwd+bwd B=%d S=%d regime=%s: %.3f ms/step  %.1f samples/s  %.1f TFLOP/s  = %.1f%% of measured bf16 peak "
          "(%.0f TF/s); %d library calls/step" % (B, S, regime, ms, B / ms * 1e3, flops / ms / 1e9,
                                                 100 * flops / ms / 1e9 / peak, peak, (L.launch_count - n0) // iters))


if __name__ == "__main__":
    main()
